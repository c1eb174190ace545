"""Proof latency vs forced window width (tables): python tools/sweep_window.py <curve> <log2> c1 c2 ..."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k = int(sys.argv[1]), int(sys.argv[2])
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
inp = bench.make_input(torch, curve, k, 5).to(dev)
for c in [int(x) for x in sys.argv[3:]]:
    b.lib().b200_msm_set_window(c)
    key = bench.make_key(b, torch, curve, k, dev)
    pre = key.precompute(0, 1)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.time()
        key.prove(inp)
        best = min(best, time.time() - t0)
    print("c=%d precompute %.1fs proof %.1f ms" % (c, pre, best * 1e3), flush=True)
    key.close()
    del key
    torch.cuda.empty_cache()
