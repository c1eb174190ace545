"""Accumulate / reduce time of one table MSM under forced window widths: python tools/window_probe.py <curve> <log2> <query> c1 c2 ..."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k, which = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
inp = bench.make_input(torch, curve, k, 5).to(dev)
n = (1 << k) + 1 if which < 3 else (1 << k) - 1
for c in [int(x) for x in sys.argv[4:]]:
    b.lib().b200_msm_set_window(c)
    key = bench.make_key(b, torch, curve, k, dev)
    key.precompute(0, 1)
    for _ in range(2):
        key.msm(which, inp, n)
    print(c, {a: round(v, 2) for a, v in b.msm_phase_ms().items()}, b.msm_last_plan(), flush=True)
    key.close()
    del key
    torch.cuda.empty_cache()
