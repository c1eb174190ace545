"""Run a few whole proofs of one curve on a synthetic key (for ncu / timing): python tools/profile_prove.py <curve> <log2> [reps] [tables]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
tables = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, curve, k, dev)
b.set_precompute(tables)
if tables:
    print("precompute s:", key.precompute(0, 1))
inp = bench.make_input(torch, curve, k, 5).to(dev)
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    proof, tm = key.prove(inp, timings=True)
    print(round((time.time() - t0) * 1e3, 2), "ms", {a: round(v, 2) for a, v in tm.items()})
    print(b.msm_phase_ms())
