"""Time one query's MSM alone (tables): python tools/time_query_msm.py <curve> <log2> <which>"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k, which = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, curve, k, dev)
key.precompute(0, 1)
inp = bench.make_input(torch, curve, k, 5).to(dev)
n = (1 << k) + 1 if which < 3 else (1 << k) - 1
for _ in range(3):
    key.msm(which, inp, n)
    print({a: round(v, 2) for a, v in b.msm_phase_ms().items()}, b.msm_last_plan(), flush=True)
