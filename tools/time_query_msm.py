"""Time single-query MSMs alone (tables): python tools/time_query_msm.py <curve> <log2> <which[,which..]> [modes]
modes: comma list of 0 (XYZZ) / 1 (batch-affine), default "0". Prints one JSON line per (mode, query, repetition)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k = int(sys.argv[1]), int(sys.argv[2])
whiches = [int(x) for x in sys.argv[3].split(",")]
modes = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "0").split(",")]
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, curve, k, dev)
key.precompute(0, 1)
inp = bench.make_input(torch, curve, k, 5).to(dev)
ref = {}
for mode in modes:
    b.set_batch_affine(mode)
    for which in whiches:
        n = (1 << k) + 1 if which < 3 else (1 << k) - 1
        for rep in range(3):
            out = b.g_to_affine(curve, 2 if which == 2 else 1, key.msm(which, inp, n))  # (the projective form varies)
            assert ref.setdefault(which, out) == out, "modes disagree"
            print(json.dumps({"lib": os.path.basename(b.LIB_PATH), "mode": mode, "which": which, "rep": rep,
                              **{a: round(v, 3) for a, v in b.msm_phase_ms().items()}, **b.msm_last_plan()}), flush=True)
