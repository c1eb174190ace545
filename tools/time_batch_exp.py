"""Throughput of the fixed-base batch exponentiation (key generation): python tools/time_batch_exp.py [log2 n]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import snark_challenge_prover_reference_b200 as b
import bench, util
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
for curve, group in ((0, 1), (0, 2), (1, 1), (1, 2)):
    n = 1 << (k if curve == 0 else min(k, 15))
    sc = bench.rand_fr(torch, n, 11).to(dev)
    out = torch.empty(n * b.affine_bytes(curve, group), dtype=torch.uint8, device=dev)
    base = util.generator_affine(curve, group)
    for rep in range(2):
        ms = b.batch_exp(curve, group, base, sc, n, out)
    tot = sum(ms.values())
    print(json.dumps({"curve": b.CURVE_NAMES[curve], "group": group, "n": n, **{a: round(v, 2) for a, v in ms.items()},
                      "total_ms": round(tot, 2), "exps_per_s": n / (tot / 1e3)}), flush=True)
