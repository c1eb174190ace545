"""Time ONE rank's share of a per-query plan on one GPU, with the phases of its last MSM:
python tools/profile_spans.py "<A>;<B1>;<B2>;<L>;<H>" [log2=20] [reps=4]     e.g. "0-0;0-0;0-19;0-0;0-0" (units of 1/64)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
spans = [tuple(int(v) for v in part.split("-")) for part in sys.argv[1].split(";")]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, 0, k, dev)
inp = bench.make_input(torch, 0, k, 5)
print("tables s:", key.precompute_queries(spans, bench.PLAN_UNITS))
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    part, tm = key.prove_partial_queries(inp, spans, bench.PLAN_UNITS, b1_scaled=True)
    print(round((time.time() - t0) * 1e3, 2), "ms", {a: round(v, 2) for a, v in tm.items()})
    print("  last msm:", {a: round(v, 2) for a, v in b.msm_phase_ms().items()}, b.msm_last_plan(), flush=True)
