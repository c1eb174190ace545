"""Emulate the ranks of a multi-GPU plan on ONE GPU: every rank's share of the MNT4753 proof is run alone, timed, and the
partial sums of all ranks are combined and compared with the unsharded proof.
python tools/profile_plan.py <world> [mode=queries|balanced] [log2=20] [reps=3]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
world = int(sys.argv[1])
mode = sys.argv[2] if len(sys.argv) > 2 else "queries"
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, 0, k, dev)
inp = bench.make_input(torch, 0, k, 5)
U = bench.PLAN_UNITS
if mode == "queries":
    spans, model = bench.query_plan(world)
else:
    _, runs, _ = bench.step_plan(world, mode)
    spans, model = [[(lo, hi)] * 5 if hi > lo else None for lo, hi in runs], [None] * world
parts, n_ranks = b"", 0
for r, sp in enumerate(spans):
    if not sp:
        print("rank", r, "no part in MNT4753")
        continue
    t0 = time.time()
    key.precompute_queries(sp, U)
    pre = time.time() - t0
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.time()
        part, tm = key.prove_partial_queries(inp, sp, U, b1_scaled=True)
        ms = (time.time() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    parts += part
    n_ranks += 1
    print("rank", r, "A/B1/B2/L/H", sp, "best %.1f ms" % best, "(model %s)" % model[r], "tables %.1f s" % pre, flush=True)
proof = b.prove_combine(0, parts, n_ranks, None)
key.precompute(0, 1)
whole = key.prove(inp)
print("combined partial sums == unsharded proof:", proof == whole)
assert proof == whole
