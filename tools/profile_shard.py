"""Time one rank's share of a sharded proof on ONE GPU (emulates rank r of `world`):
python tools/profile_shard.py <curve> <log2> <world> [reps]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
key = bench.make_key(b, torch, curve, k, dev)
print("precompute s:", key.precompute(0, world))
inp = bench.make_input(torch, curve, k, 5)
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    b.prove_timeline(begin=True)
    part, tm = key.prove_partial(inp, 0, world)
    print(round((time.time() - t0) * 1e3, 2), "ms", {a: round(v, 2) for a, v in tm.items()})
    print("  last msm:", {a: round(v, 2) for a, v in b.msm_phase_ms().items()}, b.msm_last_plan())
    print("  timeline (ms since start: accumulate starts, reduce starts, reduce ends):", b.prove_timeline())
