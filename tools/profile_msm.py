"""Run a few MSMs of one kind so that ncu can capture the kernels in isolation: python tools/profile_msm.py <curve> <group> <log2n> [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
curve, group, lg = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = 1 << lg
b.check(b.lib().b200_set_device(0))
pts = torch.empty(n * b.affine_bytes(curve, group), dtype=torch.uint8, device="cuda")
b.check(b.lib().b200_gen_points(curve, group, pts.data_ptr(), n, 12345))
g = torch.Generator(device="cpu").manual_seed(1)
sc = torch.randint(0, 256, (n, 96), dtype=torch.uint8, generator=g)
sc[:, 94:] = 0
sc = sc.cuda()
for _ in range(reps):
    b.msm(curve, group, sc, pts, n)
    print(b.msm_phase_ms())
