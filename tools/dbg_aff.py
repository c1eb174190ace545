import sys, os, random
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import torch
import snark_challenge_prover_reference_b200 as b
import mnt753 as M, util
O = util.load_oracle()
b.check(b.lib().b200_set_device(0))
for curve in (0, 1):
    c = util.curve_obj(curve)
    n = 2
    ab = b.affine_bytes(curve, 1)
    pts = torch.empty(n * ab, dtype=torch.uint8, device="cuda")
    b.check(b.lib().b200_gen_points(curve, 1, pts.data_ptr(), n, 7))
    points = b.from_device(pts)
    for scal in ((1, 0), (1, 1), (5, 3), (5, 0), (0, 3), (3, 5), (5, 5), (3, 3), (4, 3), (5, 2), (6, 3), (13, 11), (c.r - 1, 1)):
        sc = b"".join(util.fe_bytes(M.to_mont(s, c.r)) for s in scal)
        got = b.g_to_affine(curve, 1, b.msm(curve, 1, b.to_device(sc), pts, n))
        exp = util.orc_msm_affine(O, curve, 1, sc, points, n)
        print(curve, scal if scal[0] < 10**12 else "r-1", got == exp, b.msm_last_plan())
