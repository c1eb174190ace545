"""Run a few Fr NTTs of one size (for ncu / timing): python tools/profile_ntt.py <curve> <log2> [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import snark_challenge_prover_reference_b200 as b
import bench
curve, k = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b.check(b.lib().b200_set_device(0))
dev = torch.device("cuda", 0)
m = 1 << k
dom = b.Domain(curve, m)
x = bench.rand_fr(torch, m, 7).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dom.fft(x)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    dom.fft(x)
e1.record()
torch.cuda.synchronize()
print("fft 2^%d: %.3f ms" % (k, e0.elapsed_time(e1) / reps))
dom.icoset_fft(x)
torch.cuda.synchronize()
