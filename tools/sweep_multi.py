"""BASELINE.json config 5 on N GPUs: G1 / G2 MSM size sweep with every MSM sharded by contiguous point range over the
ranks (each rank owns the pre-shifted base table of its range; partial sums are all_gather'ed and added on rank 0), and
the Fr NTT sweep as independent replicas (a transform does not shard). Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py [--quick]
Prints JSON lines on rank 0. Times are device-side (CUDA events), max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import bench
import snark_challenge_prover_reference_b200 as b

FE = 96
MAC = {(0, 1): 620928, (0, 2): 1749888, (1, 1): 620928, (1, 2): 3612672}


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    b.check(b.lib().b200_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    quick = "--quick" in sys.argv

    def timed_max(fn, reps):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    peak = b.imad_peak()
    pk = max(peak["mad_wide_mac32_per_s"], peak["carry_chain_mac32_per_s"], peak["montgomery_mul_mac32_per_s"])
    sizes = {0: [14, 16, 18, 20] + ([] if quick else [22]), 1: [10, 12, 14, 15]}
    for curve in (0, 1):
        for lg in sizes[curve]:
            n = 1 << lg
            one = n // world
            lo, hi = rank * one, (n if rank == world - 1 else (rank + 1) * one)
            sc = bench.rand_fr(torch, n, 100 + lg)[lo:hi].contiguous().to(dev)
            for group in (1, 2):
                ab = b.affine_bytes(curve, group)
                pts = torch.empty((hi - lo) * ab, dtype=torch.uint8, device=dev)
                b.check(b.lib().b200_gen_points(curve, group, pts.data_ptr(), hi - lo, 7000003 + lo))
                ctx = b.MsmContext(curve, group, pts, hi - lo)

                def run():
                    part = ctx.run(sc)
                    if world == 1:
                        return part
                    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
                    allp = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(allp, mine)
                    if rank != 0:
                        return None
                    acc = bytes(allp[0].cpu().numpy().tobytes())
                    for t in allp[1:]:
                        acc = b.g_add(curve, group, acc, bytes(t.cpu().numpy().tobytes()))
                    return acc

                ms, res = timed_max(run, 3)
                plan = b.msm_last_plan()
                if rank == 0:
                    print(json.dumps({"kind": "msm", "n_gpus": world, "curve": b.CURVE_NAMES[curve], "group": "G%d" % group,
                                      "log2n": lg, "ms": round(ms, 3), "points_per_s": round(n / (ms / 1e3)),
                                      "c_per_rank": plan["c"], "windows": plan["windows"],
                                      "frac_of_imad_peak_algorithmic_per_gpu": round(MAC[(curve, group)] * n / world / (ms / 1e3) / pk, 3),
                                      "result_sha": __import__("hashlib").sha256(b.g_to_affine(curve, group, res)).hexdigest()[:16]}),
                          flush=True)
                ctx.close()
                del pts
            del sc
            torch.cuda.empty_cache()
    # NTT: independent replicas, one per GPU (aggregate throughput = N transforms per measured time)
    for curve, lgs in ((0, [14, 16, 18, 20, 22] + ([] if quick else [24])), (1, [10, 12, 14, 15])):
        for lg in lgs:
            m = 1 << lg
            dom = b.Domain(curve, m)
            x = bench.rand_fr(torch, m, 7).to(dev)
            for kind in ("fft", "icoset_fft"):
                fn = getattr(dom, kind)
                ms, _ = timed_max(lambda: fn(x), 5)
                if rank == 0:
                    print(json.dumps({"kind": "ntt", "n_gpus": world, "op": kind, "curve": b.CURVE_NAMES[curve], "log2m": lg,
                                      "ms": round(ms, 4), "elements_per_s_aggregate": round(world * m / (ms / 1e3)),
                                      "GBps_algorithmic_per_gpu": round(192 * m / ms / 1e6, 1),
                                      "frac_of_hbm_peak_per_gpu": round(192 * m / ms / 1e6 / 6548.2, 4),
                                      "frac_of_imad_peak_per_gpu": round(588 * m * lg / (ms / 1e3) / pk, 3)}), flush=True)
            dom.close()
            del x
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
