"""BASELINE.json config 5: MSM size sweep (G1, G2/Fq2, G2/Fq3; with and without the pre-shifted base tables) and Fr NTT
size sweep on one B200. Prints JSON lines; `--md` also writes a markdown table (profiles/sweep_r01.md).
Timing: CUDA-event phase times of the library for MSMs (one MSM alone), torch CUDA events for NTTs (5 runs, mean)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
import snark_challenge_prover_reference_b200 as b

FE = 96
MAC = {(0, 1): 620928, (0, 2): 1749888, (1, 1): 620928, (1, 2): 3612672}


def main():
    b.check(b.lib().b200_set_device(0))
    dev = torch.device("cuda", 0)
    peak = b.imad_peak()
    pk = max(peak["mad_wide_mac32_per_s"], peak["carry_chain_mac32_per_s"])
    rows = []
    msm_sizes = {0: [14, 16, 18, 20, 22], 1: [10, 12, 14, 15]}  # MNT6753 keys stop at 2^15 (Fr 2-adicity)
    for curve in (0, 1):
        for lg in msm_sizes[curve]:
            key = bench.make_key(b, torch, curve, lg, dev)
            n = (1 << lg) + 1
            sc = bench.rand_fr(torch, n, 100 + lg).to(dev)
            for tables in (False, True):
                b.set_precompute(tables)
                try:
                    if tables and lg >= 22:
                        raise b.B200Error("36 x 4.8 GB = 174 GB for the five queries at 2^22 exceed one GPU's HBM")
                    pre = key.precompute(0, 1) if tables else 0.0
                except b.B200Error as e:  # tables of a FULL key need 36 x the key: 174 GB at 2^22
                    print(json.dumps({"kind": "msm", "curve": b.CURVE_NAMES[curve], "log2n": lg, "tables": True,
                                      "skipped": "tables for all five queries do not fit: " + str(e)[-60:]}), flush=True)
                    continue
                for which, group in ((1, 1), (2, 2)):  # B1 and B2 queries (the A query's equal bases would be merged)
                    best = None
                    for _ in range(3):
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        key.msm(which, sc, n)
                        wall = (time.perf_counter() - t0) * 1e3
                        ph = b.msm_phase_ms()
                        if best is None or wall < best[0]:
                            best = (wall, ph, b.msm_last_plan())
                    wall, ph, plan = best
                    rec = {"kind": "msm", "curve": b.CURVE_NAMES[curve], "group": "G%d" % group, "log2n": lg, "tables": tables,
                           "ms": round(wall, 3), "points_per_s": round(n / (wall / 1e3)), "accumulate_ms": round(ph["accumulate"], 3),
                           "reduce_ms": round(ph["reduce"], 3), "host_tail_ms": round(ph["host_tail"], 3), "c": plan["c"],
                           "windows": plan["windows"],
                           "frac_of_imad_peak_algorithmic": round(MAC[(curve, group)] * n / (ph["accumulate"] / 1e3) / pk, 3),
                           "table_build_s": round(pre, 2)}
                    rows.append(rec)
                    print(json.dumps(rec), flush=True)
            b.set_precompute(True)
            key.close()
            del key, sc
            torch.cuda.empty_cache()
    for curve, sizes in ((0, list(range(14, 25, 2))), (1, [10, 12, 14, 15])):
        for lg in sizes:
            m = 1 << lg
            dom = b.Domain(curve, m)
            x = bench.rand_fr(torch, m, 7).to(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for kind in ("fft", "icoset_fft"):
                fn = getattr(dom, kind)
                fn(x)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    fn(x)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                rec = {"kind": "ntt", "op": kind, "curve": b.CURVE_NAMES[curve], "log2m": lg, "ms": round(ms, 4),
                       "elements_per_s": round(m / (ms / 1e3)), "GBps_algorithmic": round(192 * m / ms / 1e6, 1),
                       "frac_of_hbm_peak": round(192 * m / ms / 1e6 / 6548.2, 4),
                       "frac_of_imad_peak": round(588 * m * lg / (ms / 1e3) / pk, 3)}
                rows.append(rec)
                print(json.dumps(rec), flush=True)
            dom.close()
            del x
            torch.cuda.empty_cache()
    if "--md" in sys.argv:
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "sweep_r01.md")
        with open(out, "w") as f:
            f.write("# Size sweeps on 1 x B200 (BASELINE.json config 5)\n\n")
            f.write("IMAD.WIDE peak measured in this run: %.2f TMAC32/s. `frac (alg.)` = SURVEY 8d algorithmic MAC32 / accumulate time / peak "
                    "(> 1 with tables: fewer windows and 10 instead of 11 multiplications per addition).\n\n" % (pk / 1e12))
            f.write("| curve | group | log2 n | tables | c | windows | MSM ms | points/s | accumulate ms | reduce ms | host tail ms | frac (alg.) |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                if r["kind"] == "msm":
                    f.write("| %s | %s | %d | %s | %d | %d | %.2f | %.3g | %.2f | %.2f | %.2f | %.2f |\n" % (
                        r["curve"], r["group"], r["log2n"], "yes" if r["tables"] else "no", r["c"], r["windows"], r["ms"],
                        r["points_per_s"], r["accumulate_ms"], r["reduce_ms"], r["host_tail_ms"], r["frac_of_imad_peak_algorithmic"]))
            f.write("\n| curve (Fr) | op | log2 m | ms | elements/s | algorithmic GB/s (192 B/element) | frac of HBM peak (6548 GB/s) | frac of IMAD peak |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                if r["kind"] == "ntt":
                    f.write("| %s | %s | %d | %.3f | %.3g | %.1f | %.4f | %.2f |\n" % (r["curve"], r["op"], r["log2m"], r["ms"],
                                                                                  r["elements_per_s"], r["GBps_algorithmic"], r["frac_of_hbm_peak"], r["frac_of_imad_peak"]))


if __name__ == "__main__":
    main()
