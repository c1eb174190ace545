"""Quick on-GPU probe: IMAD peak microbenchmark, MSM / NTT timings by size (not a bench line; development aid)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import snark_challenge_prover_reference_b200 as b

FE = 96


def rand_fr(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw = torch.randint(0, 256, (n, FE), dtype=torch.uint8, generator=g)
    raw[:, 94:] = 0
    return raw.cuda()


def main():
    out = {}
    b.check(b.lib().b200_set_device(0))
    out["imad"] = b.imad_peak()
    print(json.dumps(out["imad"]), flush=True)
    sizes = [int(x) for x in sys.argv[1:] if not x.startswith("--")] or [14, 16, 18, 20]
    for curve, group in ((0, 1), (0, 2), (1, 1), (1, 2)):
        for lg in sizes:
            if curve == 1 and lg > 18:
                continue
            if group == 2 and lg > 18 and "--big-g2" not in sys.argv:
                continue
            n = 1 << lg
            ab = b.affine_bytes(curve, group)
            pts = torch.empty(n * ab, dtype=torch.uint8, device="cuda")
            t0 = time.time()
            b.check(b.lib().b200_gen_points(curve, group, pts.data_ptr(), n, 12345))
            torch.cuda.synchronize()
            tgen = time.time() - t0
            sc = rand_fr(n, lg)
            best = None
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.time()
                b.msm(curve, group, sc, pts, n)
                dt = time.time() - t0
                ph = b.msm_phase_ms()
                if best is None or dt < best[0]:
                    best = (dt, ph)
            rec = {"curve": curve, "group": group, "log2n": lg, "gen_s": round(tgen, 3), "msm_ms": round(best[0] * 1e3, 2),
                   "points_per_s": round(n / best[0]), "phases": {k: round(v, 2) for k, v in best[1].items()}}
            print(json.dumps(rec), flush=True)
            del pts, sc
    for curve, lg in ((0, 14), (0, 20), (1, 15)):
        m = 1 << lg
        dom = b.Domain(curve, m)
        x = rand_fr(m, 5)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for kind in ("fft", "ifft", "coset_fft", "icoset_fft"):
            fn = getattr(dom, kind)
            fn(x)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(5):
                fn(x)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 5
            print(json.dumps({"ntt": kind, "curve": curve, "log2m": lg, "ms": round(ms, 3),
                              "GBps_algorithmic": round(192 * m / ms / 1e6, 1)}), flush=True)
        ca, cb, cc = rand_fr(m, 1), rand_fr(m, 2), rand_fr(m, 3)
        o = torch.empty((m + 1) * FE, dtype=torch.uint8, device="cuda")
        dom.compute_h(ca, cb, cc, o)
        torch.cuda.synchronize()
        ev0.record()
        dom.compute_h(ca, cb, cc, o)
        ev1.record()
        torch.cuda.synchronize()
        print(json.dumps({"compute_h_ms": round(ev0.elapsed_time(ev1), 3), "curve": curve, "log2m": lg}), flush=True)
        dom.close()


if __name__ == "__main__":
    main()
