#!/usr/bin/env python
"""bench.py - Groth16 prover hot path (MNT4753 + MNT6753) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one MNT4753 proof at 2^20 constraints + one MNT6753 proof at 2^15 constraints (the challenge sizes of
generate_parameters.cpp:127), i.e. per proof 7 Fr NTTs (compute_H) + 5 MSMs (A, B1, L, H in G1; B2 in G2) + the
O(1) tail, on a synthetic proving key / witness of exactly that shape built on the device (see make_key()).
Metric: constraints proved per second (whole job). Latency per proof is in the extra key `proof_latency_s`.

  value : inputs (w, ca, cb, cc) already resident in HBM when the timed region starts.
  e2e   : the same step through the C-ABI call with the input image in pinned HOST memory: the H2D copy of the
          403 MB + 12.6 MB input images and the D2H of the partial sums are inside the timed region. The proving key
          stays resident, as in the reference, whose own timer starts after the key is loaded (main.cpp:203,270).
  N > 1 : every MSM is sharded by contiguous point range over the ranks (multiexp.tcc:417-431 does the same over
          OpenMP threads); the only exchange is an all_gather of 5 partial group elements (<= 2016 B) per proof over
          NCCL; compute_H is replicated. Total work is fixed => "scaling": "strong".

`--impl reference` times the UNMODIFIED reference CPU prover (oracle/_ref/main, OpenMP on all host cores) on a
bounded sample of the same workload (`generate_parameters fast`: MNT4753 2^14 + MNT6753 2^10).
"""
import argparse
import ctypes
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FE = 96
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CACHE = os.environ.get("B200_BENCH_CACHE", "/tmp/b200_bench_cache")
METRIC = "groth16_prove_constraints_per_s"
UNIT = "constraints/s"
# algorithmic work per MSM point: W=48 windows x mixed add (11 / 31 / 64 Fq mul) x 1176 MAC32
MAC32_PER_POINT = {"g1": 620928, "g2_fq2": 1749888, "g2_fq3": 3612672}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-mnt4", type=int, default=20)
    ap.add_argument("--log2-mnt6", type=int, default=15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ reference arm
def ensure_fast_params():
    """`generate_parameters fast` (the reference's own generator) -> CACHE/{MNT4753,MNT6753}-{parameters,input}"""
    os.makedirs(CACHE, exist_ok=True)
    names = ["MNT4753-parameters", "MNT4753-input", "MNT6753-parameters", "MNT6753-input"]
    if not all(os.path.exists(os.path.join(CACHE, n)) for n in names):
        gen = os.path.join(REF_DIR, "generate_parameters")
        if not os.path.exists(gen):
            return None
        subprocess.check_call([gen, "fast"], cwd=CACHE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return CACHE


def run_reference_once(cache, cores):
    """one sample step: ./main MNT4753 (2^14) + ./main MNT6753 (2^10); returns (seconds over the reference's own
    'input -> output' region, per-curve dict)"""
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    total = 0.0
    detail = {}
    for curve in ("MNT4753", "MNT6753"):
        out = subprocess.run([os.path.join(REF_DIR, "main"), curve, "compute", curve + "-parameters", curve + "-input",
                              curve + "-output-ref"], cwd=cache, env=env, capture_output=True, text=True, check=True).stdout
        m = re.search(r"Total time from input to output: : (\d+) ms", out)
        t = float(m.group(1)) / 1e3
        detail[curve] = t
        total += t
    return total, detail


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(REF_DIR, "main")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/main not built (oracle/build_ref.sh needs /root/reference)"}))
        return
    cores = os.cpu_count() or 1
    cache = ensure_fast_params()
    constraints = (1 << 14) + (1 << 10)
    for _ in range(args.warmup):
        run_reference_once(cache, cores)
    t0 = time.time()
    times = [run_reference_once(cache, cores) for _ in range(args.steps)]
    wall = time.time() - t0
    secs = sum(t for t, _ in times)
    value = constraints * args.steps / secs
    sample = ("`generate_parameters fast` (MNT4753 2^14 + MNT6753 2^10 constraints) through the unmodified libsnark "
              "main, OMP_NUM_THREADS=%d; timed region = the reference's own 'Total time from input to output'" % cores)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x24 (753-bit integers)", "data": "synthetic", "impl": "reference",
            "config": {"workload": "MNT4753 2^20 + MNT6753 2^15 Groth16 prove (7 NTT + 5 MSM each)",
                       "sample": sample, "wall_s": wall},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "proof_latency_s": times[-1][1], "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
class ClockSampler(threading.Thread):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 9:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][2]),
                "reasons": reasons, "samples": len(self.rows)}


def rand_fr(torch, n, seed, pinned=False):
    """n uniformly random 752-bit values: every value < 2^752 < r is the Montgomery representation of some element"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw = torch.randint(0, 256, (n, FE), dtype=torch.uint8, generator=g)
    raw[:, 94:] = 0
    return raw.pin_memory() if pinned else raw


def make_key(pkg, torch, curve, k, dev):
    """Synthetic proving key of the challenge's shape on the device: multiples of the generators (so every base is
    a valid curve point), with the structure observed in real keys (SURVEY.md 8 pitfalls): A has m/2 copies of one
    point and O at index m; B1/B2 have O at indices 0 and m and a duplicate pair."""
    m = 1 << k
    d = m - 1
    g1, g2 = pkg.affine_bytes(curve, 1), pkg.affine_bytes(curve, 2)

    def gen(group, n, first):
        t = torch.empty(n * pkg.affine_bytes(curve, group), dtype=torch.uint8, device=dev)
        pkg.check(pkg.lib().b200_gen_points(curve, group, t.data_ptr(), n, first))
        return t

    A, B1, B2 = gen(1, m + 1, 1000003), gen(1, m + 1, 2000003), gen(2, m + 1, 3000017)
    L, H = gen(1, m - 1, 4000037), gen(1, d, 5000011)
    Av = A.view(m + 1, g1)
    Av[2:m - 1:2] = Av[2].clone()
    Av[m - 1] = Av[2]
    Av[m] = 0
    for Q, sz in ((B1, g1), (B2, g2)):
        Qv = Q.view(m + 1, sz)
        Qv[0] = 0
        Qv[m] = 0
        Qv[m - 2] = Qv[m - 3]
    torch.cuda.synchronize()
    return pkg.Params.from_device(curve, d, m, A, B1, B2, L, H)


def regroup_partials(blobs, pbytes):
    """blobs[r] = rank r's partial sums of every proof of the step, concatenated (what one all_gather delivers);
    returns, per proof, the rank-major concatenation b200_prove_combine expects."""
    out, off = [], 0
    for nb in pbytes:
        out.append(b"".join(bl[off:off + nb] for bl in blobs))
        off += nb
    assert all(len(bl) == off for bl in blobs), "partial-sum blobs have unexpected length"
    return out


def make_input(torch, curve, k, seed):
    """pinned host image of an input file: w[m+1] (w[0] = 1 in Montgomery form), ca, cb, cc [d+1], r"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import mnt753 as M
    m = 1 << k
    n = (m + 1) + 3 * m + 1
    img = rand_fr(torch, n, seed, pinned=True)
    r = M.MOD_A if curve == 0 else M.MOD_B
    one = (M.R % r).to_bytes(FE, "little")
    img[0] = torch.frombuffer(bytearray(one), dtype=torch.uint8)
    return img


def parity_on_reference_sample(pkg, cache):
    """prove the reference-generated fast parameters on the GPU and compare sha256 with ./main's output"""
    res = {}
    for curve, name in ((0, "MNT4753"), (1, "MNT6753")):
        ref_out = os.path.join(cache, name + "-output-ref")
        if not os.path.exists(ref_out):
            return None
        key = pkg.Params.from_bytes(curve, open(os.path.join(cache, name + "-parameters"), "rb").read())
        proof = key.prove(open(os.path.join(cache, name + "-input"), "rb").read())
        key.close()
        res[name] = hashlib.sha256(proof).hexdigest() == hashlib.sha256(open(ref_out, "rb").read()).hexdigest()
    return res


def b200_arm(args):
    import torch
    import torch.distributed as dist
    import snark_challenge_prover_reference_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg.check(pkg.lib().b200_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shapes = ((0, args.log2_mnt4), (1, args.log2_mnt6))
    keys = [make_key(pkg, torch, c, k, dev) for c, k in shapes]
    # key-only preprocessing (tables of pre-shifted bases in the spare HBM), outside every timed region like the
    # reference's own key loading (main.cpp:200-203); `no_tables` below reports the same step without it
    pkg.set_precompute(True)
    preprocess_s = [key.precompute(rank, world) for key in keys]
    host_inputs = [make_input(torch, c, k, 77 + c) for c, k in shapes]
    dev_inputs = [h.to(dev) for h in host_inputs]
    constraints = sum(1 << k for _, k in shapes)
    pbytes = [pkg.partial_bytes(c) for c, _ in shapes]

    def prove_all(inputs, timings=None):
        """one step: both proofs IN FLIGHT TOGETHER (b200_prove_batch: the 2^15 MNT6753 proof runs underneath the
        2^20 MNT4753 one); N > 1: one all_gather of every rank's partial sums, rank 0 combines. Returns the proof
        bytes on rank 0."""
        t0 = time.perf_counter()
        if world == 1:
            proofs, tms = pkg.prove_batch([(keys[i], inputs[i]) for i in range(len(shapes))], timings=True)
        else:
            parts, tms = pkg.prove_batch([(keys[i], inputs[i], rank, world) for i in range(len(shapes))], timings=True)
            mine = torch.frombuffer(bytearray(b"".join(parts)), dtype=torch.uint8).to(dev)
            allp = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            proofs = []
            if rank == 0:
                blobs = [bytes(t.cpu().numpy().tobytes()) for t in allp]
                for i, ((curve, k), parts_i) in enumerate(zip(shapes, regroup_partials(blobs, pbytes))):
                    r_fr = bytes(host_inputs[i][-1].numpy().tobytes())
                    proofs.append(pkg.prove_combine(curve, parts_i, world, r_fr))
        if timings is not None:
            wall = time.perf_counter() - t0
            for tm in tms:
                # latency of this proof inside the concurrent step (its own call's wall clock); step_wall_s = both
                tm["wall_s"] = tm["total_ms"] / 1e3
                tm["step_wall_s"] = wall
                timings.append(tm)
        return proofs

    def timed(inputs, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tms = []
        barrier()
        ev0.record()
        for _ in range(steps):
            proofs = prove_all(inputs, tms)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), tms, proofs

    for _ in range(args.warmup):
        prove_all(dev_inputs)
    pkg.msm_phase_totals(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = pkg.launch_count()
    ms_dev, tms_dev, proofs_dev = timed(dev_inputs, args.steps)
    launches = pkg.launch_count() - launches0
    phases = pkg.msm_phase_totals(reset=True)
    ms_e2e, tms_e2e, proofs_e2e = timed(host_inputs, args.steps)
    pkg.set_precompute(False)
    prove_all(dev_inputs)
    ms_nt, _, proofs_nt = timed(dev_inputs, max(1, min(2, args.steps)))
    nt_steps = max(1, min(2, args.steps))
    pkg.set_precompute(True)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert proofs_dev == proofs_e2e, "device-resident and host-buffer runs disagree"
    assert proofs_dev == proofs_nt, "table and table-free MSM paths disagree"
    value = constraints * args.steps / (ms_dev / 1e3)
    e2e = constraints * args.steps / (ms_e2e / 1e3)
    h2d = sum(h.numel() for h in host_inputs)
    d2h = sum(pbytes)

    # Roofline of the dominant kernel, msm_accumulate_kernel<G1> (4 of the 5 MSMs of each proof). Inside a proof the
    # five MSMs run concurrently on five streams, so per-kernel event times overlap; the kernel is therefore timed
    # here in isolation (same key, same scalars, one MSM at a time through b200_params_msm, CUDA events around the
    # accumulation launches on the MSM's stream). Bound: the INT32 multiplier (IMAD.WIDE / fmaheavy) pipe.
    imad = pkg.imad_peak()
    peak = max(imad["mad_wide_mac32_per_s"], imad["carry_chain_mac32_per_s"])
    iso = {"g1": [], "g2": [], "a_merged": []}
    plans = {}
    if world == 1:
        for i, (curve, k) in enumerate(shapes):
            m = 1 << k
            d_w = dev_inputs[i]
            for which, n in ((0, m + 1), (1, m + 1), (3, m - 1), (4, m - 1), (2, m + 1)):
                keys[i].msm(which, d_w, n)  # warm
                keys[i].msm(which, d_w, n)
                ph = pkg.msm_phase_ms()
                # the A query is kept out of the roofline: the scalars of its m/2 equal bases are merged at run time
                # (DESIGN.md 4.2), so it accumulates half as many points as it is credited with
                iso["g2" if which == 2 else ("a_merged" if which == 0 else "g1")].append(
                    (curve, n, ph["accumulate"], ph["reduce"]))
                plans[(curve, which == 2)] = pkg.msm_last_plan()
    else:
        # sharded run: per-kernel event times of rank 0 inside the timed region (the five MSMs of a proof overlap on
        # five streams, so these over-state the kernel time; the isolated measurement is the N=1 line's)
        for i, (curve, k) in enumerate(shapes):
            n = (1 << k) // world
            for _ in range(4):
                iso["g1"].append((curve, n, phases["g1"]["accumulate"] / (8 * args.steps), phases["g1"]["reduce"] / (8 * args.steps)))
            iso["g2"].append((curve, n, phases["g2"]["accumulate"] / (2 * args.steps), phases["g2"]["reduce"] / (2 * args.steps)))
    g1_mac = sum(MAC32_PER_POINT["g1"] * n for _, n, _, _ in iso["g1"])
    acc_ms_g1 = sum(t for _, _, t, _ in iso["g1"])
    achieved = g1_mac / (acc_ms_g1 / 1e3)
    g1_big = [t for c, _, t, _ in iso["g1"] if c == 0]
    roofline = {"bound": "imad", "bound_note": "INT32 multiplier (IMAD.WIDE / fmaheavy) pipe; neither HBM nor tensor bound, SURVEY.md 8d",
                "kernel": "msm_accumulate_kernel<G1>", "achieved": achieved / 1e12, "peak": peak / 1e12,
                "unit": "TMAC32/s", "frac": achieved / peak if peak else None,
                "frac_issued": (sum(plans[(c, False)]["windows"] * 10 * 1152 * n for c, n, _, _ in iso["g1"]) / (acc_ms_g1 / 1e3) / peak
                                if plans else None),
                "issued_note": "IMAD.WIDE actually issued per point = windows x 10 multiplications x 1152, over the same time",
                "windows": {("MNT4753" if c == 0 else "MNT6753") + ("_g2" if g2 else "_g1"): v for (c, g2), v in plans.items()},
                "traffic": 14.65e9 + 15.78e9,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one 2^20-point launch, ncu --set full "
                                "(profiles/prof_accumulate_g1_r01_v3_raw.csv); algorithmic: 36 windows x 2^20 x 192 B = 7.2 GB "
                                "of table reads, the rest is per-thread stack traffic (field operands live in local memory); "
                                "0.6 TB/s, far below the HBM roofline",
                "launches": len(iso["g1"]), "avg_launch_ms": statistics.mean(g1_big) if g1_big else None,
                "a_query_equal_bases_merged_ms": [round(t, 2) for _, _, t, _ in iso["a_merged"]],
                "peak_source": "measured live by b200_imad_peak (IMAD.WIDE carry-chain microbenchmark on all SMs)",
                "timing": "kernel timed alone with CUDA events on its stream (inside a proof 5 MSMs overlap)",
                "note": "achieved uses SURVEY 8d's algorithmic 620928 MAC32/point (48 windows x 11 mul x 1176); with the "
                        "pre-shifted base tables and XYZZ additions the kernel issues 36-42 windows x 10 mul, so frac > 1 "
                        "means less work per point, not a faster pipe: ncu shows the fmaheavy pipe 84-94 % busy (profiles/)"}
    g2_mac = sum((MAC32_PER_POINT["g2_fq2"] if c == 0 else MAC32_PER_POINT["g2_fq3"]) * n for c, n, _, _ in iso["g2"])
    acc_ms_g2 = sum(t for _, _, t, _ in iso["g2"])
    roofline_g2 = {"kernel": "msm_accumulate_kernel<G2>", "achieved": g2_mac / (acc_ms_g2 / 1e3) / 1e12,
                   "peak": peak / 1e12, "unit": "TMAC32/s", "frac": (g2_mac / (acc_ms_g2 / 1e3)) / peak if peak else None,
                   "launch_ms": [round(t, 2) for _, _, t, _ in iso["g2"]],
                   "traffic": 147.5e9 + 228.1e9,
                   "traffic_note": "2^20-point Fq2 launch (profiles/prof_accumulate_g2_r01_v3_raw.csv): 14.5 GB algorithmic; the "
                                   "2.7 KB stack frames of 75 776 resident threads (205 MB) exceed the 126 MB L2, so operand "
                                   "round trips reach DRAM at 2.4 TB/s - the reason the pipe is 85 % busy here vs 94 % for G1"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # compute_H timed alone (inside a proof it is enqueued asynchronously under the MSMs): CUDA events on the default
    # stream, which is the stream ntt.cu launches on
    ch_ms = 0.0
    for i, (curve, k) in enumerate(shapes):
        m = 1 << k
        dom = pkg.Domain(curve, m)
        bufs = [dev_inputs[i][:m].clone() for _ in range(3)]
        out = torch.empty((m + 1) * FE, dtype=torch.uint8, device=dev)
        dom.compute_h(*bufs, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            dom.compute_h(*bufs, out)
        e1.record()
        torch.cuda.synchronize()
        ch_ms += e0.elapsed_time(e1)
        dom.close()
    # compute_H: 7 NTTs of m elements + pointwise ops; algorithmic HBM bytes 7*192*m + 4*96*m
    ch_bytes = args.steps * sum((7 * 192 + 4 * 96) * (1 << k) for _, k in shapes)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline_ntt = {"bound": "hbm", "kernel": "ntt_pass_kernel (compute_H: 7 NTT + pointwise)",
                    "achieved": ch_bytes / (ch_ms / 1e3) / 1e9 if ch_ms else None, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ch_bytes / (ch_ms / 1e3) / 1e9) / hbm_peak if ch_ms else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)",
                    "ms_per_step": ch_ms / args.steps,
                    "traffic": 717e6, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the three passes of ONE 2^20 "
                    "transform (ncu --set full, profiles/prof_ntt_r01_v3_raw.csv) vs 201 MB algorithmic per transform (603 MB "
                    "for three read-write sweeps); fmaheavy pipe 80 % busy",
                    "imad_frac": (args.steps * sum((7 * 588 * k + 4 * 1176) * (1 << k) for _, k in shapes) / (ch_ms / 1e3)) / peak if ch_ms else None,
                    "note": "753-bit butterflies are IMAD-bound (588 MAC32 per 96 B element-stage, 61 MAC32/byte): imad_frac is "
                            "the fraction of the measured IMAD.WIDE peak, the binding roofline; see DESIGN.md 4.3"}

    per_curve = {}
    for i, (curve, k) in enumerate(shapes):
        ts = [t for j, t in enumerate(tms_dev) if j % len(shapes) == i]
        per_curve[pkg.CURVE_NAMES[curve]] = {
            "log2_constraints": k, "latency_s": statistics.mean(t["wall_s"] for t in ts),
            "phases_ms": {key: statistics.mean(t[key] for t in ts) for key in ts[0] if key.endswith("_ms")}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x24 (753-bit integers)", "data": "synthetic", "impl": "b200",
            "config": {"workload": "MNT4753 2^%d + MNT6753 2^%d Groth16 prove (7 NTT + 5 MSM each), MSMs sharded over %d GPU(s) by point range"
                                   % (args.log2_mnt4, args.log2_mnt6, world),
                       "l2": "inputs larger than L2: each step streams >1.6 GB of bases and 416 MB of scalars",
                       "key": "synthetic multiples of the generators with the duplicate / infinity structure of real keys",
                       "key_preprocess": "pre-shifted base tables 2^(start_j)*P_i per MSM window, built once per key "
                                         "in %s s (MNT4753, MNT6753), outside the timed region; see no_tables" % json.dumps([round(x, 2) for x in preprocess_s])},
            "no_tables": {"value": constraints * nt_steps / (ms_nt / 1e3), "unit": UNIT, "ms_per_step": ms_nt / nt_steps,
                          "note": "same step with B200_PRECOMPUTE=0 (per-window buckets + host window combine)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "roofline_g2": roofline_g2,
            "roofline_ntt": roofline_ntt, "proof_pair_latency_s": ms_dev / args.steps / 1e3, "proof_latency_s": {n: v["latency_s"] for n, v in per_curve.items()},
            "per_curve": per_curve,
            "msm_points_per_s": {
                "note": "one MSM alone on one GPU, accumulate + reduce phases, points of this rank's slice",
                "g1_2^%d" % args.log2_mnt4: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g1"] if c == 0), default=None),
                "g2_fq2_2^%d" % args.log2_mnt4: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g2"] if c == 0), default=None),
                "g1_2^%d" % args.log2_mnt6: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g1"] if c == 1), default=None),
                "g2_fq3_2^%d" % args.log2_mnt6: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g2"] if c == 1), default=None)},
            "msm_phase_ms_per_step": {g: {k: v / args.steps for k, v in ph.items()} for g, ph in phases.items()},
            "imad_peak": imad}
    if world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(REF_DIR, "main")):
        cores = os.cpu_count() or 1
        cache = ensure_fast_params()
        secs, detail = run_reference_once(cache, cores)
        cval = ((1 << 14) + (1 << 10)) / secs
        line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": "`generate_parameters fast` (MNT4753 2^14 + MNT6753 2^10) through the unmodified "
                                          "libsnark main (Bos-Coster, OpenMP), one run; seconds per curve: %s" % json.dumps(detail)}
        line["parity"] = {"sha256_equal_to_reference_main_on_fast_params": parity_on_reference_sample(pkg, cache)}
    elif world == 1:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
