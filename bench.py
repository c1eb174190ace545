#!/usr/bin/env python
"""bench.py - Groth16 prover hot path (MNT4753 + MNT6753) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one MNT4753 proof at 2^20 constraints + one MNT6753 proof at 2^15 constraints (the challenge sizes of
generate_parameters.cpp:127), i.e. per proof 7 Fr NTTs (compute_H) + 5 MSMs (A, B1, L, H in G1; B2 in G2) + the
O(1) tail. Both arms run on THE SAME FILES: a synthetic proving key + witness of exactly that shape, written once in
the reference's own file formats by tools/synth_key (multiples of the generators with the duplicate / infinity
structure of real keys; libsnark's ./main never validates a key), cached under B200_BENCH_CACHE.
Metric: constraints proved per second (whole job). Latency per proof is in the extra key `proof_latency_s`.

  value : inputs (w, ca, cb, cc) already resident in HBM when the timed region starts.
  e2e   : the same step through the C-ABI call with the input image in pinned HOST memory: the H2D copy of the
          403 MB + 12.6 MB input images and the D2H of the partial sums are inside the timed region. The proving key
          stays resident, as in the reference, whose own timer starts after the key is loaded (main.cpp:203,270).
  N > 1 : every MNT4753 MSM is sharded by contiguous point range over the ranks (multiexp.tcc:417-431 does the same
          over OpenMP threads); the only exchange is an all_gather of the partial group elements per step over NCCL;
          compute_H is replicated. Total work is fixed => "scaling": "strong".

`--impl reference` times the UNMODIFIED reference CPU prover (oracle/_ref/main, OpenMP on all host cores) on the same
files, i.e. the full workload, ONCE (`steps_effective`: 1 - a full-size CPU proof pair takes minutes), and leaves its
output next to the files; the b200 arm compares the sha256 of its own proofs with it (`parity`).
"""
import argparse
import csv
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
CURVES = ("MNT4753", "MNT6753")
# algorithmic work per MSM point: W=48 windows x mixed add (11 / 31 / 64 Fq mul) x 1176 MAC32
MAC32_PER_POINT = {"g1": 620928, "g2_fq2": 1749888, "g2_fq3": 3612672}
# what the kernels ISSUE per bucket insertion, in Fq multiplications of 1152 IMAD.WIDE each (24 x (24 + 24), the
# interleaved CIOS of fp_ptx_gen.cuh): XYZZ mixed addition 8M + 2S; over Fq2 a multiplication is 3 and a squaring 2
# base multiplications, over Fq3 6 and 5. Batch-affine: 5M + 1S per addition (3 of them the shared inversion).
ISSUED_MULS = {"xyzz": {"g1": 10, "g2_fq2": 8 * 3 + 2 * 2, "g2_fq3": 8 * 6 + 2 * 5},
               "affine": {"g1": 6, "g2_fq2": 5 * 3 + 2, "g2_fq3": 5 * 6 + 5}}
IMAD_PER_MUL = 1152


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-mnt4", type=int, default=20)
    ap.add_argument("--log2-mnt6", type=int, default=15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-rerun", action="store_true", help="reference arm: ignore a cached measurement of this box")
    return ap.parse_args()


def workload_name(k4, k6, world=None):
    s = "MNT4753 2^%d + MNT6753 2^%d Groth16 prove (7 NTT + 5 MSM each)" % (k4, k6)
    return s if world is None else s + ", MSMs sharded over %d GPU(s) by point range" % world


# ------------------------------------------------------------------------------------------------ workload files
def synth_tool():
    exe = os.path.join(ROOT, "tools", "_bin", "synth_key")
    src = os.path.join(ROOT, "tools", "synth_key.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-mbmi2", "-pthread", "-I",
                               os.path.join(ROOT, "snark_challenge_prover_reference_b200", "csrc"), src, "-o", exe])
    return exe


def ensure_synth(k4, k6, wait_only=False):
    """The step's files: <dir>/MNT{4,6}753-{parameters,input} in the reference's formats. Written once per box by one
    process (wait_only: another local rank writes them; poll for the marker)."""
    d = os.path.join(CACHE, "synth_k%d_k%d" % (k4, k6))
    marker = os.path.join(d, ".done")
    if wait_only:
        t0 = time.time()
        while not os.path.exists(marker):
            if time.time() - t0 > 1800:
                raise SystemExit("bench.py: timed out waiting for %s" % marker)
            time.sleep(0.5)
        return d
    if not os.path.exists(marker):
        os.makedirs(d, exist_ok=True)
        tool = synth_tool()
        for name, k in zip(CURVES, (k4, k6)):
            subprocess.check_call([tool, name, str(k), os.path.join(d, name + "-parameters"),
                                   os.path.join(d, name + "-input")])
        open(marker, "w").write("ok\n")
    return d


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference_main(d, cores, out_suffix="-output-ref"):
    """./main <curve> compute on the files of directory d, both curves; returns {"secs", "detail", "phases", "sha256"}
    over the reference's own timed region ('Total time from input to output', main.cpp:203-270: key load excluded)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    res = {"secs": 0.0, "detail": {}, "phases": {}, "sha256": {}, "load_params_s": {}}
    for curve in CURVES:
        out = subprocess.run([os.path.join(REF_DIR, "main"), curve, "compute", curve + "-parameters", curve + "-input",
                              curve + out_suffix], cwd=d, env=env, capture_output=True, text=True, check=True).stdout
        t = float(re.search(r"Total time from input to output: : (\d+) ms", out).group(1)) / 1e3
        res["detail"][curve] = t
        res["secs"] += t
        res["load_params_s"][curve] = float(re.search(r"load params: (\d+) ms", out).group(1)) / 1e3
        res["phases"][curve] = {m.group(1).strip(): float(m.group(2))
                                for m in re.finditer(r"\(leave\) ([A-Za-z0-9 ]+?)\s*\t\[([0-9.]+)s", out)
                                if "multiexp" in m.group(1) or m.group(1).strip() == "Compute the polynomial H"}
        res["sha256"][curve] = sha256_file(os.path.join(d, curve + out_suffix))
    return res


def reference_result(k4, k6, rerun=False):
    """Measure (or reuse this box's earlier measurement of) the reference prover on the step's files."""
    d = ensure_synth(k4, k6)
    path = os.path.join(d, "ref_result.json")
    if os.path.exists(path) and not rerun:
        res = json.load(open(path))
        res["cached"] = True
        return res
    cores = os.cpu_count() or 1
    t0 = time.time()
    res = run_reference_main(d, cores)
    res.update({"cores": cores, "wall_s": time.time() - t0, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                "constraints": (1 << k4) + (1 << k6), "cached": False})
    json.dump(res, open(path, "w"))
    return res


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(os.path.join(REF_DIR, "main")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/main not built (oracle/build_ref.sh needs /root/reference)"}))
        return
    k4, k6 = args.log2_mnt4, args.log2_mnt6
    res = reference_result(k4, k6, rerun=args.ref_rerun)
    value = res["constraints"] / res["secs"]
    sample = ("the FULL workload, once: the same synthetic key / witness files the b200 arm proves (tools/synth_key), "
              "through the unmodified libsnark ./main, OMP_NUM_THREADS=%d; timed region = the reference's own 'Total "
              "time from input to output' (key load excluded, main.cpp:203,270)%s" %
              (res["cores"], "; measurement reused from an earlier --impl reference call on this box (%s)" % res["when"]
               if res["cached"] else ""))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_effective": 1, "warmup": args.warmup, "warmup_effective": 0,
            "ms_per_step": 1e3 * res["secs"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x24 (753-bit integers)", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(k4, k6), "sample": sample, "wall_s": res["wall_s"],
                       "same_files_as_b200_arm": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "proof_latency_s": res["detail"], "reference_phases_s": res["phases"],
            "reference_load_params_s": res["load_params_s"], "proof_sha256": res["sha256"], "gpu_launches": 0}
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
    """Device-side twin of tools/synth_key (same bases, same structure) for the profiling tools and the sweeps, which
    need keys of many sizes without going through files."""
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


def regroup_partials(blobs, pbytes):
    """blobs[r] = rank r's partial sums of every proof of the step, concatenated (what one all_gather delivers);
    returns, per proof, the rank-major concatenation b200_prove_combine expects."""
    out, off = [], 0
    for nb in pbytes:
        out.append(b"".join(bl[off:off + nb] for bl in blobs))
        off += nb
    assert all(len(bl) == off for bl in blobs), "partial-sum blobs have unexpected length"
    return out


DEFAULT_PLAN = "queries"
PLAN_UNITS = 64   # an uneven split of the large proof is expressed in runs of 1/64 slices (b200_prove_partial_span)
# model behind the "balanced" plan, from this round's single-GPU measurements (profiles/r02_summary.md): the MNT4753
# proof costs a fixed ~34 ms (replicated compute_H, bucket reductions, preparation) plus ~336 ms x the rank's share of
# the points (370 ms whole, 81.7 ms for a 1/7 share); the whole MNT6753 proof adds ~38 ms to a rank that also works on
# MNT4753 (34 ms alone; next to MNT4753's accumulations its kernels get no free multiplier cycles)
PLAN_MODEL = {"mnt4_fixed_ms": 34.0, "mnt4_per_share_ms": 336.0, "mnt6_whole_ms": 38.0}


def step_plan(world, mode=None):
    """Who proves what in one step. Only the 2^20 MNT4753 proof is worth sharding: one rank's share of the 2^15 MNT6753
    proof would be 4096 points at N = 8, all latency. Plans:
      shard     both proofs cut into `world` equal slices (round 1);
      dedicated MNT6753 whole on the last rank, MNT4753 over the other N-1 ranks in equal slices;
      balanced  (default for N >= 2) MNT6753 whole on the last rank, which ALSO takes a smaller run of MNT4753 slices,
                sized so that all ranks finish together under PLAN_MODEL.
    Returns (mode, runs, small_rank): runs[r] = (lo, hi) in units of 1/PLAN_UNITS of the MNT4753 point ranges (lo == hi:
    the rank takes no part in it); small_rank = the rank proving MNT6753 whole, or None when it is sharded too."""
    mode = mode or os.environ.get("B200_BENCH_MNT6_MODE") or (DEFAULT_PLAN if world >= 2 else "shard")
    U = PLAN_UNITS
    if world > 1 and mode == "queries":
        spans, _ = query_plan(world)
        return "queries", [(0, sum(e - b for b, e in sp)) if sp else (0, 0) for sp in spans], world - 1
    if world == 1 or mode == "shard":
        cuts = [r * U // world for r in range(world)] + [U]
        return "shard", [(cuts[r], cuts[r + 1]) for r in range(world)], None
    if mode == "dedicated":
        cuts = [r * U // (world - 1) for r in range(world - 1)] + [U]
        return "dedicated", [(cuts[r], cuts[r + 1]) for r in range(world - 1)] + [(U, U)], world - 1
    M = PLAN_MODEL
    per_other = M["mnt4_per_share_ms"] / (world - 1)
    x = max(0.0, (per_other - M["mnt6_whole_ms"]) / (M["mnt4_per_share_ms"] + per_other))
    last = int(round(x * U))
    rest = U - last
    cuts = [r * rest // (world - 1) for r in range(world - 1)] + [rest, U]
    return "balanced", [(cuts[r], cuts[r + 1]) for r in range(world)], world - 1


# ---- per-query plan ("queries"): the five MSMs of the large proof are cut independently (b200_prove_partial_queries).
# Cutting every MSM N ways makes every GPU repeat what does not shrink with its share: the witness map (13 ms), five
# bucket reductions and preparations, and narrower windows (c = 18 instead of 21 at a 1/7 share: 17 % more bucket
# insertions per point). Here a GPU gets FEW MSMs and a large share of each: e.g. the whole H MSM (one compute_H on one
# GPU), a third of B2, ... Model, ms on one B200 at full size (profiles/r02_summary.md 8): accumulation per MSM, its
# bucket reduction, and what a slice of f costs.
QUERY_ORDER = ("A", "B1", "B2", "L", "H")          # the order of b200_params_query / of the spans
QUERY_MODEL = {"acc": {"A": 24.3, "B1": 49.3, "B2": 146.0, "L": 49.3, "H": 49.3},
               "red": {"A": 5.9, "B1": 5.9, "B2": 20.5, "L": 5.9, "H": 5.9},
               # a slice of B2 costs more than its parts (single-GPU emulations of the plans, tools/profile_plan.py and
               # profile_spans.py): ~+3 ms at 19/64 and at 40/64, 0 whole: 5 ms x (1 - f)
               "sliced_extra": {"B2": 5.0},
               "compute_h": 13.0, "prep": 1.3, "rank_fixed": 3.0, "mnt6_whole": 38.0}


def slice_cost_ms(q, units, model=None):
    """modelled time of `units`/PLAN_UNITS of MSM q on one GPU: accumulation with the window width a slice of that size
    gets (one bit narrower per halving), its bucket reduction (which shrinks with the bucket count, not below 30 %),
    one preparation; H also pays the witness map"""
    M = model or QUERY_MODEL
    if units <= 0:
        return 0.0
    import math
    f = units / PLAN_UNITS
    c = max(8, round(21 + math.log2(f)))
    windows = -(-754 // c)
    t = M["acc"][q] * f * windows / 36.0 + M["red"][q] * max(0.3, 2.0 ** (c - 21)) + M["prep"]
    t += M.get("sliced_extra", {}).get(q, 0.0) * (1.0 - f)
    return t + (M["compute_h"] if q == "H" else 0.0)


def query_plan(world, model=None):
    """per rank the runs [first, end) of PLAN_UNITS slices of A, B1, B2, L, H it sums (None: no part in the large proof),
    and the modelled per-rank times. MNT6753 runs whole on the last rank. Smallest makespan T for which a greedy fill
    works: MSMs largest first, each poured into the least-loaded GPUs up to T."""
    M = model or QUERY_MODEL
    U = PLAN_UNITS
    # largest first, H counted with its witness map so that it is placed while whole GPUs are still free (every GPU with
    # a part of H repeats compute_H); the cheapest MSMs come last and are the ones cut to even the GPUs out
    order = sorted(QUERY_ORDER, key=lambda q: -(M["acc"][q] + (M["compute_h"] if q == "H" else 0.0)))

    def fill(T):
        load = [M["rank_fixed"]] * world
        load[world - 1] += M["mnt6_whole"]
        spans = [{q: (0, 0) for q in QUERY_ORDER} for _ in range(world)]
        for q in order:
            left, at = U, 0
            for r in sorted(range(world), key=lambda r: load[r]):
                if left == 0:
                    break
                u = left
                while u > 0 and load[r] + slice_cost_ms(q, u, M) > T:
                    u -= 1
                if u == 0:
                    continue
                spans[r][q] = (at, at + u)
                load[r] += slice_cost_ms(q, u, M)
                at += u
                left -= u
            if left:
                return None
        return spans, load

    lo, hi = 0.0, 2000.0
    while hi - lo > 0.25:
        mid = (lo + hi) / 2
        if fill(mid) is None:
            lo = mid
        else:
            hi = mid
    spans, load = fill(hi)
    out = []
    for r in range(world):
        pairs = [spans[r][q] for q in QUERY_ORDER]
        out.append(pairs if any(e > b for b, e in pairs) else None)
    return out, [round(v, 1) for v in load]


def rank_jobs(rank, world, mode=None):
    """this rank's share of a step: [(proof index, first slice, number of slices the proof is cut into, end slice)] -
    the rank sums slices [first, end) of every MSM of that proof; (i, 0, 1, 1) = the whole proof"""
    mode, runs, small_rank = step_plan(world, mode)
    jobs = []
    lo, hi = runs[rank]
    if hi > lo:
        # ("queries": the runs differ per MSM - rank_spans(); the tuple only says that the rank has a part in the proof)
        jobs.append(((0, 0, PLAN_UNITS, 0) if mode == "queries" else (0, lo, PLAN_UNITS, hi)) if world > 1 else (0, 0, 1, 1))
    if small_rank is None:
        if world > 1:
            jobs.append((1, rank, world, rank + 1))
        else:
            jobs.append((1, 0, 1, 1))
    elif rank == small_rank:
        jobs.append((1, 0, 1, 1))
    return jobs


def rank_spans(rank, world, mode=None):
    """{proof index: per-query runs} for the proofs this rank shards per query (plan "queries"), else {}"""
    mode = step_plan(world, mode)[0]
    if mode != "queries":
        return {}
    sp = query_plan(world)[0][rank]
    return {0: sp} if sp else {}


def pack_rank_blob(jobs, outs, slot):
    """what one rank contributes to the step's all_gather: per proof a fixed-size slot holding its partial sums (or,
    for a proof it ran whole, the finished proof), zeros where it took no part"""
    blob = bytearray(sum(slot))
    for job, o in zip(jobs, outs):
        off = sum(slot[:job[0]])
        blob[off:off + len(o)] = o
    return blob


def combine_step(pkg, blobs, world, slot, pbytes, proof_len, r_fr, mode=None):
    """rank 0: the gathered blobs -> the step's two proofs"""
    _, runs, small_rank = step_plan(world, mode)
    parts = regroup_partials(blobs, slot)
    big_ranks = [r for r in range(world) if runs[r][1] > runs[r][0]]
    big = b"".join(parts[0][r * slot[0]:r * slot[0] + pbytes[0]] for r in big_ranks)
    proofs = [pkg.prove_combine(0, big, len(big_ranks), r_fr[0])]
    if small_rank is None:
        small = b"".join(parts[1][r * slot[1]:r * slot[1] + pbytes[1]] for r in range(world))
        proofs.append(pkg.prove_combine(1, small, world, r_fr[1]))
    else:
        proofs.append(parts[1][small_rank * slot[1]:small_rank * slot[1] + proof_len[1]])
    return proofs


def drop_in_run(files, curve_name="MNT4753"):
    """What a user of the reference's `cuda_prover_piecewise` sees: the driver binaries built over the `B::` bundle
    (the reference's own UNMODIFIED cuda_prover_piecewise.cu -> oracle/_ref/piecewise_b200, and this repo's
    csrc/host/prover_main.cpp -> bin/cuda_prover_piecewise) run as separate processes on the step's files; their own
    'Total time from input to output' (key load + table build excluded, like main.cpp:203,270) and the proof's sha256."""
    out = {}
    exes = {"reference_driver_over_b200_bundle": os.path.join(REF_DIR, "piecewise_b200"),
            "repo_driver": os.path.join(ROOT, "snark_challenge_prover_reference_b200", "bin", "cuda_prover_piecewise")}
    for name, exe in exes.items():
        if not os.path.exists(exe):
            out[name] = {"unavailable": os.path.relpath(exe, ROOT) + " not built"}
            continue
        dst = os.path.join(files, "%s-output-%s" % (curve_name, name))
        t0 = time.time()
        r = subprocess.run([exe, curve_name, "compute", os.path.join(files, curve_name + "-parameters"),
                            os.path.join(files, curve_name + "-input"), dst], capture_output=True, text=True,
                           env=dict(os.environ, B200_BUNDLE_TIMING="1"))
        wall = time.time() - t0
        if r.returncode != 0 or not os.path.exists(dst):
            out[name] = {"failed": (r.stderr or r.stdout)[-300:]}
            continue
        # (the bundle's own lines come last: the repo driver prints the same two lines around them)
        m = (re.findall(r"Total time from input to output: : (\d+) ms", r.stdout) or [None])[-1]
        lp = (re.findall(r"load params: (\d+) ms", r.stdout) or [None])[0]
        out[name] = {"input_to_output_ms": float(m) if m else None,
                     "load_params_ms": float(lp) if lp else None,
                     "process_wall_s": wall, "sha256": sha256_file(dst)}
    return out


def ncu_traffic(csv_name, kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes) of the first launch of a kernel in a committed
    `ncu --page raw --csv` export under profiles/ (row 0 names, row 1 units, then one row per launch)."""
    path = os.path.join(ROOT, "profiles", csv_name)
    try:
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        for r in rows[2:]:
            if kernel_substr in r[h.index("Kernel Name")]:
                tot = 0.0
                for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = h.index(name)
                    tot += float(r[i].replace(",", "")) * scale[units[i]]
                return {"bytes": tot, "source": "profiles/" + csv_name, "launch_ms": float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
                        if units[h.index("gpu__time_duration.sum")] == "ms" else None}
    except Exception as e:  # a missing / reshaped export must not break the bench line
        return {"bytes": None, "source": "profiles/%s unreadable: %s" % (csv_name, e)}
    return {"bytes": None, "source": "profiles/%s has no launch of %s" % (csv_name, kernel_substr)}


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

    k4, k6 = args.log2_mnt4, args.log2_mnt6
    shapes = ((0, k4), (1, k6))
    files = ensure_synth(k4, k6, wait_only=local != 0)
    mode, runs, small_rank = step_plan(world)
    my_jobs = rank_jobs(rank, world)  # (index into shapes, first slice, slices per proof, end slice)
    my_spans = rank_spans(rank, world)  # plan "queries": {0: per-query runs of the large proof}
    # the drop-in path, measured before this process takes its own 100 GB of HBM (separate processes, N = 1 only)
    drop_in = drop_in_run(files) if world == 1 and not args.no_cpu_baseline else None

    # ---- key load (B::read_params) + key-only preprocessing, outside every timed region like the reference's own key
    # loading (main.cpp:200-203); both are reported
    keys, load_ms, preprocess_s = {}, {}, {}
    pkg.set_precompute(True)
    for i, r, w, e in my_jobs:
        name = CURVES[i]
        t0 = time.perf_counter()
        keys[i] = pkg.Params.from_file(shapes[i][0], os.path.join(files, name + "-parameters"))
        load_ms[name] = dict(keys[i].load_ms(), wall=1e3 * (time.perf_counter() - t0))
        preprocess_s[name] = keys[i].precompute_queries(my_spans[i], w) if i in my_spans else keys[i].precompute(r, w, e)
    import numpy as np
    host_inputs, dev_inputs = {}, {}
    for i, _, _, _ in my_jobs:
        raw = np.fromfile(os.path.join(files, CURVES[i] + "-input"), dtype=np.uint8)
        host_inputs[i] = torch.from_numpy(raw).pin_memory()
        dev_inputs[i] = host_inputs[i].to(dev)
    r_fr = [open(os.path.join(files, CURVES[i] + "-input"), "rb").read()[-FE:] for i in range(2)]
    constraints = sum(1 << k for _, k in shapes)
    pbytes = [pkg.partial_bytes(c) for c, _ in shapes]
    proof_len = [pkg.proof_bytes(c) for c, _ in shapes]
    slot = [max(pbytes[i], proof_len[i]) for i in range(2)]   # per-rank bytes exchanged per proof

    # ---- optional: the witness map of the large proof split over three ranks (B200_BENCH_SPLIT_H=1, N >= 3). Replicated,
    # compute_H costs every rank 13 ms of multiplier time; split, the first three ranks of the MNT4753 group transform a, b
    # and c (iFFT + cosetFFT each), two 100 MB vectors travel over NVLink to the first rank, which finishes
    # (a*b - c) / Z -> icosetFFT and broadcasts the coefficients; every rank then runs its MSMs on them
    # (b200_prove_partial_ext). NCCL is the plumbing, the transforms are the same kernels.
    big_ranks = [r for r in range(world) if runs[r][1] > runs[r][0]]
    split_h = world >= 3 and len(big_ranks) >= 3 and os.environ.get("B200_BENCH_SPLIT_H", "0") == "1" and mode != "queries"
    h_state = {}
    if split_h:
        grp = dist.new_group(big_ranks)   # every rank must take part in creating it
        if rank in big_ranks:
            m4 = 1 << k4
            h_state.update(group=grp, m=m4, dom=pkg.Domain(0, m4), H=torch.zeros((m4 + 1) * FE, dtype=torch.uint8, device=dev),
                           mine=torch.empty(m4 * FE, dtype=torch.uint8, device=dev),
                           other=[torch.empty(m4 * FE, dtype=torch.uint8, device=dev) for _ in range(2)] if rank == big_ranks[0] else None)

    def witness_map_split(image):
        """-> device tensor with the H coefficients (m + 1 elements, the last one zero) on every MNT4753 rank"""
        m4, dom, H, mine = h_state["m"], h_state["dom"], h_state["H"], h_state["mine"]
        idx = big_ranks.index(rank)
        if idx < 3:  # a, b or c: elements [m+1 + idx*m, +m) of the input image (host or device)
            lo = (m4 + 1 + idx * m4) * FE
            mine.copy_(image[lo:lo + m4 * FE], non_blocking=True)
            dom.ifft(mine)
            dom.coset_fft(mine)
        if idx == 0:
            reqs = [dist.irecv(h_state["other"][j], src=big_ranks[j + 1], group=h_state["group"]) for j in range(2)]
            for rq in reqs:
                rq.wait()
            pkg.check(pkg.lib().b200_fr_muleq(0, mine.data_ptr(), h_state["other"][0].data_ptr(), m4))
            pkg.check(pkg.lib().b200_fr_subeq(0, mine.data_ptr(), h_state["other"][1].data_ptr(), m4))
            dom.divide_by_z_on_coset(mine)
            dom.icoset_fft(mine)
            H[:m4 * FE].copy_(mine)
        elif idx < 3:
            dist.send(mine, dst=big_ranks[0], group=h_state["group"])
        dist.broadcast(H, src=big_ranks[0], group=h_state["group"])
        return H

    b1_scaled = os.environ.get("B200_BENCH_SCALE_AT_COMBINE", "0") != "1"

    def prove_all(inputs, timings=None):
        """one step: this rank's proofs IN FLIGHT TOGETHER (b200_prove_batch); N > 1: one all_gather of every rank's
        partial sums (or finished small proof), rank 0 combines. Returns the two proofs' bytes on rank 0."""
        t0 = time.perf_counter()
        if world == 1:
            proofs, tms = pkg.prove_batch([(keys[i], inputs[i]) for i, _, _, _ in my_jobs], timings=True)
            busy = time.perf_counter() - t0
        else:
            h_ext = witness_map_split(inputs[0]) if split_h and rank in big_ranks else None
            jobs = [(keys[i], inputs[i], r, w, e, h_ext if i == 0 else None, my_spans.get(i)) if w > 1 else (keys[i], inputs[i])
                    for i, r, w, e in my_jobs]
            # every rank multiplies its own B1 sum by r under its GPU work (b200_prove_partial_scaled): rank 0's combine is
            # then additions and three inversions, not 753 serial doublings (B200_BENCH_SCALE_AT_COMBINE=1: the old way)
            outs, tms = pkg.prove_batch(jobs, timings=True, b1_scaled=b1_scaled) if jobs else ([], [])
            busy = time.perf_counter() - t0
            mine = torch.frombuffer(pack_rank_blob(my_jobs, outs, slot), dtype=torch.uint8).to(dev)
            allp = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            proofs = []
            if rank == 0:
                blobs = [bytes(t.cpu().numpy().tobytes()) for t in allp]
                proofs = combine_step(pkg, blobs, world, slot, pbytes, proof_len, [None, None] if b1_scaled else r_fr)
        if timings is not None:
            wall = time.perf_counter() - t0
            for (i, _, _, _), tm in zip(my_jobs, tms):
                # latency of this proof inside the concurrent step (its own call's wall clock); step_wall_s = both
                tm["wall_s"] = tm["total_ms"] / 1e3
                tm["step_wall_s"] = wall
                tm["busy_s"] = busy
                tm["curve"] = CURVES[i]
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
    nt_steps = max(1, min(2, args.steps))
    ms_nt, _, proofs_nt = timed(dev_inputs, nt_steps)
    pkg.set_precompute(True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # bytes this rank moves per e2e step: its slice of w (b200_prove_partial uploads only what its MSM slices read) plus
    # ca, cb, cc in full (compute_H is replicated); back come the partial sums or the finished proof
    def slice_len(n, r, w, e):
        cut = lambda k: n if k >= w else k * (n // w)
        return cut(e) - cut(r)

    def h2d_elements(i, r, w, e):
        m_i = 1 << shapes[i][1]
        if i not in my_spans:
            return slice_len(m_i + 1, r, w, e) + 3 * m_i + 1
        # per-query runs: the union of the w ranges of A, B1, B2, L (one copy) and, with a part of H, ca / cb / cc
        sp = my_spans[i]
        cut = lambda k: m_i + 1 if k >= w else k * ((m_i + 1) // w)
        los = [cut(b) for (b, en) in sp[:4] if en > b]
        his = [cut(en) for (b, en) in sp[:4] if en > b]
        return (max(his) - min(los) if los else 0) + (3 * m_i + 1 if sp[4][1] > sp[4][0] else 0)
    my_h2d = sum(FE * h2d_elements(i, r, w, e) for i, r, w, e in my_jobs)
    my_d2h = sum(pbytes[i] if w > 1 else proof_len[i] for i, _, w, _ in my_jobs)
    h2d_list, d2h_list = [my_h2d], [my_d2h]
    # every rank's own time inside b200_prove_batch per step (what the plan tries to equalise)
    my_busy = statistics.mean(t["busy_s"] for t in tms_dev) * 1e3 if tms_dev else 0.0
    rank_busy_ms = [my_busy]
    if world > 1:
        bt = torch.zeros(world, device=dev, dtype=torch.float64)
        bt[rank] = my_busy
        dist.all_reduce(bt)
        rank_busy_ms = [round(float(v), 2) for v in bt.tolist()]
    if world > 1:
        lt = torch.tensor([launches, my_h2d, my_d2h], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches, h2d_list, d2h_list = int(lt[0].item()), [int(lt[1].item())], [int(lt[2].item())]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    assert proofs_dev == proofs_e2e, "device-resident and host-buffer runs disagree"
    assert proofs_dev == proofs_nt, "table and table-free MSM paths disagree"
    value = constraints * args.steps / (ms_dev / 1e3)
    e2e = constraints * args.steps / (ms_e2e / 1e3)
    h2d = sum(h2d_list)
    d2h = sum(d2h_list)

    # ---- parity with the reference on the step's own files (written by --impl reference on this box)
    parity = {}
    ref_path = os.path.join(files, "ref_result.json")
    ref = json.load(open(ref_path)) if os.path.exists(ref_path) else None
    mine_sha = {CURVES[i]: hashlib.sha256(proofs_dev[i]).hexdigest() for i in range(2)}
    if ref:
        parity["sha256_equal_to_reference_main_full_size"] = {c: mine_sha[c] == ref["sha256"][c] for c in CURVES}
        parity["reference_run"] = "oracle/_ref/main on the same files, %s (--impl reference on this box)" % ref["when"]
    else:
        parity["sha256_equal_to_reference_main_full_size"] = None
        parity["reference_run"] = "no reference output for these files on this box (run `bench.py --impl reference` first)"
    parity["proof_sha256"] = mine_sha
    if drop_in:
        for name, res in drop_in.items():
            if "sha256" in res:
                res["sha256_equal_to_b200_prove"] = res.pop("sha256") == mine_sha["MNT4753"]

    # ---- roofline of the dominant kernel: the G1 bucket accumulation (4 of the 5 MSMs of each proof). Inside a proof the
    # five MSMs run concurrently on five streams, so per-kernel event times overlap; the kernel is therefore timed
    # here in isolation (same key, same scalars, one MSM at a time through b200_params_msm, CUDA events around the
    # accumulation launches on the MSM's stream). Bound: the INT32 multiplier (IMAD.WIDE / fmaheavy) pipe.
    imad = pkg.imad_peak()
    peak = max(imad["mad_wide_mac32_per_s"], imad["carry_chain_mac32_per_s"], imad["montgomery_mul_mac32_per_s"])
    lib_mode = pkg.batch_affine_mode()

    def accum_of(kind, entries):
        """which accumulation kernel the library picks (msm_affine_wins in msm.cu)"""
        if lib_mode != "auto":
            return lib_mode
        return "affine" if kind == "g2_fq2" and entries >= (20 << 20) else "xyzz"
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
    roofline = roofline_g2 = None
    if world == 1:
        acc_ms_g1 = sum(t for _, _, t, _ in iso["g1"])
        g1_alg = sum(MAC32_PER_POINT["g1"] * n for _, n, _, _ in iso["g1"])
        accum_mode = accum_of("g1", plans[(0, False)]["windows"] << k4)
        g1_issued = sum(plans[(c, False)]["windows"] * ISSUED_MULS[accum_of("g1", plans[(c, False)]["windows"] * n)]["g1"] * IMAD_PER_MUL * n
                        for c, n, _, _ in iso["g1"])
        g1_big = [t for c, _, t, _ in iso["g1"] if c == 0]
        tr = ncu_traffic("prof_accumulate_g1_r02_raw.csv", "Mnt4G1")
        roofline = {"bound": "imad", "bound_note": "INT32 multiplier (IMAD.WIDE / fmaheavy) pipe; neither HBM nor tensor bound, SURVEY.md 8d",
                    "kernel": "G1 bucket accumulation (%s)" % accum_mode,
                    "achieved": g1_issued / (acc_ms_g1 / 1e3) / 1e12, "peak": peak / 1e12, "unit": "TMAC32/s",
                    "frac": g1_issued / (acc_ms_g1 / 1e3) / peak,
                    "frac_note": "ISSUED IMAD.WIDE (windows x %d Fq multiplications per bucket insertion x 1152) / time / measured peak" % ISSUED_MULS[accum_mode]["g1"],
                    "frac_algorithmic": g1_alg / (acc_ms_g1 / 1e3) / peak,
                    "algorithmic_note": "SURVEY 8d credits 620928 MAC32 per point (48 windows x 11 mul x 1176); the kernel does the same sum "
                                        "with fewer windows (pre-shifted tables) and cheaper additions, so this ratio may exceed 1: it "
                                        "measures work avoided, not pipe speed",
                    "peak_nominal": imad["nominal_mac32_per_s"] / 1e12,
                    "windows": {CURVES[c] + ("_g2" if g2 else "_g1"): v for (c, g2), v in plans.items()},
                    "traffic": tr["bytes"], "traffic_source": tr["source"],
                    "traffic_algorithmic": plans[(0, False)]["windows"] * ((1 << k4) - 1) * 192.0,
                    "launches": len(iso["g1"]), "avg_launch_ms": statistics.mean(g1_big) if g1_big else None,
                    "a_query_equal_bases_merged_ms": [round(t, 2) for _, _, t, _ in iso["a_merged"]],
                    "peak_source": "measured live by b200_imad_peak: best of three microbenchmarks (independent IMAD.WIDE chains, the multiplier's "
                                   "carry-chain pattern, the generated Montgomery multiply in a register-resident loop), >= 50 ms runs on all SMs; "
                                   "peak_nominal = 32 IMAD.WIDE/clk/SM x SMs x max SM clock",
                    "timing": "kernel timed alone with CUDA events on its stream (inside a proof 5 MSMs overlap)"}
        acc_ms_g2 = sum(t for _, _, t, _ in iso["g2"])
        g2_alg = sum((MAC32_PER_POINT["g2_fq2"] if c == 0 else MAC32_PER_POINT["g2_fq3"]) * n for c, n, _, _ in iso["g2"])
        g2_kind = lambda c: "g2_fq2" if c == 0 else "g2_fq3"
        g2_issued = sum(plans[(c, True)]["windows"] * ISSUED_MULS[accum_of(g2_kind(c), plans[(c, True)]["windows"] * n)][g2_kind(c)] * IMAD_PER_MUL * n
                        for c, n, _, _ in iso["g2"])
        g2_mode = accum_of("g2_fq2", plans[(0, True)]["windows"] << k4)
        # G2 at 2^20 runs the batch-affine rounds (automatic mode): the committed capture is the FIRST round (half of the
        # additions); the XYZZ kernel's capture is prof_accumulate_g2_r02_raw.csv
        tr2 = ncu_traffic("prof_affine_round_g2_r02_raw.csv" if g2_mode == "affine" else "prof_accumulate_g2_r02_raw.csv", "Mnt4G2")
        roofline_g2 = {"kernel": "G2 bucket accumulation (MNT4753: %s)" % g2_mode, "achieved": g2_issued / (acc_ms_g2 / 1e3) / 1e12,
                       "peak": peak / 1e12, "unit": "TMAC32/s", "frac": g2_issued / (acc_ms_g2 / 1e3) / peak,
                       "frac_algorithmic": g2_alg / (acc_ms_g2 / 1e3) / peak,
                       "launch_ms": [round(t, 2) for _, _, t, _ in iso["g2"]],
                       "traffic": tr2["bytes"], "traffic_source": tr2["source"] + (" (first of 7 rounds: 18.8 M of the 37.7 M additions)" if g2_mode == "affine" else ""),
                       "traffic_algorithmic": plans[(0, True)]["windows"] * ((1 << k4) + 1) * 384.0}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # compute_H timed alone (inside a proof it is enqueued asynchronously under the MSMs): CUDA events on the default
    # stream, which is the stream ntt.cu launches on
    roofline_ntt = None
    if world == 1:
        ch_ms = 0.0
        for i, (curve, k) in enumerate(shapes):
            m = 1 << k
            dom = pkg.Domain(curve, m)
            bufs = [dev_inputs[i][:m * FE].clone() for _ in range(3)]
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
        trn = ncu_traffic("prof_ntt_r01_v3_raw.csv", "ntt_pass_kernel")
        roofline_ntt = {"bound": "hbm", "kernel": "ntt_pass_kernel (compute_H: 7 NTT + pointwise)",
                        "achieved": ch_bytes / (ch_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (ch_bytes / (ch_ms / 1e3) / 1e9) / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)",
                        "ms_per_step": ch_ms / args.steps,
                        "traffic_first_pass": trn["bytes"], "traffic_source": trn["source"],
                        "imad_frac": (args.steps * sum((7 * 588 * k + 4 * 1176) * (1 << k) for _, k in shapes) / (ch_ms / 1e3)) / peak,
                        "note": "753-bit butterflies are IMAD-bound (588 MAC32 per 96 B element-stage, 61 MAC32/byte): imad_frac is "
                                "the fraction of the measured IMAD.WIDE peak, the binding roofline; see DESIGN.md 4.3"}

    per_curve = {}
    for i, (curve, k) in enumerate(shapes):
        ts = [t for t in tms_dev if t["curve"] == CURVES[i]]
        if ts:
            per_curve[CURVES[i]] = {
                "log2_constraints": k, "latency_s": statistics.mean(t["wall_s"] for t in ts),
                "phases_ms": {key: statistics.mean(t[key] for t in ts) for key in ts[0] if key.endswith("_ms")}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x24 (753-bit integers)", "data": "synthetic", "impl": "b200",
            "config": {"workload": workload_name(k4, k6, world),
                       "files": "tools/synth_key output in the reference's formats; the reference arm proves the same files",
                       "l2": "inputs larger than L2: each step streams >1.6 GB of bases and 416 MB of scalars",
                       "key": "synthetic multiples of the generators with the duplicate / infinity structure of real keys",
                       "multi_gpu": {"mode": mode, "mnt4753_runs_of_%d" % PLAN_UNITS: None if mode == "queries" else runs, "mnt6753_rank": small_rank,
                                     "rank_busy_ms_per_step": rank_busy_ms,
                                     "per_query_runs_A_B1_B2_L_H": query_plan(world)[0] if mode == "queries" else None,
                                     "witness_map": ("on the ranks with a part of H" if mode == "queries" else
                                                     "split over 3 ranks + broadcast" if split_h else "replicated")},
                       "accumulation": lib_mode + " (auto = batched affine additions for large G2/Fq2 MSMs, XYZZ mixed additions otherwise)",
                       "key_preprocess": "pre-shifted base tables 2^(start_j)*P_i per MSM window, built once per key, "
                                         "outside the timed region (see key_load); see no_tables"},
            "key_load": {"from_file_ms": load_ms, "precompute_s": preprocess_s,
                         "note": "rank 0's keys; B::read_params = chunked pinned reads + async H2D (reference: 8.0 s parse "
                                 "for the MNT4753 key, BASELINE.md 2); precompute = base tables + equal-base grouping"},
            "no_tables": {"value": constraints * nt_steps / (ms_nt / 1e3), "unit": UNIT, "ms_per_step": ms_nt / nt_steps,
                          "note": "same step with B200_PRECOMPUTE=0 (per-window buckets + host window combine)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "roofline_g2": roofline_g2,
            "roofline_ntt": roofline_ntt, "proof_pair_latency_s": ms_dev / args.steps / 1e3,
            "proof_latency_s": {n: v["latency_s"] for n, v in per_curve.items()},
            "per_curve": per_curve, "parity": parity,
            "drop_in": {"MNT4753": drop_in, "note": "the B:: bundle path driven by cuda_prover_piecewise (asynchronous multiexps resolved at "
                        "G1_scale / G1_add / groth16_output_write), full size, separate processes"} if drop_in else None,
            "msm_phase_ms_per_step": {g: {k: v / args.steps for k, v in ph.items()} for g, ph in phases.items()},
            "imad_peak": imad}
    if world == 1:
        line["msm_points_per_s"] = {
            "note": "one MSM alone on one GPU, accumulate + reduce phases",
            "g1_2^%d" % k4: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g1"] if c == 0), default=None),
            "g2_fq2_2^%d" % k4: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g2"] if c == 0), default=None),
            "g1_2^%d" % k6: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g1"] if c == 1), default=None),
            "g2_fq3_2^%d" % k6: max((n / ((a + r) / 1e3) for c, n, a, r in iso["g2"] if c == 1), default=None)}
    if world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(REF_DIR, "main")):
        if ref:
            cval = ref["constraints"] / ref["secs"]
            line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": ref["cores"], "kind": "reference",
                                    "sample": "the full workload (same files), measured once by `bench.py --impl reference` on this "
                                              "box at %s; seconds per curve: %s" % (ref["when"], json.dumps(ref["detail"]))}
        else:
            # no full-size reference run on this box yet: a bounded sample, the same generator at 1/16 of the size
            ks = (max(k4 - 4, 4), max(k6 - 4, 4))
            sd = ensure_synth(*ks)
            cores = os.cpu_count() or 1
            res = run_reference_main(sd, cores)
            line["cpu_baseline"] = {"value": ((1 << ks[0]) + (1 << ks[1])) / res["secs"], "unit": UNIT, "cores": cores,
                                    "kind": "reference",
                                    "sample": "BOUNDED SAMPLE: MNT4753 2^%d + MNT6753 2^%d (1/16 of the workload, same generator) through "
                                              "the unmodified libsnark main, one run; `bench.py --impl reference` measures the full "
                                              "workload; seconds per curve: %s" % (ks[0], ks[1], json.dumps(res["detail"]))}
            # and the sample's proofs must agree too
            ok = {}
            for i, name in enumerate(CURVES):
                key = pkg.Params.from_file(i, os.path.join(sd, name + "-parameters"))
                proof = key.prove(open(os.path.join(sd, name + "-input"), "rb").read())
                key.close()
                ok[name] = hashlib.sha256(proof).hexdigest() == res["sha256"][name]
            parity["sha256_equal_to_reference_main_on_sample"] = ok
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
