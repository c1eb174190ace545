"""CPU suite, part 4: the N>1 protocol of bench.py / b200_prove_partial on world_size 2 over gloo.
Each rank owns the contiguous point range [rank*n/world, ...) of every MSM (last rank takes the remainder, as
multiexp.tcc:417-431), the 5 partial group elements per rank are all_gather'ed, rank 0 combines them with the product's
host tail (b200_prove_combine) and must reproduce the reference's proof. The per-rank partial sums are computed by
the oracle here (no GPU on this box); on the GPU box test_gpu_parity.py::test_sharded_prover_* does the same with
the CUDA path."""
import ctypes
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util


def shard_range(n, rank, world):
    one = n // world
    return rank * one, (n if rank == world - 1 else (rank + 1) * one)


def _worker(rank, world, port, curve, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import snark_challenge_prover_reference_b200 as b200
    O = util.load_oracle()
    params, inp, expected = util.golden(curve, k)
    d, m, q = util.split_params(curve, params)
    x = util.split_input(inp, d, m)
    H = util.orc_compute_h(O, curve, d, x["ca"], x["cb"], x["cc"])  # replicated on every rank
    jobs = [(1, x["w"], q["A"], m + 1), (1, x["w"], q["B1"], m + 1), (2, x["w"], q["B2"], m + 1),
            (1, H, q["H"], d), (1, x["w"][2 * 96:], q["L"], m - 1)]
    part = b""
    for group, sc, pts, n in jobs:
        lo, hi = shard_range(n, rank, world)
        ab = b200.affine_bytes(curve, group)
        out = ctypes.create_string_buffer(b200.proj_bytes(curve, group))
        sb, pb = util.buf(sc[lo * 96:hi * 96]), util.buf(pts[lo * ab:hi * ab])
        O.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), hi - lo, ctypes.addressof(out), 1)
        part += out.raw
    assert len(part) == b200.partial_bytes(curve)
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        allp = b"".join(t.numpy().tobytes() for t in gathered)
        proof = b200.prove_combine(curve, allp, world, x["r"])
        ret["ok"] = proof == expected
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("curve", [0, 1])
def test_two_rank_sharded_proof_over_gloo(curve):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + curve + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, curve, 5, ret), nprocs=world, join=True)
    assert ret.get("ok") is True


def _oracle_partials(b200, O, curve, k, rank, world):
    params, inp, expected = util.golden(curve, k)
    d, m, q = util.split_params(curve, params)
    x = util.split_input(inp, d, m)
    H = util.orc_compute_h(O, curve, d, x["ca"], x["cb"], x["cc"])
    jobs = [(1, x["w"], q["A"], m + 1), (1, x["w"], q["B1"], m + 1), (2, x["w"], q["B2"], m + 1),
            (1, H, q["H"], d), (1, x["w"][2 * 96:], q["L"], m - 1)]
    part = b""
    for group, sc, pts, n in jobs:
        lo, hi = shard_range(n, rank, world)
        ab = b200.affine_bytes(curve, group)
        out = ctypes.create_string_buffer(b200.proj_bytes(curve, group))
        sb, pb = util.buf(sc[lo * 96:hi * 96]), util.buf(pts[lo * ab:hi * ab])
        O.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), hi - lo, ctypes.addressof(out), 1)
        part += out.raw
    return part, x["r"], expected


def _worker_step(rank, world, port, ret):
    """bench.py's N>1 step: BOTH curves' partial sums in one blob per rank, one all_gather, regroup, combine."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import snark_challenge_prover_reference_b200 as b200
    O = util.load_oracle()
    cases = [(0, 5), (1, 5)]
    parts = [_oracle_partials(b200, O, c, k, rank, world) for c, k in cases]
    mine = torch.frombuffer(bytearray(b"".join(p[0] for p in parts)), dtype=torch.uint8)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        blobs = [t.numpy().tobytes() for t in gathered]
        grouped = bench.regroup_partials(blobs, [b200.partial_bytes(c) for c, _ in cases])
        ret["ok"] = all(b200.prove_combine(c, grouped[i], world, parts[i][1]) == parts[i][2]
                        for i, (c, _) in enumerate(cases))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_with_both_curves_in_one_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29850 + (os.getpid() % 100)
    mp.spawn(_worker_step, args=(world, port, ret), nprocs=world, join=True)
    assert ret.get("ok") is True


def _worker_plan(rank, world, port, mode, ret):
    """bench.py's multi-GPU plans at world 2: the small proof WHOLE on the last rank, the large one in runs of 1/64
    slices - none on the last rank ("dedicated") or a shorter run there ("balanced"); one all_gather of fixed-size
    slots; rank 0 combines."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import snark_challenge_prover_reference_b200 as b200
    O = util.load_oracle()
    cases = [(0, 8), (1, 5)]
    pbytes = [b200.partial_bytes(c) for c, _ in cases]
    proof_len = [b200.proof_bytes(c) for c, _ in cases]
    slot = [max(a, b) for a, b in zip(pbytes, proof_len)]
    jobs = bench.rank_jobs(rank, world, mode)
    spans = bench.rank_spans(rank, world, mode)
    outs, r_fr = [], [util.golden(c, k)[1][-96:] for c, k in cases]
    for i, first, units, end in jobs:
        curve, k = cases[i]
        if i == 0 and i in spans:
            outs.append(_oracle_partials_queries(b200, O, curve, k, spans[i], units))
        elif i == 0:
            outs.append(_oracle_partials_span(b200, O, curve, k, first, end, units))
        else:
            params, inp, _ = util.golden(curve, k)
            outs.append(util.orc_prove(O, curve, params, inp))
    mine = torch.frombuffer(bench.pack_rank_blob(jobs, outs, slot), dtype=torch.uint8)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        blobs = [t.numpy().tobytes() for t in gathered]
        proofs = bench.combine_step(b200, blobs, world, slot, pbytes, proof_len, r_fr, mode)
        ret["ok"] = proofs == [util.golden(c, k)[2] for c, k in cases]
    dist.barrier()
    dist.destroy_process_group()


def _oracle_partials_span(b200, O, curve, k, first, end, units):
    """oracle partial sums over the run [first, end) of `units` slices, cut like b200_prove_partial_span (capi.cu
    query_slice): A / B1 / B2 by m+1, L aligned to the same scalars, H by d"""
    params, inp, _ = util.golden(curve, k)
    d, m, q = util.split_params(curve, params)
    x = util.split_input(inp, d, m)
    H = util.orc_compute_h(O, curve, d, x["ca"], x["cb"], x["cc"])
    cut = lambda n, r: n if r >= units else r * (n // units)
    lo1, hi1 = cut(m + 1, first), cut(m + 1, end)
    lo3, hi3 = max(lo1 - 2, 0), min(max(hi1 - 2, 0), m - 1)
    jobs = [(1, x["w"], q["A"], lo1, hi1, 0), (1, x["w"], q["B1"], lo1, hi1, 0), (2, x["w"], q["B2"], lo1, hi1, 0),
            (1, H, q["H"], cut(d, first), cut(d, end), 0), (1, x["w"], q["L"], lo3, max(hi3, lo3), 2)]
    part = b""
    for group, sc, pts, lo, hi, shift in jobs:
        ab = b200.affine_bytes(curve, group)
        out = ctypes.create_string_buffer(b200.proj_bytes(curve, group))
        sb, pb = util.buf(sc[(lo + shift) * 96:(hi + shift) * 96]), util.buf(pts[lo * ab:hi * ab])
        O.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), hi - lo, ctypes.addressof(out), 1)
        part += out.raw
    return part


def _oracle_partials_queries(b200, O, curve, k, spans, units):
    """oracle partial sums with one run of slices PER QUERY (b200_prove_partial_queries): spans in the order A, B1, B2, L,
    H; an empty run gives O; slots in the order of the partial-sum blob (A, B1, B2, H, L)"""
    params, inp, _ = util.golden(curve, k)
    d, m, q = util.split_params(curve, params)
    x = util.split_input(inp, d, m)
    H = util.orc_compute_h(O, curve, d, x["ca"], x["cb"], x["cc"])
    cut = lambda n, r: n if r >= units else r * (n // units)

    def w_range(span):
        return cut(m + 1, span[0]), cut(m + 1, span[1])
    lo3, hi3 = w_range(spans[3])
    lo3, hi3 = max(lo3 - 2, 0), min(max(hi3 - 2, 0), m - 1)
    jobs = [(1, x["w"], q["A"], *w_range(spans[0]), 0), (1, x["w"], q["B1"], *w_range(spans[1]), 0),
            (2, x["w"], q["B2"], *w_range(spans[2]), 0), (1, H, q["H"], cut(d, spans[4][0]), cut(d, spans[4][1]), 0),
            (1, x["w"], q["L"], lo3, max(hi3, lo3), 2)]
    part = b""
    for group, sc, pts, lo, hi, shift in jobs:
        ab = b200.affine_bytes(curve, group)
        if hi <= lo:
            part += b200.g_from_affine(curve, group, bytes(ab))   # O
            continue
        out = ctypes.create_string_buffer(b200.proj_bytes(curve, group))
        sb, pb = util.buf(sc[(lo + shift) * 96:(hi + shift) * 96]), util.buf(pts[lo * ab:hi * ab])
        O.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), hi - lo, ctypes.addressof(out), 1)
        part += out.raw
    return part


def test_per_query_plans_tile_every_msm():
    import bench
    U = bench.PLAN_UNITS
    for world in range(2, 9):
        spans, load = bench.query_plan(world)
        assert len(spans) == world and len(load) == world
        for qi in range(5):
            runs = sorted(sp[qi] for sp in spans if sp and sp[qi][1] > sp[qi][0])
            assert runs[0][0] == 0 and runs[-1][1] == U and all(a[1] == b[0] for a, b in zip(runs, runs[1:])), (world, qi, runs)
        mode, runs, small = bench.step_plan(world, "queries")
        assert mode == "queries" and small == world - 1
        assert [r for r in range(world) if runs[r][1] > runs[r][0]] == [r for r in range(world) if spans[r]]
        # under its own model the per-query plan must not be worse than cutting every MSM into `world` even slices
        # (the rank that also proves MNT6753 whole carries that on top)
        even = sum(bench.slice_cost_ms(q, U // world) for q in bench.QUERY_ORDER) + bench.QUERY_MODEL["rank_fixed"]
        assert max(load) <= even + bench.QUERY_MODEL["mnt6_whole"]


@pytest.mark.parametrize("mode", ["dedicated", "balanced", "queries"])
def test_two_rank_step_plans(mode):
    import bench
    U = bench.PLAN_UNITS
    for world in (2, 4, 8):
        for md in ("shard", "dedicated", "balanced"):
            m, runs, small = bench.step_plan(world, md)
            assert runs[0][0] == 0 and runs[-1][1] == U and all(runs[r][1] == runs[r + 1][0] for r in range(world - 1)), (world, md, runs)
            assert (small is None) == (md == "shard")
    assert bench.step_plan(8, "dedicated")[1][-1] == (U, U) and bench.step_plan(1)[:1] == ("shard",)
    b2 = bench.step_plan(2, "balanced")[1]
    assert 0 < b2[1][1] - b2[1][0] < b2[0][1] - b2[0][0]   # the rank that also proves MNT6753 takes the shorter run
    assert bench.rank_jobs(7, 8, "dedicated") == [(1, 0, 1, 1)] and bench.rank_jobs(1, 2, "shard") == [(0, U // 2, U, U), (1, 1, 2, 2)]
    assert bench.rank_jobs(0, 1) == [(0, 0, 1, 1), (1, 0, 1, 1)]
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29950 + (os.getpid() % 40) + {"dedicated": 0, "balanced": 41, "queries": 82}[mode]
    mp.spawn(_worker_plan, args=(world, port, mode, ret), nprocs=world, join=True)
    assert ret.get("ok") is True
