"""CPU suite, part 3: the C-ABI library loads and exports every symbol include/b200_groth16.h declares, its host-side
math (shared formulas with the device code) agrees with the oracle, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import random
import re

import pytest

import mnt753 as M
import util


def test_every_declared_symbol_is_exported(b200):
    header = open(os.path.join(util.ROOT, "include", "b200_groth16.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(b200.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(b200.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("tag", [0, 1])
def test_host_field_ops(b200, oracle, tag):
    p = M.PRIMES["AB"[tag]]
    rng = random.Random(tag)
    for _ in range(50):
        a, b = util.fe_bytes(rng.randrange(p)), util.fe_bytes(rng.randrange(p))
        for op in (0, 1, 2):
            assert b200.host_fp_op(tag, op, a, b) == util.orc_fp(oracle, tag, op, a, b)
    a = util.fe_bytes(rng.randrange(1, p))
    assert b200.host_fp_op(tag, 3, a) == util.orc_fp(oracle, tag, 3, a)
    # three independent inversion routines (bitwise binary gcd = the device's, batched binary gcd, Fermat) agree with
    # the oracle's xgcd on random and edge values
    rng2 = random.Random(77 + tag)
    for x in [1, 2, p - 1, p - 2, (p + 1) // 2, M.R % p, 1 << 31, (1 << 64) + 1] + [rng2.randrange(1, p) for _ in range(40)]:
        xb = util.fe_bytes(x)
        want = util.orc_fp(oracle, tag, 3, xb)
        assert b200.host_fp_op(tag, 7, xb) == want
        assert b200.host_fp_op(tag, 8, xb) == want
    for x in [3, p - 5] + [rng2.randrange(1, p) for _ in range(6)]:
        xb = util.fe_bytes(x)
        assert b200.host_fp_op(tag, 6, xb) == util.orc_fp(oracle, tag, 3, xb)
    assert b200.host_fp_op(tag, 4, a) == util.orc_fp(oracle, tag, 4, a)
    assert b200.host_fp_op(tag, 5, a) == util.orc_fp(oracle, tag, 5, a)


@pytest.mark.parametrize("curve,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_host_group_ops(b200, oracle, curve, group):
    c = util.curve_obj(curve)
    rng = random.Random(curve * 2 + group)
    G = b200.g_from_affine(curve, group, util.generator_affine(curve, group))
    k = util.fe_bytes(M.to_mont(rng.randrange(c.r), c.r))
    P = b200.g_scale(curve, group, k, G)
    assert b200.g_to_affine(curve, group, P) == util.orc_to_affine(oracle, curve, group,
                                                                 util.orc_group(oracle, curve, group, 3, G, k))
    S = b200.g_add(curve, group, P, G)
    assert b200.g_to_affine(curve, group, S) == util.orc_to_affine(oracle, curve, group,
                                                                 util.orc_group(oracle, curve, group, 0, P, G))
    D = b200.g_add(curve, group, P, P)  # doubling branch
    assert b200.g_to_affine(curve, group, D) == util.orc_to_affine(oracle, curve, group,
                                                                 util.orc_group(oracle, curve, group, 1, P))
    # infinity round trip: y == 0 on the wire <-> (0:1:0)
    zero = bytes(b200.affine_bytes(curve, group))
    Z = b200.g_from_affine(curve, group, zero)
    assert b200.g_to_affine(curve, group, Z) == zero
    assert b200.g_to_affine(curve, group, b200.g_add(curve, group, Z, P)) == b200.g_to_affine(curve, group, P)


def test_combine_of_oracle_partials_reproduces_golden(b200, oracle):
    """b200_prove_combine (host tail: sum of per-rank partials, C = H + L + r*B1, to-affine) fed with partial sums
    computed by the ORACLE over two point ranges must give the reference's proof bytes."""
    for curve in (0, 1):
        params, inp, expected = util.golden(curve, 5)
        d, m, q = util.split_params(curve, params)
        x = util.split_input(inp, d, m)
        H = util.orc_compute_h(oracle, curve, d, x["ca"], x["cb"], x["cc"])
        jobs = [(1, x["w"], q["A"], m + 1), (1, x["w"], q["B1"], m + 1), (2, x["w"], q["B2"], m + 1),
                (1, H, q["H"], d), (1, x["w"][2 * 96:], q["L"], m - 1)]
        world = 2
        partials = b""
        for rank in range(world):
            for group, sc, pts, n in jobs:
                one = n // world
                lo, hi = rank * one, (n if rank == world - 1 else (rank + 1) * one)
                ab = b200.affine_bytes(curve, group)
                out = ctypes.create_string_buffer(b200.proj_bytes(curve, group))
                sb, pb = util.buf(sc[lo * 96:hi * 96]), util.buf(pts[lo * ab:hi * ab])
                oracle.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), hi - lo, ctypes.addressof(out), 1)
                partials += out.raw
        assert b200.prove_combine(curve, partials, world, x["r"]) == expected
        # ranks that multiplied their own B1 sum by r (b200_prove_partial_scaled) are combined without r
        g1p, ps = b200.proj_bytes(curve, 1), len(partials) // world
        scaled = b""
        for rank in range(world):
            part = partials[rank * ps:(rank + 1) * ps]
            scaled += part[:g1p] + b200.g_scale(curve, 1, x["r"], part[g1p:2 * g1p]) + part[2 * g1p:]
        assert b200.prove_combine(curve, scaled, world, None) == expected


def test_compute_fails_loudly_without_gpu(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error) as e:
        b200.Domain(0, 16)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(b200.B200Error):
        b200.msm(0, 1, 0, 0, 4)
    # every entry point that would compute: key loading, whole proofs, batches
    import util
    params, inp, _ = util.golden(0, 5)
    with pytest.raises(b200.B200Error):
        b200.Params.from_bytes(0, params)
    arr = (b200.ProofJob * 1)()
    assert b200.lib().b200_prove_batch(ctypes.addressof(arr), 1) != 0
    assert b"CUDA" in b200.lib().b200_last_error() or b"fallback" in b200.lib().b200_last_error()


def test_equal_bases_are_grouped_on_the_host(b200):
    """Key-load-time grouping behind the MSM's equal-base merging (host only): one big group spanning several
    1024-member segments, a pair, points at infinity that must stay ungrouped, everything else distinct."""
    import random
    rng = random.Random(5)
    pb, n = 192, 5000
    pts = [bytes(rng.randrange(256) for _ in range(pb)) for _ in range(n)]
    big = list(range(100, 3000, 1)) + [4000, 4999]
    for i in big:
        pts[i] = pts[7]
    pts[3500] = pts[3600]
    inf = pts[1][:pb // 2] + bytes(pb // 2)          # y == 0: the point at infinity, twice
    pts[1] = inf
    pts[2] = inf
    merged, groups = b200.host_equal_bases(b"".join(pts), n, pb)
    as_sets = sorted((sorted(g) for g in groups), key=len)
    assert as_sets == [[3500, 3600], sorted([7] + big)]
    assert merged == len(big) + 1
    assert all(g[0] == min(g) or g[0] in g for g in groups)
    # nothing to merge
    merged, groups = b200.host_equal_bases(b"".join(pts[3601:3700]), 99, pb)
    assert merged == 0 and groups == []


def test_jacobian_doubling_of_the_table_builder_matches_oracle(b200, oracle):
    """jac_dbl (curve.cuh: the doubling msm_precompute_kernel runs 753 times per base) against the oracle's projective
    doubling, all four groups, host instantiation of the template the device compiles too."""
    import ctypes
    import random
    import mnt753 as M
    for curve in (0, 1):
        c = util.curve_obj(curve)
        for group in (1, 2):
            rng = random.Random(50 + 2 * curve + group)
            G = b200.g_from_affine(curve, group, util.generator_affine(curve, group))
            k = util.fe_bytes(M.to_mont(rng.randrange(2, c.r), c.r))
            P = util.orc_group(oracle, curve, group, 3, G, k)
            xy = util.orc_to_affine(oracle, curve, group, P)
            for nd in (1, 2, 21):
                out = ctypes.create_string_buffer(len(xy))
                src = ctypes.create_string_buffer(xy, len(xy))
                b200.check(b200.lib().b200_host_jacobian_doublings(curve, group, ctypes.addressof(src), nd, ctypes.addressof(out)))
                Q = P
                for _ in range(nd):
                    Q = util.orc_group(oracle, curve, group, 1, Q)
                assert out.raw == util.orc_to_affine(oracle, curve, group, Q), (curve, group, nd)
