"""GPU suite (-m gpu): every CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs,
against the committed golden proofs of the reference, and - at sizes the oracle cannot reach - through
size-independent properties (linearity of the MSM, transform round trips). Integer work: the bar is bit-exact."""
import ctypes
import random

import pytest

import mnt753 as M
import util

pytestmark = pytest.mark.gpu
FE = 96


@pytest.fixture(scope="module")
def dev(b200):
    import torch
    assert torch.cuda.is_available()
    b200.check(b200.lib().b200_set_device(0))
    return torch.device("cuda:0")


def _edge_values(p):
    return [0, 1, 2, p - 1, p - 2, (1 << 752) % p, M.R % p, (p + 1) // 2, 0xFFFFFFFF, 1 << 32, (1 << 736) - 1]


# ------------------------------------------------------------------------------------------------ field layer
@pytest.mark.parametrize("tag", [0, 1])
def test_fp_ops_vs_python_ints(b200, dev, tag):
    p = M.PRIMES["AB"[tag]]
    rng = random.Random(1000 + tag)
    ev = _edge_values(p)
    xs = [a for a in ev for _ in ev] + [rng.randrange(p) for _ in range(3000)]
    ys = [b for _ in ev for b in ev] + [rng.randrange(p) for _ in range(3000)]
    n = len(xs)
    da = b200.to_device(b"".join(util.fe_bytes(v) for v in xs))
    db = b200.to_device(b"".join(util.fe_bytes(v) for v in ys))
    import torch
    dr = torch.empty(n * FE, dtype=torch.uint8, device=dev)
    rinv = pow(M.R, -1, p)
    expect = {0: lambda a, b: (a + b) % p, 1: lambda a, b: (a - b) % p, 2: lambda a, b: a * b * rinv % p,
              3: lambda a, b: a * a * rinv % p, 4: lambda a, b: a * rinv % p, 5: lambda a, b: a * M.R % p}
    for op, f in expect.items():
        b200.check(b200.lib().b200_dev_fp_op(tag, op, da.data_ptr(), db.data_ptr(), dr.data_ptr(), n))
        out = b200.from_device(dr)
        for i in range(n):
            assert util.fe_int(out[i * FE:(i + 1) * FE]) == f(xs[i], ys[i]), (op, i)
    # inversion on a few elements
    k = 64
    b200.check(b200.lib().b200_dev_fp_op(tag, 6, da.data_ptr() + 20 * FE * 11, None, dr.data_ptr(), k))
    out = b200.from_device(dr)
    for i in range(k):
        a = xs[220 + i]
        if a:
            assert M.from_mont(util.fe_int(out[i * FE:(i + 1) * FE]), p) * M.from_mont(a, p) % p == 1


@pytest.mark.parametrize("curve", [0, 1])
def test_tower_ops_vs_oracle(b200, oracle, dev, curve):
    import torch
    c = util.curve_obj(curve)
    d = c.ext_deg
    rng = random.Random(2000 + curve)
    n = 96
    a = b"".join(util.rand_fe_bytes(rng, c.q, d) for _ in range(n))
    bb = b"".join(util.rand_fe_bytes(rng, c.q, d) for _ in range(n))
    da, db = b200.to_device(a), b200.to_device(bb)
    dr = torch.empty(n * d * FE, dtype=torch.uint8, device=dev)
    for op in (0, 1, 2, 3, 4):  # add sub mul sqr inv (Fp2::inv / Fp3::inv called directly, not through to_affine)
        b200.check(b200.lib().b200_dev_fqe_op(curve, op, da.data_ptr(), db.data_ptr(), dr.data_ptr(), n))
        out = b200.from_device(dr)
        sz = d * FE
        for i in range(n):
            exp = util.orc_fqe(oracle, curve, 2, op, a[i * sz:(i + 1) * sz], bb[i * sz:(i + 1) * sz])
            assert out[i * sz:(i + 1) * sz] == exp, (op, i)


# ------------------------------------------------------------------------------------------------ group layer
def _some_points(b200, oracle, curve, group, n, seed):
    """n projective points (random multiples of the generator, randomised Z) + their affine forms, via the oracle"""
    c = util.curve_obj(curve)
    rng = random.Random(seed)
    G = b200.g_from_affine(curve, group, util.generator_affine(curve, group))
    proj, aff = [], []
    for _ in range(n):
        k = util.fe_bytes(M.to_mont(rng.randrange(1, c.r), c.r))
        P = util.orc_group(oracle, curve, group, 3, G, k)
        proj.append(P)
        aff.append(util.orc_to_affine(oracle, curve, group, P))
    return proj, aff


@pytest.mark.parametrize("curve,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_group_ops_vs_oracle(b200, oracle, dev, curve, group):
    import torch
    proj, aff = _some_points(b200, oracle, curve, group, 12, 3000 + 10 * curve + group)
    pb, ab = b200.proj_bytes(curve, group), b200.affine_bytes(curve, group)
    zero_p = b200.g_from_affine(curve, group, bytes(ab))
    neg = lambda P: util.orc_group(oracle, curve, group, 0, zero_p, P)  # placeholder, replaced below
    # build operand lists with the special cases the reference handles: O+P, P+O, P+P, P+(-P), O+O
    negs = []
    for a in aff:
        d = util.deg(curve, group)
        x, y = a[:d * FE], a[d * FE:]
        q = util.curve_obj(curve).q
        ny = b"".join(util.fe_bytes((q - util.fe_int(y[i * FE:(i + 1) * FE])) % q) for i in range(d))
        negs.append(x + ny)
    P = proj + [zero_p, proj[0], proj[1], proj[2], zero_p]
    Qp = proj[1:] + proj[:1] + [proj[3], zero_p, proj[1], b200.g_from_affine(curve, group, negs[2]), zero_p]
    Qa = aff[1:] + aff[:1] + [aff[3], bytes(ab), aff[1], negs[2], bytes(ab)]
    n = len(P)
    dP, dQp, dQa = b200.to_device(b"".join(P)), b200.to_device(b"".join(Qp)), b200.to_device(b"".join(Qa))
    dR = torch.empty(n * pb, dtype=torch.uint8, device=dev)
    dA = torch.empty(n * ab, dtype=torch.uint8, device=dev)

    def affine_of(dproj):
        b200.check(b200.lib().b200_dev_group_op(curve, group, 3, dproj.data_ptr(), None, dA.data_ptr(), n))
        out = b200.from_device(dA)
        return [out[i * ab:(i + 1) * ab] for i in range(n)]

    b200.check(b200.lib().b200_dev_group_op(curve, group, 0, dP.data_ptr(), dQp.data_ptr(), dR.data_ptr(), n))
    got = affine_of(dR)
    for i in range(n):
        exp = util.orc_to_affine(oracle, curve, group, util.orc_group(oracle, curve, group, 0, P[i], Qp[i]))
        assert got[i] == exp, ("add", i)
    b200.check(b200.lib().b200_dev_group_op(curve, group, 1, dP.data_ptr(), None, dR.data_ptr(), n))
    got = affine_of(dR)
    for i in range(n):
        exp = util.orc_to_affine(oracle, curve, group, util.orc_group(oracle, curve, group, 1, P[i]))
        assert got[i] == exp, ("dbl", i)
    b200.check(b200.lib().b200_dev_group_op(curve, group, 2, dP.data_ptr(), dQa.data_ptr(), dR.data_ptr(), n))
    got = affine_of(dR)
    for i in range(n):
        exp = util.orc_to_affine(oracle, curve, group, util.orc_group(oracle, curve, group, 2, P[i], Qa[i]))
        assert got[i] == exp, ("mixed_add", i)


@pytest.mark.parametrize("curve,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_gen_points_are_multiples_of_generator(b200, dev, curve, group):
    import torch
    n, first = 21, 5
    ab = b200.affine_bytes(curve, group)
    out = torch.empty(n * ab, dtype=torch.uint8, device=dev)
    b200.check(b200.lib().b200_gen_points(curve, group, out.data_ptr(), n, first))
    got = b200.from_device(out)
    F, a, b, G = util.group_params(curve, group)
    P = M.ec_mul(F, a, first, G)
    for i in range(n):
        assert got[i * ab:(i + 1) * ab] == util.encode_affine(curve, P, group), i
        P = M.ec_add(F, a, P, G)


# ------------------------------------------------------------------------------------------------ MSM
def _msm_case(b200, oracle, dev, curve, group, n, seed, special=True, window=0):
    import torch
    c = util.curve_obj(curve)
    rng = random.Random(seed)
    ab = b200.affine_bytes(curve, group)
    pts = torch.empty(max(n, 1) * ab, dtype=torch.uint8, device=dev)
    b200.check(b200.lib().b200_gen_points(curve, group, pts.data_ptr(), n, 1 + rng.randrange(1000)))
    points = bytearray(b200.from_device(pts)[:n * ab])
    scalars = [rng.randrange(c.r) for _ in range(n)]
    if special and n >= 16:
        # the structure observed in real keys (SURVEY.md 8 pitfalls): O entries, one point repeated many times,
        # a duplicate pair, P / -P with equal scalars; scalars 0, 1 (Montgomery one), r-1
        points[0:ab] = bytes(ab)
        points[(n - 1) * ab:n * ab] = bytes(ab)
        for i in range(2, n - 2, 2):
            points[i * ab:(i + 1) * ab] = points[2 * ab:3 * ab]
        d = util.deg(curve, group)
        x, y = points[5 * ab:5 * ab + d * FE], points[5 * ab + d * FE:6 * ab]
        ny = b"".join(util.fe_bytes((c.q - util.fe_int(y[i * FE:(i + 1) * FE])) % c.q) for i in range(d))
        points[7 * ab:8 * ab] = bytes(x) + ny
        scalars[7] = scalars[5]
        scalars[1] = 0
        scalars[3] = 1
        scalars[9] = c.r - 1
        scalars[11] = 1
    sc = b"".join(util.fe_bytes(M.to_mont(s, c.r)) for s in scalars)
    d_s, d_p = b200.to_device(sc), b200.to_device(bytes(points))
    b200.lib().b200_msm_set_window(window)
    try:
        got = b200.g_to_affine(curve, group, b200.msm(curve, group, d_s, d_p, n))
    finally:
        b200.lib().b200_msm_set_window(0)
    exp = util.orc_msm_affine(oracle, curve, group, sc, bytes(points), n)
    assert got == exp, (curve, group, n, window)


@pytest.fixture(params=[0, 1], ids=["xyzz", "batch-affine"])
def accum(request, b200):
    """Run the test under both bucket-accumulation modes of the MSM (the results must be identical bit for bit)."""
    b200.set_batch_affine(request.param)
    yield request.param
    b200.set_batch_affine(2)  # the library's default: automatic


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 1000])
def test_msm_g1_vs_oracle(b200, oracle, dev, curve, n, accum):
    _msm_case(b200, oracle, dev, curve, 1, n, 4000 + n + curve)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [1, 5, 64, 300])
def test_msm_g2_vs_oracle(b200, oracle, dev, curve, n, accum):
    _msm_case(b200, oracle, dev, curve, 2, n, 5000 + n + curve)


@pytest.mark.parametrize("window", [3, 8, 13, 16])
def test_msm_window_widths(b200, oracle, dev, window, accum):
    _msm_case(b200, oracle, dev, 0, 1, 200, 6000 + window, window=window)
    _msm_case(b200, oracle, dev, 1, 2, 40, 6100 + window, window=window)


def test_msm_degenerate_scalars(b200, oracle, dev, accum):
    import torch
    for curve in (0, 1):
        c = util.curve_obj(curve)
        n = 64
        ab = b200.affine_bytes(curve, 1)
        pts = torch.empty(n * ab, dtype=torch.uint8, device=dev)
        b200.check(b200.lib().b200_gen_points(curve, 1, pts.data_ptr(), n, 3))
        points = b200.from_device(pts)
        for val in (0, 1, c.r - 1):
            sc = util.fe_bytes(M.to_mont(val, c.r)) * n
            got = b200.g_to_affine(curve, 1, b200.msm(curve, 1, b200.to_device(sc), pts, n))
            assert got == util.orc_msm_affine(oracle, curve, 1, sc, points, n), val
        # empty input -> O
        assert b200.g_to_affine(curve, 1, b200.msm(curve, 1, pts, pts, 0)) == bytes(ab)


def test_msm_skewed_scalars_fold_path(b200, oracle, dev, accum):
    """All scalars equal: every window puts all n points into ONE bucket, i.e. thousands of task sums per bucket, which
    exercises the parallel folding of task sums (two levels at n = 2^14). msm(1,...,1) is checked against the oracle
    (which adds the n points directly, multiexp.tcc:468-479) and msm(s,...,s) against s * msm(1,...,1)."""
    import torch
    for curve, group, n in ((0, 1, 1 << 14), (1, 2, 1 << 12)):
        c = util.curve_obj(curve)
        ab = b200.affine_bytes(curve, group)
        pts = torch.empty(n * ab, dtype=torch.uint8, device=dev)
        b200.check(b200.lib().b200_gen_points(curve, group, pts.data_ptr(), n, 11))
        one = util.fe_bytes(M.to_mont(1, c.r))
        s_val = random.Random(99 + curve).randrange(2, c.r)
        s = util.fe_bytes(M.to_mont(s_val, c.r))
        sum_ones = b200.msm(curve, group, b200.to_device(one * n), pts, n)
        sum_s = b200.msm(curve, group, b200.to_device(s * n), pts, n)
        assert b200.g_to_affine(curve, group, sum_ones) == util.orc_msm_affine(oracle, curve, group, one * n,
                                                                              b200.from_device(pts), n)
        assert b200.g_to_affine(curve, group, sum_s) == b200.g_to_affine(curve, group, b200.g_scale(curve, group, s, sum_ones))


def test_msm_linearity_large(b200, dev, accum):
    """n = 2^15 (beyond what the oracle does in seconds): msm(s,P) + msm(t,P) == msm(s+t,P), and the same sum from
    two different window widths."""
    import torch
    curve, n = 0, 1 << 15
    c = util.curve_obj(curve)
    ab = b200.affine_bytes(curve, 1)
    pts = torch.empty(n * ab, dtype=torch.uint8, device=dev)
    b200.check(b200.lib().b200_gen_points(curve, 1, pts.data_ptr(), n, 77))
    g = torch.Generator(device="cpu").manual_seed(1234)
    raw = torch.randint(0, 256, (2, n, FE), dtype=torch.uint8, generator=g)
    raw[:, :, 94:] = 0  # < 2^752 < r: any such value is a valid Montgomery representation
    s, t = raw[0].contiguous().to(dev), raw[1].contiguous().to(dev)
    st = torch.empty_like(s)
    b200.check(b200.lib().b200_dev_fp_op(0, 0, s.data_ptr(), t.data_ptr(), st.data_ptr(), n))
    a = b200.msm(curve, 1, s, pts, n)
    b = b200.msm(curve, 1, t, pts, n)
    b200.lib().b200_msm_set_window(11)
    try:
        cc = b200.msm(curve, 1, st, pts, n)
    finally:
        b200.lib().b200_msm_set_window(0)
    assert b200.g_to_affine(curve, 1, b200.g_add(curve, 1, a, b)) == b200.g_to_affine(curve, 1, cc)


# ------------------------------------------------------------------------------------------------ NTT / compute_H
@pytest.mark.parametrize("curve,logm", [(0, 1), (0, 2), (0, 5), (0, 8), (0, 9), (0, 13), (0, 17), (1, 1), (1, 3), (1, 8), (1, 12),
                                        (1, 15)])
def test_domain_ops_vs_oracle(b200, oracle, dev, curve, logm):
    c = util.curve_obj(curve)
    m = 1 << logm
    rng = random.Random(7000 + 100 * curve + logm)
    data = util.rand_fe_bytes(rng, c.r, m)
    dom = b200.Domain(curve, m)
    for kind, fn in ((0, dom.fft), (1, dom.ifft), (2, dom.coset_fft), (3, dom.icoset_fft), (4, dom.divide_by_z_on_coset)):
        d = b200.to_device(data)
        fn(d)
        assert b200.from_device(d) == util.orc_domain(oracle, curve, kind, data, m), (kind, logm)
    dom.close()


def test_domain_rejects_bad_sizes(b200, dev):
    for curve, m in ((0, 3), (0, 1), (1, 1 << 16), (0, 100)):
        with pytest.raises(b200.B200Error):
            b200.Domain(curve, m)


# logm = 17: three passes (6 + 6 + 5 stages) - the middle pass has s0 > 0 and neither a pre- nor a post-table, like the
# 2^20 transform of the challenge size; 2^15 is MNT6753's challenge size
@pytest.mark.parametrize("curve,logm", [(0, 6), (1, 6), (0, 11), (1, 10), (0, 17), (1, 15)])
def test_compute_h_vs_oracle(b200, oracle, dev, curve, logm):
    import torch
    c = util.curve_obj(curve)
    m = 1 << logm
    rng = random.Random(8000 + curve + logm)
    ca, cb, cc = (util.rand_fe_bytes(rng, c.r, m) for _ in range(3))
    dom = b200.Domain(curve, m)
    out = torch.empty((m + 1) * FE, dtype=torch.uint8, device=dev)
    dom.compute_h(b200.to_device(ca), b200.to_device(cb), b200.to_device(cc), out)
    assert b200.from_device(out) == util.orc_compute_h(oracle, curve, m - 1, ca, cb, cc)
    dom.close()


def test_vector_ops_vs_oracle(b200, oracle, dev):
    for curve in (0, 1):
        c = util.curve_obj(curve)
        tag = 0 if curve == 0 else 1
        rng = random.Random(8100 + curve)
        n = 333
        a, b = util.rand_fe_bytes(rng, c.r, n), util.rand_fe_bytes(rng, c.r, n)
        da, db = b200.to_device(a), b200.to_device(b)
        b200.check(b200.lib().b200_fr_muleq(curve, da.data_ptr(), db.data_ptr(), n))
        out = b200.from_device(da)
        for i in range(n):
            assert out[i * FE:(i + 1) * FE] == util.orc_fp(oracle, tag, 2, a[i * FE:(i + 1) * FE], b[i * FE:(i + 1) * FE])
        da = b200.to_device(a)
        b200.check(b200.lib().b200_fr_subeq(curve, da.data_ptr(), db.data_ptr(), n))
        out = b200.from_device(da)
        for i in range(n):
            assert out[i * FE:(i + 1) * FE] == util.orc_fp(oracle, tag, 1, a[i * FE:(i + 1) * FE], b[i * FE:(i + 1) * FE])


def test_ntt_roundtrip_full_size(b200, dev):
    """2^20 (MNT4753 challenge size) and 2^15 (MNT6753): icosetFFT(cosetFFT(x)) == x and iFFT(FFT(x)) == x."""
    import torch
    for curve, logm in ((0, 20), (1, 15)):
        m = 1 << logm
        g = torch.Generator(device="cpu").manual_seed(99 + curve)
        raw = torch.randint(0, 256, (m, FE), dtype=torch.uint8, generator=g)
        raw[:, 94:] = 0
        x = raw.to(dev)
        y = x.clone()
        dom = b200.Domain(curve, m)
        dom.coset_fft(y)
        assert not torch.equal(x, y)
        dom.icoset_fft(y)
        assert torch.equal(x, y)
        dom.fft(y)
        dom.ifft(y)
        assert torch.equal(x, y)
        dom.close()


# ------------------------------------------------------------------------------------------------ whole prover
@pytest.fixture(params=[True, False], ids=["tables", "no-tables"])
def precompute(request, b200):
    """run with and without the pre-shifted base tables (b200_params_precompute)"""
    b200.set_precompute(request.param)
    yield request.param
    b200.set_precompute(True)


@pytest.mark.parametrize("curve,k", [(0, 5), (1, 5), (0, 8), (1, 8)])
def test_prover_matches_reference_golden(b200, dev, precompute, curve, k, accum):
    params, inp, expected = util.golden(curve, k)
    P = b200.Params.from_bytes(curve, params)
    if precompute:
        assert P.precompute() > 0
    assert (P.d, P.m) == ((1 << k) - 1, 1 << k)
    got = P.prove(inp)
    assert got == expected
    again, tm = P.prove(inp, timings=True)  # second proof on the same resident key
    assert again == expected and tm["total_ms"] > 0
    P.close()


@pytest.mark.parametrize("curve", [0, 1])
def test_sharded_prover_matches_reference_golden(b200, dev, precompute, curve, accum):
    """MSMs split by point range over `world` ranks (run sequentially on one GPU here) + host combine."""
    params, inp, expected = util.golden(curve, 8)
    P = b200.Params.from_bytes(curve, params)
    for world in (2, 3):
        parts = b"".join(P.prove_partial(inp, r, world)[0] for r in range(world))
        assert b200.prove_combine(curve, parts, world, inp[-FE:]) == expected
    # uneven sharding: runs of 1/64 slices (b200_prove_partial_span), one of them empty
    runs = [(0, 36), (36, 36), (36, 61), (61, 64)]
    parts = b"".join(P.prove_partial(inp, lo, 64, hi)[0] for lo, hi in runs)
    assert b200.prove_combine(curve, parts, len(runs), inp[-FE:]) == expected
    # per-query sharding (b200_prove_partial_queries): every MSM cut on its own, some GPUs without a part in some MSMs
    # (A, B1, B2, L, H): whole MSMs, B2 in three pieces, L cut off the slice grid of the others, H whole on one rank,
    # a rank with nothing at all
    plans = [[(0, 64), (0, 0), (0, 20), (50, 64), (0, 0)],
             [(0, 0), (0, 64), (20, 41), (0, 0), (0, 64)],
             [(0, 0), (0, 0), (41, 64), (0, 50), (0, 0)],
             [(0, 0), (0, 0), (0, 0), (0, 0), (0, 0)]]
    parts = b"".join(P.prove_partial_queries(inp, sp, 64)[0] for sp in plans)
    assert b200.prove_combine(curve, parts, len(plans), inp[-FE:]) == expected
    parts = b"".join(P.prove_partial_queries(inp, sp, 64, b1_scaled=True)[0] for sp in plans)
    assert b200.prove_combine(curve, parts, len(plans), None) == expected
    # the same slicing through b200_prove_batch jobs, and a second time on the cached tables
    for _ in range(2):
        parts = b"".join(b200.prove_batch([(P, inp, 0, 64, 0, None, sp)], b1_scaled=True)[0] for sp in plans[:3])
        assert b200.prove_combine(curve, parts, 3, None) == expected
    with pytest.raises(b200.B200Error, match="bad slice run"):
        P.prove_partial_queries(inp, [(0, 65)] + [(0, 0)] * 4, 64)
    # a key whose A query has no equal bases (here: a copy of B1), so that A takes part in the shared preparation, and a
    # rank without a part in B2: B1 then makes the preparation that A and L reuse. Expected proof from the oracle.
    d0, m0, q = util.split_params(curve, params)
    g1 = 2 * FE
    plain = params[:16] + q["B1"] + params[16 + (m0 + 1) * g1:]
    assert len(plain) == len(params)
    want = util.orc_prove(util.load_oracle(), curve, plain, inp)
    Q = b200.Params.from_bytes(curve, plain)
    assert Q.prove(inp) == want
    plans = [[(0, 64), (0, 64), (0, 0), (0, 64), (0, 0)], [(0, 0), (0, 0), (0, 64), (0, 0), (0, 64)]]
    parts = b"".join(Q.prove_partial_queries(inp, sp, 64, b1_scaled=True)[0] for sp in plans)
    assert b200.prove_combine(curve, parts, 2, None) == want
    Q.close()
    # the witness map computed outside the call (what bench.py does when it splits compute_H over three ranks)
    import torch
    d, m = P.d, P.m
    vecs = [b200.to_device(inp[(m + 1 + i * (d + 1)) * FE:(m + 1 + (i + 1) * (d + 1)) * FE]) for i in range(3)]
    h = torch.zeros((d + 2) * FE, dtype=torch.uint8, device=dev)
    dom = b200.Domain(curve, d + 1)
    dom.compute_h(vecs[0], vecs[1], vecs[2], h)
    garbage = bytearray(inp)
    garbage[(m + 1) * FE:(m + 1 + 3 * (d + 1)) * FE] = bytes(3 * (d + 1) * FE)   # ca / cb / cc must not be read
    parts = b"".join(P.prove_partial(bytes(garbage), r, 2, d_h=h)[0] for r in range(2))
    assert b200.prove_combine(curve, parts, 2, inp[-FE:]) == expected
    dom.close()
    P.close()


def test_concurrent_proofs_match_reference_golden(b200, dev):
    """b200_prove_batch: the MNT4753 and MNT6753 proofs (and two sizes of each) in flight at the same time, several
    batches in a row on the same persistent worker threads; then the sharded variant of the same call."""
    cases = [(0, 8), (1, 8), (0, 5), (1, 5)]
    keys, inputs, expected = [], [], []
    for curve, k in cases:
        params, inp, exp = util.golden(curve, k)
        keys.append(b200.Params.from_bytes(curve, params))
        inputs.append(inp)
        expected.append(exp)
    for _ in range(3):
        got = b200.prove_batch(list(zip(keys, inputs)))
        assert got == expected
    got, tms = b200.prove_batch(list(zip(keys[:2], inputs[:2])), timings=True)
    assert got == expected[:2] and all(t["total_ms"] > 0 for t in tms)
    world = 2
    parts = [b200.prove_batch([(keys[0], inputs[0], r, world), (keys[1], inputs[1], r, world)]) for r in range(world)]
    for j in (0, 1):
        allp = b"".join(parts[r][j] for r in range(world))
        assert b200.prove_combine(cases[j][0], allp, world, inputs[j][-FE:]) == expected[j]
    # the same with every rank multiplying its own B1 sum by r (b200_prove_partial_scaled): combine without r
    parts = [b200.prove_batch([(keys[0], inputs[0], r, world), (keys[1], inputs[1], r, world)], b1_scaled=True)
             for r in range(world)]
    for j in (0, 1):
        allp = b"".join(parts[r][j] for r in range(world))
        assert b200.prove_combine(cases[j][0], allp, world, None) == expected[j]
    one = b"".join(keys[0].prove_partial(inputs[0], r, 3, b1_scaled=True)[0] for r in range(3))
    assert b200.prove_combine(0, one, 3, None) == expected[0]
    # a failing job is reported with its index and does not wedge the workers
    with pytest.raises(b200.B200Error, match="proof job 1"):
        b200.prove_batch([(keys[0], inputs[0]), (keys[1], inputs[1][:-FE])])
    assert b200.prove_batch(list(zip(keys[:2], inputs[:2]))) == expected[:2]
    with pytest.raises(b200.B200Error, match="same key"):
        b200.prove_batch([(keys[0], inputs[0]), (keys[0], inputs[0])])
    for P in keys:
        P.close()


@pytest.mark.parametrize("curve", [0, 1])
def test_equal_bases_are_merged_bit_exactly(b200, oracle, dev, curve):
    """Key-load time merging of equal bases (MsmDedup): a query with one group larger than a segment (1024), a small
    group, a pair, a duplicated point at infinity and scalars that cancel; table MSM (merging on) == table-free MSM."""
    import torch
    k = 12
    m = 1 << k
    g1, g2 = b200.affine_bytes(curve, 1), b200.affine_bytes(curve, 2)

    def gen(group, n, first):
        t = torch.empty(n * b200.affine_bytes(curve, group), dtype=torch.uint8, device=dev)
        b200.check(b200.lib().b200_gen_points(curve, group, t.data_ptr(), n, first))
        return t

    A, B1, B2 = gen(1, m + 1, 11), gen(1, m + 1, 5000), gen(2, m + 1, 9000)
    L, H = gen(1, m - 1, 13000), gen(1, m - 1, 17000)
    Av = A.view(m + 1, g1)
    Av[100:3100] = Av[7].clone()          # 3001 copies: three segments
    Av[3200:3205] = Av[3300].clone()      # group of 6
    Av[4000] = Av[4001].clone()           # a pair
    Av[50] = 0
    Av[51] = 0                            # two points at infinity
    torch.cuda.synchronize()
    key = b200.Params.from_device(curve, m - 1, m, A, B1, B2, L, H)
    assert key.precompute() > 0
    c = util.curve_obj(curve)
    rng = random.Random(99 + curve)
    scalars = [rng.randrange(c.r) for _ in range(m + 1)]
    scalars[7] = (-sum(scalars[100:3100])) % c.r   # the big group sums to zero
    scalars[3300] = 0
    scalars[4000] = c.r - 1
    sc = b200.to_device(b"".join(util.fe_bytes(M.to_mont(s, c.r)) for s in scalars))
    got = b200.g_to_affine(curve, 1, key.msm(0, sc, m + 1))
    exp = b200.g_to_affine(curve, 1, b200.msm(curve, 1, sc, A, m + 1))
    assert got == exp
    assert got == util.orc_msm_affine(oracle, curve, 1, b200.from_device(sc), b200.from_device(A), m + 1, chunks=8)
    # and through a whole proof: tables (merging) vs table-free
    g = torch.Generator(device="cpu").manual_seed(4321 + curve)
    n_in = (m + 1) + 3 * m + 1
    img = torch.randint(0, 256, (n_in, FE), dtype=torch.uint8, generator=g)
    img[:, 94:] = 0
    p1 = key.prove(img)
    b200.set_precompute(False)
    try:
        p2 = key.prove(img)
    finally:
        b200.set_precompute(True)
    assert p1 == p2
    key.close()


# ------------------------------------------------------------------------------------------------ table-mode MSM vs oracle
def _structured_key(b200, dev, curve, k):
    """device-side synthetic key of 2^k constraints with the structure of real keys (SURVEY.md 8 pitfalls 1-2)"""
    import torch
    m = 1 << k
    g1, g2 = b200.affine_bytes(curve, 1), b200.affine_bytes(curve, 2)

    def gen(group, n, first):
        t = torch.empty(n * b200.affine_bytes(curve, group), dtype=torch.uint8, device=dev)
        b200.check(b200.lib().b200_gen_points(curve, group, t.data_ptr(), n, first))
        return t

    A, B1, B2 = gen(1, m + 1, 1000003), gen(1, m + 1, 2000003), gen(2, m + 1, 3000017)
    L, H = gen(1, m - 1, 4000037), gen(1, m - 1, 5000011)
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
    return b200.Params.from_device(curve, m - 1, m, A, B1, B2, L, H), (A, B1, B2, L, H)


def _scalars_with_specials(curve, n, seed):
    c = util.curve_obj(curve)
    rng = random.Random(seed)
    scalars = [rng.randrange(c.r) for _ in range(n)]
    scalars[0] = 1
    scalars[1] = 0
    scalars[5] = c.r - 1
    scalars[7] = 1 << 752  # exercises the carry into the top window
    return b"".join(util.fe_bytes(M.to_mont(s, c.r)) for s in scalars)


@pytest.mark.parametrize("curve,window", [(0, 21), (1, 17), (0, 19)])
def test_table_msm_forced_wide_windows_vs_oracle(b200, oracle, dev, curve, window, accum):
    """The headline schedule - pre-shifted base tables, ONE merged bucket set, window widths 17..21 (36-45 windows) -
    against the oracle, for all five queries of a key (A with its m/2 equal bases merged, B2 in G2): the window width
    the 2^20 proof picks is forced on a 2^12 key, so the oracle finishes in seconds."""
    k = 12
    m = 1 << k
    b200.set_precompute(True)
    b200.lib().b200_msm_set_window(window)
    try:
        key, qs = _structured_key(b200, dev, curve, k)
        assert key.precompute() > 0
    finally:
        b200.lib().b200_msm_set_window(0)
    sc = _scalars_with_specials(curve, m + 1, 9100 + curve + window)
    d_sc = b200.to_device(sc)
    which = (0, 1, 2, 3, 4) if window != 19 else (1, 4)
    for w in which:
        n = m + 1 if w < 3 else m - 1
        group = 2 if w == 2 else 1
        got = b200.g_to_affine(curve, group, key.msm(w, d_sc, n))
        plan = b200.msm_last_plan()
        assert plan["c"] == window and plan["windows"] == (754 + window - 1) // window, plan
        exp = util.orc_msm_affine(oracle, curve, group, sc[:n * FE], b200.from_device(qs[w]), n, chunks=8)
        assert got == exp, (curve, window, w)
    key.close()


def test_table_msm_natural_window_2_16_vs_oracle(b200, oracle, dev, accum):
    """n = 2^16 + 1 G1 points: the window width chosen by the library itself is >= 17 here (table mode)."""
    import torch
    curve, k = 0, 16
    m = 1 << k
    key, qs = _structured_key(b200, dev, curve, k)
    assert key.precompute() > 0
    sc = _scalars_with_specials(curve, m + 1, 9200)
    got = b200.g_to_affine(curve, 1, key.msm(1, b200.to_device(sc), m + 1))
    assert b200.msm_last_plan()["c"] >= 17
    assert got == util.orc_msm_affine(oracle, curve, 1, sc, b200.from_device(qs[1]), m + 1, chunks=16)
    key.close()


@pytest.mark.parametrize("curve,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_host_synth_key_matches_device_generator(b200, dev, curve, group, tmp_path):
    """tools/synth_key (host, the bench's file generator) and gen_points_kernel produce the same multiples of the
    generator."""
    import subprocess, sys, os, torch
    sys.path.insert(0, util.ROOT)
    import bench
    k = 6
    m = 1 << k
    pf, inf = str(tmp_path / "p"), str(tmp_path / "i")
    subprocess.check_call([bench.synth_tool(), bench.CURVES[curve], str(k), pf, inf])
    _, _, q = util.split_params(curve, open(pf, "rb").read())
    name, first, n = ("L", 4000037, m - 1) if group == 1 else ("B2", 3000017, m + 1)
    ab = b200.affine_bytes(curve, group)
    out = torch.empty(n * ab, dtype=torch.uint8, device=dev)
    b200.check(b200.lib().b200_gen_points(curve, group, out.data_ptr(), n, first))
    got = b200.from_device(out)
    lo, hi = (0, n) if group == 1 else (1, m - 2)   # B2 carries O at 0 and m and a duplicate pair at m-2
    assert got[lo * ab:hi * ab] == q[name][lo * ab:hi * ab]


# ------------------------------------------------------------------------------------------------ key generation: batch_exp
@pytest.mark.parametrize("curve,group,window", [(0, 1, 0), (0, 1, 16), (0, 2, 11), (1, 1, 0), (1, 2, 7)])
def test_batch_exp_vs_oracle(b200, oracle, dev, curve, group, window):
    """fixed-base windowed exponentiation (libff::batch_exp, multiexp.tcc:547-645) - out[i] =
    s_i * g, affine wire format - against the oracle's scalar multiplication, including scalars 0, 1, r-1 and a base
    that is not the generator."""
    import torch
    c = util.curve_obj(curve)
    rng = random.Random(9300 + 10 * curve + group + window)
    n = 150
    scalars = [rng.randrange(c.r) for _ in range(n)]
    scalars[0], scalars[1], scalars[2], scalars[3] = 0, 1, c.r - 1, 1 << 752
    sc = b"".join(util.fe_bytes(M.to_mont(s, c.r)) for s in scalars)
    ab = b200.affine_bytes(curve, group)
    G = b200.g_from_affine(curve, group, util.generator_affine(curve, group))
    k = util.fe_bytes(M.to_mont(rng.randrange(2, c.r), c.r))
    base_proj = util.orc_group(oracle, curve, group, 3, G, k)
    base = util.orc_to_affine(oracle, curve, group, base_proj)
    out = torch.empty(n * ab, dtype=torch.uint8, device=dev)
    ms = b200.batch_exp(curve, group, base, b200.to_device(sc), n, out, window)
    assert all(v >= 0 for v in ms.values())
    got = b200.from_device(out)
    for i in range(n):
        exp = util.orc_to_affine(oracle, curve, group, util.orc_group(oracle, curve, group, 3, base_proj, sc[i * FE:(i + 1) * FE]))
        assert got[i * ab:(i + 1) * ab] == exp, (curve, group, window, i)


@pytest.mark.parametrize("curve,group", [(0, 1), (1, 2)])
def test_msm_context_vs_oracle(b200, oracle, dev, curve, group, accum):
    """b200_msm_ctx_*: pre-shifted base tables for an arbitrary point set (what the sharded sweeps use per rank)"""
    import torch
    c = util.curve_obj(curve)
    n = 257
    ab = b200.affine_bytes(curve, group)
    pts = torch.empty(n * ab, dtype=torch.uint8, device=dev)
    b200.check(b200.lib().b200_gen_points(curve, group, pts.data_ptr(), n, 4242))
    points = bytearray(b200.from_device(pts))
    points[3 * ab:4 * ab] = bytes(ab)             # a point at infinity
    points[9 * ab:10 * ab] = points[8 * ab:9 * ab]  # a duplicate pair
    sc = _scalars_with_specials(curve, n, 9400 + curve)
    ctx = b200.MsmContext(curve, group, b200.to_device(bytes(points)), n)
    for _ in range(2):
        got = b200.g_to_affine(curve, group, ctx.run(b200.to_device(sc)))
        assert got == util.orc_msm_affine(oracle, curve, group, sc, bytes(points), n)
    ctx.close()
