"""CPU suite, part 1: pins the oracle (oracle/oracle.c) before anything trusts it.

 * byte-for-byte against the committed golden vectors, which are outputs of the UNMODIFIED reference prover
   (libsnark/main.cpp built by oracle/build_ref.sh) on files made by the reference generator - see
   tests/golden/make_golden.sh;
 * against plain Python integers for the field / tower / curve layers (tools/mnt753.py).
"""
import hashlib
import os
import random

import pytest

import mnt753 as M
import util


@pytest.mark.parametrize("curve,k", [(0, 5), (1, 5), (0, 8), (1, 8)])
def test_oracle_prover_matches_reference_golden(oracle, curve, k):
    params, inp, expected = util.golden(curve, k)
    got = util.orc_prove(oracle, curve, params, inp)
    assert got == expected
    sums = open(os.path.join(util.GOLDEN, "SHA256SUMS")).read()
    assert hashlib.sha256(got).hexdigest() in sums


def test_oracle_prover_independent_of_chunking(oracle):
    params, inp, expected = util.golden(0, 5)
    for chunks in (1, 3):
        assert util.orc_prove(oracle, 0, params, inp, chunks=chunks) == expected


@pytest.mark.parametrize("tag", [0, 1])
def test_oracle_field_vs_python(oracle, tag):
    p = M.PRIMES["AB"[tag]]
    rng = random.Random(100 + tag)
    rinv = pow(M.R, -1, p)
    cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (p - 1, 1)] + [(rng.randrange(p), rng.randrange(p)) for _ in range(100)]
    for a, b in cases:
        A, B = util.fe_bytes(a), util.fe_bytes(b)
        assert util.fe_int(util.orc_fp(oracle, tag, 0, A, B)) == (a + b) % p
        assert util.fe_int(util.orc_fp(oracle, tag, 1, A, B)) == (a - b) % p
        assert util.fe_int(util.orc_fp(oracle, tag, 2, A, B)) == a * b * rinv % p
    a = rng.randrange(1, p)
    inv = util.fe_int(util.orc_fp(oracle, tag, 3, util.fe_bytes(a)))
    assert M.from_mont(inv, p) * M.from_mont(a, p) % p == 1
    # constants derived inside the oracle agree with the generator's
    import ctypes
    out = ctypes.create_string_buffer(96)
    oracle.orc_fp_const(tag, 4, ctypes.addressof(out))
    assert util.fe_int(out.raw) == M.to_mont(M.root_of_unity("AB"[tag]), p)


@pytest.mark.parametrize("curve", [0, 1])
def test_oracle_tower_vs_python(oracle, curve):
    c = util.curve_obj(curve)
    F = M.g2_field(c)
    rng = random.Random(7 + curve)
    d = c.ext_deg
    enc = lambda x: b"".join(util.fe_bytes(M.to_mont(v, c.q)) for v in x)
    for _ in range(20):
        a = tuple(rng.randrange(c.q) for _ in range(d))
        b = tuple(rng.randrange(c.q) for _ in range(d))
        assert util.orc_fqe(oracle, curve, 2, 2, enc(a), enc(b)) == enc(F.mul(a, b))
        assert util.orc_fqe(oracle, curve, 2, 3, enc(a)) == enc(F.sqr(a))
        assert util.orc_fqe(oracle, curve, 2, 4, enc(a)) == enc(F.inv(a))


@pytest.mark.parametrize("curve,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_oracle_group_vs_python(oracle, curve, group):
    F, a, b, G = util.group_params(curve, group)
    rng = random.Random(11 * curve + group)
    k1, k2 = rng.randrange(1, 1 << 64), rng.randrange(1, 1 << 64)
    P1, P2 = M.ec_mul(F, a, k1, G), M.ec_mul(F, a, k2, G)
    one = util.fe_bytes(M.to_mont(1, util.curve_obj(curve).q))
    proj = lambda P: util.encode_affine(curve, P, group) + one + bytes((util.deg(curve, group) - 1) * 96)
    s = util.orc_group(oracle, curve, group, 0, proj(P1), proj(P2))
    assert util.orc_to_affine(oracle, curve, group, s) == util.encode_affine(curve, M.ec_add(F, a, P1, P2), group)
    dd = util.orc_group(oracle, curve, group, 1, proj(P1))
    assert util.orc_to_affine(oracle, curve, group, dd) == util.encode_affine(curve, M.ec_add(F, a, P1, P1), group)
    # P + P through the addition entry point takes the doubling branch; P + (-P) gives O -> zero bytes
    s = util.orc_group(oracle, curve, group, 0, proj(P1), proj(P1))
    assert util.orc_to_affine(oracle, curve, group, s) == util.encode_affine(curve, M.ec_add(F, a, P1, P1), group)
    s = util.orc_group(oracle, curve, group, 0, proj(P1), proj(M.ec_neg(F, P1)))
    assert util.orc_to_affine(oracle, curve, group, s) == bytes(2 * util.deg(curve, group) * 96)


def test_oracle_fft_roundtrip_and_naive(oracle):
    # pattern of libfqfft/tests/evaluation_domain_test.cpp:40-154: FFT == naive evaluation, iFFT(FFT) == id
    for curve in (0, 1):
        c = util.curve_obj(curve)
        r = c.r
        m = 8
        rng = random.Random(5)
        vals = [rng.randrange(r) for _ in range(m)]
        data = b"".join(util.fe_bytes(M.to_mont(v, r)) for v in vals)
        out = util.orc_domain(oracle, curve, 0, data, m)
        tag = "A" if curve == 0 else "B"
        w = pow(M.root_of_unity(tag), 1 << (M.TWO_ADICITY[tag] - 3), r)
        for j in range(m):
            exp = sum(vals[i] * pow(w, i * j, r) for i in range(m)) % r
            assert util.fe_int(out[96 * j:96 * j + 96]) == M.to_mont(exp, r)
        assert util.orc_domain(oracle, curve, 1, out, m) == data
        cos = util.orc_domain(oracle, curve, 2, data, m)
        assert util.orc_domain(oracle, curve, 3, cos, m) == data
