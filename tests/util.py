"""Shared helpers for the tests: oracle loader (ctypes), encoders, seeded random inputs."""
import ctypes
import os
import random
import subprocess

import mnt753 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FE = 96


def load_oracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    L = ctypes.CDLL(path)
    vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.orc_init.restype = i
    L.orc_fp_op.argtypes = [i, i, vp, vp, vp]
    L.orc_fp_const.argtypes = [i, i, vp]
    L.orc_fqe_op.argtypes = [i, i, i, vp, vp, vp]
    L.orc_group_op.argtypes = [i, i, i, vp, vp, vp]
    L.orc_to_affine.argtypes = [i, i, vp, vp]
    L.orc_msm.argtypes = [i, i, vp, vp, sz, vp, i]
    L.orc_domain_op.argtypes = [i, i, vp, sz]
    L.orc_domain_op.restype = i
    L.orc_compute_h.argtypes = [i, sz, vp, vp, vp, vp]
    L.orc_compute_h.restype = i
    L.orc_prove.argtypes = [i, ctypes.c_char_p, sz, ctypes.c_char_p, sz, ctypes.c_char_p, i]
    L.orc_prove.restype = ctypes.c_long
    L.orc_init()
    return L


def buf(data):
    return ctypes.create_string_buffer(bytes(data), len(data))


def curve_obj(curve):
    return M.MNT4753 if curve == 0 else M.MNT6753


def deg(curve, group):
    return 1 if group == 1 else curve_obj(curve).ext_deg


def fe_bytes(x):
    return int(x).to_bytes(FE, "little")


def fe_int(b):
    return int.from_bytes(b, "little")


def rand_fe_bytes(rng, p, n):
    return b"".join(fe_bytes(rng.randrange(p)) for _ in range(n))


def generator_affine(curve, group):
    c = curve_obj(curve)
    if group == 1:
        coords = [(c.g1[0],), (c.g1[1],)]
    else:
        coords = [tuple(c.g2[0]), tuple(c.g2[1])]
    return b"".join(fe_bytes(M.to_mont(v, c.q)) for co in coords for v in co)


def encode_affine(curve, P, group):
    """affine python point (tuple of tuples of ints, non-Montgomery) or None -> wire bytes"""
    c = curve_obj(curve)
    d = deg(curve, group)
    if P is None:
        return bytes(2 * d * FE)
    return b"".join(fe_bytes(M.to_mont(v, c.q)) for co in P for v in co)


def group_params(curve, group):
    c = curve_obj(curve)
    return M.g1_params(c) if group == 1 else M.g2_params(c)


# ---- oracle wrappers --------------------------------------------------------------------------------------------
def orc_fp(O, tag, op, a, b=None):
    out = ctypes.create_string_buffer(FE)
    ab = buf(a)
    bb = buf(b) if b is not None else None
    O.orc_fp_op(tag, op, ctypes.addressof(ab), ctypes.addressof(bb) if bb is not None else None, ctypes.addressof(out))
    return out.raw


def orc_fqe(O, curve, group, op, a, b=None):
    n = deg(curve, group) * FE
    out = ctypes.create_string_buffer(n)
    ab = buf(a)
    bb = buf(b) if b is not None else None
    O.orc_fqe_op(curve, group, op, ctypes.addressof(ab), ctypes.addressof(bb) if bb else None, ctypes.addressof(out))
    return out.raw


def orc_group(O, curve, group, op, p, q=None):
    n = 3 * deg(curve, group) * FE
    out = ctypes.create_string_buffer(n)
    pb = buf(p)
    qb = buf(q) if q is not None else None
    O.orc_group_op(curve, group, op, ctypes.addressof(pb), ctypes.addressof(qb) if qb else None, ctypes.addressof(out))
    return out.raw


def orc_to_affine(O, curve, group, p):
    out = ctypes.create_string_buffer(2 * deg(curve, group) * FE)
    pb = buf(p)
    O.orc_to_affine(curve, group, ctypes.addressof(pb), ctypes.addressof(out))
    return out.raw


def orc_msm_affine(O, curve, group, scalars, points, n, chunks=4):
    """oracle multi-exponentiation -> affine wire bytes"""
    out = ctypes.create_string_buffer(3 * deg(curve, group) * FE)
    sb, pb = buf(scalars), buf(points)
    O.orc_msm(curve, group, ctypes.addressof(sb), ctypes.addressof(pb), n, ctypes.addressof(out), chunks)
    return orc_to_affine(O, curve, group, out.raw)


def orc_domain(O, curve, kind, data, m):
    b = buf(data)
    rc = O.orc_domain_op(curve, kind, ctypes.addressof(b), m)
    assert rc == 0
    return b.raw


def orc_compute_h(O, curve, d, ca, cb, cc):
    out = ctypes.create_string_buffer((d + 2) * FE)
    a, b, c = buf(ca), buf(cb), buf(cc)
    rc = O.orc_compute_h(curve, d, ctypes.addressof(a), ctypes.addressof(b), ctypes.addressof(c), ctypes.addressof(out))
    assert rc == 0
    return out.raw


def orc_prove(O, curve, params, inp, chunks=8):
    out = ctypes.create_string_buffer(4096)
    n = O.orc_prove(curve, params, len(params), inp, len(inp), out, chunks)
    assert n > 0, n
    return out.raw[:n]


def golden(curve, k):
    name = "%s_k%d" % ("MNT4753" if curve == 0 else "MNT6753", k)
    rd = lambda ext: open(os.path.join(GOLDEN, name + ext), "rb").read()
    return rd(".params"), rd(".input"), rd(".output")


def split_params(curve, image):
    """-> d, m, dict of byte strings A, B1, B2, L, H"""
    d = int.from_bytes(image[0:8], "little")
    m = int.from_bytes(image[8:16], "little")
    g1, g2 = 2 * FE, 2 * FE * deg(curve, 2)
    o = 16
    out = {}
    for name, cnt, sz in (("A", m + 1, g1), ("B1", m + 1, g1), ("B2", m + 1, g2), ("L", m - 1, g1), ("H", d, g1)):
        out[name] = image[o:o + cnt * sz]
        o += cnt * sz
    assert o == len(image)
    return d, m, out


def split_input(image, d, m):
    o = 0
    out = {}
    for name, cnt in (("w", m + 1), ("ca", d + 1), ("cb", d + 1), ("cc", d + 1), ("r", 1)):
        out[name] = image[o:o + cnt * FE]
        o += cnt * FE
    assert o == len(image)
    return out
