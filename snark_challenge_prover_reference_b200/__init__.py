"""B200-native Groth16 prover hot path for MNT4753 / MNT6753 (sm_100a CUDA kernels behind a C ABI).

Python here is plumbing only: it loads ``libb200groth16.so`` (built in-tree by ``build.py``) with ctypes and exposes
the C entry points of ``include/b200_groth16.h`` one-to-one, plus a few helpers that use torch for device memory.
There is no CPU fallback: importing works anywhere, but every compute call fails loudly without the CUDA library
or without a GPU.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB selects another build of the same library (kernel variants compiled with other switches, for timing runs)
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(HERE, "libb200groth16.so")

MNT4753, MNT6753 = 0, 1
CURVE_NAMES = {MNT4753: "MNT4753", MNT6753: "MNT6753"}
FE = 96  # bytes per field element (12 LE u64 limbs, Montgomery form)


def g2_degree(curve):
    return 2 if curve == MNT4753 else 3


def affine_bytes(curve, group):
    return 2 * FE * (1 if group == 1 else g2_degree(curve))


def proj_bytes(curve, group):
    return 3 * FE * (1 if group == 1 else g2_degree(curve))


def proof_bytes(curve):
    return 2 * affine_bytes(curve, 1) + affine_bytes(curve, 2)


def partial_bytes(curve):
    return 4 * proj_bytes(curve, 1) + proj_bytes(curve, 2)


class B200Error(RuntimeError):
    pass


class ProveTimings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("h2d_ms", "compute_h_ms", "msm_a_ms", "msm_b1_ms", "msm_b2_ms",
                                               "msm_h_ms", "msm_l_ms", "tail_ms", "total_ms")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ProofJob(ctypes.Structure):
    """b200_proof_job of include/b200_groth16.h"""
    _fields_ = [("key", ctypes.c_void_p), ("h_input", ctypes.c_void_p), ("input_bytes", ctypes.c_size_t),
                ("h_out", ctypes.c_void_p), ("out_bytes", ctypes.c_size_t), ("rank", ctypes.c_int),
                ("world", ctypes.c_int), ("rank_end", ctypes.c_int), ("d_h_coefficients", ctypes.c_void_p),
                ("b1_scaled", ctypes.c_int), ("query_spans", ctypes.POINTER(ctypes.c_int)), ("status", ctypes.c_int), ("timings", ProveTimings)]


_lib = None

_vp, _sz, _i, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
_SIGNATURES = {
    "b200_version": (ctypes.c_char_p, []),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_device_count": (_i, []),
    "b200_set_device": (_i, [_i]),
    "b200_sync": (_i, []),
    "b200_malloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "b200_free": (_i, [_vp]),
    "b200_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "b200_host_free": (_i, [_vp]),
    "b200_memcpy_h2d": (_i, [_vp, _vp, _sz]),
    "b200_memcpy_d2h": (_i, [_vp, _vp, _sz]),
    "b200_memcpy_d2d": (_i, [_vp, _vp, _sz]),
    "b200_memset_zero": (_i, [_vp, _sz]),
    "b200_fr_muleq": (_i, [_i, _vp, _vp, _sz]),
    "b200_fr_subeq": (_i, [_i, _vp, _vp, _sz]),
    "b200_domain_create": (_i, [_i, _sz, ctypes.POINTER(_vp)]),
    "b200_domain_destroy": (_i, [_vp]),
    "b200_domain_size": (_sz, [_vp]),
    "b200_domain_table": (_i, [_vp, _i, _vp, _sz]),
    "b200_domain_fft": (_i, [_vp, _vp]),
    "b200_domain_ifft": (_i, [_vp, _vp]),
    "b200_domain_coset_fft": (_i, [_vp, _vp]),
    "b200_domain_icoset_fft": (_i, [_vp, _vp]),
    "b200_domain_divide_by_z_on_coset": (_i, [_vp, _vp]),
    "b200_compute_h": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200_msm_g1": (_i, [_i, _vp, _vp, _sz, _vp]),
    "b200_msm_g2": (_i, [_i, _vp, _vp, _sz, _vp]),
    "b200_msm_ctx_create": (_i, [_i, _i, _vp, _sz, ctypes.POINTER(_vp)]),
    "b200_msm_ctx_run": (_i, [_vp, _vp, _vp]),
    "b200_msm_ctx_destroy": (_i, [_vp]),
    "b200_prove_batch": (_i, [_vp, _i]),
    "b200_prove_timeline": (_i, [_i, _vp]),
    "b200_host_equal_bases": (_i, [_vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b200_msm_set_window": (_i, [_i]),
    "b200_msm_set_batch_affine": (_i, [_i]),
    "b200_msm_get_batch_affine": (_i, []),
    "b200_msm_last_phase_ms": (_i, [ctypes.POINTER(ctypes.c_double)]),
    "b200_msm_last_plan": (_i, [ctypes.POINTER(ctypes.c_int)]),
    "b200_msm_phase_totals": (_i, [ctypes.POINTER(ctypes.c_double), _i]),
    "b200_launch_count": (ctypes.c_ulonglong, []),
    "b200_g1_add": (_i, [_i, _vp, _vp, _vp]),
    "b200_g2_add": (_i, [_i, _vp, _vp, _vp]),
    "b200_g1_scale": (_i, [_i, _vp, _vp, _vp]),
    "b200_g2_scale": (_i, [_i, _vp, _vp, _vp]),
    "b200_g1_to_affine": (_i, [_i, _vp, _vp]),
    "b200_g2_to_affine": (_i, [_i, _vp, _vp]),
    "b200_g1_from_affine": (_i, [_i, _vp, _vp]),
    "b200_g2_from_affine": (_i, [_i, _vp, _vp]),
    "b200_host_jacobian_doublings": (_i, [_i, _i, _vp, _i, _vp]),
    "b200_host_fp_op": (_i, [_i, _i, _vp, _vp, _vp]),
    "b200_params_from_host": (_i, [_i, _vp, _sz, ctypes.POINTER(_vp)]),
    "b200_params_from_file": (_i, [_i, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "b200_params_load_ms": (_i, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "b200_file_to_device": (_i, [ctypes.c_char_p, _sz, _vp, _sz]),
    "b200_params_domain": (_vp, [_vp]),
    "b200_params_from_device": (_i, [_i, _sz, _sz, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp)]),
    "b200_params_destroy": (_i, [_vp]),
    "b200_params_precompute": (_i, [_vp, _i, _i]),
    "b200_params_precompute_ms": (ctypes.c_double, [_vp]),
    "b200_params_warmup": (_i, [_vp]),
    "b200_set_precompute": (_i, [_i]),
    "b200_params_msm": (_i, [_vp, _i, _vp, _sz, _vp]),
    "b200_params_msm_async": (_i, [_vp, _i, _vp, _sz, _vp, ctypes.POINTER(_vp)]),
    "b200_msm_wait": (_i, [_vp]),
    "b200_params_d": (_sz, [_vp]),
    "b200_params_m": (_sz, [_vp]),
    "b200_params_query": (_vp, [_vp, _i]),
    "b200_prove": (_i, [_vp, _vp, _sz, _vp, ctypes.POINTER(_sz), ctypes.POINTER(ProveTimings)]),
    "b200_prove_partial": (_i, [_vp, _vp, _sz, _i, _i, _vp, ctypes.POINTER(_sz), ctypes.POINTER(ProveTimings)]),
    "b200_groth16_finalize": (_i, [_i, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_sz)]),
    "b200_prove_full": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, ctypes.POINTER(_sz)]),
    "b200_prove_partial_span": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, ctypes.POINTER(_sz), ctypes.POINTER(ProveTimings)]),
    "b200_params_precompute_span": (_i, [_vp, _i, _i, _i]),
    "b200_prove_partial_ext": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _vp, ctypes.POINTER(_sz), ctypes.POINTER(ProveTimings)]),
    "b200_prove_partial_scaled": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _vp, ctypes.POINTER(_sz), ctypes.POINTER(ProveTimings)]),
    "b200_prove_partial_queries": (_i, [_vp, _vp, _sz, ctypes.POINTER(_i), _i, _i, _vp, _vp, ctypes.POINTER(_sz),
                                        ctypes.POINTER(ProveTimings)]),
    "b200_params_precompute_queries": (_i, [_vp, ctypes.POINTER(_i), _i]),
    "b200_prove_combine": (_i, [_i, _vp, _i, _vp, _vp, ctypes.POINTER(_sz)]),
    "b200_dev_fp_op": (_i, [_i, _i, _vp, _vp, _vp, _sz]),
    "b200_dev_fqe_op": (_i, [_i, _i, _vp, _vp, _vp, _sz]),
    "b200_dev_group_op": (_i, [_i, _i, _i, _vp, _vp, _vp, _sz]),
    "b200_gen_points": (_i, [_i, _i, _vp, _sz, _u64]),
    "b200_batch_exp": (_i, [_i, _i, _vp, _vp, _sz, _vp, _i, ctypes.POINTER(ctypes.c_double)]),
    "b200_imad_peak": (_i, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load the CUDA library (once). Raises B200Error if it has not been built - there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error("%s is missing: run `python -m snark_challenge_prover_reference_b200.build` "
                            "(the CUDA library is the product; there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error("b200 call failed (%d): %s" % (rc, lib().b200_last_error().decode()))


def _ptr(x):
    """Address of a bytes-like / ctypes buffer / torch tensor / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(x)) if isinstance(x, bytes) else (ctypes.c_char * len(x)).from_buffer(x),
                           ctypes.c_void_p).value
    return ctypes.addressof(x)


# ---------------------------------------------------------------------------------------------------------------
# host helpers (no GPU needed): the serial tail of the prover
def host_fp_op(tag, op, a, b=None):
    out = ctypes.create_string_buffer(FE)
    ab = ctypes.create_string_buffer(bytes(a), FE)
    bb = ctypes.create_string_buffer(bytes(b), FE) if b is not None else None
    check(lib().b200_host_fp_op(tag, op, ctypes.addressof(ab), ctypes.addressof(bb) if bb else None,
                                ctypes.addressof(out)))
    return out.raw


def _host_group(fname, curve, group, *bufs, out_len):
    out = ctypes.create_string_buffer(out_len)
    keep = [ctypes.create_string_buffer(bytes(b), len(b)) for b in bufs]
    fn = getattr(lib(), fname % ("g1" if group == 1 else "g2"))
    check(fn(curve, *[ctypes.addressof(k) for k in keep], ctypes.addressof(out)))
    return out.raw


def g_add(curve, group, p, q):
    return _host_group("b200_%s_add", curve, group, p, q, out_len=proj_bytes(curve, group))


def g_scale(curve, group, fr, p):
    return _host_group("b200_%s_scale", curve, group, fr, p, out_len=proj_bytes(curve, group))


def g_to_affine(curve, group, p):
    return _host_group("b200_%s_to_affine", curve, group, p, out_len=affine_bytes(curve, group))


def g_from_affine(curve, group, xy):
    return _host_group("b200_%s_from_affine", curve, group, xy, out_len=proj_bytes(curve, group))


# ---------------------------------------------------------------------------------------------------------------
# device helpers (torch uint8 tensors as device memory)
def to_device(data, device="cuda:0"):
    import torch
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8) if len(data) else torch.zeros(0, dtype=torch.uint8)
    return t.to(device)


def from_device(t):
    return t.cpu().numpy().tobytes()


def msm(curve, group, d_scalars, d_points, n):
    """sum_i scalars[i]*points[i] -> projective point bytes (host). d_* are torch CUDA tensors or raw addresses."""
    out = ctypes.create_string_buffer(proj_bytes(curve, group))
    fn = lib().b200_msm_g1 if group == 1 else lib().b200_msm_g2
    check(fn(curve, _ptr(d_scalars), _ptr(d_points), n, ctypes.addressof(out)))
    return out.raw


def msm_phase_ms():
    arr = (ctypes.c_double * 5)()
    check(lib().b200_msm_last_phase_ms(arr))
    return dict(zip(("digits", "sort", "accumulate", "reduce", "host_tail"), list(arr)))


def msm_last_plan():
    arr = (ctypes.c_int * 3)()
    check(lib().b200_msm_last_plan(arr))
    return {"c": arr[0], "windows": arr[1], "task_len": arr[2]}


def msm_phase_totals(reset=False):
    arr = (ctypes.c_double * 10)()
    check(lib().b200_msm_phase_totals(arr, 1 if reset else 0))
    names = ("digits", "sort", "accumulate", "reduce", "host_tail")
    return {"g1": dict(zip(names, list(arr)[:5])), "g2": dict(zip(names, list(arr)[5:]))}


def launch_count():
    return int(lib().b200_launch_count())


class MsmContext:
    """Pre-shifted base tables for an arbitrary point set (b200_msm_ctx_*): create once, run per scalar vector."""

    def __init__(self, curve, group, d_points, n):
        self.curve, self.group, self.n = curve, group, n
        self.h = ctypes.c_void_p()
        check(lib().b200_msm_ctx_create(curve, group, _ptr(d_points), n, ctypes.byref(self.h)))

    def run(self, d_scalars):
        out = ctypes.create_string_buffer(proj_bytes(self.curve, self.group))
        check(lib().b200_msm_ctx_run(self.h, _ptr(d_scalars), ctypes.addressof(out)))
        return out.raw

    def close(self):
        if self.h and _lib is not None:
            _lib.b200_msm_ctx_destroy(self.h)
        self.h = None

    __del__ = close


class Domain:
    """basic radix-2 evaluation domain over Fr of `curve` (B::get_evaluation_domain)."""

    h = None

    def __init__(self, curve, m):
        h = ctypes.c_void_p()
        check(lib().b200_domain_create(curve, m, ctypes.byref(h)))
        self.h, self.curve, self.m = h, curve, m

    def close(self):
        if self.h and lib is not None and _lib is not None:
            _lib.b200_domain_destroy(self.h)
        self.h = None

    __del__ = close

    def table(self, which, count):
        out = ctypes.create_string_buffer(count * FE)
        check(lib().b200_domain_table(self.h, which, ctypes.addressof(out), count))
        return out.raw

    def fft(self, a):
        check(lib().b200_domain_fft(self.h, _ptr(a)))

    def ifft(self, a):
        check(lib().b200_domain_ifft(self.h, _ptr(a)))

    def coset_fft(self, a):
        check(lib().b200_domain_coset_fft(self.h, _ptr(a)))

    def icoset_fft(self, a):
        check(lib().b200_domain_icoset_fft(self.h, _ptr(a)))

    def divide_by_z_on_coset(self, a):
        check(lib().b200_domain_divide_by_z_on_coset(self.h, _ptr(a)))

    def compute_h(self, ca, cb, cc, out):
        check(lib().b200_compute_h(self.h, _ptr(ca), _ptr(cb), _ptr(cc), _ptr(out)))


class Params:
    """Proving key resident on the current device (B::read_params)."""

    def __init__(self, curve, handle, keep=None):
        self.curve, self.h, self._keep = curve, handle, keep

    @classmethod
    def from_bytes(cls, curve, image):
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(bytes(image), len(image))
        check(lib().b200_params_from_host(curve, ctypes.addressof(buf), len(image), ctypes.byref(h)))
        return cls(curve, h)

    @classmethod
    def from_file(cls, curve, path):
        """B::read_params: load a parameter file (chunked reads into pinned memory, asynchronous H2D)"""
        h = ctypes.c_void_p()
        check(lib().b200_params_from_file(curve, os.fsencode(path), ctypes.byref(h)))
        return cls(curve, h)

    def load_ms(self):
        """{read, copy_wait, total} milliseconds of from_file"""
        out = (ctypes.c_double * 3)()
        check(lib().b200_params_load_ms(self.h, out))
        return {"read": out[0], "copy_wait": out[1], "total": out[2]}

    @classmethod
    def from_device(cls, curve, d, m, A, B1, B2, L, H):
        h = ctypes.c_void_p()
        check(lib().b200_params_from_device(curve, d, m, _ptr(A), _ptr(B1), _ptr(B2), _ptr(L), _ptr(H), ctypes.byref(h)))
        return cls(curve, h, keep=(A, B1, B2, L, H))

    @property
    def d(self):
        return lib().b200_params_d(self.h)

    @property
    def m(self):
        return lib().b200_params_m(self.h)

    def close(self):
        if self.h and lib is not None and _lib is not None:
            _lib.b200_params_destroy(self.h)
        self.h = None

    __del__ = close

    def precompute(self, rank=0, world=1, rank_end=None):
        """build the pre-shifted base tables for this rank's slice - or run of slices [rank, rank_end) - (key-only
        preprocessing); returns seconds"""
        check(lib().b200_params_precompute_span(self.h, rank, rank + 1 if rank_end is None else rank_end, world))
        return lib().b200_params_precompute_ms(self.h) / 1e3

    def msm(self, which, d_scalars, n):
        """MSM over one whole query (0 A, 1 B1, 2 B2, 3 L, 4 H) -> projective point bytes"""
        out = ctypes.create_string_buffer(proj_bytes(self.curve, 2 if which == 2 else 1))
        check(lib().b200_params_msm(self.h, which, _ptr(d_scalars), n, ctypes.addressof(out)))
        return out.raw

    def prove(self, input_image, timings=False):
        """One whole proof from a HOST input image; returns the proof bytes (A | B | C, wire format)."""
        out = ctypes.create_string_buffer(proof_bytes(self.curve))
        n = ctypes.c_size_t()
        tm = ProveTimings()
        check(lib().b200_prove(self.h, _ptr(input_image), _len(input_image), ctypes.addressof(out), ctypes.byref(n),
                               ctypes.byref(tm)))
        return (out.raw[:n.value], tm.as_dict()) if timings else out.raw[:n.value]

    def prove_full(self, input_image, s_fr, extras):
        """one proof with the complete Groth16 terms (b200_prove_full)"""
        out = ctypes.create_string_buffer(proof_bytes(self.curve))
        n = ctypes.c_size_t()
        sb, eb = ctypes.create_string_buffer(bytes(s_fr), len(s_fr)), ctypes.create_string_buffer(bytes(extras), len(extras))
        check(lib().b200_prove_full(self.h, _ptr(input_image), _len(input_image), ctypes.addressof(sb), ctypes.addressof(eb),
                                    ctypes.addressof(out), ctypes.byref(n)))
        return out.raw[:n.value]

    def prove_partial_queries(self, input_image, spans, world, d_h=None, b1_scaled=False):
        """partial sums with a run of slices PER QUERY: spans = [(first, end)] * 5 for A, B1, B2, L, H in units of 1/world
        (b200_prove_partial_queries); an empty run leaves O in that slot"""
        out = ctypes.create_string_buffer(partial_bytes(self.curve))
        n = ctypes.c_size_t()
        tm = ProveTimings()
        check(lib().b200_prove_partial_queries(self.h, _ptr(input_image), _len(input_image), _spans(spans), world,
                                               1 if b1_scaled else 0, _ptr(d_h), ctypes.addressof(out), ctypes.byref(n),
                                               ctypes.byref(tm)))
        return out.raw[:n.value], tm.as_dict()

    def precompute_queries(self, spans, world):
        check(lib().b200_params_precompute_queries(self.h, _spans(spans), world))
        return lib().b200_params_precompute_ms(self.h) / 1e3

    def prove_partial(self, input_image, rank, world, rank_end=None, d_h=None, b1_scaled=False):
        """partial sums over slice `rank` - or the run of slices [rank, rank_end) - of `world`; d_h: device vector of
        H coefficients computed elsewhere (b200_prove_partial_ext); b1_scaled: the B1 slot holds r * (the rank's B1 sum)
        (b200_prove_partial_scaled; combine with r_fr=None)"""
        out = ctypes.create_string_buffer(partial_bytes(self.curve))
        n = ctypes.c_size_t()
        tm = ProveTimings()
        fn = lib().b200_prove_partial_scaled if b1_scaled else lib().b200_prove_partial_ext
        check(fn(self.h, _ptr(input_image), _len(input_image), rank,
                                           rank + 1 if rank_end is None else rank_end, world, _ptr(d_h),
                                           ctypes.addressof(out), ctypes.byref(n), ctypes.byref(tm)))
        return out.raw[:n.value], tm.as_dict()


def _spans(spans):
    flat = [int(v) for pair in spans for v in pair]
    assert len(flat) == 10, "spans: five (first, end) pairs, for A, B1, B2, L, H"
    return (ctypes.c_int * 10)(*flat)


def _len(x):
    if hasattr(x, "numel"):
        return x.numel() * x.element_size()
    return len(x)


def prove_combine(curve, partials_all, world, r_fr):
    out = ctypes.create_string_buffer(proof_bytes(curve))
    n = ctypes.c_size_t()
    pb = ctypes.create_string_buffer(bytes(partials_all), len(partials_all))
    rb = ctypes.create_string_buffer(bytes(r_fr), FE) if r_fr is not None else None   # None: B1 slots already scaled
    check(lib().b200_prove_combine(curve, ctypes.addressof(pb), world, ctypes.addressof(rb) if rb else None,
                                   ctypes.addressof(out), ctypes.byref(n)))
    return out.raw[:n.value]


def batch_exp(curve, group, base_affine, d_scalars, n, d_out, window=0):
    """out[i] = scalars[i] * base (fixed-base windowed exponentiation, libff::batch_exp); returns phase times in ms"""
    ms = (ctypes.c_double * 3)()
    b = ctypes.create_string_buffer(bytes(base_affine), len(base_affine))
    check(lib().b200_batch_exp(curve, group, ctypes.addressof(b), _ptr(d_scalars), n, _ptr(d_out), window, ms))
    return {"table": ms[0], "exp": ms[1], "to_affine": ms[2]}


def groth16_finalize(curve, proof, r_fr, s_fr, extras):
    """challenge proof (A | B | C) -> complete Groth16 proof with the alpha / beta / delta / s terms (host only)"""
    out = ctypes.create_string_buffer(proof_bytes(curve))
    n = ctypes.c_size_t()
    bufs = [ctypes.create_string_buffer(bytes(x), len(x)) for x in (proof, r_fr, s_fr, extras)]
    check(lib().b200_groth16_finalize(curve, *[ctypes.addressof(b) for b in bufs], ctypes.addressof(out), ctypes.byref(n)))
    return out.raw[:n.value]


def prove_batch(jobs, timings=False, b1_scaled=False):
    """Several proofs in flight at once on the current device (b200_prove_batch). jobs: sequence of
    (Params, host_input_image) for whole proofs or (Params, host_input_image, rank, world) for one rank's partial
    sums. Returns the list of proof (or partial-sum) byte strings, in job order."""
    arr = (ProofJob * len(jobs))()
    outs, keep = [], []
    for a, job in zip(arr, jobs):
        key, image = job[0], job[1]
        rank, world = (job[2], job[3]) if len(job) > 2 else (0, 1)
        a.rank_end = job[4] if len(job) > 4 else 0   # (key, image, rank, world, rank_end): a run of slices
        a.d_h_coefficients = _ptr(job[5]) if len(job) > 5 and job[5] is not None else None   # witness map done elsewhere
        out = ctypes.create_string_buffer(max(proof_bytes(key.curve), partial_bytes(key.curve)))
        outs.append(out)
        a.key = key.h.value if hasattr(key.h, "value") else key.h
        a.h_input, a.input_bytes = _ptr(image), _len(image)
        a.h_out, a.rank, a.world = ctypes.addressof(out), rank, world
        a.b1_scaled = 1 if b1_scaled else 0   # partial sums with r * B1 in the B1 slot (b200_prove_partial_scaled)
        if len(job) > 6 and job[6] is not None:   # (.., rank_end, d_h, spans): per-query runs
            keep.append(_spans(job[6]))
            a.query_spans = keep[-1]
    check(lib().b200_prove_batch(ctypes.addressof(arr), len(jobs)))
    res = [o.raw[:a.out_bytes] for o, a in zip(outs, arr)]
    return (res, [a.timings.as_dict() for a in arr]) if timings else res


def host_equal_bases(points, n, point_bytes):
    """Host-side grouping of equal bases (test hook, needs no GPU): returns (merged, [[rep, member, ...], ...])."""
    cap = ctypes.c_size_t(n)
    gcap = ctypes.c_size_t(3 * n)
    members = (ctypes.c_uint32 * max(n, 1))()
    groups = (ctypes.c_uint32 * max(3 * n, 1))()
    merged = ctypes.c_size_t()
    buf = ctypes.create_string_buffer(bytes(points), len(points))
    check(lib().b200_host_equal_bases(ctypes.addressof(buf), n, point_bytes, ctypes.addressof(members), ctypes.byref(cap),
                                      ctypes.addressof(groups), ctypes.byref(gcap), ctypes.byref(merged)))
    # every group's members are contiguous in `members`, representative first, groups in the order of `groups`
    reps = [groups[3 * g] for g in range(gcap.value // 3)]
    out = []
    for i in members[:cap.value]:
        if len(out) < len(reps) and i == reps[len(out)]:
            out.append([i])
        else:
            out[-1].append(i)
    return merged.value, out


def prove_timeline(begin=False):
    """Diagnostics (b200_prove_timeline): begin=True marks t = 0; afterwards returns, per MSM in issue order
    (B2, A, B1, L, H), the ms at which accumulation started, reduction started and reduction ended."""
    if begin:
        check(lib().b200_prove_timeline(1, None))
        return None
    out = (ctypes.c_double * 15)()
    check(lib().b200_prove_timeline(0, ctypes.addressof(out)))
    return {name: [round(out[i * 3 + k], 2) for k in range(3)] for i, name in enumerate(("B2", "A", "B1", "L", "H"))}


def set_batch_affine(mode):
    """Select the bucket accumulation of the MSMs: 0 / False XYZZ mixed additions, 1 / True batched affine additions,
    2 automatic (the library's default)."""
    check(lib().b200_msm_set_batch_affine(int(mode)))


def set_precompute(on):
    check(lib().b200_set_precompute(1 if on else 0))


def batch_affine_mode():
    return {0: "xyzz", 1: "affine", 2: "auto"}[lib().b200_msm_get_batch_affine()]


def imad_peak():
    v = (ctypes.c_double * 4)()
    ms = (ctypes.c_double * 4)()
    check(lib().b200_imad_peak(v, ms))
    return {"mad_wide_mac32_per_s": v[0], "carry_chain_mac32_per_s": v[1], "nominal_mac32_per_s": v[2],
            "montgomery_mul_mac32_per_s": v[3], "ms": [ms[0], ms[1], ms[3]], "sm_clock_mhz": ms[2]}
