"""Generator + CPU interpreter for the 753-bit Montgomery multiplication PTX used by the sm_100a kernels.

The product a*b*2^-768 mod p is computed with 24 x 32-bit limbs as an interleaved (CIOS-style) Montgomery
multiplication in which every 32x32->64 partial product is accumulated into a 64-bit-aligned register pair by a
(mad.lo.cc, madc.hi.cc) pair - the pattern ptxas fuses into ONE IMAD.WIDE.U32(.X) - using two accumulator arrays:
X covers limb positions (0,1)(2,3)... and Y covers (1,2)(3,4)...; after each row the arrays swap roles, which makes
the divide-by-2^32 of the Montgomery step free (SURVEY.md 8a1 says what is computed: fp.tcc:161-186).

Because there is no GPU in the build container, this module also contains a small interpreter for exactly the PTX
subset it emits, so tests/test_ptx_model.py can run the *generated text* on random inputs against Python ints.

    python tools/gen_fp_ptx.py            -> writes csrc/fp_ptx_gen.cuh
"""
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
from mnt753 import PRIMES, LIMBS32, to_limbs32  # noqa: E402

N = LIMBS32
MASK = 0xFFFFFFFF


def _inv32(p):
    return (-pow(p, -1, 1 << 32)) & MASK


class Emitter:
    def __init__(self):
        self.lines = []

    def op(self, s):
        self.lines.append(s)


def gen_mul_body(p, sqr=False):
    """Return list of PTX lines. Inputs: a0..a23, b0..b23 (for sqr b aliases a). Outputs r0..r23.
    Scratch: P0..P23, Q0..Q23, m, z (zero), t0..t23, brw, pr (pred)."""
    pl = to_limbs32(p)
    inv = _inv32(p)
    e = Emitter()
    A = [f"a{j}" for j in range(N)]
    B = [f"a{j}" for j in range(N)] if sqr else [f"b{j}" for j in range(N)]
    P = [f"P{j}" for j in range(N)]
    Q = [f"Q{j}" for j in range(N)]
    e.op("mov.u32 z, 0;")

    def reduce_row(X, Y):
        # m = X[0] * inv ; Y += p_odd*m ; X += p_even*m ; carry of X chain -> Y[N-1]
        e.op(f"mul.lo.u32 m, {X[0]}, {inv};")
        for k in range(N // 2):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            hi = "madc.hi.cc.u32" if k < N // 2 - 1 else "madc.hi.u32"
            e.op(f"{lo} {Y[2*k]}, m, {pl[2*k+1]}, {Y[2*k]};")
            e.op(f"{hi} {Y[2*k+1]}, m, {pl[2*k+1]}, {Y[2*k+1]};")
        for k in range(N // 2):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            e.op(f"{lo} {X[2*k]}, m, {pl[2*k]}, {X[2*k]};")
            e.op(f"madc.hi.cc.u32 {X[2*k+1]}, m, {pl[2*k]}, {X[2*k+1]};")
        e.op(f"addc.u32 {Y[N-1]}, {Y[N-1]}, 0;")

    # ---- row 0: X = a_even*b0, Y = a_odd*b0
    X, Y = P, Q
    for k in range(N // 2):
        e.op(f"mul.lo.u32 {X[2*k]}, {A[2*k]}, {B[0]};")
        e.op(f"mul.hi.u32 {X[2*k+1]}, {A[2*k]}, {B[0]};")
    for k in range(N // 2):
        e.op(f"mul.lo.u32 {Y[2*k]}, {A[2*k+1]}, {B[0]};")
        e.op(f"mul.hi.u32 {Y[2*k+1]}, {A[2*k+1]}, {B[0]};")
    reduce_row(X, Y)

    # ---- rows 1..N-1
    for i in range(1, N):
        Xo, Yo = X, Y
        X, Y = Yo, Xo  # role swap == divide by 2^32 ; new Y is built in place from old X shifted down one pair
        e.op(f"add.cc.u32 {X[0]}, {X[0]}, {Xo[1]};")
        for k in range(N // 2 - 1):
            e.op(f"madc.lo.cc.u32 {Y[2*k]}, {A[2*k+1]}, {B[i]}, {Xo[2*k+2]};")
            e.op(f"madc.hi.cc.u32 {Y[2*k+1]}, {A[2*k+1]}, {B[i]}, {Xo[2*k+3]};")
        e.op(f"madc.lo.cc.u32 {Y[N-2]}, {A[N-1]}, {B[i]}, z;")
        e.op(f"madc.hi.u32 {Y[N-1]}, {A[N-1]}, {B[i]}, z;")
        for k in range(N // 2):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            e.op(f"{lo} {X[2*k]}, {A[2*k]}, {B[i]}, {X[2*k]};")
            e.op(f"madc.hi.cc.u32 {X[2*k+1]}, {A[2*k]}, {B[i]}, {X[2*k+1]};")
        e.op(f"addc.u32 {Y[N-1]}, {Y[N-1]}, 0;")
        reduce_row(X, Y)

    # ---- merge: value/2^32 = Y[j] + X[j+1]
    T = [f"t{j}" for j in range(N)]
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        src = X[j + 1] if j + 1 < N else "z"
        e.op(f"{opn} {Y[j]}, {Y[j]}, {src};")
    # ---- conditional subtract p
    for j in range(N):
        opn = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        e.op(f"{opn} {T[j]}, {Y[j]}, {pl[j]};")
    e.op("subc.u32 brw, z, z;")
    e.op("setp.ne.u32 pr, brw, 0;")
    for j in range(N):
        e.op(f"selp.u32 r{j}, {Y[j]}, {T[j]}, pr;")
    return e.lines



def gen_add_body(p):
    """r = a + b mod p (inputs canonical)."""
    pl = to_limbs32(p)
    L = ["mov.u32 z, 0;"]
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        L.append(f"{opn} t{j}, a{j}, b{j};")
    for j in range(N):
        opn = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        L.append(f"{opn} P{j}, t{j}, {pl[j]};")
    L.append("subc.u32 brw, z, z;")
    L.append("setp.ne.u32 pr, brw, 0;")
    for j in range(N):
        L.append(f"selp.u32 r{j}, t{j}, P{j}, pr;")
    return L


def gen_sub_body(p):
    """r = a - b mod p (inputs canonical)."""
    pl = to_limbs32(p)
    L = ["mov.u32 z, 0;"]
    for j in range(N):
        opn = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        L.append(f"{opn} t{j}, a{j}, b{j};")
    L.append("subc.u32 brw, z, z;")
    for j in range(N):
        L.append(f"and.b32 P{j}, brw, {pl[j]};")
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        L.append(f"{opn} r{j}, t{j}, P{j};")
    return L


# --------------------------------------------------------------------------- interpreter (subset)
def run_ptx(lines, regs):
    """Interpret the emitted PTX subset. regs: dict name->int (inputs). Returns regs. One CC.CF flag, as in PTX."""
    cf = 0
    R = dict(regs)

    def val(tok):
        tok = tok.strip()
        if tok[0].isdigit():
            return int(tok) & MASK
        return R[tok]

    for ln in lines:
        ln = ln.strip().rstrip(";")
        opc, rest = ln.split(None, 1)
        args = [x.strip() for x in rest.split(",")]
        d = args[0]
        if opc == "mov.u32":
            R[d] = val(args[1])
        elif opc == "mul.lo.u32":
            R[d] = (val(args[1]) * val(args[2])) & MASK
        elif opc == "mul.hi.u32":
            R[d] = (val(args[1]) * val(args[2])) >> 32
        elif opc in ("mad.lo.cc.u32", "madc.lo.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "mad.hi.cc.u32",
                     "madc.lo.u32", "mad.lo.u32", "mad.hi.u32"):
            prod = val(args[1]) * val(args[2])
            part = (prod & MASK) if ".lo" in opc else (prod >> 32)
            cin = cf if opc.startswith("madc") else 0
            s = part + val(args[3]) + cin
            R[d] = s & MASK
            if ".cc" in opc:
                cf = s >> 32
        elif opc in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
            cin = cf if opc.startswith("addc") else 0
            s = val(args[1]) + val(args[2]) + cin
            R[d] = s & MASK
            if ".cc" in opc:
                cf = s >> 32
        elif opc in ("sub.cc.u32", "subc.cc.u32", "subc.u32", "sub.u32"):
            cin = cf if opc.startswith("subc") else 0
            s = val(args[1]) - val(args[2]) - cin
            R[d] = s & MASK
            if ".cc" in opc:
                cf = 1 if s < 0 else 0
        elif opc == "and.b32":
            R[d] = val(args[1]) & val(args[2])
        elif opc == "setp.ne.u32":
            R[d] = 1 if val(args[1]) != val(args[2]) else 0
        elif opc == "selp.u32":
            R[d] = val(args[1]) if R[args[3]] else val(args[2])
        else:
            raise ValueError("unsupported op " + opc)
    return R


def model_mul(lines, a, b):
    regs = {}
    for j, v in enumerate(to_limbs32(a)):
        regs[f"a{j}"] = v
    for j, v in enumerate(to_limbs32(b)):
        regs[f"b{j}"] = v
    out = run_ptx(lines, regs)
    return sum(out[f"r{j}"] << (32 * j) for j in range(N))


# --------------------------------------------------------------------------- C++ emission
def emit_function(name, lines, sqr=False, nk=0):
    """One asm statement; operands: %0..%23 = r (out), %24..%47 = a, %48..%71 = b."""
    sub = {}
    for j in range(N):
        sub[f"r{j}"] = f"%{j}"
        sub[f"a{j}"] = f"%{N + j}"
        sub[f"b{j}"] = f"%{2 * N + j}"

    def tr(ln):
        opc, rest = ln.rstrip(";").split(None, 1)
        args = [x.strip() for x in rest.split(",")]
        args = [sub.get(x, x) for x in args]
        return f"{opc} {', '.join(args)};"

    decl = ".reg .u32 P<24>, Q<24>, t<24>, m, z, brw;" if nk == 0 else ".reg .u32 k<%d>, m, brw;" % nk
    body = ["{", decl, ".reg .pred pr;"] + [tr(l) for l in lines] + ["}"]
    s = []
    if sqr:
        s.append(f"__device__ __forceinline__ void {name}(uint32_t (&r)[24], const uint32_t (&a)[24]) {{")
    else:
        s.append(f"__device__ __forceinline__ void {name}(uint32_t (&r)[24], const uint32_t (&a)[24], const uint32_t (&b)[24]) {{")
    s.append("  asm(")
    for l in body:
        s.append(f'    "{l}\\n\\t"')
    outs = ", ".join(f'"=r"(r[{j}])' for j in range(N))
    ins = ", ".join(f'"r"(a[{j}])' for j in range(N))
    if not sqr:
        ins += ", " + ", ".join(f'"r"(b[{j}])' for j in range(N))
    s.append(f"    : {outs}")
    s.append(f"    : {ins});")
    s.append("}")
    return "\n".join(s)


def main():
    out = os.path.join(os.path.dirname(__file__), "..", "snark_challenge_prover_reference_b200", "csrc", "fp_ptx_gen.cuh")
    parts = ["// GENERATED by tools/gen_fp_ptx.py - do not edit. 753-bit Montgomery multiplication, 24x32-bit limbs,",
             "// IMAD.WIDE-friendly (mad.lo.cc/madc.hi.cc pairs) with moduli and -p^-1 as immediates.",
             "#pragma once", "#include <stdint.h>", ""]
    for tag, p in PRIMES.items():
        parts.append(emit_function(f"fp_mul_ptx_{tag}", gen_mul_body(p)))
        parts.append("")
        if "--experimental" in sys.argv:
            # measured and rejected (kept in the generator with its interpreter tests, not emitted): one-level Karatsuba
            # (round 1: -12 % wide MACs, multiplier pipe 94 % -> 84 % busy, no net gain)
            kl, nk = gen_mul_karatsuba_body(p)
            parts.append(emit_function(f"fp_mulk_ptx_{tag}", kl, nk=nk))
            parts.append("")
        # dedicated squaring, 876 instead of 1152 wide MACs. NOT used by the bucket accumulation (round 2: squarings are
        # 2 of its 10 multiplications; G1 49.5 -> 49.0 ms, G2 160.7 -> 162.0 ms, noise), but by the base-table builder,
        # whose Jacobian doublings are 8 squarings + 1 multiplication (Fp::sqr_fast).
        sl, ns = gen_sqr_body(p)
        parts.append(emit_function(f"fp_sqr_ptx_{tag}", sl, sqr=True, nk=ns))
        parts.append("")
        parts.append(emit_function(f"fp_add_ptx_{tag}", gen_add_body(p)))
        parts.append("")
        parts.append(emit_function(f"fp_sub_ptx_{tag}", gen_sub_body(p)))
        parts.append("")
    with open(out, "w") as f:
        f.write("\n".join(parts))
    print("wrote", os.path.normpath(out))




# --------------------------------------------------------------------------- Karatsuba variant
def gen_mul_karatsuba_body(p):
    """a*b*2^-768 mod p with a one-level Karatsuba product (3 x 12x12 limbs = 432 wide MACs instead of 576) followed
    by a Montgomery reduction that runs the CIOS row machinery on the LOW half of the product only and adds the high
    half at the end:  T = a*b (48 limbs);  U = (T_lo + sum_i m_i p 2^(32 i)) / 2^768;  r = U + T_hi  (< 2p).
    Inputs a0..a23, b0..b23; outputs r0..r23. All intermediate names are PTX virtual registers (the caller declares
    them with .reg .u32 k<400>)."""
    pl = to_limbs32(p)
    inv = _inv32(p)
    L = []
    cnt = [0]

    def new(n=1):
        regs = [f"k{cnt[0] + i}" for i in range(n)]
        cnt[0] += n
        return regs if n > 1 else regs[0]

    z = new()
    L.append(f"mov.u32 {z}, 0;")
    H = N // 2  # 12

    def mul_half(u, v, hu=H):
        """full product of two hu-limb numbers -> 2*hu limbs, even/odd accumulation (sppark-style wide multiply)."""
        n2 = 2 * hu
        E = [None] * (n2 + 2)  # value at limb position k (even-aligned pairs (0,1),(2,3),...)
        O = [None] * (n2 + 2)  # odd-aligned pairs (1,2),(3,4),... : O[k] holds limb position k
        for i in range(hu):
            for par in (0, 1):  # par 0: u_even * v_i ; par 1: u_odd * v_i
                idxs = list(range(par, hu, 2))
                first = idxs[0] + i
                arr = E if first % 2 == 0 else O
                started = False
                for j in idxs:
                    pos = i + j
                    lo_init = arr[pos] is None
                    hi_init = arr[pos + 1] is None
                    if lo_init:
                        arr[pos] = new()
                    if hi_init:
                        arr[pos + 1] = new()
                    # low word
                    if lo_init and not started:
                        L.append(f"mul.lo.u32 {arr[pos]}, {u[j]}, {v[i]};")
                        carry_live = False
                    elif lo_init:
                        L.append(f"madc.lo.cc.u32 {arr[pos]}, {u[j]}, {v[i]}, {z};")
                        carry_live = True
                    elif not started:
                        L.append(f"mad.lo.cc.u32 {arr[pos]}, {u[j]}, {v[i]}, {arr[pos]};")
                        carry_live = True
                    else:
                        L.append(f"madc.lo.cc.u32 {arr[pos]}, {u[j]}, {v[i]}, {arr[pos]};")
                        carry_live = True
                    # high word
                    addend = z if hi_init else arr[pos + 1]
                    if carry_live:
                        L.append(f"madc.hi.cc.u32 {arr[pos+1]}, {u[j]}, {v[i]}, {addend};")
                    elif hi_init:
                        L.append(f"mul.hi.u32 {arr[pos+1]}, {u[j]}, {v[i]};")
                        # no carry defined yet: make the flag well defined for the next madc
                        L.append(f"add.cc.u32 {arr[pos+1]}, {arr[pos+1]}, 0;")
                    else:
                        L.append(f"mad.hi.cc.u32 {arr[pos+1]}, {u[j]}, {v[i]}, {arr[pos+1]};")
                    started = True
                # carry out of the chain goes to the next word of the same array
                top = i + idxs[-1] + 2
                if top < n2 + 1:
                    if arr[top] is None:
                        arr[top] = new()
                        L.append(f"addc.u32 {arr[top]}, {z}, 0;")
                    else:
                        L.append(f"addc.u32 {arr[top]}, {arr[top]}, 0;")
        # merge: result = E + O
        res = new(n2)
        for k in range(n2):
            ek = E[k] if E[k] is not None else z
            ok = O[k] if O[k] is not None else z
            opn = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n2 - 1 else "addc.u32")
            L.append(f"{opn} {res[k]}, {ek}, {ok};")
        return res

    aL, aH = [f"a{j}" for j in range(H)], [f"a{j}" for j in range(H, N)]
    bL, bH = [f"b{j}" for j in range(H)], [f"b{j}" for j in range(H, N)]
    # sa = aL + aH, sb = bL + bH (12 limbs + carry bit as a 0 / 0xffffffff mask)
    sa, sb = new(H), new(H)
    ca, cb = new(), new()
    for (s, lo, hi, c) in ((sa, aL, aH, ca), (sb, bL, bH, cb)):
        for k in range(H):
            opn = "add.cc.u32" if k == 0 else "addc.cc.u32"
            L.append(f"{opn} {s[k]}, {lo[k]}, {hi[k]};")
        L.append(f"addc.u32 {c}, {z}, 0;")
        L.append(f"sub.u32 {c}, {z}, {c};")  # 0 -> 0, 1 -> 0xffffffff
    zm24 = mul_half(sa, sb)
    # zm (25 limbs) = sa_lo*sb_lo + (ca ? sb_lo : 0)<<384 + (cb ? sa_lo : 0)<<384 + (ca&cb)<<768
    zm = zm24 + [new()]
    t = new(H)
    for k in range(H):
        L.append(f"and.b32 {t[k]}, {sb[k]}, {ca};")
    for k in range(H):
        opn = "add.cc.u32" if k == 0 else "addc.cc.u32"
        L.append(f"{opn} {zm[H+k]}, {zm[H+k]}, {t[k]};")
    L.append(f"addc.u32 {zm[2*H]}, {z}, 0;")
    for k in range(H):
        L.append(f"and.b32 {t[k]}, {sa[k]}, {cb};")
    for k in range(H):
        opn = "add.cc.u32" if k == 0 else "addc.cc.u32"
        L.append(f"{opn} {zm[H+k]}, {zm[H+k]}, {t[k]};")
    L.append(f"addc.u32 {zm[2*H]}, {zm[2*H]}, 0;")
    cc = new()
    L.append(f"and.b32 {cc}, {ca}, {cb};")
    L.append(f"and.b32 {cc}, {cc}, 1;")
    L.append(f"add.u32 {zm[2*H]}, {zm[2*H]}, {cc};")
    z0 = mul_half(aL, bL)
    z2 = mul_half(aH, bH)
    # z1 = zm - z0 - z2 (25 limbs, non-negative)
    for sub in (z0, z2):
        for k in range(2 * H):
            opn = "sub.cc.u32" if k == 0 else "subc.cc.u32"
            L.append(f"{opn} {zm[k]}, {zm[k]}, {sub[k]};")
        L.append(f"subc.u32 {zm[2*H]}, {zm[2*H]}, 0;")
    # T = z0 + z1 << 384 + z2 << 768  (48 limbs): T[0..11] = z0[0..11]; T[12..35] = (z0[12..23] | z2[0..11]) + z1[0..23];
    # T[36..47] = z2[12..23] + z1[24] + carry
    T = z0[:H] + [None] * (3 * H)
    mid = z0[H:] + z2[:H]
    for k in range(2 * H):
        opn = "add.cc.u32" if k == 0 else "addc.cc.u32"
        L.append(f"{opn} {mid[k]}, {mid[k]}, {zm[k]};")
        T[H + k] = mid[k]
    for k in range(H):
        src = zm[2 * H] if k == 0 else z
        opn = "addc.cc.u32" if k < H - 1 else "addc.u32"
        L.append(f"{opn} {z2[H+k]}, {z2[H+k]}, {src};")
        T[3 * H + k] = z2[H + k]
    # ---- Montgomery reduction of T_lo with the even/odd row machinery; window X (positions 0..23), Y (1..24)
    X = list(T[:N])
    Y = new(N)
    for k in range(N):
        L.append(f"mov.u32 {Y[k]}, 0;")
    for i in range(N):
        if i > 0:
            Xo, Yo = X, Y
            X = Yo
            fresh = new(2)
            L.append(f"mov.u32 {fresh[0]}, 0;")
            L.append(f"mov.u32 {fresh[1]}, 0;")
            Y = Xo[2:] + fresh  # shift down one 64-bit pair: pure renaming
            L.append(f"add.cc.u32 {X[0]}, {X[0]}, {Xo[1]};")
        L.append(f"mul.lo.u32 m, {X[0]}, {inv};")
        for k in range(N // 2):
            lo = ("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32") if k == 0 else "madc.lo.cc.u32"
            hi = "madc.hi.cc.u32" if k < N // 2 - 1 else "madc.hi.u32"
            L.append(f"{lo} {Y[2*k]}, m, {pl[2*k+1]}, {Y[2*k]};")
            L.append(f"{hi} {Y[2*k+1]}, m, {pl[2*k+1]}, {Y[2*k+1]};")
        for k in range(N // 2):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            L.append(f"{lo} {X[2*k]}, m, {pl[2*k]}, {X[2*k]};")
            L.append(f"madc.hi.cc.u32 {X[2*k+1]}, m, {pl[2*k]}, {X[2*k+1]};")
        L.append(f"addc.u32 {Y[N-1]}, {Y[N-1]}, 0;")
    # U = Y[j] + X[j+1]; r' = U + T_hi
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        src = X[j + 1] if j + 1 < N else z
        L.append(f"{opn} {Y[j]}, {Y[j]}, {src};")
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        L.append(f"{opn} {Y[j]}, {Y[j]}, {T[N + j]};")
    Tm = new(N)
    for j in range(N):
        opn = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        L.append(f"{opn} {Tm[j]}, {Y[j]}, {pl[j]};")
    L.append(f"subc.u32 brw, {z}, {z};")
    L.append("setp.ne.u32 pr, brw, 0;")
    for j in range(N):
        L.append(f"selp.u32 r{j}, {Y[j]}, {Tm[j]}, pr;")
    return L, cnt[0]


# --------------------------------------------------------------------------- dedicated squaring
def gen_sqr_body(p):
    """a*a*2^-768 mod p with 876 wide MACs instead of 1152: the 276 products a_i*a_j (i < j) are accumulated once
    (even/odd aligned chains, as in the Karatsuba variant's mul_half), doubled, the 24 squares a_i^2 are added on the
    diagonal, and the 48-limb result goes through the same Montgomery reduction as the Karatsuba variant.
    Inputs a0..a23; outputs r0..r23; intermediates are PTX virtual registers k<n> (count returned)."""
    pl = to_limbs32(p)
    inv = _inv32(p)
    L = []
    cnt = [0]

    def new(n=1):
        regs = [f"k{cnt[0] + i}" for i in range(n)]
        cnt[0] += n
        return regs if n > 1 else regs[0]

    z = new()
    L.append(f"mov.u32 {z}, 0;")
    a = [f"a{j}" for j in range(N)]
    n2 = 2 * N
    E = [None] * (n2 + 2)  # even-aligned pairs (0,1),(2,3),... ; E[k] holds limb position k
    O = [None] * (n2 + 2)  # odd-aligned pairs (1,2),(3,4),...
    for i in range(N - 1):
        for par in (0, 1):
            idxs = [j for j in range(i + 1, N) if j % 2 == par]
            if not idxs:
                continue
            first = idxs[0] + i
            arr = E if first % 2 == 0 else O
            started = False
            carry_live = False
            for j in idxs:
                pos = i + j
                lo_init = arr[pos] is None
                hi_init = arr[pos + 1] is None
                if lo_init:
                    arr[pos] = new()
                if hi_init:
                    arr[pos + 1] = new()
                if lo_init and not started:
                    L.append(f"mul.lo.u32 {arr[pos]}, {a[j]}, {a[i]};")
                    carry_live = False
                elif lo_init:
                    L.append(f"madc.lo.cc.u32 {arr[pos]}, {a[j]}, {a[i]}, {z};")
                    carry_live = True
                elif not started:
                    L.append(f"mad.lo.cc.u32 {arr[pos]}, {a[j]}, {a[i]}, {arr[pos]};")
                    carry_live = True
                else:
                    L.append(f"madc.lo.cc.u32 {arr[pos]}, {a[j]}, {a[i]}, {arr[pos]};")
                    carry_live = True
                addend = z if hi_init else arr[pos + 1]
                if carry_live:
                    L.append(f"madc.hi.cc.u32 {arr[pos+1]}, {a[j]}, {a[i]}, {addend};")
                elif hi_init:
                    L.append(f"mul.hi.u32 {arr[pos+1]}, {a[j]}, {a[i]};")
                    L.append(f"add.cc.u32 {arr[pos+1]}, {arr[pos+1]}, 0;")  # defines the carry flag (= 0)
                else:
                    L.append(f"mad.hi.cc.u32 {arr[pos+1]}, {a[j]}, {a[i]}, {arr[pos+1]};")
                started = True
            top = i + idxs[-1] + 2
            if top < n2 + 1:
                if arr[top] is None:
                    arr[top] = new()
                    L.append(f"addc.u32 {arr[top]}, {z}, 0;")
                else:
                    L.append(f"addc.u32 {arr[top]}, {arr[top]}, 0;")
    # T = 2 * (E + O) + sum_i a_i^2 2^(64 i)
    T = new(n2)
    for k in range(n2):
        ek = E[k] if E[k] is not None else z
        ok = O[k] if O[k] is not None else z
        opn = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n2 - 1 else "addc.u32")
        L.append(f"{opn} {T[k]}, {ek}, {ok};")
    for k in range(n2):
        opn = "add.cc.u32" if k == 0 else ("addc.cc.u32" if k < n2 - 1 else "addc.u32")
        L.append(f"{opn} {T[k]}, {T[k]}, {T[k]};")
    for i in range(N):
        lo = "mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32"
        hi = "madc.hi.cc.u32" if i < N - 1 else "madc.hi.u32"
        L.append(f"{lo} {T[2*i]}, {a[i]}, {a[i]}, {T[2*i]};")
        L.append(f"{hi} {T[2*i+1]}, {a[i]}, {a[i]}, {T[2*i+1]};")
    # ---- Montgomery reduction of T_lo with the even/odd row machinery (same as gen_mul_karatsuba_body)
    X = list(T[:N])
    Y = new(N)
    for k in range(N):
        L.append(f"mov.u32 {Y[k]}, 0;")
    for i in range(N):
        if i > 0:
            Xo, Yo = X, Y
            X = Yo
            fresh = new(2)
            L.append(f"mov.u32 {fresh[0]}, 0;")
            L.append(f"mov.u32 {fresh[1]}, 0;")
            Y = Xo[2:] + fresh
            L.append(f"add.cc.u32 {X[0]}, {X[0]}, {Xo[1]};")
        L.append(f"mul.lo.u32 m, {X[0]}, {inv};")
        for k in range(N // 2):
            lo = ("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32") if k == 0 else "madc.lo.cc.u32"
            hi = "madc.hi.cc.u32" if k < N // 2 - 1 else "madc.hi.u32"
            L.append(f"{lo} {Y[2*k]}, m, {pl[2*k+1]}, {Y[2*k]};")
            L.append(f"{hi} {Y[2*k+1]}, m, {pl[2*k+1]}, {Y[2*k+1]};")
        for k in range(N // 2):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            L.append(f"{lo} {X[2*k]}, m, {pl[2*k]}, {X[2*k]};")
            L.append(f"madc.hi.cc.u32 {X[2*k+1]}, m, {pl[2*k]}, {X[2*k+1]};")
        L.append(f"addc.u32 {Y[N-1]}, {Y[N-1]}, 0;")
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        src = X[j + 1] if j + 1 < N else z
        L.append(f"{opn} {Y[j]}, {Y[j]}, {src};")
    for j in range(N):
        opn = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < N - 1 else "addc.u32")
        L.append(f"{opn} {Y[j]}, {Y[j]}, {T[N + j]};")
    Tm = new(N)
    for j in range(N):
        opn = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        L.append(f"{opn} {Tm[j]}, {Y[j]}, {pl[j]};")
    L.append(f"subc.u32 brw, {z}, {z};")
    L.append("setp.ne.u32 pr, brw, 0;")
    for j in range(N):
        L.append(f"selp.u32 r{j}, {Y[j]}, {Tm[j]}, pr;")
    return L, cnt[0]


if __name__ == "__main__":
    main()
