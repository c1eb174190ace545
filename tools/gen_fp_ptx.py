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
def emit_function(name, lines, sqr=False):
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

    body = ["{", ".reg .u32 P<24>, Q<24>, t<24>, m, z, brw;", ".reg .pred pr;"] + [tr(l) for l in lines] + ["}"]
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
        parts.append(emit_function(f"fp_add_ptx_{tag}", gen_add_body(p)))
        parts.append("")
        parts.append(emit_function(f"fp_sub_ptx_{tag}", gen_sub_body(p)))
        parts.append("")
    with open(out, "w") as f:
        f.write("\n".join(parts))
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    main()
