"""Generator + CPU simulator for the LANE-COOPERATIVE point arithmetic used by the latency-bound MSM phases.

A projective point addition over Fq3 is 14 tower multiplications = 84 base-field multiplications plus ~200 additions; run
by ONE thread (round 1) that is a chain of ~120 K dependent instructions, 0.4 ms. But the formula is wide: its
multiplications come in 4 waves of up to 5, and each tower multiplication is 3 (Fq2) / 6 (Fq3) independent base
multiplications. This script expands the reference's formulas (homogeneous projective add-1998-cmo-2 / dbl-2007-bl, the
ones libff uses: mnt4753_g1.cpp:134-207, 315-346; Fq2 / Fq3 Karatsuba, fp2.tcc:78-126, fp3.tcc:82-123) into a DAG of
BASE-FIELD operations, levelises it (ASAP), allocates shared-memory slots by liveness and emits, per group, a static
schedule: level by level, lane l of a lane group executes operation l of the level on slots of the group's scratchpad
(csrc/coop_sched_gen.h, interpreted by csrc/coop.cuh).

Because the build container has no GPU, the emitted tables are also executed here on Python integers
(simulate()) and compared with the curve model (tools/mnt753.py) by tests/test_coop_sched.py.

    python tools/gen_coop_sched.py            -> writes csrc/coop_sched_gen.h
"""
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
import mnt753 as M  # noqa: E402

MUL, ADD, SUB, MULK = 0, 1, 2, 3
IN_A, IN_B, OUT = 0x4000, 0x8000, 0xC000

GROUPS = {
    # name: (curve, degree, non-residue, lanes per group)
    "Mnt4G1": (M.MNT4753, 1, 0, 4),
    "Mnt4G2": (M.MNT4753, 2, 13, 16),
    "Mnt6G1": (M.MNT6753, 1, 0, 4),
    "Mnt6G2": (M.MNT6753, 3, 11, 32),
}


class Builder:
    def __init__(self):
        self.nodes = []  # (kind, a, b, k) ; inputs: ("in", code)

    def inp(self, code):
        self.nodes.append(("in", code, None, 0))
        return len(self.nodes) - 1

    def op(self, kind, a, b=None, k=0):
        self.nodes.append((kind, a, b, k))
        return len(self.nodes) - 1


class Tower:
    """tower arithmetic on tuples of node ids, mirroring csrc/field.cuh"""

    def __init__(self, bld, deg, nr):
        self.b, self.deg, self.nr = bld, deg, nr

    def add(self, x, y):
        return tuple(self.b.op(ADD, p, q) for p, q in zip(x, y))

    def sub(self, x, y):
        return tuple(self.b.op(SUB, p, q) for p, q in zip(x, y))

    def dbl(self, x):
        return self.add(x, x)

    def mul(self, x, y):
        b, nr = self.b, self.nr
        if self.deg == 1:
            return (b.op(MUL, x[0], y[0]),)
        if self.deg == 2:  # Karatsuba, fp2.tcc:78-90
            aA, bB = b.op(MUL, x[0], y[0]), b.op(MUL, x[1], y[1])
            s = b.op(MUL, b.op(ADD, x[0], x[1]), b.op(ADD, y[0], y[1]))
            c1 = b.op(SUB, b.op(SUB, s, aA), bB)
            c0 = b.op(ADD, aA, b.op(MULK, bB, k=nr))
            return (c0, c1)
        a0, a1, a2 = x
        b0, b1, b2 = y  # fp3.tcc:82-96
        aA, bB, cC = b.op(MUL, a0, b0), b.op(MUL, a1, b1), b.op(MUL, a2, b2)
        s = b.op(MUL, b.op(ADD, a1, a2), b.op(ADD, b1, b2))
        s = b.op(MULK, b.op(SUB, b.op(SUB, s, bB), cC), k=nr)
        t = b.op(MUL, b.op(ADD, a0, a1), b.op(ADD, b0, b1))
        t = b.op(ADD, b.op(SUB, b.op(SUB, t, aA), bB), b.op(MULK, cC, k=nr))
        u = b.op(MUL, b.op(ADD, a0, a2), b.op(ADD, b0, b2))
        u = b.op(SUB, b.op(ADD, b.op(SUB, u, aA), bB), cC)
        return (b.op(ADD, aA, s), t, u)

    def sqr(self, x):
        b, nr = self.b, self.nr
        if self.deg == 1:
            return (b.op(MUL, x[0], x[0]),)
        if self.deg == 2:  # complex squaring, fp2.tcc:117-126
            ab = b.op(MUL, x[0], x[1])
            s = b.op(MUL, b.op(ADD, x[0], x[1]), b.op(ADD, x[0], b.op(MULK, x[1], k=nr)))
            c0 = b.op(SUB, b.op(SUB, s, ab), b.op(MULK, ab, k=nr))
            return (c0, b.op(ADD, ab, ab))
        a0, a1, a2 = x  # CH-SQR2, fp3.tcc:106-123
        s0 = b.op(MUL, a0, a0)
        s1 = b.op(MUL, a0, a1)
        s1 = b.op(ADD, s1, s1)
        t = b.op(ADD, b.op(SUB, a0, a1), a2)
        s2 = b.op(MUL, t, t)
        s3 = b.op(MUL, a1, a2)
        s3 = b.op(ADD, s3, s3)
        s4 = b.op(MUL, a2, a2)
        c0 = b.op(ADD, s0, b.op(MULK, s3, k=nr))
        c1 = b.op(ADD, s1, b.op(MULK, s4, k=nr))
        c2 = b.op(SUB, b.op(SUB, b.op(ADD, b.op(ADD, s1, s2), s3), s0), s4)
        return (c0, c1, c2)


def mul_by_a(T, name, x):
    b = T.b
    if name == "Mnt4G1":   # a = 2
        return T.dbl(x)
    if name == "Mnt6G1":   # a = 11
        return (b.op(MULK, x[0], k=11),)
    if name == "Mnt4G2":   # a' = (2 * 13, 0): componentwise times 26 (mnt4753_g2.cpp:31-34)
        return (b.op(MULK, x[0], k=26), b.op(MULK, x[1], k=26))
    # Mnt6G2: a' = (0, 0, 11): (c0, c1, c2) -> (121 c1, 121 c2, 11 c0) (mnt6753_g2.cpp:38-41)
    return (b.op(MULK, x[1], k=121), b.op(MULK, x[2], k=121), b.op(MULK, x[0], k=11))


def build_add(name):
    """general addition of two finite points; returns (builder, outputs, cross products for the P == Q test)"""
    curve, deg, nr, lanes = GROUPS[name]
    bld = Builder()
    T = Tower(bld, deg, nr)
    comp = lambda base, c: tuple(bld.inp(base + c * deg + i) for i in range(deg))
    X1, Y1, Z1 = comp(IN_A, 0), comp(IN_A, 1), comp(IN_A, 2)
    X2, Y2, Z2 = comp(IN_B, 0), comp(IN_B, 1), comp(IN_B, 2)
    X1Z2, X2Z1 = T.mul(X1, Z2), T.mul(Z1, X2)
    Y1Z2, Y2Z1 = T.mul(Y1, Z2), T.mul(Z1, Y2)
    Z1Z2 = T.mul(Z1, Z2)
    u, v = T.sub(Y2Z1, Y1Z2), T.sub(X2Z1, X1Z2)
    uu, vv = T.sqr(u), T.sqr(v)
    t2 = T.mul(uu, Z1Z2)
    R = T.mul(vv, X1Z2)
    vvv = T.mul(v, vv)
    A = T.sub(T.sub(T.sub(t2, vvv), R), R)
    X3 = T.mul(v, A)
    Y3 = T.sub(T.mul(u, T.sub(R, A)), T.mul(vvv, Y1Z2))
    Z3 = T.mul(vvv, Z1Z2)
    return bld, X3 + Y3 + Z3, X1Z2 + X2Z1 + Y1Z2 + Y2Z1


def build_dbl(name):
    curve, deg, nr, lanes = GROUPS[name]
    bld = Builder()
    T = Tower(bld, deg, nr)
    comp = lambda base, c: tuple(bld.inp(base + c * deg + i) for i in range(deg))
    X, Y, Z = comp(IN_A, 0), comp(IN_A, 1), comp(IN_A, 2)
    XX = T.sqr(X)
    ZZ = T.sqr(Z)
    w = mul_by_a(T, name, ZZ)
    w = T.add(T.add(T.add(w, XX), XX), XX)
    s = T.dbl(T.mul(Y, Z))
    Rr = T.mul(Y, s)
    RR = T.sqr(Rr)
    Bv = T.sub(T.sub(T.sqr(T.add(X, Rr)), XX), RR)
    h = T.sub(T.sub(T.sqr(w), Bv), Bv)
    X3 = T.mul(h, s)
    Y3 = T.sub(T.sub(T.mul(w, T.sub(Bv, h)), RR), RR)
    Z3 = T.mul(s, T.sqr(s))
    return bld, X3 + Y3 + Z3, ()


def schedule(bld, outputs, keep, lanes):
    """-> dict(levels=[[(kind,k,dst,a,b)...]...], ntemps, keep_slots, checkpoint_level)"""
    nodes = bld.nodes
    n = len(nodes)
    # dead-code elimination from the outputs (+ the kept cross products)
    live = [False] * n
    stack = list(outputs) + list(keep)
    while stack:
        i = stack.pop()
        if live[i]:
            continue
        live[i] = True
        kind, a, b, k = nodes[i]
        if kind != "in":
            stack.append(a)
            if b is not None:
                stack.append(b)
    # Levels are ordered by (multiplication depth, addition depth inside it): all multiplications with the same number of
    # multiplications on their longest input path share ONE level - a level that contains a multiplication costs a
    # multiplication (~4 us) however many lanes multiply, so what matters is the number of such levels (4 for an
    # addition, whatever the tower) - and the cheap additions / subtractions / small-constant products fill sub-levels
    # between them.
    forced = {}  # multiplication -> minimum depth (a multiplication with slack deferred out of an over-full level)

    def depths():
        md, ad = [0] * n, [0] * n
        for i, (kind, a, b, k) in enumerate(nodes):
            if kind == "in" or not live[i]:
                continue
            ops_ = [a] + ([b] if b is not None else [])
            m = max(md[o] for o in ops_)
            if kind == MUL:
                md[i], ad[i] = max(m + 1, forced.get(i, 0)), 0
            else:
                md[i] = m
                ad[i] = 1 + max([ad[o] for o in ops_ if md[o] == m and nodes[o][0] != "in"] + [0])
        return md, ad

    md, ad = depths()
    # A level holds at most `lanes` multiplications. When a depth has more, multiplications whose results are not needed
    # at the next depth (slack) move one depth down instead of opening an extra multiplication level: with 4-lane groups
    # the G1 addition's 5 + 2 + 3 + 4 multiplications become 4 + 3 + 3 + 4.
    for _ in range(64):
        top_md = max(md)
        alap = [top_md] * n
        for i in range(n - 1, -1, -1):
            kind, a, b, k = nodes[i]
            if kind == "in" or not live[i]:
                continue
            need = alap[i] - (1 if kind == MUL else 0)
            for o in (a, b):
                if o is not None:
                    alap[o] = min(alap[o], need)
        moved = False
        for m in range(1, top_md):
            here = [i for i in range(n) if live[i] and nodes[i][0] == MUL and md[i] == m]
            if len(here) > lanes:
                slack = sorted((i for i in here if alap[i] > m), key=lambda i: -alap[i])
                for i in slack[:len(here) - lanes]:
                    forced[i] = m + 1
                    moved = True
                if moved:
                    break
        if not moved:
            break
        md, ad = depths()
    keys = sorted({(md[i], ad[i]) for i in range(n) if live[i] and nodes[i][0] != "in"})
    rank = {kk: r + 1 for r, kk in enumerate(keys)}
    level = [0] * n
    for i in range(n):
        if live[i] and nodes[i][0] != "in":
            level[i] = rank[(md[i], ad[i])]
    # an output node must be computed by an operation (never a bare input) and each output slot is written once
    assert all(nodes[o][0] != "in" for o in outputs) and len(set(outputs)) == len(outputs)
    # The destination point may alias an input point: nothing may be written while an input can still be read. An output
    # is therefore placed after the last level that reads an input (and no output feeds another node).
    users = set()
    last_input_read = 0
    for i, (kind, a, b, k) in enumerate(nodes):
        if kind != "in" and live[i]:
            for o in (a, b):
                if o is not None:
                    users.add(o)
                    if nodes[o][0] == "in":
                        last_input_read = max(last_input_read, level[i])
    assert not (users & set(outputs))
    for o in outputs:
        level[o] = max(level[o], last_input_read + 1)
    by_level = {}
    for i in range(n):
        if live[i] and nodes[i][0] != "in":
            by_level.setdefault(level[i], []).append(i)
    # split levels wider than the lane group (multiplications first, so that a level's slow operations share a sub-level)
    levels = []
    for L in sorted(by_level):
        ops = sorted(by_level[L], key=lambda i: (nodes[i][0] != MUL, i))
        for s in range(0, len(ops), lanes):
            levels.append(ops[s:s + lanes])
    final_level = {}
    for li, ops in enumerate(levels):
        for i in ops:
            final_level[i] = li
    last_use = {}
    for li, ops in enumerate(levels):
        for i in ops:
            kind, a, b, k = nodes[i]
            for o in (a, b):
                if o is not None:
                    last_use[o] = max(last_use.get(o, -1), li)
    for o in keep:
        last_use[o] = max(last_use.get(o, -1), len(levels))  # (their own uses already extend them past the checkpoint)
    out_index = {o: j for j, o in enumerate(outputs)}
    # in-place safety: the destination point may alias an input point, so every input must be read strictly before the
    # first output is written
    first_out = min(final_level[o] for o in outputs)
    for i, (kind, a, b, k) in enumerate(nodes):
        if kind == "in":
            assert last_use.get(i, -1) < first_out, "an input is read after an output has been written"
    slot = {}
    free, ntemps = [], 0
    expiring = {}
    table = []
    for li, ops in enumerate(levels):
        row = []
        for i in ops:
            kind, a, b, k = nodes[i]
            if i in out_index:
                dst = OUT + out_index[i]
            else:
                if free:
                    dst = free.pop()
                else:
                    dst = ntemps
                    ntemps += 1
                slot[i] = dst
                expiring.setdefault(last_use.get(i, li), []).append(dst)

            def ref(o):
                if nodes[o][0] == "in":
                    return nodes[o][1]
                if o in out_index:
                    return OUT + out_index[o]
                return slot[o]
            row.append((kind, k, dst, ref(a), ref(b) if b is not None else ref(a)))
        table.append(row)
        for s in expiring.pop(li, []):  # slots whose last reader was this level are free from the next level on
            free.append(s)
    # outputs that are read again later (e.g. none here) would be read through OUT + j: fine, they are only read after
    # they were written. Kept cross products: their slots and the level after which all of them exist.
    keep_slots = [slot[o] for o in keep]
    checkpoint = (max(final_level[o] for o in keep) + 1) if keep else 0
    return {"levels": table, "ntemps": ntemps, "keep": keep_slots, "checkpoint": checkpoint}


def simulate(sched, p, nr_unused, a_vals, b_vals):
    """run a schedule on integers mod p; a_vals / b_vals: component lists of the input points; returns output list"""
    temps = {}
    out = {}

    def get(r):
        if r >= OUT:
            return out[r - OUT]
        if r >= IN_B:
            return b_vals[r - IN_B]
        if r >= IN_A:
            return a_vals[r - IN_A]
        return temps[r]
    for row in sched["levels"]:
        vals = []
        for kind, k, dst, a, b in row:  # all operands are read before any result of the level is written
            x, y = get(a), get(b)
            vals.append((dst, x * y % p if kind == MUL else (x + y) % p if kind == ADD else (x - y) % p if kind == SUB else x * k % p))
        for dst, v in vals:
            if dst >= OUT:
                out[dst - OUT] = v
            else:
                temps[dst] = v
    return [out[j] for j in range(len(out))]


def all_schedules():
    res = {}
    for name, (curve, deg, nr, lanes) in GROUPS.items():
        for opname, build in (("add", build_add), ("dbl", build_dbl)):
            bld, outs, keep = build(name)
            res[(name, opname)] = schedule(bld, list(outs), list(keep), lanes)
    return res


def emit():
    S = all_schedules()
    L = ["// GENERATED by tools/gen_coop_sched.py - do not edit. Static schedules of the lane-cooperative point addition and",
         "// doubling (base-field operations per level; lane l of a group runs operation l of the level). See csrc/coop.cuh.",
         "#pragma once", "#include <stdint.h>", "", "namespace b200 {",
         "struct CoopOp { uint8_t kind, k; uint16_t dst, a, b; };  // kind: 0 mul, 1 add, 2 sub, 3 multiply by the small constant k",
         "constexpr uint16_t kCoopInA = 0x4000, kCoopInB = 0x8000, kCoopOut = 0xC000;",
         "struct CoopSchedule { const CoopOp *ops; const uint16_t *level_start; int nlevels, ntemps, checkpoint; const uint16_t *keep; };", ""]
    for (name, opname), sc in S.items():
        flat, starts = [], [0]
        for row in sc["levels"]:
            flat += row
            starts.append(len(flat))
        tag = "%s_%s" % (name, opname)
        L.append("// %s %s: %d operations (%d multiplications) in %d levels, %d scratch slots" % (
            name, opname, len(flat), sum(1 for o in flat if o[0] == MUL), len(sc["levels"]), sc["ntemps"]))
        L.append("__device__ const CoopOp kCoopOps_%s[%d] = {" % (tag, len(flat)))
        for i in range(0, len(flat), 6):
            L.append("    " + " ".join("{%d, %d, 0x%04x, 0x%04x, 0x%04x}," % o for o in flat[i:i + 6]))
        L.append("};")
        L.append("__device__ const uint16_t kCoopLevels_%s[%d] = {%s};" % (tag, len(starts), ", ".join(map(str, starts))))
        keep = sc["keep"] or [0]
        L.append("__device__ const uint16_t kCoopKeep_%s[%d] = {%s};" % (tag, len(keep), ", ".join(map(str, keep))))
        L.append("constexpr int kCoopNLevels_%s = %d, kCoopNTemps_%s = %d, kCoopCheckpoint_%s = %d;" % (
            tag, len(sc["levels"]), tag, sc["ntemps"], tag, sc["checkpoint"]))
        L.append("")
    L.append("}  // namespace b200")
    return "\n".join(L) + "\n"


def main():
    out = os.path.join(os.path.dirname(__file__), "..", "snark_challenge_prover_reference_b200", "csrc", "coop_sched_gen.h")
    with open(out, "w") as f:
        f.write(emit())
    for (name, opname), sc in all_schedules().items():
        nops = sum(len(r) for r in sc["levels"])
        print("%s %s: %d ops, %d levels (%d with multiplications), %d temps, checkpoint after level %d" % (
            name, opname, nops, len(sc["levels"]), sum(1 for r in sc["levels"] if any(o[0] == MUL for o in r)), sc["ntemps"], sc["checkpoint"]))
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    main()
