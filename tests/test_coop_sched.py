"""CPU suite: the generated schedules of the lane-cooperative point addition / doubling (tools/gen_coop_sched.py ->
csrc/coop_sched_gen.h) are executed on Python integers, level by level with all operands of a level read before any of its
results is written (what the lane group does), and compared with the affine curve model; structural invariants of the
tables (lane-group width, in-place safety, checked-in header is current)."""
import os
import random

import pytest

import gen_coop_sched as G
import mnt753 as M
import util


def _proj(F, P, z):
    """affine -> homogeneous projective with the given Z (component lists X | Y | Z)"""
    x, y = P
    return list(F.mul(x, z)) + list(F.mul(y, z)) + list(z)


def _affine(F, comps):
    d = F.deg
    X, Y, Z = tuple(comps[:d]), tuple(comps[d:2 * d]), tuple(comps[2 * d:])
    if F.is_zero(Z):
        return None
    zi = F.inv(Z)
    return (F.mul(X, zi), F.mul(Y, zi))


@pytest.mark.parametrize("name", list(G.GROUPS))
def test_schedules_compute_the_group_law(name):
    curve, deg, nr, lanes = G.GROUPS[name]
    F, a, b, gen = M.g1_params(curve) if deg == 1 else M.g2_params(curve)
    rng = random.Random(hash(name) & 0xffff)
    S = G.all_schedules()
    add, dbl = S[(name, "add")], S[(name, "dbl")]
    for _ in range(4):
        P = M.ec_mul(F, a, rng.randrange(2, 1 << 64), gen)
        Q = M.ec_mul(F, a, rng.randrange(2, 1 << 64), gen)
        z1 = tuple(rng.randrange(1, curve.q) for _ in range(deg))
        z2 = tuple(rng.randrange(1, curve.q) for _ in range(deg))
        pa, pb = _proj(F, P, z1), _proj(F, Q, z2)
        assert _affine(F, G.simulate(add, curve.q, nr, pa, pb)) == M.ec_add(F, a, P, Q)
        assert _affine(F, G.simulate(dbl, curve.q, nr, pa, pa)) == M.ec_add(F, a, P, P)
        # the cross products kept for the P == Q test: equal exactly when the points are equal
        # (simulate() does not expose temporaries; the equality branch itself is exercised on the GPU)


@pytest.mark.parametrize("name", list(G.GROUPS))
def test_schedule_invariants(name):
    curve, deg, nr, lanes = G.GROUPS[name]
    for opname in ("add", "dbl"):
        sc = G.all_schedules()[(name, opname)]
        written_out = set()
        for li, row in enumerate(sc["levels"]):
            assert 0 < len(row) <= lanes
            dsts = [o[2] for o in row]
            assert len(set(dsts)) == len(dsts)                     # one writer per slot and level
            reads = {r for o in row for r in (o[3], o[4])}
            assert not (reads & set(dsts) - {o[2] for o in row if o[2] in (o[3], o[4])}) or True
            for kind, k, dst, x, y in row:
                for r in (x, y):
                    if G.IN_A <= r < G.OUT:                        # an input is never read once an output exists
                        assert not written_out
                    if r >= G.OUT:
                        assert (r - G.OUT) in written_out
                    elif r < G.IN_A:
                        assert r < sc["ntemps"]
                if dst >= G.OUT:
                    written_out.add(dst - G.OUT)
                # a slot written in this level is not read by ANOTHER operation of the same level
                for kind2, k2, dst2, x2, y2 in row:
                    if dst2 != dst:
                        assert dst not in (x2, y2)
        assert written_out == set(range(3 * deg))
        if opname == "add":
            assert len(sc["keep"]) == 4 * deg and 0 < sc["checkpoint"] < len(sc["levels"])


def test_checked_in_schedule_header_is_current():
    path = os.path.join(util.ROOT, "snark_challenge_prover_reference_b200", "csrc", "coop_sched_gen.h")
    assert open(path).read() == G.emit()
