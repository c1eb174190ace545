"""CPU suite: the bench's workload-file generator (tools/synth_key) writes well-formed reference-format files whose
bases are the advertised multiples of the generators, and the oracle restatement and the UNMODIFIED reference prover
(oracle/_ref/main, when it has been built) produce the same proof from them."""
import hashlib
import os
import subprocess
import sys

import pytest

import mnt753 as M
import util

sys.path.insert(0, util.ROOT)
import bench  # noqa: E402

FE = 96


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    out = {}
    for curve, name in enumerate(bench.CURVES):
        pf, inf = str(d / (name + "-parameters")), str(d / (name + "-input"))
        subprocess.check_call([bench.synth_tool(), name, "5", pf, inf])
        out[curve] = (pf, inf)
    return d, out


@pytest.mark.parametrize("curve", [0, 1])
def test_synth_key_contents(synth, curve):
    _, files = synth
    image = open(files[curve][0], "rb").read()
    d, m, q = util.split_params(curve, image)
    assert (d, m) == (31, 32)
    for name, group, first, idx in (("A", 1, 1000003, (0, 1, 3, 2)), ("B1", 1, 2000003, (1, 7, m - 1)),
                                    ("B2", 2, 3000017, (1, 2, m - 1)), ("L", 1, 4000037, (0, m - 2)),
                                    ("H", 1, 5000011, (0, 5, d - 1))):
        F, a, b, G = util.group_params(curve, group)
        ab = 2 * FE * util.deg(curve, group)
        for i in idx:
            assert q[name][i * ab:(i + 1) * ab] == util.encode_affine(curve, M.ec_mul(F, a, first + i, G), group), (name, i)
    g1 = 2 * FE
    A = q["A"]
    assert all(A[i * g1:(i + 1) * g1] == A[2 * g1:3 * g1] for i in list(range(2, m - 1, 2)) + [m - 1])
    assert A[m * g1:] == bytes(g1) and q["B1"][:g1] == bytes(g1) and q["B1"][m * g1:] == bytes(g1)
    assert q["B1"][(m - 2) * g1:(m - 1) * g1] == q["B1"][(m - 3) * g1:(m - 2) * g1]
    inp = util.split_input(open(files[curve][1], "rb").read(), d, m)
    r = util.curve_obj(curve).r
    assert util.fe_int(inp["w"][:FE]) == M.R % r
    assert all(util.fe_int(inp["w"][i * FE:(i + 1) * FE]) < (1 << 752) for i in range(m + 1))


@pytest.mark.parametrize("curve", [0, 1])
def test_oracle_and_reference_agree_on_synth_files(synth, oracle, curve):
    d, files = synth
    params, inp = open(files[curve][0], "rb").read(), open(files[curve][1], "rb").read()
    proof = util.orc_prove(oracle, curve, params, inp)
    main = os.path.join(bench.REF_DIR, "main")
    if not os.path.exists(main):
        pytest.skip("oracle/_ref/main not built in this checkout")
    name = bench.CURVES[curve]
    subprocess.run([main, name, "compute", files[curve][0], files[curve][1], str(d / (name + "-out"))], check=True,
                   capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert hashlib.sha256(open(d / (name + "-out"), "rb").read()).hexdigest() == hashlib.sha256(proof).hexdigest()


def test_reference_arm_line_is_same_config(tmp_path):
    """bench.py --impl reference at a small size: one JSON line, measured on the step's own files, reused on a second call."""
    if not os.path.exists(os.path.join(bench.REF_DIR, "main")):
        pytest.skip("oracle/_ref/main not built in this checkout")
    import json
    env = dict(os.environ, B200_BENCH_CACHE=str(tmp_path))
    cmd = [sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--log2-mnt4", "6", "--log2-mnt6", "5"]
    first = json.loads(subprocess.run(cmd, env=env, check=True, capture_output=True, text=True).stdout)
    again = json.loads(subprocess.run(cmd, env=env, check=True, capture_output=True, text=True).stdout)
    assert first["impl"] == "reference" and first["steps_effective"] == 1
    assert first["config"]["workload"].startswith("MNT4753 2^6 + MNT6753 2^5") and first["config"]["same_files_as_b200_arm"]
    assert first["value"] == again["value"] and "reused" in again["config"]["sample"]
    assert set(first["proof_sha256"]) == set(bench.CURVES)
