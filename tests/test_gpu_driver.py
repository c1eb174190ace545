"""GPU suite: the command-line drivers. `bin/cuda_prover_piecewise` (this repo's host driver over the B:: bundle) and,
when it was built in the container that has /root/reference, `oracle/_ref/piecewise_b200` (the reference's OWN
unmodified cuda_prover_piecewise.cu compiled against this repo's prover_reference_functions.hpp) must write exactly
the reference prover's proof bytes for the committed golden key/witness files - the README.md:47-58 sha256 recipe."""
import os
import subprocess

import pytest

import util

pytestmark = pytest.mark.gpu
PKG = os.path.join(util.ROOT, "snark_challenge_prover_reference_b200")
DRIVERS = [os.path.join(PKG, "bin", "cuda_prover_piecewise"), os.path.join(util.ROOT, "oracle", "_ref", "piecewise_b200")]


@pytest.mark.parametrize("driver", DRIVERS, ids=["repo-driver", "reference-driver"])
@pytest.mark.parametrize("tables", ["1", "0"])
@pytest.mark.parametrize("curve,k", [(0, 8), (1, 8), (0, 5)])
def test_driver_writes_reference_proof(tmp_path, driver, tables, curve, k):
    if not os.path.exists(driver):
        pytest.skip("%s not built" % driver)
    name = "MNT4753" if curve == 0 else "MNT6753"
    base = os.path.join(util.GOLDEN, "%s_k%d" % (name, k))
    out = tmp_path / "proof"
    env = dict(os.environ, B200_PRECOMPUTE=tables, LD_LIBRARY_PATH=PKG + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([driver, name, "compute", base + ".params", base + ".input", str(out)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == open(base + ".output", "rb").read()


def test_driver_reports_missing_file(tmp_path):
    driver = DRIVERS[0]
    if not os.path.exists(driver):
        pytest.skip("driver not built")
    r = subprocess.run([driver, "MNT4753", "compute", "/nonexistent", "/nonexistent", str(tmp_path / "o")],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "cannot open" in r.stderr
