"""complete Groth16 proofs checked by the REFERENCE's pairing-based verifier.

oracle/_ref/groth16_tool (reference-linked test infrastructure, oracle/groth16_tool.cpp) generates a real key with the
reference's generator - parameter / input files exactly as generate_parameters writes them, plus the alpha / beta /
delta elements and the verification key that the challenge's files leave out - and verifies complete proofs with
r1cs_gg_ppzksnark_verifier_strong_IC. The product adds the missing proof terms (b200_groth16_finalize: host group
operations) to the challenge proof A | B | C.

CPU suite: the challenge proof comes from the oracle restatement. GPU suite: from the CUDA prover (b200_prove_full)."""
import os
import random
import subprocess

import pytest

import mnt753 as M
import util

TOOL = os.path.join(util.ROOT, "oracle", "_ref", "groth16_tool")
NAMES = ("MNT4753", "MNT6753")
FE = 96


@pytest.fixture(scope="module")
def keys(tmp_path_factory):
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/groth16_tool not built in this checkout (oracle/build_ref.sh needs /root/reference)")
    d = tmp_path_factory.mktemp("groth16")
    out = {}
    for curve, name in enumerate(NAMES):
        f = {k: str(d / ("%s.%s" % (name, k))) for k in ("params", "input", "extras", "vk", "proof")}
        subprocess.run([TOOL, "gen", name, "5", f["params"], f["input"], f["extras"], f["vk"]], check=True,
                       capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="4"))
        out[curve] = f
    return out


def _verify(curve, f, proof):
    open(f["proof"], "wb").write(proof)
    r = subprocess.run([TOOL, "verify", NAMES[curve], f["vk"], f["input"], f["proof"]], capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    return r.returncode == 0


def _random_s(curve, seed):
    r = util.curve_obj(curve).r
    return util.fe_bytes(M.to_mont(random.Random(seed).randrange(1, r), r))


@pytest.mark.parametrize("curve", [0, 1])
def test_complete_proof_from_oracle_verifies(b200, oracle, keys, curve):
    f = keys[curve]
    params, inp, extras = (open(f[k], "rb").read() for k in ("params", "input", "extras"))
    abc = util.orc_prove(oracle, curve, params, inp)
    full = b200.groth16_finalize(curve, abc, inp[-FE:], _random_s(curve, 5), extras)
    assert len(full) == len(abc) and full != abc
    assert _verify(curve, f, full)
    # the bare challenge proof is NOT a Groth16 proof, and a tampered complete proof is rejected
    assert not _verify(curve, f, abc)
    bad = bytearray(full)
    bad[10] ^= 1
    assert not _verify(curve, f, bytes(bad))
    # another randomiser s gives another valid proof (zero-knowledge re-randomisation)
    other = b200.groth16_finalize(curve, abc, inp[-FE:], _random_s(curve, 6), extras)
    assert other != full and _verify(curve, f, other)


@pytest.mark.gpu
@pytest.mark.parametrize("curve", [0, 1])
def test_complete_proof_from_gpu_verifies(b200, keys, curve):
    import torch
    assert torch.cuda.is_available()
    b200.check(b200.lib().b200_set_device(0))
    f = keys[curve]
    inp, extras = open(f["input"], "rb").read(), open(f["extras"], "rb").read()
    key = b200.Params.from_file(curve, f["params"])
    full = key.prove_full(inp, _random_s(curve, 7), extras)
    assert full == b200.groth16_finalize(curve, key.prove(inp), inp[-FE:], _random_s(curve, 7), extras)
    key.close()
    assert _verify(curve, f, full)
