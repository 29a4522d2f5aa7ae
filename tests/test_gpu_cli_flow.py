"""The file-to-file flow of `co-circom split-witness` + `generate-proof` (tools/co_circom.py over the C ABI) on the reference's
fixtures: witness.wtns + circuit.r1cs -> three .shared files -> REP3 proof on the GPU -> proof.json + public.json, accepted by the
oracle's verifier with the fixture's verification_key.json -- the acceptance criterion of the reference's own end-to-end tests
(tests/tests/circom/e2e_tests/mod.rs:24-110) and examples (co-circom/co-circom/examples/groth16/run_full_*.sh)."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import cref, formats, groth16
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
CURVES = {"bn254": (BN254, "BN254"), "bls12_381": (BLS12_381, "BLS12-381")}


def _cli():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import co_circom
    return co_circom


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bn254", "poseidon"), ("bls12_381", "poseidon")])
def test_split_witness_then_generate_proof(cocg, tmp_path, curve, circ):
    c, cname = CURVES[curve]
    d = os.path.join(G, "groth16", curve, circ)
    cli = _cli()
    cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "REP3",
              "--curve", cname, "--out-dir", str(tmp_path)])
    shares = [str(tmp_path / f"witness.wtns.{i}.shared") for i in range(3)]
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    public = [int(x) for x in json.load(open(os.path.join(d, "public.json")))]
    # the three files are a replicated sharing of the witness: a_0 + a_1 + a_2 = w, b_i = a_(i-1), same public inputs everywhere
    dec = [cocg.shared_witness_decode(cocg.BN254 if c is BN254 else cocg.BLS12_381, open(p, "rb").read(), 2) for p in shares]
    a = [cref.fr_from_mont(c, x[1][0]) for x in dec]
    b = [cref.fr_from_mont(c, x[1][1]) for x in dec]
    ell = len(public)
    assert [(x + y + z) % c.r for x, y, z in zip(*a)] == [v % c.r for v in wt[ell + 1:]]
    for i in range(3):
        assert b[i] == a[(i + 2) % 3]
        assert cref.fr_from_mont(c, dec[i][0]) == [v % c.r for v in wt[:ell + 1]]
    assert a[0] != a[1] and len(set(a[0])) > len(a[0]) // 2      # actually random
    out, pub_out = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    cli.main(["generate-proof", "groth16", "--witness", *shares, "--zkey", os.path.join(d, "circuit.zkey"), "--protocol", "REP3", "--curve", cname,
              "--out", out, "--public-input", pub_out])
    assert json.load(open(pub_out)) == json.load(open(os.path.join(d, "public.json")))
    _, A, B, C = formats.proof_from_json(open(out).read())
    vk = formats.vk_from_json(open(os.path.join(d, "verification_key.json")).read())
    assert groth16.verify(vk, A, B, C, public)
    assert not groth16.verify(vk, A, B, C, [(public[0] + 1) % c.r] + public[1:])
    # and `co-circom verify` (the product's own host pairing) reaches the same verdicts
    cli.main(["verify", "groth16", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", pub_out, "--curve", cname])
    bad_pub = str(tmp_path / "bad_public.json")
    json.dump([str((public[0] + 1) % c.r)] + [str(v) for v in public[1:]], open(bad_pub, "w"))
    with pytest.raises(SystemExit) as e:
        cli.main(["verify", "groth16", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", bad_pub, "--curve", cname])
    assert e.value.code == 1


def test_split_witness_shamir_files_reconstruct(cocg, tmp_path):
    """SHAMIR -t 2 -n 5: any 3 of the 5 share files interpolate to the witness (shamir/shamir_core.rs:8-118); 2 do not determine it."""
    c = BN254
    d = os.path.join(G, "groth16", "bn254", "poseidon")
    cli = _cli()
    cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "SHAMIR",
              "--curve", "BN254", "--out-dir", str(tmp_path), "-t", "2", "-n", "5"])
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    info = cocg.r1cs_info(os.path.join(d, "circuit.r1cs"))
    sh = []
    for i in range(5):
        pub, (a,) = cocg.shared_witness_decode(cocg.BN254, open(tmp_path / f"witness.wtns.{i}.shared", "rb").read(), 1)
        assert cref.fr_from_mont(c, pub) == [v % c.r for v in wt[:info["num_inputs"]]]
        sh.append(cref.fr_from_mont(c, a))
    want = [v % c.r for v in wt[info["num_inputs"]:]]

    def interpolate(parties):
        xs = [p + 1 for p in parties]
        lag = []
        for i, xi in enumerate(xs):
            num = den = 1
            for j, xj in enumerate(xs):
                if i != j:
                    num = num * xj % c.r
                    den = den * (xj - xi) % c.r
            lag.append(num * pow(den, -1, c.r) % c.r)
        return [sum(l * sh[p][k] for l, p in zip(lag, parties)) % c.r for k in range(len(want))]

    assert interpolate([0, 2, 4]) == want
    assert interpolate([1, 2, 3]) == want
    assert interpolate([0, 1]) != want


def test_shamir_files_to_verified_proof(cocg, tmp_path):
    """SHAMIR (3, 1): split-witness -> generate-proof -> verify, all through the CLI; and the oracle's verifier agrees."""
    d = os.path.join(G, "groth16", "bn254", "poseidon")
    cli = _cli()
    cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "SHAMIR",
              "--curve", "BN254", "--out-dir", str(tmp_path), "-t", "1", "-n", "3"])
    shares = [str(tmp_path / f"witness.wtns.{i}.shared") for i in range(3)]
    out, pub_out = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    cli.main(["generate-proof", "groth16", "--witness", *shares, "--zkey", os.path.join(d, "circuit.zkey"), "--protocol", "SHAMIR", "-t", "1",
              "--curve", "BN254", "--out", out, "--public-input", pub_out])
    cli.main(["verify", "groth16", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", pub_out, "--curve", "BN254"])
    _, A, B, C = formats.proof_from_json(open(out).read())
    vk = formats.vk_from_json(open(os.path.join(d, "verification_key.json")).read())
    assert groth16.verify(vk, A, B, C, [int(x) for x in json.load(open(pub_out))])
    assert json.load(open(pub_out)) == json.load(open(os.path.join(d, "public.json")))


def test_cli_rejects_bad_arguments(cocg, tmp_path):
    cli = _cli()
    d = os.path.join(G, "groth16", "bn254", "multiplier2")
    with pytest.raises(SystemExit, match="threshold to be 1"):
        cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "REP3",
                  "--curve", "BN254", "--out-dir", str(tmp_path), "-t", "2"])
    with pytest.raises(SystemExit, match="different curve"):
        cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "REP3",
                  "--curve", "BLS12-381", "--out-dir", str(tmp_path)])
    with pytest.raises(SystemExit, match="does not exist"):
        cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "REP3",
                  "--curve", "BN254", "--out-dir", str(tmp_path / "nope")])


def test_plonk_files_to_verified_proof(cocg, tmp_path):
    """`generate-proof plonk` (co-circom.rs:455-636 with the Plonk proof system): witness.wtns -> three REP3 share files -> CoPlonk::prove
    on the GPU -> proof.json accepted by `verify plonk` with the fixture's snarkjs verification key; a changed public input is rejected."""
    d = os.path.join(G, "plonk", "bn254", "multiplier2")
    r1cs = os.path.join(d, "circuit.r1cs")
    cli = _cli()
    cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", r1cs, "--protocol", "REP3", "--curve", "BN254",
              "--out-dir", str(tmp_path)])
    shares = [str(tmp_path / f"witness.wtns.{i}.shared") for i in range(3)]
    out, pub_out = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    cli.main(["generate-proof", "plonk", "--witness", *shares, "--zkey", os.path.join(d, "circuit.zkey"), "--protocol", "REP3", "--curve", "BN254",
              "--out", out, "--public-input", pub_out])
    assert json.load(open(pub_out)) == json.load(open(os.path.join(d, "public.json")))
    assert json.load(open(out))["protocol"] == "plonk"
    cli.main(["verify", "plonk", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", pub_out, "--curve", "BN254"])
    public = json.load(open(pub_out))
    bad_pub = str(tmp_path / "bad_public.json")
    json.dump([str(int(public[0]) + 1)] + public[1:], open(bad_pub, "w"))
    with pytest.raises(SystemExit) as e:
        cli.main(["verify", "plonk", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", bad_pub, "--curve", "BN254"])
    assert e.value.code == 1


def test_plonk_shamir_files_to_verified_proof(cocg, tmp_path):
    """The same flow over Shamir (5, 2) share files (`--protocol SHAMIR -t 2 -n 5`, co-circom.rs:455-636)."""
    d = os.path.join(G, "plonk", "bn254", "multiplier2")
    cli = _cli()
    cli.main(["split-witness", "--witness", os.path.join(d, "witness.wtns"), "--r1cs", os.path.join(d, "circuit.r1cs"), "--protocol", "SHAMIR",
              "--curve", "BN254", "--out-dir", str(tmp_path), "-t", "2", "-n", "5"])
    shares = [str(tmp_path / f"witness.wtns.{i}.shared") for i in range(5)]
    out, pub_out = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    cli.main(["generate-proof", "plonk", "--witness", *shares, "--zkey", os.path.join(d, "circuit.zkey"), "--protocol", "SHAMIR", "-t", "2",
              "--curve", "BN254", "--out", out, "--public-input", pub_out])
    assert json.load(open(pub_out)) == json.load(open(os.path.join(d, "public.json")))
    cli.main(["verify", "plonk", "--proof", out, "--vk", os.path.join(d, "verification_key.json"), "--public-input", pub_out, "--curve", "BN254"])
    with pytest.raises(SystemExit, match="n > 2t"):
        cli.main(["generate-proof", "plonk", "--witness", *shares[:4], "--zkey", os.path.join(d, "circuit.zkey"), "--protocol", "SHAMIR", "-t", "2",
                  "--curve", "BN254", "--out", out])
