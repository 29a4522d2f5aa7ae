"""CPU-only: the product's Groth16 verifier (collaborative-circom_b200/host/pairing.hpp, verify_json.hpp through the C ABI) on the
reference's snarkjs fixtures -- the same known-answer test that pins the oracle's pairing:

  test_vectors/Groth16/{bn254,bls12_381}/{multiplier2,poseidon}/{circom.proof, public.json, verification_key.json}
  co-groth16/src/lib.rs:26-206 (verify-only tests), co-circom/src/bin/co-circom.rs:640-720 (`co-circom verify`)

Accept the four shipped proofs; reject a changed public input, a changed proof element, a proof under the other circuit's key;
fail loudly (non-zero return, not "rejected") on malformed input, as the reference does at deserialisation."""
import json
import os
import time

import numpy as np
import pytest

from oracle import cref, formats, groth16
from oracle.curves import BN254, BLS12_381

G = os.path.join(os.path.dirname(__file__), "golden")
CURVES = {"bn254": BN254, "bls12_381": BLS12_381}
CASES = [("bn254", "multiplier2"), ("bn254", "poseidon"), ("bls12_381", "multiplier2"), ("bls12_381", "poseidon")]


def _load(curve, circ):
    d = os.path.join(G, "groth16", curve, circ)
    return tuple(open(os.path.join(d, f)).read() for f in ("verification_key.json", "circom.proof", "public.json"))


@pytest.mark.parametrize("curve,circ", CASES)
def test_accepts_the_snarkjs_proofs_and_rejects_changed_inputs(cocg, curve, circ):
    c = CURVES[curve]
    vk, proof, pub = _load(curve, circ)
    t0 = time.perf_counter()
    assert cocg.groth16_verify_json(vk, proof, pub) is True
    assert time.perf_counter() - t0 < 5.0
    p = json.loads(pub)
    bad = json.dumps([str((int(p[0]) + 1) % c.r)] + p[1:])
    assert cocg.groth16_verify_json(vk, proof, bad) is False
    # a different, valid group element in the proof: C := A
    pj = json.loads(proof)
    pj["pi_c"] = pj["pi_a"]
    assert cocg.groth16_verify_json(vk, json.dumps(pj), pub) is False
    # the oracle agrees on all three verdicts
    ovk = formats.vk_from_json(vk)
    _, A, B, C = formats.proof_from_json(proof)
    assert groth16.verify(ovk, A, B, C, [int(x) for x in p])
    assert not groth16.verify(ovk, A, B, A, [int(x) for x in p])


def test_proof_under_another_circuits_key_is_rejected(cocg):
    vk_m, _, _ = _load("bn254", "multiplier2")
    _, proof_p, pub_p = _load("bn254", "poseidon")
    vk_p, _, _ = _load("bn254", "poseidon")
    if len(json.loads(vk_m)["IC"]) == len(json.loads(vk_p)["IC"]):
        assert cocg.groth16_verify_json(vk_m, proof_p, pub_p) is False
    else:
        with pytest.raises(cocg.CocgError, match="number of public inputs"):
            cocg.groth16_verify_json(vk_m, proof_p, pub_p)
    vk_b, _, _ = _load("bls12_381", "poseidon")
    with pytest.raises(cocg.CocgError, match="different curves"):
        cocg.groth16_verify_json(vk_b, proof_p, pub_p)


def test_malformed_inputs_fail_loudly(cocg):
    c = BN254
    vk, proof, pub = _load("bn254", "multiplier2")
    pj = json.loads(proof)
    off = dict(pj)
    off["pi_a"] = [pj["pi_a"][0], str((int(pj["pi_a"][1]) + 1) % c.q), "1"]
    with pytest.raises(cocg.CocgError, match="not on the curve"):
        cocg.groth16_verify_json(vk, json.dumps(off), pub)
    big = dict(pj)
    big["pi_a"] = [str(c.q), pj["pi_a"][1], "1"]
    with pytest.raises(cocg.CocgError, match="larger than the modulus"):
        cocg.groth16_verify_json(vk, json.dumps(big), pub)
    with pytest.raises(cocg.CocgError, match="number of public inputs"):
        cocg.groth16_verify_json(vk, proof, "[]")
    with pytest.raises(cocg.CocgError, match="json"):
        cocg.groth16_verify_json(vk, proof[:-3], pub)
    with pytest.raises(cocg.CocgError, match="missing key"):
        cocg.groth16_verify_json(vk, json.dumps({k: v for k, v in pj.items() if k != "pi_b"}), pub)
    with pytest.raises(cocg.CocgError, match="decimal"):
        cocg.groth16_verify_json(vk, proof, '["0x21"]')
    # a G2 point on the twist but outside the prime-order subgroup (the BN254 twist has a cofactor of ~2^254)
    q = c.q

    def fq_sqrt(v):
        r = pow(v, (q + 1) // 4, q)          # q = 3 mod 4
        return r if r * r % q == v % q else None

    def f2_sqrt(a):
        a0, a1 = a
        alpha = fq_sqrt((a0 * a0 + a1 * a1) % q)
        if alpha is None:
            return None
        for sign in (1, -1):
            delta = (a0 + sign * alpha) * pow(2, -1, q) % q
            x0 = fq_sqrt(delta)
            if x0:
                x1 = a1 * pow(2 * x0, -1, q) % q
                if c.f2_sqr((x0, x1)) == (a0 % q, a1 % q):
                    return (x0, x1)
        return None

    found = None
    for k in range(1, 400):
        x = (k, 1)
        y = f2_sqrt(c.f2_add(c.f2_mul(c.f2_sqr(x), x), c.b2))
        # in the subgroup iff (r - 1) P == -P   (Curve.mul reduces its scalar mod r, so r itself cannot be used)
        if y is not None and c.mul((x, y), c.r - 1, 2) != (x, c.f2_neg(y)):
            found = (x, y)
            break
    assert found is not None
    sub = dict(pj)
    sub["pi_b"] = [[str(found[0][0]), str(found[0][1])], [str(found[1][0]), str(found[1][1])], ["1", "0"]]
    with pytest.raises(cocg.CocgError, match="subgroup"):
        cocg.groth16_verify_json(vk, json.dumps(sub), pub)


@pytest.mark.parametrize("curve,circ", [("bn254", "poseidon"), ("bls12_381", "multiplier2")])
def test_binary_entry_point_matches_json(cocg, curve, circ):
    c = CURVES[curve]
    vk, proof, pub = _load(curve, circ)
    ovk = formats.vk_from_json(vk)
    _, A, B, C = formats.proof_from_json(proof)
    vkb = np.concatenate([cref.g_to_mont(c, [ovk.alpha_g1], 1).ravel(), cref.g_to_mont(c, [ovk.beta_g2], 2).ravel(),
                          cref.g_to_mont(c, [ovk.gamma_g2], 2).ravel(), cref.g_to_mont(c, [ovk.delta_g2], 2).ravel()])
    ic = cref.g_to_mont(c, ovk.ic, 1)
    block = np.concatenate([cref.g_to_mont(c, [A], 1).ravel(), cref.g_to_mont(c, [B], 2).ravel(), cref.g_to_mont(c, [C], 1).ravel()])
    p = [int(x) for x in json.loads(pub)]
    cid = cocg.BN254 if c is BN254 else cocg.BLS12_381
    assert cocg.groth16_verify(cid, vkb, ic, block, cref.fr_to_mont(c, p)) is True
    assert cocg.groth16_verify(cid, vkb, ic, block, cref.fr_to_mont(c, [(p[0] + 5) % c.r] + p[1:])) is False
    # our own JSON writer round-trips into the verifier
    assert cocg.groth16_verify_json(vk, cocg.proof_to_json(cid, block), cocg.public_inputs_to_json(cid, cref.fr_to_mont(c, [1] + p))) is True


# ------------------------------------------------------------------------------------------------ Plonk
PLONK_CASES = [("bn254", "multiplier2"), ("bn254", "poseidon"), ("bls12_381", "multiplier2"), ("bls12_381", "poseidon")]


def _load_plonk(curve, circ):
    d = os.path.join(G, "plonk", curve, circ)
    return tuple(open(os.path.join(d, f)).read() for f in ("verification_key.json", "circom.proof", "public.json"))


@pytest.mark.parametrize("curve,circ", PLONK_CASES)
def test_plonk_verifier_accepts_snarkjs_proofs_and_rejects_changes(cocg, curve, circ):
    """co-plonk/src/lib.rs:255-275 (`Plonk::verify` on the shipped snarkjs proofs), here through the product's host verifier."""
    c = CURVES[curve]
    vk, proof, pub = _load_plonk(curve, circ)
    assert cocg.plonk_verify_json(vk, proof, pub) is True
    p = json.loads(pub)
    bad = json.dumps([str((int(p[0]) + 1) % c.r)] + p[1:])
    assert cocg.plonk_verify_json(vk, proof, bad) is False
    pj = json.loads(proof)
    for key in ("eval_a", "eval_zw"):
        t = dict(pj)
        t[key] = str((int(pj[key]) + 1) % c.r)
        assert cocg.plonk_verify_json(vk, json.dumps(t), pub) is False
    t = dict(pj)
    t["Wxi"] = pj["Wxiw"]
    assert cocg.plonk_verify_json(vk, json.dumps(t), pub) is False
    with pytest.raises(cocg.CocgError, match="number of public inputs"):
        cocg.plonk_verify_json(vk, proof, json.dumps(p + ["1"]))
    off = dict(pj)
    off["Z"] = [pj["Z"][0], str((int(pj["Z"][1]) + 1) % c.q), "1"]
    with pytest.raises(cocg.CocgError, match="not on the curve"):
        cocg.plonk_verify_json(vk, json.dumps(off), pub)


def test_plonk_verifier_challenges_match_reference_kat(cocg):
    """co-plonk/src/plonk.rs:285-350: alpha, beta, gamma, xi, v, u for the multiplier2 proof -- pins the product's Keccak-256
    transcript (C++) literally."""
    kat = json.load(open(os.path.join(G, "plonk_round2_kats.json")))["verifier_challenges"]
    vk, proof, pub = _load_plonk("bn254", "multiplier2")
    ok, ch = cocg.plonk_verify_json(vk, proof, pub, want_challenges=True)
    assert ok
    got = cref.fr_from_mont(BN254, ch)
    assert got[:4] == [int(kat[k]) for k in ("alpha", "beta", "gamma", "xi")]
    assert got[4] == int(kat["v"][0]) and got[5] == int(kat["u"])
    # a groth16 key is not a plonk key
    gvk, gproof, gpub = _load("bn254", "multiplier2")
    with pytest.raises(cocg.CocgError, match="plonk|missing key"):
        cocg.plonk_verify_json(gvk, gproof, gpub)


def test_json_readers_survive_mutated_inputs(cocg):
    """The verifiers parse untrusted text: truncations and byte flips of the fixture files must end in a verdict or a CocgError,
    never in a crash (the C ABI promises 0 / non-zero + cohost_last_error, nothing throws or aborts)."""
    import random
    rng = random.Random(2024)
    g = _load("bn254", "multiplier2")
    p = _load_plonk("bn254", "multiplier2")
    outcomes = {"ok": 0, "rejected": 0, "error": 0}
    for fn, files in ((cocg.groth16_verify_json, g), (cocg.plonk_verify_json, p)):
        for trial in range(120):
            docs = list(files)
            k = rng.randrange(3)
            text = docs[k]
            mode = trial % 4
            if mode == 0:
                text = text[:rng.randrange(len(text))]
            elif mode == 1:
                i = rng.randrange(len(text))
                text = text[:i] + rng.choice('{}[],:"0123456789x \\') + text[i + 1:]
            elif mode == 2:
                i = rng.randrange(len(text))
                text = text[:i] + text[i + rng.randrange(1, 40):]
            else:
                text = text.replace('"1"', rng.choice(['"0"', '"2"', '1', '[]', '{}']), rng.randrange(1, 4))
            docs[k] = text
            try:
                outcomes["ok" if fn(*docs) else "rejected"] += 1
            except cocg.CocgError:
                outcomes["error"] += 1
    assert outcomes["error"] > 50 and sum(outcomes.values()) == 240


def test_cli_verify_subcommand_on_fixtures(cocg, tmp_path, capsys):
    """`co-circom verify groth16|plonk --proof --vk --public-input --curve` (co-circom/src/bin/co-circom.rs:640-720) needs no GPU."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import co_circom
    for system, sub in (("groth16", "groth16"), ("plonk", "plonk")):
        for curve, cname in (("bn254", "BN254"), ("bls12_381", "BLS12-381")):
            d = os.path.join(G, sub, curve, "multiplier2")
            args = ["verify", system, "--proof", os.path.join(d, "circom.proof"), "--vk", os.path.join(d, "verification_key.json"),
                    "--public-input", os.path.join(d, "public.json"), "--curve", cname]
            co_circom.main(args)
            assert "verified successfully" in capsys.readouterr().out
            bad = tmp_path / f"{system}_{curve}_public.json"
            pub = json.load(open(os.path.join(d, "public.json")))
            bad.write_text(json.dumps([str(int(pub[0]) + 1)] + pub[1:]))
            args[args.index("--public-input") + 1] = str(bad)
            with pytest.raises(SystemExit) as e:
                co_circom.main(args)
            assert e.value.code == 1
            other = "BLS12-381" if cname == "BN254" else "BN254"
            with pytest.raises(SystemExit, match="different curve"):
                co_circom.main(args[:-1] + [other])
