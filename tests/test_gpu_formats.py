"""The product's own snarkjs readers (C++ host layer, collaborative-circom_b200/host/formats.hpp) against the reference's zkey
known-answer tests and against whole proofs.

  circom-types/src/groth16/zkey.rs:335-585   every point + the matrices of the multiplier2 zkeys (BN254, BLS12-381)
  circom-types/src/witness.rs:94-135          witness values
  co-groth16/src/lib.rs:26-206                zkey + wtns from files -> prove -> verify
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import cref, formats, groth16
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CURVES = {"bn254": BN254, "bls12_381": BLS12_381}


def _pt(v):
    if v is None:
        return None
    if isinstance(v[0], list):
        return ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))
    return (int(v[0]), int(v[1]))


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_zkey_reader_matches_reference_literals(cocg, curve):
    c = CURVES[curve]
    kat = json.load(open(os.path.join(G, "zkey_kats.json")))[curve]
    zk = cocg.Groth16ZKey.from_file(os.path.join(G, "groth16", curve, "multiplier2", "circuit.zkey"))
    assert (zk.n_public, zk.n_vars, zk.num_constraints, 1 << zk.pow) == (1, 4, 1, 4)
    for name, (_, group) in zk.QUERIES.items():
        assert cref.g_from_mont(c, zk.query(name), group) == [_pt(p) for p in kat[name]], name
    vk = zk.vk()
    for name, group in (("alpha_g1", 1), ("beta_g1", 1), ("delta_g1", 1), ("beta_g2", 2), ("delta_g2", 2)):
        assert cref.g_from_mont(c, vk[name], group)[0] == _pt(kat[name]), name
    # matrices (zkey.rs:568-584): A = [[(r - 1, 2)]], B = [[(1, 3)]]; the public-input rows are dropped
    ra, ca, va = zk.matrix(0)
    rb, cb, vb = zk.matrix(1)
    assert list(ra) == [0, 1] and list(ca) == [2] and cref.fr_from_mont(c, va) == [c.r - 1]
    assert list(rb) == [0, 1] and list(cb) == [3] and cref.fr_from_mont(c, vb) == [1]
    zk.close()


@pytest.mark.parametrize("curve,circ", [("bn254", "poseidon"), ("bls12_381", "poseidon"), ("bn254", "multiplier2")])
def test_files_to_proof(cocg, curve, circ):
    """zkey + wtns read by the product, proof by the GPU path, checked by the oracle (its own readers, its own pairing)."""
    d = os.path.join(G, "groth16", curve, circ)
    zk = cocg.Groth16ZKey.from_file(os.path.join(d, "circuit.zkey"))
    ozk = formats.parse_groth16_zkey(open(os.path.join(d, "circuit.zkey"), "rb").read())
    c = ozk.curve
    _, owt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    wt = zk.load_witness(os.path.join(d, "witness.wtns"))
    assert cref.fr_from_mont(c, wt) == [v % c.r for v in owt]
    # every query and both matrices equal the oracle's parse
    for name, (_, group) in zk.QUERIES.items():
        assert cref.g_from_mont(c, zk.query(name), group) == getattr(ozk, name), name
    for which, rows in ((0, ozk.a_rows), (1, ozk.b_rows)):
        rp, col, val = zk.matrix(which)
        vals = cref.fr_from_mont(c, val)
        got = [[(vals[k], int(col[k])) for k in range(rp[i], rp[i + 1])] for i in range(zk.num_constraints)]
        assert got == rows
    vk = formats.vk_from_json(open(os.path.join(d, "verification_key.json")).read())
    public = [int(x) for x in json.load(open(os.path.join(d, "public.json")))]
    sess = cocg.PlainSession(zk)
    rng = random.Random(9)
    r, s = rng.randrange(c.r), rng.randrange(c.r)
    ell = zk.n_public
    proof = sess.prove(wt[:ell + 1], wt[ell + 1:], cref.fr_to_mont(c, [r]), cref.fr_to_mont(c, [s]))
    lq = cref.lq(c)
    A = cref.g_from_mont(c, proof[:2 * lq], 1)[0]
    B = cref.g_from_mont(c, proof[2 * lq:6 * lq], 2)[0]
    C = cref.g_from_mont(c, proof[6 * lq:], 1)[0]
    assert (A, B, C) == groth16.prove_plain(ozk, owt, r, s)
    assert groth16.verify(vk, A, B, C, public)
    sess.close()
    zk.close()


def test_reader_rejects_malformed_files(cocg, tmp_path):
    good = open(os.path.join(G, "groth16", "bn254", "multiplier2", "circuit.zkey"), "rb").read()
    for name, blob in (("magic", b"zkex" + good[4:]), ("truncated", good[:200]), ("empty", b"")):
        p = tmp_path / (name + ".zkey")
        p.write_bytes(blob)
        with pytest.raises(cocg.CocgError):
            cocg.Groth16ZKey.from_file(str(p))
    with pytest.raises(cocg.CocgError, match="cannot open"):
        cocg.Groth16ZKey.from_file(str(tmp_path / "missing.zkey"))
    # a plonk zkey (protocol id 2) is not a groth16 key
    with pytest.raises(cocg.CocgError, match="groth16"):
        cocg.Groth16ZKey.from_file(os.path.join(G, "plonk", "bn254", "multiplier2", "circuit.round1.zkey"))
