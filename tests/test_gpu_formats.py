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


def test_zkey_load_rejects_points_off_curve_or_outside_the_subgroup(cocg, bn, bls, tmp_path):
    """circom-types/src/traits.rs:107-155 (g1_from_bytes / g2_from_bytes): a point that is off the curve, or on it but outside the
    prime-order subgroup, makes the parser fail.  Here the check is a kernel over the resident query (cocg_bases_check): the genuine
    fixtures load, a corrupted coordinate is refused, and a curve point outside the subgroup is refused where a cofactor exists
    (BLS12-381 G1: h = 0x396c8c005555e1568c00aaab0000aaab; BN254 G1 has cofactor 1)."""
    from oracle import formats
    for curve, c, ctx in (("bn254", BN254, bn), ("bls12_381", BLS12_381, bls)):
        path = os.path.join(G, "groth16", curve, "multiplier2", "circuit.zkey")
        raw = bytearray(open(path, "rb").read())
        zk = cocg.Groth16ZKey.from_file(path)           # genuine file: accepted
        zk.close()
        ozk = formats.parse_groth16_zkey(bytes(raw))
        # flip one byte inside the first point of the H section (section 9): almost surely off the curve
        off = _section_offset(bytes(raw), 9)
        bad = bytearray(raw)
        bad[off + 3] ^= 0x55
        p = tmp_path / f"bad_{curve}.zkey"
        p.write_bytes(bytes(bad))
        with pytest.raises(cocg.CocgError, match="InvalidData|not on the curve"):
            cocg.Groth16ZKey.from_file(str(p))
        # kernel-level: genuine points pass, a curve point of the wrong order fails where the group has a cofactor
        pts = cref.g_to_mont(c, ozk.h_query, 1)
        h = ctx.bases_upload(1, pts)
        assert ctx.bases_check(h) == (0, None)
        ctx.bases_free(h)
        if c is BLS12_381:
            x = 1
            while True:  # a point of E(Fq) that is not in G1: take any curve point and keep it unless it happens to have order r
                rhs = (x * x * x + 4) % c.q
                y = pow(rhs, (c.q + 1) // 4, c.q)
                if y * y % c.q == rhs and c.to_affine(c.jac_mul(c.to_jac((x, y), 1), c.r, 1), 1) is not None:  # (Curve.mul reduces the scalar mod r)
                    break
                x += 1
            pts2 = pts.copy()
            pts2[1] = cref.g_to_mont(c, [(x, y)], 1)[0]
            h2 = ctx.bases_upload(1, pts2)
            assert ctx.bases_check(h2, subgroup=True) == (1, 1)
            assert ctx.bases_check(h2, subgroup=False) == (0, None)    # it IS on the curve
            ctx.bases_free(h2)


def _section_offset(data, want):
    """byte offset of the payload of section `want` in a snarkjs binfile (circom-types/src/binfile.rs:52-105)"""
    import struct
    nsec = struct.unpack_from("<I", data, 8)[0]
    off = 12
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", data, off)
        off += 12
        if sid == want:
            return off
        off += ln
    raise KeyError(want)
