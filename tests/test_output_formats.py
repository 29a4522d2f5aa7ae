"""CPU-only: the product's output-side formats (collaborative-circom_b200/host/serialize.hpp through the C ABI of include/cohost.h)
against the reference's own fixtures.

  circom-types/src/groth16/proof.rs:31-99      the snarkjs circom.proof files (BN254 and BLS12-381): our writer, fed the fixture's
                                               points, must reproduce the fixture's JSON value exactly (keys, order, decimal strings)
  co-circom/src/bin/co-circom.rs:611-629       public.json: decimal strings of the public inputs without the leading 1
  co-circom-snarks/src/lib.rs:24-41,           SharedWitness files: bincode + ark-serialize layout, checked byte-for-byte against an
  serde_compat.rs:5-24                         independent Python restatement and by decode(encode(x)) == x; error paths
"""
import json
import os
import random
import struct

import numpy as np
import pytest

from oracle import cref, formats
from oracle.curves import BN254, BLS12_381

G = os.path.join(os.path.dirname(__file__), "golden")
CURVES = {"bn254": BN254, "bls12_381": BLS12_381}


def _cid(cocg, c):
    return cocg.BN254 if c is BN254 else cocg.BLS12_381


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bn254", "poseidon"), ("bls12_381", "multiplier2"), ("bls12_381", "poseidon")])
def test_proof_json_reproduces_the_snarkjs_fixture(cocg, curve, circ):
    c = CURVES[curve]
    text = open(os.path.join(G, "groth16", curve, circ, "circom.proof")).read()
    _, A, B, C = formats.proof_from_json(text)
    block = np.concatenate([cref.g_to_mont(c, [A], 1).ravel(), cref.g_to_mont(c, [B], 2).ravel(), cref.g_to_mont(c, [C], 1).ravel()])
    ours = cocg.proof_to_json(_cid(cocg, c), block)
    assert " " not in ours and "\n" not in ours                     # serde_json::to_writer: compact
    want, got = json.loads(text), json.loads(ours)
    assert list(got.keys()) == ["pi_a", "pi_b", "pi_c", "protocol", "curve"]  # struct field order of Groth16Proof
    assert got == want
    assert formats.proof_from_json(ours)[1:] == (A, B, C)


def test_proof_json_infinity_and_errors(cocg):
    c = BN254
    g2 = cref.g_to_mont(c, [c.gen(2)], 2).ravel()
    block = np.concatenate([np.zeros(8, dtype=np.uint64), g2, cref.g_to_mont(c, [c.gen(1)], 1).ravel()])
    got = json.loads(cocg.proof_to_json(cocg.BN254, block))
    assert got["pi_a"] == ["0", "1", "0"]                           # g1_to_strings_projective, traits.rs:186-193
    assert got["pi_c"] == ["1", "2", "1"]
    block[8:24] = 0
    with pytest.raises(cocg.CocgError, match="infinity"):           # serialize_g2 unwraps xy(): no encoding exists
        cocg.proof_to_json(cocg.BN254, block)
    with pytest.raises(cocg.CocgError, match="curve"):
        cocg.proof_to_json(7, block)


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bn254", "poseidon"), ("bls12_381", "poseidon")])
def test_public_inputs_json_matches_fixture(cocg, curve, circ):
    c = CURVES[curve]
    d = os.path.join(G, "groth16", curve, circ)
    want = json.load(open(os.path.join(d, "public.json")))
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    pub = cref.fr_to_mont(c, wt[:1 + len(want)])
    ours = cocg.public_inputs_to_json(_cid(cocg, c), pub)
    assert json.loads(ours) == want and " " not in ours
    assert cocg.public_inputs_to_json(_cid(cocg, c), pub[:1]) == "[]"
    zero = cref.fr_to_mont(c, [1, 0, c.r - 1])
    assert json.loads(cocg.public_inputs_to_json(_cid(cocg, c), zero)) == ["0", str(c.r - 1)]


def _ark_vec(vals):
    return struct.pack("<Q", len(vals)) + b"".join(int(v).to_bytes(32, "little") for v in vals)


def _bincode_shared_witness(pub, comps):
    """Independent restatement: bincode(serialize_bytes) = u64 length + bytes; ark compressed Vec<F> = u64 count + LE elements."""
    p = _ark_vec(pub)
    w = b"".join(_ark_vec(c) for c in comps)
    return struct.pack("<Q", len(p)) + p + struct.pack("<Q", len(w)) + w


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
@pytest.mark.parametrize("k", [2, 1])
def test_shared_witness_file_layout_and_round_trip(cocg, curve, k):
    c = CURVES[curve]
    rng = random.Random(5 + k)
    for n_pub, n in ((2, 7), (1, 0), (3, 300)):
        pub = [1] + [rng.randrange(c.r) for _ in range(n_pub - 1)]
        comps = [[rng.randrange(c.r) for _ in range(n)] for _ in range(k)]
        if n:
            comps[0][0], comps[-1][-1] = 0, c.r - 1
        mp, mc = cref.fr_to_mont(c, pub), [cref.fr_to_mont(c, x) if n else np.zeros((0, 4), dtype=np.uint64) for x in comps]
        img = cocg.shared_witness_encode(_cid(cocg, c), mp, mc)
        assert img == _bincode_shared_witness(pub, comps)
        dp, dc = cocg.shared_witness_decode(_cid(cocg, c), img, k)
        assert np.array_equal(dp, mp)
        for a, b in zip(dc, mc):
            assert np.array_equal(a, b)


def test_shared_witness_decode_rejects_bad_files(cocg):
    c = BN254
    pub, comps = [1, 5], [[3, 4], [6, 7]]
    good = _bincode_shared_witness(pub, comps)
    with pytest.raises(cocg.CocgError, match="truncated"):
        cocg.shared_witness_decode(cocg.BN254, good[:-5], 2)
    with pytest.raises(cocg.CocgError, match="trailing|wrong protocol"):       # a REP3 file read as a Shamir share
        cocg.shared_witness_decode(cocg.BN254, good, 1)
    bad = _bincode_shared_witness(pub, [[c.r, 4], [6, 7]])                      # element == modulus: Validate::Yes rejects
    with pytest.raises(cocg.CocgError, match="canonical"):
        cocg.shared_witness_decode(cocg.BN254, bad, 2)
    ragged = _bincode_shared_witness(pub, [[3, 4], [6]])
    with pytest.raises(cocg.CocgError, match="length"):
        cocg.shared_witness_decode(cocg.BN254, ragged, 2)
    huge = struct.pack("<Q", 1 << 60) + good[8:]
    with pytest.raises(cocg.CocgError, match="truncated"):
        cocg.shared_witness_decode(cocg.BN254, huge, 2)


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_r1cs_header_matches_reference_kat(cocg, curve):
    """circom-types/src/r1cs.rs:280-330: multiplier2 has num_inputs 2 (n_pub_out 1, n_pub_in 0), 4 wires, 1 constraint."""
    info = cocg.r1cs_info(os.path.join(G, "groth16", curve, "multiplier2", "circuit.r1cs"))
    assert info == {"curve": _cid(cocg, CURVES[curve]), "n_wires": 4, "n_pub_out": 1, "n_pub_in": 0, "n_constraints": 1, "num_inputs": 2}
    zk = formats.parse_groth16_zkey(open(os.path.join(G, "groth16", curve, "poseidon", "circuit.zkey"), "rb").read(), check_points=False)
    info = cocg.r1cs_info(os.path.join(G, "groth16", curve, "poseidon", "circuit.r1cs"))
    assert (info["n_wires"], info["num_inputs"]) == (zk.n_vars, zk.n_public + 1)
    with pytest.raises(cocg.CocgError, match="r1cs"):
        cocg.r1cs_info(os.path.join(G, "groth16", curve, "multiplier2", "witness.wtns"))


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_plonk_zkey_header_reader_agrees_with_verification_key_json(cocg, curve):
    """The product's C++ Plonk zkey reader (host/plonk.hpp, circom-types/src/plonk/zkey.rs:329-420) on the full multiplier2 keys: the
    verifying-key tail of the header equals the snarkjs verification_key.json; the trimmed round-1 fixture reports no optional parts."""
    c = CURVES[curve]
    d = os.path.join(G, "plonk", curve, "multiplier2")
    h = cocg.plonk_zkey_header(os.path.join(d, "circuit.zkey"))
    vk = json.load(open(os.path.join(d, "verification_key.json")))
    assert (h["curve"], h["n_public"], h["domain_size"], h["parts"]) == (_cid(cocg, c), vk["nPublic"], 1 << vk["power"], 15)
    assert cref.fr_from_mont(c, h["k"]) == [int(vk["k1"]), int(vk["k2"])]
    got = cref.g_from_mont(c, h["vk_g1"], 1)
    for P, name in zip(got, ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3")):
        v = vk[name]
        assert P == (None if v[2] == "0" else (int(v[0]), int(v[1]))), name
    x2 = vk["X_2"]
    assert cref.g_from_mont(c, h["x_2"], 2)[0] == ((int(x2[0][0]), int(x2[0][1])), (int(x2[1][0]), int(x2[1][1])))
    if curve == "bn254":
        t = cocg.plonk_zkey_header(os.path.join(d, "circuit.round1.zkey"))
        assert t["parts"] & 14 == 0 and t["n_constraints"] == h["n_constraints"]


def test_shared_witness_bytes_hand_derived(cocg):
    """A `.shared` file image derived BY HAND from the reference's code, as a literal -- not produced by any encoder:

    * co-circom-snarks/src/lib.rs:24-41: `SharedWitness { public_inputs: Vec<F>, witness: FieldShareVec }`, both fields with
      `serialize_with = ark_se`; files are written with `bincode::serialize_into(out_file, share)` (co-circom/src/bin/co-circom.rs:215, 244).
    * serde_compat.rs:5-13 (ark_se): `a.serialize_with_mode(&mut bytes, Compress::Yes)` then `s.serialize_bytes(&bytes)`.
    * bincode 1.x default options: fixed-width little-endian integers; a struct is its fields in order with no framing;
      `serialize_bytes` = u64 length, then the bytes.
    * ark-serialize 0.4: `Vec<T>` = u64 length (LE) then the elements; a prime-field element of a 254 / 255-bit modulus = 32 bytes,
      little-endian, canonical (non-Montgomery) value; `#[derive(CanonicalSerialize)]` on `Rep3PrimeFieldShareVec { a, b }`
      (mpc-core/src/protocols/rep3/fieldshare.rs:231-236) = a then b.

    public_inputs = [1, 5], witness a = [7], b = [9] (BN254):
      field 1: u64 72 | [u64 2 | 1 as 32 LE bytes | 5 as 32 LE bytes]                          (8 + 72 bytes)
      field 2: u64 80 | [u64 1 | 7 as 32 LE bytes | u64 1 | 9 as 32 LE bytes]                   (8 + 80 bytes)"""
    le32 = lambda v: "%02x" % v + "00" * 31
    u64 = lambda v: "%02x" % v + "00" * 7
    literal = bytes.fromhex(u64(72) + u64(2) + le32(1) + le32(5) + u64(80) + u64(1) + le32(7) + u64(1) + le32(9))
    assert len(literal) == 168
    c = BN254
    img = cocg.shared_witness_encode(cocg.BN254, cref.fr_to_mont(c, [1, 5]), [cref.fr_to_mont(c, [7]), cref.fr_to_mont(c, [9])])
    assert img == literal
    pub, comps = cocg.shared_witness_decode(cocg.BN254, literal, 2)
    assert cref.fr_from_mont(c, pub) == [1, 5] and cref.fr_from_mont(c, comps[0]) == [7] and cref.fr_from_mont(c, comps[1]) == [9]
    # Shamir: ShamirPrimeFieldShareVec { a } (shamir/fieldshare.rs:152-155) -- one vector
    literal_shamir = bytes.fromhex(u64(72) + u64(2) + le32(1) + le32(5) + u64(40) + u64(1) + le32(7))
    assert cocg.shared_witness_encode(cocg.BN254, cref.fr_to_mont(c, [1, 5]), [cref.fr_to_mont(c, [7])]) == literal_shamir


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_plonk_proof_json_writer_reproduces_snarkjs_fixture(cocg, curve):
    """PlonkProof serde layout (circom-types/src/plonk/proof.rs:7-87; its own test reads the same fixtures, :97-190): the product's writer,
    fed the points and evaluations of the snarkjs `circom.proof`, reproduces the fixture's JSON value."""
    c = BN254 if curve == "bn254" else BLS12_381
    want = json.load(open(os.path.join(G, "plonk", curve, "multiplier2", "circom.proof")))
    pts = [None if want[k][2] == "0" else (int(want[k][0]), int(want[k][1])) for k in ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")]
    evs = [int(want[k]) for k in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw")]
    block = np.concatenate([cref.g_to_mont(c, pts, 1).reshape(-1), cref.fr_to_mont(c, evs).reshape(-1)])
    got = json.loads(cocg.plonk_proof_to_json(_cid(cocg, c), block))
    assert got == want


def test_block_plan_covers_every_block_once(cocg):
    """Multi-GPU block mode (host/types.hpp BlockPlan): the 3 witness maps, 6 b_g2 MSMs and 18 G1 MSMs of a proof are all placed, on valid
    ranks, the modelled load of the slowest rank stays within 6 % of the mean up to 8 ranks, and with three or more ranks the three
    witness maps run on three different GPUs."""
    for world in (1, 2, 3, 4, 5, 8, 16):
        p = cocg.block_plan(world)
        g1 = [r for party in p["g1"] for comp in party for r in comp]
        g2 = [r for pair in p["g2"] for r in pair]
        assert len(p["wm"]) == 3 and len(g2) == 6 and len(g1) == 18
        assert all(0 <= r < world for r in p["wm"] + g1 + g2)
        load = [9.7 * p["wm"].count(r) + 8.3 * g2.count(r) + 2.6 * g1.count(r) for r in range(world)]
        if world <= 8:
            assert max(load) <= 1.06 * sum(load) / world, (world, load)
        if world >= 3:
            assert len(set(p["wm"])) == 3
