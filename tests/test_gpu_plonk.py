"""Round 1 of the Plonk prover through the C++ host layer (host/plonk.hpp) against the reference's bit-exact KATs
(co-plonk/src/round1.rs:344-427: PlainDriver, deterministic blinders), and the same commitments from three REP3 parties."""
import json
import os
import random

import numpy as np
import pytest

from oracle import cref, formats, groth16
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bls12_381", "poseidon")])
def test_round1_kat_plain_and_rep3(cocg, curve, circ):
    d = os.path.join(G, "plonk", curve, circ)
    kat = json.load(open(os.path.join(G, "plonk_round1_kats.json")))[curve + "/" + circ]
    want = [(int(kat[k][0]), int(kat[k][1])) for k in ("commit_a", "commit_b", "commit_c")]
    zk = cocg.PlonkZKey(os.path.join(d, "circuit.round1.zkey"))
    ozk = formats.parse_plonk_zkey(open(os.path.join(d, "circuit.round1.zkey"), "rb").read())
    assert (zk.n_vars, zk.n_public, zk.domain_size, zk.n_additions, zk.n_constraints) == \
        (ozk.n_vars, ozk.n_public, ozk.domain_size, ozk.n_additions, ozk.n_constraints)
    c = ozk.curve
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    got = zk.round1_plain(pub, cref.fr_to_mont(c, wit))
    assert [cref.g_from_mont(c, g, 1)[0] for g in got] == want
    # three REP3 parties, trivial shares of the deterministic blinders: every party opens the same literal commitments
    rng = random.Random(3)
    shares = groth16.share_rep3(wit, rng, c.r)
    out = zk.round1_rep3(pub, [cref.fr_to_mont(c, s[0]) for s in shares], [cref.fr_to_mont(c, s[1]) for s in shares])
    for party in range(3):
        assert [cref.g_from_mont(c, g, 1)[0] for g in out[party]] == want
    # random blinders: still agree between parties, differ from the KAT
    out2 = zk.round1_rep3(pub, [cref.fr_to_mont(c, s[0]) for s in shares], [cref.fr_to_mont(c, s[1]) for s in shares], deterministic=False)
    assert np.array_equal(out2[0], out2[1]) and np.array_equal(out2[1], out2[2]) and not np.array_equal(out2[0], out[0])
    zk.close()


# ================================================================================================ rounds 2-5, whole proofs
def _full_fixture(curve, circ, tmp_path):
    """(path of the full zkey, oracle zkey, witness values).  The 6.3 MB poseidon key is stored xz-compressed."""
    import lzma
    d = os.path.join(G, "plonk", curve, circ)
    path = os.path.join(d, "circuit.zkey")
    if not os.path.exists(path):
        raw = lzma.open(path + ".xz").read()
        path = str(tmp_path / "circuit.zkey")
        with open(path, "wb") as f:
            f.write(raw)
    ozk = formats.parse_plonk_zkey(open(path, "rb").read())
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    return path, ozk, wt


def _proof_dict(c, zk, block):
    """proof block (9 packed affine points | 6 Montgomery Fr) -> the oracle's dict"""
    lq = zk.lq
    out = {}
    for i, k in enumerate(("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")):
        out[k] = cref.g_from_mont(c, block[i * 2 * lq:(i + 1) * 2 * lq], 1)[0]
    ev = cref.fr_from_mont(c, block[18 * lq:].reshape(6, 4))
    out.update(dict(zip(("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"), ev)))
    return out


def _rep3_shares(c, wit, seed):
    shares = groth16.share_rep3(wit, random.Random(seed), c.r)
    return [cref.fr_to_mont(c, s[0]) for s in shares], [cref.fr_to_mont(c, s[1]) for s in shares]


def test_plonk_full_proof_reproduces_reference_round_kats(cocg, tmp_path):
    """BN254 multiplier2 with the reference's deterministic blinders b_i = i: [a] [b] [c] (round1.rs:344-386), [z] (round2.rs:326-355),
    [t1] [t2] [t3] (round3.rs:553-596), the six evaluations (round4.rs:181-247), [Wxi] [Wxiw] (round5.rs:391-429) -- literally, from
    CoPlonk<PlainDriver> and from each of the three CoPlonk<Rep3Protocol> parties."""
    path, ozk, wt = _full_fixture("bn254", "multiplier2", tmp_path)
    c = ozk.curve
    k1 = json.load(open(os.path.join(G, "plonk_round1_kats.json")))["bn254/multiplier2"]
    kat = json.load(open(os.path.join(G, "plonk_round2_kats.json")))
    pt = lambda v: (int(v[0]), int(v[1]))
    want = {"A": pt(k1["commit_a"]), "B": pt(k1["commit_b"]), "C": pt(k1["commit_c"]), "Z": pt(kat["commit_z"]),
            "T1": pt(kat["commit_t"][0]), "T2": pt(kat["commit_t"][1]), "T3": pt(kat["commit_t"][2]),
            "Wxi": pt(kat["commit_w"][0]), "Wxiw": pt(kat["commit_w"][1])}
    want.update({k: int(v) for k, v in kat["evals"].items()})
    zk = cocg.PlonkZKey(path)
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    plain = cocg.PlonkSession(zk, "plain", seeds=bytes(32))
    got = _proof_dict(c, zk, plain.prove(pub, [cref.fr_to_mont(c, wit)], deterministic=True)[0])
    assert got == want
    plain.close()
    wa, wb = _rep3_shares(c, wit, 5)
    rep3 = cocg.PlonkSession(zk, "rep3", seeds=bytes(range(96)))
    for mode in ("host", "device"):
        rep3.set_mpc_exchange(mode)
        out = rep3.prove(pub, wa, wb, deterministic=True)
        for party in range(3):
            assert _proof_dict(c, zk, out[party]) == want, (mode, party)
    rep3.close()
    zk.close()


@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bls12_381", "multiplier2"), ("bn254", "poseidon")])
def test_plonk_random_blinders_verify(cocg, tmp_path, curve, circ):
    """co-plonk/src/lib.rs:203-275 (prove with PlainDriver, then Plonk::verify) and tests/tests/circom/e2e_tests (three REP3 parties
    output the same proof, which verifies): random blinders, the product's own verifier and the proof JSON writer; a changed public
    input is rejected."""
    path, ozk, wt = _full_fixture(curve, circ, tmp_path)
    c = ozk.curve
    d = os.path.join(G, "plonk", curve, circ)
    vk = open(os.path.join(d, "verification_key.json")).read()
    public = json.load(open(os.path.join(d, "public.json")))
    zk = cocg.PlonkZKey(path)
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    plain = cocg.PlonkSession(zk, "plain")
    p1 = plain.prove(pub, [cref.fr_to_mont(c, wit)])[0]
    p2 = plain.prove(pub, [cref.fr_to_mont(c, wit)])[0]
    assert not np.array_equal(p1, p2)                                  # fresh blinders per proof
    for p in (p1, p2):
        pj = cocg.plonk_proof_to_json(zk.curve, p)
        assert json.loads(pj)["protocol"] == "plonk"
        assert cocg.plonk_verify_json(vk, pj, json.dumps(public)) is True
    bad = [str((int(public[0]) + 1) % c.r)] + public[1:]
    assert cocg.plonk_verify_json(vk, cocg.plonk_proof_to_json(zk.curve, p1), json.dumps(bad)) is False
    plain.close()
    wa, wb = _rep3_shares(c, wit, 6)
    rep3 = cocg.PlonkSession(zk, "rep3")
    for _ in range(2):                                                  # the session is reusable
        out = rep3.prove(pub, wa, wb)
        assert np.array_equal(out[0], out[1]) and np.array_equal(out[1], out[2])
        assert cocg.plonk_verify_json(vk, cocg.plonk_proof_to_json(zk.curve, out[0]), json.dumps(public)) is True
    assert rep3.launch_count() > 0
    rep3.close()
    zk.close()


def _shamir_shares(c, wit, n, t, seed):
    """shamir/utils share_field_elements: party i holds p(i + 1), p random of degree t with p(0) = value."""
    rng = random.Random(seed)
    out = [[] for _ in range(n)]
    for v in wit:
        coeffs = [rng.randrange(c.r) for _ in range(t)]
        for p in range(n):
            out[p].append((v + sum(cf * pow(p + 1, k + 1, c.r) for k, cf in enumerate(coeffs))) % c.r)
    return [cref.fr_to_mont(c, o) for o in out]


@pytest.mark.parametrize("n,t", [(3, 1), (5, 2), (4, 1)])
def test_plonk_shamir_reproduces_reference_round_kats(cocg, tmp_path, n, t):
    """CoPlonk<ShamirProtocol> (co-plonk is generic over the MPC protocol, plonk.rs:50-77; driver mpc-core/src/protocols/shamir.rs):
    with the deterministic blinders every party of a (n, t) sharing outputs the reference's round KATs literally."""
    path, ozk, wt = _full_fixture("bn254", "multiplier2", tmp_path)
    c = ozk.curve
    k1 = json.load(open(os.path.join(G, "plonk_round1_kats.json")))["bn254/multiplier2"]
    kat = json.load(open(os.path.join(G, "plonk_round2_kats.json")))
    pt = lambda v: (int(v[0]), int(v[1]))
    want = {"A": pt(k1["commit_a"]), "B": pt(k1["commit_b"]), "C": pt(k1["commit_c"]), "Z": pt(kat["commit_z"]),
            "T1": pt(kat["commit_t"][0]), "T2": pt(kat["commit_t"][1]), "T3": pt(kat["commit_t"][2]),
            "Wxi": pt(kat["commit_w"][0]), "Wxiw": pt(kat["commit_w"][1])}
    want.update({k: int(v) for k, v in kat["evals"].items()})
    zk = cocg.PlonkZKey(path)
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    sh = cocg.PlonkSession(zk, "shamir", seeds=bytes(range(32 * n)), num_parties=n, threshold=t)
    for mode in ("host", "device"):
        sh.set_mpc_exchange(mode)
        out = sh.prove(pub, _shamir_shares(c, wit, n, t, 9), deterministic=True)
        for party in range(n):
            assert _proof_dict(c, zk, out[party]) == want, (mode, party)
    sh.close()
    zk.close()


@pytest.mark.parametrize("curve,circ,n,t", [("bn254", "poseidon", 3, 1), ("bls12_381", "multiplier2", 5, 2)])
def test_plonk_shamir_random_blinders_verify(cocg, tmp_path, curve, circ, n, t):
    """Random blinders (degree-t halves of the double-random pairs): all parties open the same proof, it verifies, a second proof differs."""
    path, ozk, wt = _full_fixture(curve, circ, tmp_path)
    c = ozk.curve
    d = os.path.join(G, "plonk", curve, circ)
    vk = open(os.path.join(d, "verification_key.json")).read()
    public = json.load(open(os.path.join(d, "public.json")))
    zk = cocg.PlonkZKey(path)
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    sh = cocg.PlonkSession(zk, "shamir", num_parties=n, threshold=t)
    shares = _shamir_shares(c, wit, n, t, 10)
    prev = None
    for _ in range(2):
        out = sh.prove(pub, shares)
        assert all(np.array_equal(out[0], out[i]) for i in range(1, n))
        assert cocg.plonk_verify_json(vk, cocg.plonk_proof_to_json(zk.curve, out[0]), json.dumps(public)) is True
        assert prev is None or not np.array_equal(prev, out[0])
        prev = out[0]
    sh.close()
    zk.close()


def test_plonk_poseidon_intermediates_match_oracle(cocg, tmp_path):
    """n = 4096, 2228 additions (multi-block scans, multi-pass NTTs): every intermediate vector of rounds 2-5 -- the rotated grand
    product, z(X), the quotient evaluations t / tz on the 4n domain, t1 t2 t3, r(X), W_xi -- equals the oracle's (which reproduces the
    reference's KATs), for the plain driver and, summed over the parties, for REP3; the proofs are equal as well."""
    from oracle import plonk
    path, ozk, wt = _full_fixture("bn254", "poseidon", tmp_path)
    c = ozk.curve
    trace = {}
    want = plonk.prove_plain(ozk, wt, trace=trace)
    zk = cocg.PlonkZKey(path)
    ell = zk.n_public
    pub = cref.fr_to_mont(c, wt[:ell + 1])
    wit = [v % c.r for v in wt[ell + 1:]]
    names = ("buffer_a", "poly_a", "buffer_z", "poly_z", "t_evals", "tz_evals", "t1", "t2", "t3", "poly_r", "wxi")
    plain = cocg.PlonkSession(zk, "plain", seeds=bytes(32))
    plain.trace(True)
    got = _proof_dict(c, zk, plain.prove(pub, [cref.fr_to_mont(c, wit)], deterministic=True)[0])
    for nm in names:
        assert cref.fr_from_mont(c, plain.trace_get(0, nm)) == [v % c.r for v in trace[nm]], nm
    assert got == want
    plain.close()
    wa, wb = _rep3_shares(c, wit, 8)
    rep3 = cocg.PlonkSession(zk, "rep3", seeds=bytes(range(96)))
    rep3.trace(True)
    out = rep3.prove(pub, wa, wb, deterministic=True)
    for nm in names:
        parts = [cref.fr_from_mont(c, rep3.trace_get(p, nm)) for p in range(3)]
        assert [(x + y + z) % c.r for x, y, z in zip(*parts)] == [v % c.r for v in trace[nm]], nm
    for party in range(3):
        assert _proof_dict(c, zk, out[party]) == want
    rep3.close()
    zk.close()


def test_plonk_synthetic_key_runs_all_rounds(cocg):
    """The shape-faithful benchmark key (bench.py --workload plonk) at a test size: REP3 parties agree with each other and -- with the
    deterministic blinders -- with the plain driver on the reconstructed witness."""
    c = BN254
    log_n = 12
    n = 1 << log_n
    rng = np.random.default_rng(11)
    n_public, n_vars = 2, n - 5
    maps = [rng.integers(0, n_vars, size=n - 3).astype(np.uint32) for _ in range(3)]
    zk = cocg.PlonkZKey.synthetic(cocg.BN254, log_n, n_public, n_vars, maps, bytes(range(32)))
    wit = [int(x) for x in rng.integers(1, 2**62, size=zk.n_witness)]
    pub = cref.fr_to_mont(c, [1, 7, 9])
    plain = cocg.PlonkSession(zk, "plain", seeds=bytes(32))
    want = plain.prove(pub, [cref.fr_to_mont(c, wit)], deterministic=True)[0]
    plain.close()
    wa, wb = _rep3_shares(c, wit, 9)
    rep3 = cocg.PlonkSession(zk, "rep3", seeds=bytes(range(96)))
    rep3.set_mpc_exchange("device")
    out = rep3.prove(pub, wa, wb, deterministic=True)
    for party in range(3):
        assert np.array_equal(out[party], want)
    rep3.close()
    zk.close()


# ================================================================================================ vector primitives (csrc/poly.cu)
@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 4097, (1 << 16) + 3])
def test_scan_inverse_eval_primitives(cocg, bn, bls, curve, n):
    ctx = bn if curve is BN254 else bls
    r = curve.r
    rng = random.Random(n)
    vals = [rng.randrange(1, r) for _ in range(n)]
    x = ctx.upload(cref.fr_to_mont(curve, vals))
    # prefix product / prefix sum (array_prod_mul round2.rs:33-35; div_by_zerofier round5.rs:97-115)
    acc, want_mul, s, want_add = 1, [], 0, []
    for v in vals:
        acc = acc * v % r
        s = (s + v) % r
        want_mul.append(acc)
        want_add.append(s)
    assert cref.fr_from_mont(curve, ctx.vec_scan(cocg.OP_MUL, x).to_host()) == want_mul
    assert cref.fr_from_mont(curve, ctx.vec_scan(cocg.OP_ADD, x).to_host()) == want_add
    # batched inversion with a zero in the batch (inv_many: "cannot compute inverse of zero", rep3.rs:549-554)
    with_zero = list(vals)
    with_zero[n // 2] = 0
    inv, zeros = ctx.vec_inv(ctx.upload(cref.fr_to_mont(curve, with_zero)))
    assert zeros == 1
    assert cref.fr_from_mont(curve, inv.to_host()) == [pow(v, -1, r) if v else 0 for v in with_zero]
    # evaluate_poly_public (rep3.rs:923-928)
    pt = rng.randrange(r)
    want = 0
    for v in reversed(vals):
        want = (want * pt + v) % r
    assert cref.fr_from_mont(curve, ctx.poly_eval(x, cref.fr_to_mont(curve, [pt])[0]).reshape(1, 4)) == [want]
    # scale by powers without a cached table: x[i] * c * g^i
    g, cc = rng.randrange(1, r), rng.randrange(1, r)
    y = ctx.upload(cref.fr_to_mont(curve, vals))
    ctx.scale_powers(y, cref.fr_to_mont(curve, [g]), cref.fr_to_mont(curve, [cc]))
    assert cref.fr_from_mont(curve, y.to_host()) == [v * cc % r * pow(g, i, r) % r for i, v in enumerate(vals)]


def test_gather_lincomb_fill(cocg, bn):
    c = BN254
    r = c.r
    rng = random.Random(2)
    src = [rng.randrange(r) for _ in range(500)]
    idx = np.array([rng.randrange(500) for _ in range(2000)] + [0xFFFFFFFF, 500, 499], dtype=np.uint32)
    d_idx = bn.upload_u32(idx)
    got = cref.fr_from_mont(c, bn.vec_gather(bn.upload(cref.fr_to_mont(c, src)), d_idx, len(idx)).to_host())
    assert got == [src[i] if i < 500 else 0 for i in idx]
    bn.free_raw(d_idx)
    vecs = [[rng.randrange(r) for _ in range(m)] for m in (100, 37, 64)]
    f = [rng.randrange(r) for _ in range(3)]
    out = bn.vec_lincomb([bn.upload(cref.fr_to_mont(c, v)) for v in vecs], cref.fr_to_mont(c, f), 110)
    want = [sum(f[k] * vecs[k][i] for k in range(3) if i < len(vecs[k])) % r for i in range(110)]
    assert cref.fr_from_mont(c, out.to_host()) == want
    assert cref.fr_from_mont(c, bn.vec_fill(33, cref.fr_to_mont(c, [12345])[0]).to_host()) == [12345] * 33
