"""GPU parity: every kernel behind the C ABI against the CPU oracle, bit-exact (integer path: no tolerance).

Reference behaviour being checked (paths under /root/reference):
  element-wise / mul_vec local   mpc-core/src/protocols/rep3.rs:581-688, tests mpc-core/tests/protocols/rep3.rs:242-350
  fft / ifft (+ coset scaling)   mpc-core/src/protocols/rep3.rs:880-921, co-groth16/src/groth16.rs:57-77, 175-200
  msm_public_points              mpc-core/src/protocols/rep3.rs:934-947 (msm_unchecked truncates to min(len))
  evaluate_constraint            mpc-core/src/protocols/rep3.rs:690-708
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import cref, ntt as ontt
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rand_fr(n, seed):
    """n uniformly random 252-bit values as (n,4) limbs: valid residues for both scalar fields."""
    a = np.random.default_rng(seed).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def edge_fr(curve):
    r = curve.r
    return cref.ints_to_limbs([0, 1, r - 1, r - 2, curve.Rr % r, (r - 1) // 2, 2**252, 2**253 % r], 4)


def ctx_for(curve, bn, bls):
    return bn if curve is BN254 else bls


# ------------------------------------------------------------------------------------------------ element-wise
@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
@pytest.mark.parametrize("n", [1, 31, 1000, (1 << 16) + 3])
def test_vec_ops(cocg, bn, bls, curve, n):
    ctx = ctx_for(curve, bn, bls)
    a, b = rand_fr(n, 1), rand_fr(n, 2)
    e = edge_fr(curve)
    k = min(n, e.shape[0])
    a[:k] = e[:k]
    b[:k] = e[:k][::-1]
    da, db = ctx.upload(a), ctx.upload(b)
    for op, oop in ((cocg.OP_MUL, cref.OP_MUL), (cocg.OP_ADD, cref.OP_ADD), (cocg.OP_SUB, cref.OP_SUB), (cocg.OP_NEG, cref.OP_NEG),
                    (cocg.OP_TO_MONT, cref.OP_TO_MONT), (cocg.OP_FROM_MONT, cref.OP_FROM_MONT)):
        got = ctx.vec_op(op, da, db).to_host()
        want = cref.fr_vec_op(curve, oop, a, b)
        assert np.array_equal(got, want), f"op {op}"
    # in place (out aliases a), as sub_assign_vec does
    ctx.vec_op(cocg.OP_SUB, da, db, out=da)
    assert np.array_equal(da.to_host(), cref.fr_vec_op(curve, cref.OP_SUB, a, b))


def test_vec_empty(cocg, bn):
    v = bn.zeros(0)
    bn.vec_op(cocg.OP_ADD, v, v, out=v)
    bn.sync()


@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
def test_rep3_mul_local_and_scale_powers(cocg, bn, bls, curve):
    ctx = ctx_for(curve, bn, bls)
    n = 5000
    v = [rand_fr(n, s) for s in range(10, 15)]
    d = [ctx.upload(x) for x in v]
    got = ctx.rep3_mul_local(d[0], d[1], d[2], d[3], d[4]).to_host()
    assert np.array_equal(got, cref.rep3_mul_local(curve, *v))
    got = ctx.rep3_mul_local(d[0], d[1], d[2], d[3]).to_host()
    assert np.array_equal(got, cref.rep3_mul_local(curve, v[0], v[1], v[2], v[3], None))
    _, g = ontt.groth16_roots(curve, 13)
    gm, cm = cref.fr_to_mont(curve, [g]), cref.fr_to_mont(curve, [12345])
    ctx.scale_powers(d[0], gm, cm)
    assert np.array_equal(d[0].to_host(), cref.distribute_powers(curve, v[0], gm, cm))


def test_rep3_mul_vec_bn_reference_kat(cocg, bn):
    """The literal vectors of mpc-core/tests/protocols/rep3.rs:242-350, run through the GPU local step."""
    kat = json.load(open(os.path.join(G, "rep3_mul_vec_bn.json")))
    c = BN254
    x, y, xy = ([int(t) for t in kat[k]] for k in ("x", "y", "xy"))
    rng = random.Random(3)

    def share(vals):
        s0 = [rng.randrange(c.r) for _ in vals]
        s1 = [rng.randrange(c.r) for _ in vals]
        s2 = [(v - p - q) % c.r for v, p, q in zip(vals, s0, s1)]
        return [(s0, s2), (s1, s0), (s2, s1)]

    xs, ys = share(x), share(y)
    keys = [[rng.randrange(c.r) for _ in range(4)] for _ in range(3)]
    total = [0] * 4
    for i in range(3):
        mask = [(keys[i][j] - keys[(i - 1) % 3][j]) % c.r for j in range(4)]
        dv = [bn.upload(cref.fr_to_mont(c, t)) for t in (xs[i][0], xs[i][1], ys[i][0], ys[i][1], mask)]
        loc = cref.fr_from_mont(c, bn.rep3_mul_local(*dv).to_host())
        total = [(t + l) % c.r for t, l in zip(total, loc)]
    assert total == xy


# ------------------------------------------------------------------------------------------------ NTT
@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 9, 10, 12, 13, 16, 17])
def test_ntt_matches_oracle(cocg, bn, bls, curve, logn):
    ctx = ctx_for(curve, bn, bls)
    n = 1 << logn
    omega, g = ontt.groth16_roots(curve, logn)
    om = cref.fr_to_mont(curve, [omega])
    omi = cref.fr_to_mont(curve, [pow(omega, -1, curve.r)])
    gm = cref.fr_to_mont(curve, [g])
    one = cref.fr_to_mont(curve, [1])
    a, b = rand_fr(n, 100 + logn), rand_fr(n, 200 + logn)
    # forward, two share components at once
    da, db = ctx.upload(a), ctx.upload(b)
    ctx.ntt([da, db], logn, om)
    fa = cref.ntt(curve, a, om) if n > 1 else a
    fb = cref.ntt(curve, b, om) if n > 1 else b
    assert np.array_equal(da.to_host(), fa) and np.array_equal(db.to_host(), fb)
    # inverse brings it back
    ctx.ntt([da, db], logn, om, inverse=True)
    assert np.array_equal(da.to_host(), a) and np.array_equal(db.to_host(), b)
    # inverse + fused coset scaling == ifft; distribute_powers (groth16.rs:175-186)
    dc = ctx.upload(a)
    ctx.ntt([dc], logn, om, inverse=True, coset_g=gm)
    want = cref.ntt(curve, a, omi, inverse=True) if n > 1 else a
    want = cref.distribute_powers(curve, want, gm, one)
    assert np.array_equal(dc.to_host(), want)
    # forward with pre-scaling
    dd = ctx.upload(a)
    ctx.ntt([dd], logn, om, coset_g=gm)
    pre = cref.distribute_powers(curve, a, gm, one)
    assert np.array_equal(dd.to_host(), cref.ntt(curve, pre, om) if n > 1 else pre)


def test_ntt_many_vectors_on_a_priority_stream(cocg):
    """The witness map sends a, b, c (two share components each: 6 vectors) through a transform as one launch sequence, and in the
    multi-GPU block mode it does so on a stream re-created with the highest priority (cocg_set_stream_priority): 9 vectors (more than
    one batch of 8) at 2^13 (two passes) and 2^9, forward, inverse with the fused coset scaling, against the C oracle."""
    c = BN254
    ctx = cocg.Context(cocg.BN254, 0)
    ctx.set_stream_priority(True)
    for logn in (13, 9):
        n = 1 << logn
        omega, g = ontt.groth16_roots(c, logn)
        om, omi, gm, one = (cref.fr_to_mont(c, [v]) for v in (omega, pow(omega, -1, c.r), g, 1))
        hosts = [rand_fr(n, 300 + i) for i in range(9)]
        dev = [ctx.upload(h) for h in hosts]
        ctx.ntt(dev, logn, om)
        for h, d in zip(hosts, dev):
            assert np.array_equal(d.to_host(), cref.ntt(c, h, om))
        dev = [ctx.upload(h) for h in hosts]
        ctx.ntt(dev, logn, om, inverse=True, coset_g=gm)
        for h, d in zip(hosts, dev):
            assert np.array_equal(d.to_host(), cref.distribute_powers(c, cref.ntt(c, h, omi, inverse=True), gm, one))
    ctx.set_stream_priority(False)
    d = ctx.upload(hosts[0])
    ctx.ntt([d], 9, om)
    assert np.array_equal(d.to_host(), cref.ntt(c, hosts[0], om))
    ctx.close()


def test_ntt_small_against_naive_dft(cocg, bn):
    c = BN254
    for logn in (1, 2, 4):
        n = 1 << logn
        omega, _ = ontt.groth16_roots(c, logn)
        vals = [random.Random(logn).randrange(c.r) for _ in range(n)]
        d = bn.upload(cref.fr_to_mont(c, vals))
        bn.ntt([d], logn, cref.fr_to_mont(c, [omega]))
        assert cref.fr_from_mont(c, d.to_host()) == ontt.dft_naive(vals, omega, c.r)


def test_ntt_linearity_and_roundtrip_full_size(cocg, bn):
    """Size-independent properties at the benchmark size 2^20: NTT(a)+NTT(b) == NTT(a+b); iNTT(NTT(a)) == a."""
    c = BN254
    logn = 20
    n = 1 << logn
    omega, _ = ontt.groth16_roots(c, logn)
    om = cref.fr_to_mont(c, [omega])
    a, b = rand_fr(n, 7), rand_fr(n, 8)
    da, db = bn.upload(a), bn.upload(b)
    ds = bn.vec_op(cocg.OP_ADD, da, db)
    bn.ntt([da, db, ds], logn, om)
    lhs = bn.vec_op(cocg.OP_ADD, da, db).to_host()
    assert np.array_equal(lhs, ds.to_host())
    bn.ntt([da], logn, om, inverse=True)
    assert np.array_equal(da.to_host(), a)
    # spot-check one output against the definition: X[1] = sum_j a[j] w^j, via the C oracle's transform of the same data
    assert np.array_equal(ds.to_host()[:4], cref.ntt(c, cref.fr_vec_op(c, cref.OP_ADD, a, b), om)[:4])


# ------------------------------------------------------------------------------------------------ MSM
def make_bases(curve, group, n, seed=1):
    rng = random.Random(seed)
    p0 = cref.g_to_mont(curve, [curve.mul(curve.gen(group), rng.randrange(1, curve.r), group)], group)
    q = cref.g_to_mont(curve, [curve.mul(curve.gen(group), rng.randrange(1, curve.r), group)], group)
    return cref.gen_chain(curve, group, p0[0], q[0], n)


def same_point(curve, group, jac_a, jac_b):
    return cref.jac_from_mont(curve, jac_a, group) == cref.jac_from_mont(curve, jac_b, group)


@pytest.mark.parametrize("curve,group,n", [
    (BN254, 1, 1), (BN254, 1, 2), (BN254, 1, 31), (BN254, 1, 33), (BN254, 1, 1000), (BN254, 1, 1 << 14), (BN254, 1, (1 << 16) + 5),
    (BN254, 2, 1), (BN254, 2, 100), (BN254, 2, 1 << 12),
    (BLS12_381, 1, 3), (BLS12_381, 1, 1 << 12), (BLS12_381, 2, 2), (BLS12_381, 2, 1 << 10),
], ids=lambda v: getattr(v, "name", str(v)))
def test_msm_matches_oracle(cocg, bn, bls, curve, group, n):
    ctx = ctx_for(curve, bn, bls)
    pts = make_bases(curve, group, n)
    h = ctx.bases_upload(group, pts)
    sa, sb = rand_fr(n, 31), rand_fr(n, 32)
    e = edge_fr(curve)
    k = min(n, e.shape[0])
    sa[:k] = e[:k]
    out = ctx.msm(h, [ctx.upload(sa), ctx.upload(sb)])
    assert same_point(curve, group, out[0], cref.msm(curve, group, pts, sa))
    assert same_point(curve, group, out[1], cref.msm(curve, group, pts, sb))
    # host-pointer entry point and canonical (non-Montgomery) scalars
    can = cref.fr_vec_op(curve, cref.OP_FROM_MONT, sa)
    out2 = ctx.msm_host(h, [can], mont=False)
    assert same_point(curve, group, out2[0], out[0])
    # sub-range: off / n, as calculate_coeff slices query[1+l..] (groth16.rs:221-225)
    if n > 8:
        out3 = ctx.msm(h, [ctx.upload(sa[: n - 5])], off=3, n=n - 5)
        assert same_point(curve, group, out3[0], cref.msm(curve, group, pts[3:n - 2], sa[: n - 5]))
    ctx.bases_free(h)


def test_msm_empty_and_degenerate(cocg, bn):
    c = BN254
    pts = make_bases(c, 1, 64)
    h = bn.bases_upload(1, pts)
    out = bn.msm(h, [bn.zeros(0)], n=0)
    assert cref.jac_from_mont(c, out[0], 1) is None
    # all-zero scalars -> infinity
    out = bn.msm(h, [bn.zeros(64)])
    assert cref.jac_from_mont(c, out[0], 1) is None
    # bases containing infinity, repeated points with equal scalars (exercises the doubling branch), P + (-P)
    pts2 = pts.copy()
    pts2[5] = 0
    pts2[7] = pts2[6]
    pts2[9] = pts2[8]
    sc = rand_fr(64, 3)
    sc[7] = sc[6]
    sc[9] = cref.fr_vec_op(c, cref.OP_NEG, sc[8:9])[0]
    h2 = bn.bases_upload(1, pts2)
    out = bn.msm(h2, [bn.upload(sc)])
    assert same_point(c, 1, out[0], cref.msm(c, 1, pts2, sc))


def test_msm_skewed_scalars_heavy_buckets(cocg, bn):
    """Plain-driver witnesses are mostly 0/1/small: one bucket receives thousands of points (warp-per-bucket path)."""
    c = BN254
    n = 20000
    pts = make_bases(c, 1, n, seed=9)
    vals = [(i % 3) for i in range(n)]
    vals[17] = c.r - 1
    sc = cref.fr_to_mont(c, vals)
    h = bn.bases_upload(1, pts)
    out = bn.msm(h, [bn.upload(sc)])
    assert same_point(c, 1, out[0], cref.msm(c, 1, pts, sc))


def test_msm_linearity_full_size(cocg, bn):
    """2^20 terms (BASELINE config 2): MSM(a) + MSM(b) == MSM(a + b), and agreement with the C oracle."""
    c = BN254
    n = 1 << 20
    pts = make_bases(c, 1, n, seed=4)
    h = bn.bases_upload(1, pts)
    a, b = rand_fr(n, 41), rand_fr(n, 42)
    da, db = bn.upload(a), bn.upload(b)
    ds = bn.vec_op(cocg.OP_ADD, da, db)
    out = bn.msm(h, [da, db, ds])
    lhs = bn.ec_op(1, cocg.EC_ADD, out[0], out[1])
    assert same_point(c, 1, lhs, out[2])
    assert same_point(c, 1, out[0], cref.msm(c, 1, pts, a))


def test_bases_upload_strided_and_canonical(cocg, bn):
    """arkworks' Affine<..> is 72 bytes (x, y, infinity flag + padding); canonical coordinates are converted on upload."""
    c = BN254
    n = 50
    pts = make_bases(c, 1, n)
    sc = rand_fr(n, 5)
    want = cref.msm(c, 1, pts, sc)
    ark = np.zeros((n, 9), dtype=np.uint64)
    ark[:, :8] = pts
    h = bn.bases_upload(1, ark, stride=72)
    assert same_point(c, 1, bn.msm(h, [bn.upload(sc)])[0], want)
    qi = pow(c.Rq, -1, c.q)
    canon = cref.ints_to_limbs([(v * qi) % c.q for v in cref.limbs_to_ints(pts.reshape(-1, 4))], 4).reshape(n, 8)
    h2 = bn.bases_upload(1, canon, mont=False)
    assert same_point(c, 1, bn.msm(h2, [bn.upload(sc)])[0], want)


# ------------------------------------------------------------------------------------------------ SpMV
@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
def test_spmv_matches_oracle(cocg, bn, bls, curve):
    ctx = ctx_for(curve, bn, bls)
    rng = np.random.default_rng(6)
    rows, npub, nwit = 3000, 3, 2500
    nnz_per = rng.integers(0, 6, size=rows)
    rowptr = np.zeros(rows + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum(nnz_per)
    nnz = int(rowptr[-1])
    col = rng.integers(0, npub + nwit, size=nnz).astype(np.uint32)
    coeff = rand_fr(nnz, 60)
    zp, zw = rand_fr(npub, 61), rand_fr(nwit, 62)
    hdl = ctx.csr_upload(rowptr, col, coeff)
    got = ctx.spmv(hdl, ctx.upload(zp), npub, ctx.upload(zw)).to_host()
    assert np.array_equal(got, cref.spmv(curve, rowptr, col, coeff, np.concatenate([zp, zw])))
    # component that must not see the public inputs (rep3.rs:600-608)
    got = ctx.spmv(hdl, None, npub, ctx.upload(zw)).to_host()
    assert np.array_equal(got, cref.spmv(curve, rowptr, col, coeff, np.concatenate([np.zeros_like(zp), zw])))


# ------------------------------------------------------------------------------------------------ single-point ops
@pytest.mark.parametrize("curve,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)], ids=lambda v: getattr(v, "name", str(v)))
def test_ec_ops_match_oracle(cocg, bn, bls, curve, group):
    ctx = ctx_for(curve, bn, bls)
    rng = random.Random(8)
    P = cref.g_to_mont(curve, [curve.mul(curve.gen(group), rng.randrange(1, curve.r), group)], group)[0]
    Q = cref.g_to_mont(curve, [curve.mul(curve.gen(group), rng.randrange(1, curve.r), group)], group)[0]
    jp, jq = ctx.ec_op(group, cocg.EC_FROM_AFFINE, P), ctx.ec_op(group, cocg.EC_FROM_AFFINE, Q)
    assert np.array_equal(jp, cref.ec_op(curve, group, 3, P))
    k = cref.ints_to_limbs([rng.randrange(curve.r)], 4)[0]
    for op, args in ((cocg.EC_ADD, (jp, jq)), (cocg.EC_MUL, (jp, k)), (cocg.EC_NEG, (jp, None)), (cocg.EC_DBL, (jp, None)), (cocg.EC_ADD, (jp, jp))):
        got = ctx.ec_op(group, op, *args)
        want = cref.ec_op(curve, group, op, *args)
        assert same_point(curve, group, got, want), op
    s = ctx.ec_op(group, cocg.EC_MUL, jp, k)
    assert np.array_equal(ctx.ec_op(group, cocg.EC_TO_AFFINE, s), cref.ec_op(curve, group, 2, cref.ec_op(curve, group, 1, jp, k)))
    inf = ctx.ec_op(group, cocg.EC_ADD, jp, ctx.ec_op(group, cocg.EC_NEG, jp))
    assert cref.jac_from_mont(curve, inf, group) is None


# ------------------------------------------------------------------------------------------------ reference KAT on the GPU
@pytest.mark.parametrize("curve,circ", [("bn254", "multiplier2"), ("bls12_381", "poseidon")])
def test_plonk_round1_kat_on_gpu(cocg, bn, bls, curve, circ):
    """co-plonk/src/round1.rs:344-427 (bit-exact commitments with blinders b_i = i): wire buffers -> cocg_ntt (inverse, snarkjs
    root) -> blinding -> cocg_msm over p_tau must give the reference's literal points."""
    from oracle import formats, plonk
    d = os.path.join(G, "plonk", curve, circ)
    zk = formats.parse_plonk_zkey(open(os.path.join(d, "circuit.round1.zkey"), "rb").read())
    _, wt = formats.parse_wtns(open(os.path.join(d, "witness.wtns"), "rb").read())
    kat = json.load(open(os.path.join(G, "plonk_round1_kats.json")))[curve + "/" + circ]
    c = zk.curve
    ctx = ctx_for(c, bn, bls)
    n = zk.domain_size
    get = plonk.witness_with_additions(zk, wt)
    _, roots = ontt.roots_of_unity(c)
    om = cref.fr_to_mont(c, [roots[zk.pow]])
    h = ctx.bases_upload(1, cref.g_to_mont(c, zk.p_tau[:n + 2], 1))
    for k, (m, name) in enumerate(((zk.map_a, "commit_a"), (zk.map_b, "commit_b"), (zk.map_c, "commit_c"))):
        buf = cref.fr_to_mont(c, [get(i) for i in m] + [0] * (n - zk.n_constraints))
        dv = ctx.upload(buf)
        ctx.ntt([dv], zk.pow, om, inverse=True)
        coeffs = cref.fr_from_mont(c, dv.to_host())
        b_lo, b_hi = 2 * k, 2 * k + 1                       # Round1Challenges::deterministic, blind_coefficients
        coeffs[0] = (coeffs[0] - b_hi) % c.r
        coeffs[1] = (coeffs[1] - b_lo) % c.r
        poly = cref.fr_to_mont(c, coeffs + [b_hi, b_lo])
        out = ctx.msm(h, [ctx.upload(poly)], n=n + 2)
        assert cref.jac_from_mont(c, out[0], 1) == (int(kat[name][0]), int(kat[name][1])), name
    ctx.bases_free(h)


def test_msm_multi_shares_one_sort(cocg, bn):
    """cocg_msm_multi: the Groth16 shape -- l_query (n_aux points), a/b_g1 (m points, offset 1 + l), b_g2 (G2) times the same
    two share components -- equals separate MSMs / the oracle."""
    c = BN254
    m, ell = 3000, 1
    n_aux = m - ell - 1
    g1a, g1b, g1l = make_bases(c, 1, m, seed=21), make_bases(c, 1, m, seed=22), make_bases(c, 1, n_aux, seed=23)
    g2b = make_bases(c, 2, m, seed=24)
    hs = [bn.bases_upload(1, g1l), bn.bases_upload(1, g1a), bn.bases_upload(1, g1b), bn.bases_upload(2, g2b)]
    sa, sb = rand_fr(n_aux, 71), rand_fr(n_aux, 72)
    outs = bn.msm_multi(hs, [0, 1 + ell, 1 + ell, 1 + ell], [bn.upload(sa), bn.upload(sb)], n=n_aux)
    for out, (pts, group, off) in zip(outs, ((g1l, 1, 0), (g1a, 1, 1 + ell), (g1b, 1, 1 + ell), (g2b, 2, 1 + ell))):
        for j, s in enumerate((sa, sb)):
            assert same_point(c, group, out[j], cref.msm(c, group, pts[off:off + n_aux], s))
    for h in hs:
        bn.bases_free(h)


def test_msm_multi_batched_reduction_flushes(cocg, bn):
    """The bucket reductions of one cocg_msm_multi call are batched, 8 bucket sets per launch (msm_impl.cuh kMaxSets): 5 G1
    queries x 3 components = 15 sets (one flush in the middle, sets of later components landing in recycled arenas), a G2 query
    beside them, and a query of a different size (other window width, separate sort group) -- every result vs the oracle."""
    c = BN254
    n = 700
    g1 = [make_bases(c, 1, n + 3, seed=40 + i) for i in range(5)]
    small = make_bases(c, 1, 40, seed=46)       # window width differs from the n = 703 tables
    g2 = make_bases(c, 2, n + 3, seed=47)
    hs = [bn.bases_upload(1, p) for p in g1[:3]] + [bn.bases_upload(1, small), bn.bases_upload(2, g2)] + [bn.bases_upload(1, p) for p in g1[3:]]
    offs = [0, 1, 2, 0, 3, 3, 0]
    sc = [rand_fr(40, 90 + j) for j in range(3)]
    outs = bn.msm_multi(hs, offs, [bn.upload(s) for s in sc], n=40)   # n bounded by the small query
    cases = [(g1[0], 1), (g1[1], 1), (g1[2], 1), (small, 1), (g2, 2), (g1[3], 1), (g1[4], 1)]
    for out, (pts, group), off in zip(outs, cases, offs):
        for j, s in enumerate(sc):
            assert same_point(c, group, out[j], cref.msm(c, group, pts[off:off + 40], s))
    bn.bases_free(hs[3])
    hs6 = hs[:3] + hs[4:]
    offs6 = [0, 1, 2, 3, 3, 0]
    sc = [rand_fr(n, 95 + j) for j in range(3)]
    outs = bn.msm_multi(hs6, offs6, [bn.upload(s) for s in sc], n=n)
    cases = [(g1[0], 1), (g1[1], 1), (g1[2], 1), (g2, 2), (g1[3], 1), (g1[4], 1)]
    for out, (pts, group), off in zip(outs, cases, offs6):
        for j, s in enumerate(sc):
            assert same_point(c, group, out[j], cref.msm(c, group, pts[off:off + n], s))
    for h in hs6:
        bn.bases_free(h)
    # tiny window plans (tables of <= 200 bases: c <= 8, fewer partial marginals than one block of the weighting kernel) with
    # adjacent bucket sets, repeated: a set's reduction must not touch its neighbour's scratch
    for nb_pts in (9, 40, 200):
        small_cases = [(make_bases(c, 1, nb_pts, seed=60 + i), 1) for i in range(3)] + [(make_bases(c, 2, nb_pts, seed=63), 2)]
        hsm = [bn.bases_upload(g, p) for p, g in small_cases]
        sc = [rand_fr(nb_pts, 120 + j) for j in range(4)]
        want = [[cref.msm(c, group, pts, s) for s in sc] for pts, group in small_cases]
        dsc = [bn.upload(s) for s in sc]
        for _ in range(5):
            outs = bn.msm_multi(hsm, [0, 0, 0, 0], dsc, n=nb_pts)
            for out, w, (pts, group) in zip(outs, want, small_cases):
                for k in range(4):
                    assert same_point(c, group, out[k], w[k])
        for h in hsm:
            bn.bases_free(h)


def test_bls12_381_full_size_properties(cocg, bls):
    """BLS12-381 at the size of BASELINE config 5's shards (2^20 terms per GPU at 2^22 over 4+ GPUs): MSM linearity on generated
    bases, a prefix against the oracle, and an NTT round trip at 2^22."""
    c = BLS12_381
    n = 1 << 20
    h = bls.bases_generate(1, n, bytes([7] * 32))
    a, b = rand_fr(n, 81), rand_fr(n, 82)
    da, db = bls.upload(a), bls.upload(b)
    ds = bls.vec_op(cocg.OP_ADD, da, db)
    out = bls.msm(h, [da, db, ds])
    assert same_point(c, 1, bls.ec_op(1, cocg.EC_ADD, out[0], out[1]), out[2])
    m = 5000
    pts = bls.bases_download(h, 0, m)
    assert same_point(c, 1, bls.msm(h, [bls.upload(a[:m])], n=m)[0], cref.msm(c, 1, pts, a[:m]))
    for v in (da, db, ds):
        v.free()
    bls.bases_free(h)
    logn = 22
    omega, _ = ontt.groth16_roots(c, logn)
    om = cref.fr_to_mont(c, [omega])
    x = rand_fr(1 << logn, 83)
    dx = bls.upload(x)
    bls.ntt([dx], logn, om)
    bls.ntt([dx], logn, om, inverse=True)
    assert np.array_equal(dx.to_host(), x)
    dx.free()
