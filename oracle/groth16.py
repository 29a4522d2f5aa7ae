"""Groth16 prover/verifier restated on Python ints (oracle, test infrastructure only).

Follows /root/reference/co-circom/co-groth16/src/groth16.rs:113-326 (prove,
witness_map_from_matrices, calculate_coeff, create_proof_with_assignment) over
 * the plain driver  (/root/reference/mpc-core/src/protocols/plain.rs:111-416) and
 * the REP3 driver   (/root/reference/mpc-core/src/protocols/rep3.rs:581-947),
with every random value (r, s, zero-masks) INJECTED, because the reference draws them from
entropy (plain.rs:202-205, rep3.rs:595-598) and therefore pins no proof bytes.

REP3 is simulated in lock-step: three `party` states, each computing exactly what
`Rep3Protocol` computes locally, with the send_next/recv_prev exchange done by the
orchestrator (party i receives from party (i-1) mod 3, rep3.rs:661-662).
"""
from __future__ import annotations

from .curves import Curve
from .formats import Groth16ZKey, VerifyingKey
from .ntt import groth16_roots, ntt, intt, distribute_powers
from .pairing import engine


# ----------------------------------------------------------------------------
# verifier (co-groth16/src/verifier.rs:23-43 -> ark_groth16::verify_proof)
# ----------------------------------------------------------------------------
def verify(vk: VerifyingKey, A, B, C, public_inputs) -> bool:
    c = vk.curve
    if len(public_inputs) + 1 != len(vk.ic):
        return False
    acc = c.to_jac(vk.ic[0], 1)
    for x, P in zip(public_inputs, vk.ic[1:]):
        acc = c.jac_add(acc, c.jac_mul(c.to_jac(P, 1), x % c.r, 1), 1)
    vkx = c.to_affine(acc, 1)
    for P, g in ((A, 1), (C, 1), (B, 2)):
        if not c.is_on_curve(P, g):
            return False
    e = engine(c)
    return e.product_is_one([(c.neg(A, 1), B), (vk.alpha_g1, vk.beta_g2), (vkx, vk.gamma_g2), (C, vk.delta_g2)])


# ----------------------------------------------------------------------------
# plain driver prove
# ----------------------------------------------------------------------------
def _eval_rows(rows, z, r):
    return [sum(cf * z[idx] for cf, idx in row) % r for row in rows]


def witness_map_plain(zk: Groth16ZKey, public_inputs, witness):
    """groth16.rs:141-204 with PlainDriver.  Returns h (length domain_size)."""
    c = zk.curve
    r = c.r
    n = zk.domain_size
    assert (zk.num_constraints + zk.num_inputs - 1).bit_length() == zk.pow or n == 1
    omega, g = groth16_roots(c, zk.pow)
    z = list(public_inputs) + list(witness)
    a = _eval_rows(zk.a_rows, z, r) + [0] * (n - zk.num_constraints)
    b = _eval_rows(zk.b_rows, z, r) + [0] * (n - zk.num_constraints)
    for i in range(zk.num_inputs):                        # groth16.rs:169-171
        a[zk.num_constraints + i] = public_inputs[i] % r
    cc = [(x * y) % r for x, y in zip(a, b)]

    def coset(v):
        return ntt(distribute_powers(intt(v, omega, r), g, 1, r), omega, r)

    a = coset(a)
    b = coset(b)
    ab = [(x * y) % r for x, y in zip(a, b)]
    cc = coset(cc)
    return [(x - y) % r for x, y in zip(ab, cc)]


def prove_plain(zk: Groth16ZKey, wtns_values, r_rand: int, s_rand: int):
    """groth16.rs:113-139 + 237-326 with PlainDriver; returns affine (A, B, C)."""
    c = zk.curve
    ell = zk.n_public
    public_inputs = [v % c.r for v in wtns_values[:ell + 1]]
    witness = [v % c.r for v in wtns_values[ell + 1:]]
    h = witness_map_plain(zk, public_inputs, witness)
    inp = public_inputs[1:]
    aux = witness
    J = lambda P, g=1: c.to_jac(P, g)
    add = c.jac_add

    h_acc = c.msm(zk.h_query, h, 1)
    l_acc = c.msm(zk.l_query, aux, 1)
    delta1 = J(zk.delta_g1)
    rs = (r_rand * s_rand) % c.r
    rs_delta = c.jac_mul(delta1, rs, 1)

    def coeff(initial, query, vk_param, g):
        pub = c.msm(query[1:1 + ell], inp, g)
        priv = c.msm(query[1 + ell:], aux, g)
        res = add(initial, J(query[0], g), g)
        res = add(res, J(vk_param, g), g)
        res = add(res, pub, g)
        return add(res, priv, g)

    g_a = coeff(c.jac_mul(delta1, r_rand, 1), zk.a_query, zk.alpha_g1, 1)
    s_g_a = c.jac_mul(g_a, s_rand, 1)
    g1_b = coeff(c.jac_mul(delta1, s_rand, 1), zk.b_g1_query, zk.beta_g1, 1)
    r_g1_b = c.jac_mul(g1_b, r_rand, 1)
    g2_b = coeff(c.jac_mul(J(zk.delta_g2, 2), s_rand, 2), zk.b_g2_query, zk.beta_g2, 2)
    g_c = add(s_g_a, r_g1_b, 1)
    g_c = add(g_c, c.jac_neg(rs_delta, 1), 1)
    g_c = add(g_c, l_acc, 1)
    g_c = add(g_c, h_acc, 1)
    return c.to_affine(g_a, 1), c.to_affine(g2_b, 2), c.to_affine(g_c, 1)


# ----------------------------------------------------------------------------
# REP3 (3 parties, lock-step)
# ----------------------------------------------------------------------------
def share_rep3(values, rng, r):
    """rep3.rs:124-150: party i holds (x_i, x_{i-1}); returns [(a_vec, b_vec)] * 3."""
    x0 = [rng.randrange(r) for _ in values]
    x1 = [rng.randrange(r) for _ in values]
    x2 = [(v - p - q) % r for v, p, q in zip(values, x0, x1)]
    return [(x0, x2), (x1, x0), (x2, x1)]


def rep3_zero_masks(n, rng, r):
    """Three vectors summing to zero element-wise: m_i = F(k_i) - F(k_{i-1}) (rngs.rs:37-46)."""
    k = [[rng.randrange(r) for _ in range(n)] for _ in range(3)]
    return [[(k[i][j] - k[(i - 1) % 3][j]) % r for j in range(n)] for i in range(3)]


def rep3_eval_rows(pid, rows, public_inputs, wit_a, wit_b, r):
    """evaluate_constraint, rep3.rs:690-708 (+ add_with_public :600-608)."""
    npub = len(public_inputs)
    out_a, out_b = [], []
    for row in rows:
        aa = bb = 0
        for cf, idx in row:
            if idx < npub:
                v = public_inputs[idx] * cf
                if pid == 0:
                    aa += v
                elif pid == 1:
                    bb += v
            else:
                aa += cf * wit_a[idx - npub]
                bb += cf * wit_b[idx - npub]
        out_a.append(aa % r)
        out_b.append(bb % r)
    return out_a, out_b


def rep3_mul_local(xa, xb, ya, yb, mask, r):
    """mul_vec local step, rep3.rs:656-660."""
    return [(p * s + p * t + q * s + m) % r for p, q, s, t, m in zip(xa, xb, ya, yb, mask)]


def rep3_witness_map(zk: Groth16ZKey, public_inputs, shares, masks1, masks2):
    """groth16.rs:141-204 with Rep3Protocol for all three parties; returns [(h_a, h_b)]*3.
    Also returns the trace of intermediate vectors (for kernel-level parity tests)."""
    c = zk.curve
    r = c.r
    n = zk.domain_size
    omega, g = groth16_roots(c, zk.pow)
    A, B = [], []
    for pid in range(3):
        wa, wb = shares[pid]
        a_a, a_b = rep3_eval_rows(pid, zk.a_rows, public_inputs, wa, wb, r)
        b_a, b_b = rep3_eval_rows(pid, zk.b_rows, public_inputs, wa, wb, r)
        pad = [0] * (n - zk.num_constraints)
        a_a, a_b, b_a, b_b = a_a + pad, a_b + pad, b_a + pad, b_b + pad
        for i in range(zk.num_inputs):                    # promote_to_trivial_shares + clone_from_slice
            a_a[zk.num_constraints + i] = public_inputs[i] % r if pid == 0 else 0
            a_b[zk.num_constraints + i] = public_inputs[i] % r if pid == 1 else 0
        A.append((a_a, a_b))
        B.append((b_a, b_b))

    def mul_vec(X, Y, masks):
        loc = [rep3_mul_local(X[i][0], X[i][1], Y[i][0], Y[i][1], masks[i], r) for i in range(3)]
        return [(loc[i], loc[(i - 1) % 3]) for i in range(3)]

    def coset(v):
        return ntt(distribute_powers(intt(v, omega, r), g, 1, r), omega, r)

    Cc = mul_vec(A, B, masks1)
    A = [(coset(a), coset(b)) for a, b in A]
    B = [(coset(a), coset(b)) for a, b in B]
    AB = mul_vec(A, B, masks2)
    Cc = [(coset(a), coset(b)) for a, b in Cc]
    H = []
    for i in range(3):
        H.append(([(x - y) % r for x, y in zip(AB[i][0], Cc[i][0])],
                  [(x - y) % r for x, y in zip(AB[i][1], Cc[i][1])]))
    return H


def prove_rep3(zk: Groth16ZKey, public_inputs, shares, rnd):
    """Full 3-party prove.  `rnd` carries the injected randomness:
       r, s          : [(a,b)]*3 replicated shares of r and s  (rand(), rep3.rs:595-598)
       masks1/masks2 : zero-masks for the two mul_vec rounds
       mask_rs       : 3 field elements summing to zero (mul r*s, rep3.rs:503-511)
       mask_pt       : 3 G1 Jacobian points summing to zero (scalar_mul, rep3.rs:835-847)
    Returns ([(A,B,C)]*3 affine, H shares)."""
    c = zk.curve
    r = c.r
    ell = zk.n_public
    inp = [v % r for v in public_inputs[1:]]
    H = rep3_witness_map(zk, public_inputs, shares, rnd["masks1"], rnd["masks2"])
    J = lambda P, g=1: c.to_jac(P, g)
    add = c.jac_add
    inf = lambda g: J(None, g)

    def msm2(points, sh, g=1):                    # msm_public_points, rep3.rs:934-947
        return [c.msm(points, sh[0], g), c.msm(points, sh[1], g)]

    def add_pub(pid, ps, P, g):                   # add_assign_points_public*, rep3.rs:788-810
        if pid == 0:
            ps[0] = add(ps[0], P, g)
        elif pid == 1:
            ps[1] = add(ps[1], P, g)

    def padd(x, y, g=1):
        return [add(x[0], y[0], g), add(x[1], y[1], g)]

    def smul_pub(P, sh, g=1):                     # scalar_mul_public_point, rep3.rs:820-825
        return [c.jac_mul(P, sh[0], g), c.jac_mul(P, sh[1], g)]

    delta1 = J(zk.delta_g1)
    # network round: rs = mul(r, s)
    rs_loc = [rep3_mul_local([rnd["r"][i][0]], [rnd["r"][i][1]], [rnd["s"][i][0]], [rnd["s"][i][1]],
                             [rnd["mask_rs"][i]], r)[0] for i in range(3)]
    rs = [(rs_loc[i], rs_loc[(i - 1) % 3]) for i in range(3)]

    def coeff(pid, initial, query, vk_param, aux, g):
        pub = c.msm(query[1:1 + ell], inp, g)
        priv = msm2(query[1 + ell:], aux, g)
        res = list(initial)
        add_pub(pid, res, J(query[0], g), g)
        add_pub(pid, res, J(vk_param, g), g)
        add_pub(pid, res, pub, g)
        return padd(res, priv, g)

    st = []
    for pid in range(3):
        aux = shares[pid]
        d = {}
        d["h_acc"] = msm2(zk.h_query, H[pid])
        d["l_acc"] = msm2(zk.l_query, aux)
        d["rs_delta"] = smul_pub(delta1, rs[pid])
        d["g_a"] = coeff(pid, smul_pub(delta1, rnd["r"][pid]), zk.a_query, zk.alpha_g1, aux, 1)
        st.append(d)
    # open_point(g_a): send b to next, recv c from prev, a+b+c  (rep3.rs:849-853)
    g_a_open = []
    for pid in range(3):
        cpt = st[(pid - 1) % 3]["g_a"][1]
        g_a_open.append(add(add(st[pid]["g_a"][0], st[pid]["g_a"][1], 1), cpt, 1))
    # scalar_mul(g1_b, r): local_a = b.a*a.a + b.a*a.b + b.b*a.a + mask  (pointshare Mul, rep3.rs:835-847)
    loc = []
    for pid in range(3):
        d = st[pid]
        aux = shares[pid]
        d["s_g_a"] = smul_pub(g_a_open[pid], rnd["s"][pid])
        d["g1_b"] = coeff(pid, smul_pub(delta1, rnd["s"][pid]), zk.b_g1_query, zk.beta_g1, aux, 1)
        ra, rb = rnd["r"][pid]
        t = c.jac_mul(d["g1_b"][0], (ra + rb) % r, 1)
        t = add(t, c.jac_mul(d["g1_b"][1], ra, 1), 1)
        loc.append(add(t, rnd["mask_pt"][pid], 1))
        d["g2_b"] = coeff(pid, smul_pub(J(zk.delta_g2, 2), rnd["s"][pid], 2), zk.b_g2_query, zk.beta_g2, aux, 2)
    proofs = []
    gcs = []
    for pid in range(3):
        d = st[pid]
        r_g1_b = [loc[pid], loc[(pid - 1) % 3]]
        g_c = padd(d["s_g_a"], r_g1_b)
        g_c = padd(g_c, [c.jac_neg(d["rs_delta"][0]), c.jac_neg(d["rs_delta"][1])])
        g_c = padd(g_c, d["l_acc"])
        g_c = padd(g_c, d["h_acc"])
        gcs.append(g_c)
    for pid in range(3):                           # open_two_points, rep3.rs:864-878
        prev = (pid - 1) % 3
        gc = add(add(gcs[pid][0], gcs[pid][1], 1), gcs[prev][1], 1)
        g2 = add(add(st[pid]["g2_b"][0], st[pid]["g2_b"][1], 2), st[prev]["g2_b"][1], 2)
        proofs.append((c.to_affine(g_a_open[pid], 1), c.to_affine(g2, 2), c.to_affine(gc, 1)))
    return proofs, H
