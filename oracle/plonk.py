"""Round 1 of the Plonk prover restated on Python ints (oracle, test infrastructure only).

Follows /root/reference/co-circom/co-plonk/src/round1.rs:121-300 with PlainDriver and the reference's own deterministic
blinding hook (`Round1Challenges::deterministic`, b_i = i, round1.rs:101-108) -- the configuration of its bit-exact KATs
(round1.rs:344-427), which pin iNTT (snarkjs roots, co-plonk/src/types.rs:59-99) + MSM results exactly.
"""
from __future__ import annotations

from .formats import PlonkZKey
from .keccak import keccak256
from .ntt import roots_of_unity, intt, ntt


def witness_with_additions(zk: PlonkZKey, values):
    """calculate_additions (round1.rs:213-242) + get_witness (co-plonk/src/lib.rs:113-137): signal index -> value."""
    r = zk.curve.r
    vals = [v % r for v in values]
    vals[0] = 0  # PlonkWitness::new writes zero over the leading one of the Groth16-style witness to mirror snarkjs (co-plonk/src/types.rs:105-108)
    base = zk.n_vars - zk.n_additions
    add = []

    def get(i):
        if i < base:
            return vals[i]
        if i < zk.n_vars:
            return add[i - base]
        raise ValueError("corrupted witness index %d" % i)

    for s1, s2, f1, f2 in zk.additions:
        add.append((f1 * get(s1) + f2 * get(s2)) % r)
    return get


def wire_buffers(zk: PlonkZKey, values):
    """buffer_a/b/c (round1.rs:121-166): the wire values per gate, zero-padded to the domain."""
    get = witness_with_additions(zk, values)
    return [[get(i) for i in m] + [0] * (zk.domain_size - zk.n_constraints) for m in (zk.map_a, zk.map_b, zk.map_c)]


def wire_polynomials(zk: PlonkZKey, values, blinders):
    """compute_wire_polynomials (round1.rs:121-209): coefficients of the blinded a(X), b(X), c(X), length n + 2 each."""
    r = zk.curve.r
    _, roots = roots_of_unity(zk.curve)
    omega = roots[zk.pow]
    out = []
    for k, buf in enumerate(wire_buffers(zk, values)):
        poly = intt(buf, omega, r)
        b_lo, b_hi = blinders[2 * k], blinders[2 * k + 1]      # blind_coefficients with coeff_rev = b[2k..2k+2] (lib.rs:140-158)
        poly[0] = (poly[0] - b_hi) % r
        poly[1] = (poly[1] - b_lo) % r
        out.append(poly + [b_hi % r, b_lo % r])
    return out


def round1_commitments(zk: PlonkZKey, values, blinders=tuple(range(11))):
    """[a]_1, [b]_1, [c]_1 (round1.rs:263-291), affine."""
    c = zk.curve
    polys = wire_polynomials(zk, values, blinders)
    return [c.to_affine(c.msm(zk.p_tau[:len(p)], p, 1), 1) for p in polys]


# ------------------------------------------------------------------------------------------------ transcript + round 2
class Keccak256Transcript:
    """co-plonk/src/types.rs:125-176: big-endian field bytes into Keccak-256; infinity = 2 x byte_len zero bytes."""

    def __init__(self, curve):
        self.c = curve
        self.buf = bytearray()
        self.qlen = (curve.q.bit_length() + 7) // 8
        self.rlen = (curve.r.bit_length() + 7) // 8

    def add_scalar(self, v):
        self.buf += int(v % self.c.r).to_bytes(self.rlen, "big")

    def add_point(self, P):
        if P is None:
            self.buf += bytes(2 * self.qlen)
        else:
            self.buf += int(P[0]).to_bytes(self.qlen, "big") + int(P[1]).to_bytes(self.qlen, "big")

    def get_challenge(self):
        return int.from_bytes(keccak256(bytes(self.buf)), "big") % self.c.r   # from_be_bytes_mod_order


def round2_challenges(zk: PlonkZKey, values, commits):
    """beta, gamma (round2.rs:251-276; the verifier derives them the same way, plonk.rs:52-76): vk points, public inputs, round-1 commitments."""
    c = zk.curve
    t = Keccak256Transcript(c)
    for name in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3"):
        t.add_point(zk.vk_points[name])
    for v in values[1:zk.n_public + 1]:                     # the leading zero is dropped after round 1 (round1.rs:42-46)
        t.add_scalar(v % c.r)
    for P in commits:
        t.add_point(P)
    beta = t.get_challenge()
    t = Keccak256Transcript(c)
    t.add_scalar(beta)
    return beta, t.get_challenge()


def z_polynomial(zk: PlonkZKey, values, beta, gamma, blinders=tuple(range(11)), trace=None):
    """compute_z (round2.rs:146-236) with the plain driver: grand product of the permutation argument over the domain, rotated by
    one, iFFT, blinded with b[6..9]: n + 3 coefficients."""
    c = zk.curve
    r, n = c.r, zk.domain_size
    a, b, cc = wire_buffers(zk, values)
    _, roots = roots_of_unity(c)
    omega = roots[zk.pow]
    num, den = [], []
    w = 1
    for i in range(n):
        betaw = beta * w % r
        n_i = (a[i] + betaw + gamma) * (b[i] + zk.k1 * betaw + gamma) % r * (cc[i] + zk.k2 * betaw + gamma) % r
        d_i = (a[i] + beta * zk.sigma[0][1][4 * i] + gamma) * (b[i] + beta * zk.sigma[1][1][4 * i] + gamma) % r \
            * (cc[i] + beta * zk.sigma[2][1][4 * i] + gamma) % r
        num.append(n_i)
        den.append(d_i)
        w = w * omega % r
    for i in range(1, n):                                  # array_prod_mul: prefix products
        num[i] = num[i] * num[i - 1] % r
        den[i] = den[i] * den[i - 1] % r
    buf = [x * pow(d, -1, r) % r for x, d in zip(num, den)]
    buf = buf[-1:] + buf[:-1]                              # rotate_right(1)
    if trace is not None:
        trace["buffer_z"] = list(buf)
    poly = intt(buf, omega, r)
    b6, b7, b8 = blinders[6:9]                             # blind_coefficients(poly, b[6..9]): coefficients given in reverse
    poly[0] = (poly[0] - b8) % r
    poly[1] = (poly[1] - b7) % r
    poly[2] = (poly[2] - b6) % r
    return poly + [b8 % r, b7 % r, b6 % r]


def round2_commitment(zk: PlonkZKey, values, blinders=tuple(range(11))):
    """[z]_1 (round2.rs:277-289) after round 1 with the same blinders; returns (beta, gamma, commit_z affine)."""
    c = zk.curve
    commits = round1_commitments(zk, values, blinders)
    beta, gamma = round2_challenges(zk, values, commits)
    z = z_polynomial(zk, values, beta, gamma, blinders)
    return beta, gamma, c.to_affine(c.msm(zk.p_tau[:len(z)], z, 1), 1)


# ------------------------------------------------------------------------------------------------ round 3
def _extended_evals(zk, coeffs):
    """driver.fft(poly, extended_domain): zero-padded to 4n, root roots[pow + 2]."""
    c = zk.curve
    _, roots = roots_of_unity(c)
    return ntt(list(coeffs) + [0] * (4 * zk.domain_size - len(coeffs)), roots[zk.pow + 2], c.r)


def quotient_polynomials(zk: PlonkZKey, values, beta, gamma, alpha, blinders=tuple(range(11)), trace=None):
    """compute_t (round3.rs:237-470) with the plain driver: T(X) on the 4n coset structure snarkjs uses -- every wire / z evaluation
    carries its blinding part separately (the *p vectors, multiplied by Z_H, Z_H^2, Z_H^3 through z1, z2, z3), the unblinded part is
    divided by Z_H in coefficient form.  Returns the coefficient lists of t1, t2, t3 (n + 1, n + 1, n + 6)."""
    c = zk.curve
    r, n = c.r, zk.domain_size
    n4 = 4 * n
    b = [x % r for x in blinders]
    _, roots = roots_of_unity(c)
    w_n, w_4n = roots[zk.pow], roots[zk.pow + 2]
    bufs = wire_buffers(zk, values)
    ev_a, ev_b, ev_c = (_extended_evals(zk, intt(buf, w_n, r)) for buf in bufs)
    zfull = z_polynomial(zk, values, beta, gamma, blinders=(0,) * 11)[:n]        # unblinded z coefficients
    ev_z = _extended_evals(zk, zfull)
    i4 = roots[2]                                        # Domains::root_of_unity_2: the primitive 4th root (types.rs:93); Z_H on the
    z1 = [0, (-1 + i4) % r, (-2) % r, (-1 - i4) % r]     # 4n domain takes the four values i4^k - 1, get_z1..3 (round3.rs:203-234)
    z2 = [0, (-2 * i4) % r, 4, (2 * i4) % r]
    z3 = [0, (2 + 2 * i4) % r, (-8) % r, (2 - 2 * i4) % r]
    sel = {k: v[1] for k, v in zk.selectors.items()}
    s1, s2, s3 = (zk.sigma[k][1] for k in range(3))
    alpha2 = alpha * alpha % r

    def mul4(a, bb, cc, d, ap, bp, cp, dp):     # mul4vec! (round3.rs:14-52): coefficients of prod (x + xp Z) in Z
        r0 = a * bb % r * cc % r * d % r
        a0 = (ap * bb * cc * d + a * bp * cc * d + a * bb * cp * d + a * bb * cc * dp) % r
        a1 = (ap * bp * cc * d + ap * bb * cp * d + ap * bb * cc * dp + a * bp * cp * d + a * bp * cc * dp + a * bb * cp * dp) % r
        a2 = (a * bp * cp * dp + ap * bb * cp * dp + ap * bp * cc * dp + ap * bp * cp * d) % r
        a3 = ap * bp * cp * dp % r
        return r0, a0, a1, a2, a3

    def post(a0, a1, a2, a3, i):                # mul4vec_post!
        m = i % 4
        return a0 if m == 0 else (a0 + z1[m] * a1 + z2[m] * a2 + z3[m] * a3) % r

    t_vec, tz_vec = [], []
    w = 1
    for i in range(n4):
        a, bb, cc, z = ev_a[i], ev_b[i], ev_c[i], ev_z[i]
        ap, bp, cp = (b[1] + b[0] * w) % r, (b[3] + b[2] * w) % r, (b[5] + b[4] * w) % r
        w2 = w * w % r
        zp = (b[6] * w2 + b[7] * w + b[8]) % r
        ww = w * w_n % r
        zw = ev_z[(n4 + 4 + i) % n4]
        zwp = (b[6] * ww * ww + b[7] * ww + b[8]) % r
        m = i % 4
        a0 = (a * bp + ap * bb) % r
        if m:
            a0 = (a0 + z1[m] * ap * bp) % r
        e1 = (sel["qm"][i] * a * bb + a * sel["ql"][i] + bb * sel["qr"][i] + cc * sel["qo"][i]) % r
        e1z = (sel["qm"][i] * a0 + ap * sel["ql"][i] + bp * sel["qr"][i] + cp * sel["qo"][i]) % r
        pi = 0
        for j, lag in enumerate(zk.lagrange):
            pi = (pi - lag[1][i] * bufs[0][j]) % r
        e1 = (e1 + pi + sel["qc"][i]) % r
        betaw = beta * w % r
        e2, *e2z = mul4((a + betaw + gamma) % r, (bb + betaw * zk.k1 + gamma) % r, (cc + betaw * zk.k2 + gamma) % r, z, ap, bp, cp, zp)
        e3, *e3z = mul4((a + s1[i] * beta + gamma) % r, (bb + s2[i] * beta + gamma) % r, (cc + s3[i] * beta + gamma) % r, zw, ap, bp, cp, zwp)
        e2, e2zv = alpha * e2 % r, alpha * post(*e2z, i) % r
        e3, e3zv = alpha * e3 % r, alpha * post(*e3z, i) % r
        l0 = zk.lagrange[0][1][i]
        e4 = alpha2 * l0 % r * ((z - 1) % r) % r
        e4z = alpha2 * l0 % r * zp % r
        t_vec.append((e1 + e2 - e3 + e4) % r)
        tz_vec.append((e1z + e2zv - e3zv + e4z) % r)
        w = w * w_4n % r
    if trace is not None:
        trace["t_evals"], trace["tz_evals"] = list(t_vec), list(tz_vec)
    ct = intt(t_vec, w_4n, r)
    for i in range(n):                                   # neg_vec_in_place_limit
        ct[i] = (-ct[i]) % r
    for i in range(n, n4):                               # division by Z_H = X^n - 1 in coefficient form
        ct[i] = (ct[i - n] - ct[i]) % r
    ctz = intt(tz_vec, w_4n, r)
    tf = [(x + y) % r for x, y in zip(ct, ctz)]
    t1, t2, t3 = tf[:n], tf[n:2 * n], tf[2 * n:3 * n + 6]
    t1 = t1 + [b[9]]
    t2[0] = (t2[0] - b[9]) % r
    t2 = t2 + [b[10]]
    t3[0] = (t3[0] - b[10]) % r
    return t1, t2, t3


def round3_commitments(zk: PlonkZKey, values, blinders=tuple(range(11))):
    """[t1]_1, [t2]_1, [t3]_1 (round3.rs:473-520) after rounds 1-2 with the same blinders; returns (alpha, commitments)."""
    c = zk.curve
    beta, gamma, cz = round2_commitment(zk, values, blinders)
    t = Keccak256Transcript(c)
    t.add_scalar(beta)
    t.add_scalar(gamma)
    t.add_point(cz)
    alpha = t.get_challenge()
    polys = quotient_polynomials(zk, values, beta, gamma, alpha, blinders)
    return alpha, [c.to_affine(c.msm(zk.p_tau[:len(p)], p, 1), 1) for p in polys]


# ------------------------------------------------------------------------------------------------ rounds 4 and 5, whole proof
def _horner(coeffs, x, r):
    acc = 0
    for cf in reversed(coeffs):
        acc = (acc * x + cf) % r
    return acc


def _div_by_linear(poly, beta, r):
    """div_by_zerofier(inout, 1, beta) (round5.rs:96-114): quotient of poly by (X - beta), remainder assumed zero."""
    inv = pow(beta, -1, r)
    res = list(poly)
    res[0] = (-inv * res[0]) % r
    for i in range(1, len(res)):
        res[i] = (res[i - 1] - res[i]) * inv % r
    return res[:-1]


def prove_plain(zk: PlonkZKey, values, blinders=tuple(range(11)), trace=None):
    """CoPlonk::prove with the plain driver and fixed blinders (co-plonk/src/lib.rs:80-99, round1.rs .. round5.rs): the snarkjs proof
    object -- commitments affine, evaluations ints."""
    c = zk.curve
    r, n = c.r, zk.domain_size
    _, roots = roots_of_unity(c)
    w_n = roots[zk.pow]
    pa, pb, pc = wire_polynomials(zk, values, blinders)
    commit = lambda p: c.to_affine(c.msm(zk.p_tau[:len(p)], p, 1), 1)
    A, B, C = commit(pa), commit(pb), commit(pc)
    beta, gamma = round2_challenges(zk, values, [A, B, C])
    pz = z_polynomial(zk, values, beta, gamma, blinders, trace=trace)
    Z = commit(pz)
    t = Keccak256Transcript(c)
    t.add_scalar(beta); t.add_scalar(gamma); t.add_point(Z)
    alpha = t.get_challenge()
    t1, t2, t3 = quotient_polynomials(zk, values, beta, gamma, alpha, blinders, trace=trace)
    T1, T2, T3 = commit(t1), commit(t2), commit(t3)
    if trace is not None:
        trace.update(poly_a=list(pa), poly_z=list(pz), t1=list(t1), t2=list(t2), t3=list(t3), buffer_a=wire_buffers(zk, values)[0])
    # round 4 (round4.rs:114-165)
    t = Keccak256Transcript(c)
    t.add_scalar(alpha); t.add_point(T1); t.add_point(T2); t.add_point(T3)
    xi = t.get_challenge()
    xiw = xi * w_n % r
    ev = {"eval_a": _horner(pa, xi, r), "eval_b": _horner(pb, xi, r), "eval_c": _horner(pc, xi, r), "eval_zw": _horner(pz, xiw, r),
          "eval_s1": _horner(zk.sigma[0][0], xi, r), "eval_s2": _horner(zk.sigma[1][0], xi, r)}
    # round 5 (round5.rs:140-380)
    t = Keccak256Transcript(c)
    for v_ in (xi, ev["eval_a"], ev["eval_b"], ev["eval_c"], ev["eval_s1"], ev["eval_s2"], ev["eval_zw"]):
        t.add_scalar(v_)
    v = [t.get_challenge()]
    for _ in range(4):
        v.append(v[-1] * v[0] % r)
    xin = pow(xi, n, r)
    zh = (xin - 1) % r
    pub = [x % r for x in values[1:zk.n_public + 1]]
    lag, w = [], 1
    for _ in range(max(1, zk.n_public)):               # calculate_lagrange_evaluations (lib.rs:160-185)
        lag.append(w * zh % r * pow(n * (xi - w) % r, -1, r) % r)
        w = w * w_n % r
    eval_pi = (-sum(l * p for l, p in zip(lag, pub))) % r
    betaxi = beta * xi % r
    e2 = (ev["eval_a"] + betaxi + gamma) * (ev["eval_b"] + betaxi * zk.k1 + gamma) % r * (ev["eval_c"] + betaxi * zk.k2 + gamma) % r * alpha % r
    e3 = (ev["eval_a"] + beta * ev["eval_s1"] + gamma) * (ev["eval_b"] + beta * ev["eval_s2"] + gamma) % r * ev["eval_zw"] % r * alpha % r
    e4 = alpha * alpha % r * lag[0] % r
    e24 = (e2 + e4) % r
    sel = {k: x[0] for k, x in zk.selectors.items()}
    ln = n + 6
    pr = [0] * ln
    for i in range(n):
        pr[i] = (sel["qm"][i] * ev["eval_a"] % r * ev["eval_b"] + sel["ql"][i] * ev["eval_a"] + sel["qr"][i] * ev["eval_b"]
                 + sel["qo"][i] * ev["eval_c"] + sel["qc"][i] - zk.sigma[2][0][i] * e3 % r * beta) % r
    for i, zc in enumerate(pz):
        pr[i] = (pr[i] + e24 * zc) % r
    xin2 = xin * xin % r
    tmp = [0] * ln
    for i, x in enumerate(t3):
        tmp[i] = xin2 * x % r
    for i, x in enumerate(t2):
        tmp[i] = (tmp[i] + xin * x) % r
    for i, x in enumerate(t1):
        tmp[i] = (tmp[i] + x) % r
    pr = [(x - zh * y) % r for x, y in zip(pr, tmp)]
    r0 = (eval_pi - e3 * (ev["eval_c"] + gamma) - e4) % r
    pr[0] = (pr[0] + r0) % r
    res = list(pr)
    for coeffs, f in ((pa, v[0]), (pb, v[1]), (pc, v[2]), (zk.sigma[0][0], v[3]), (zk.sigma[1][0], v[4])):
        for i, x in enumerate(coeffs):
            res[i] = (res[i] + f * x) % r
    res[0] = (res[0] - v[0] * ev["eval_a"] - v[1] * ev["eval_b"] - v[2] * ev["eval_c"] - v[3] * ev["eval_s1"] - v[4] * ev["eval_s2"]) % r
    wxi = _div_by_linear(res, xi, r)
    if trace is not None:
        trace.update(poly_r=list(pr), wxi=list(wxi))
    zq = list(pz)
    zq[0] = (zq[0] - ev["eval_zw"]) % r
    wxiw = _div_by_linear(zq, xiw, r)
    proof = {"A": A, "B": B, "C": C, "Z": Z, "T1": T1, "T2": T2, "T3": T3, "Wxi": commit(wxi), "Wxiw": commit(wxiw)}
    proof.update(ev)
    return proof


def proof_to_json(curve, proof):
    """snarkjs layout of PlonkProof (circom-types/src/plonk/proof.rs): points [x, y, "1"], evaluations as decimal strings."""
    import json
    out = {}
    for k in ("A", "B", "C", "Z", "T1", "T2", "T3"):
        out[k] = [str(proof[k][0]), str(proof[k][1]), "1"] if proof[k] is not None else ["0", "1", "0"]
    for k in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw"):
        out[k] = str(proof[k])
    for k in ("Wxi", "Wxiw"):
        out[k] = [str(proof[k][0]), str(proof[k][1]), "1"] if proof[k] is not None else ["0", "1", "0"]
    out["protocol"] = "plonk"
    out["curve"] = curve.circom_name
    return json.dumps(out)
