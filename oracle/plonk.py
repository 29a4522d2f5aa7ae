"""Round 1 of the Plonk prover restated on Python ints (oracle, test infrastructure only).

Follows /root/reference/co-circom/co-plonk/src/round1.rs:121-300 with PlainDriver and the reference's own deterministic
blinding hook (`Round1Challenges::deterministic`, b_i = i, round1.rs:101-108) -- the configuration of its bit-exact KATs
(round1.rs:344-427), which pin iNTT (snarkjs roots, co-plonk/src/types.rs:59-99) + MSM results exactly.
"""
from __future__ import annotations

from .formats import PlonkZKey
from .keccak import keccak256
from .ntt import roots_of_unity, intt


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


def z_polynomial(zk: PlonkZKey, values, beta, gamma, blinders=tuple(range(11))):
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
