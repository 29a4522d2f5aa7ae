"""Big-int field / curve arithmetic for BN254 and BLS12-381 (oracle, test infrastructure only).

Restates the arithmetic the reference gets from arkworks 0.4.x (ark-bn254 0.4.0,
ark-bls12-381 0.4.0, ark-ec/ark-ff 0.4.2 -- /root/reference/Cargo.toml:33-40; the
crates are not vendored).  All values are canonical (non-Montgomery) Python ints.

Points:  G1 affine = (x, y) | None (infinity);  G2 affine = ((x0,x1),(y0,y1)) | None.
Jacobian = (X, Y, Z) with Z == 0 (or (0,0)) for infinity -- arkworks' `Projective`
(short_weierstrass Jacobian) is the type `msm_unchecked` returns
(/root/reference/mpc-core/src/protocols/rep3.rs:934-947).
"""
from __future__ import annotations


class Curve:
    """Parameters of one pairing-friendly curve."""

    def __init__(self, name, circom_name, q, r, b1, b2, g1, g2, two_adicity):
        self.name = name
        self.circom_name = circom_name  # circom-types/src/traits.rs:18,24-32
        self.q = q
        self.r = r
        self.b1 = b1           # G1: y^2 = x^3 + b1
        self.b2 = b2           # G2: y^2 = x^3 + b2, b2 in Fq2
        self.g1 = g1
        self.g2 = g2
        self.two_adicity = two_adicity
        self.n8q = (q.bit_length() + 7) // 8
        self.n8r = (r.bit_length() + 7) // 8
        # Montgomery radix used by snarkjs / arkworks: 2^(64*limbs)
        self.Rq = 1 << (64 * ((q.bit_length() + 63) // 64))
        self.Rr = 1 << (64 * ((r.bit_length() + 63) // 64))

    # ---------------- Fq2 = Fq[u]/(u^2+1) ----------------
    def f2_add(self, a, b):
        q = self.q
        return ((a[0] + b[0]) % q, (a[1] + b[1]) % q)

    def f2_sub(self, a, b):
        q = self.q
        return ((a[0] - b[0]) % q, (a[1] - b[1]) % q)

    def f2_neg(self, a):
        q = self.q
        return ((-a[0]) % q, (-a[1]) % q)

    def f2_mul(self, a, b):
        q = self.q
        return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def f2_sqr(self, a):
        return self.f2_mul(a, a)

    def f2_muls(self, a, s):
        q = self.q
        return ((a[0] * s) % q, (a[1] * s) % q)

    def f2_inv(self, a):
        q = self.q
        n = pow((a[0] * a[0] + a[1] * a[1]) % q, q - 2, q)
        return ((a[0] * n) % q, (-a[1] * n) % q)

    def f2_conj(self, a):
        return (a[0], (-a[1]) % self.q)

    def f2_pow(self, a, e):
        res = (1, 0)
        base = a
        while e:
            if e & 1:
                res = self.f2_mul(res, base)
            base = self.f2_sqr(base)
            e >>= 1
        return res

    # ---------------- field "vtables" so G1/G2 share the group law ----------------
    def field(self, group):
        return _FQ(self) if group == 1 else _FQ2(self)

    def b(self, group):
        return self.b1 if group == 1 else self.b2

    def gen(self, group):
        return self.g1 if group == 1 else self.g2

    # ---------------- group law (generic over G1/G2) ----------------
    def is_on_curve(self, P, group=1):
        if P is None:
            return True
        F = self.field(group)
        x, y = P
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b(group))

    def to_jac(self, P, group=1):
        F = self.field(group)
        if P is None:
            return (F.one, F.one, F.zero)
        return (P[0], P[1], F.one)

    def to_affine(self, J, group=1):
        F = self.field(group)
        X, Y, Z = J
        if Z == F.zero:
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_double(self, J, group=1):
        F = self.field(group)
        X, Y, Z = J
        if Z == F.zero:
            return J
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jac_add(self, P, Q, group=1):
        F = self.field(group)
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        if Z1 == F.zero:
            return Q
        if Z2 == F.zero:
            return P
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jac_double(P, group)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jac_neg(self, P, group=1):
        F = self.field(group)
        return (P[0], F.neg(P[1]), P[2])

    def jac_mul(self, P, k, group=1):
        F = self.field(group)
        acc = (F.one, F.one, F.zero)
        if k < 0:
            return self.jac_neg(self.jac_mul(P, -k, group), group)
        for bit in bin(k)[2:] if k else "":
            acc = self.jac_double(acc, group)
            if bit == "1":
                acc = self.jac_add(acc, P, group)
        return acc

    def jac_eq(self, P, Q, group=1):
        return self.to_affine(P, group) == self.to_affine(Q, group)

    def add(self, P, Q, group=1):
        return self.to_affine(self.jac_add(self.to_jac(P, group), self.to_jac(Q, group), group), group)

    def neg(self, P, group=1):
        if P is None:
            return None
        return (P[0], self.field(group).neg(P[1]))

    def mul(self, P, k, group=1):
        return self.to_affine(self.jac_mul(self.to_jac(P, group), k % self.r, group), group)

    def msm(self, points, scalars, group=1):
        """Sum_i scalars[i]*points[i], truncating to min(len) like arkworks'
        `msm_unchecked` (plain.rs:408-416).  Simple windowed bucket method; returns Jacobian."""
        F = self.field(group)
        n = min(len(points), len(scalars))
        inf = (F.one, F.one, F.zero)
        if n == 0:
            return inf
        c = 4 if n < 32 else 8
        nbits = self.r.bit_length()
        nwin = (nbits + c - 1) // c
        total = inf
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jac_double(total, group)
            buckets = [None] * (1 << c)
            for i in range(n):
                d = ((scalars[i] % self.r) >> (w * c)) & ((1 << c) - 1)
                if d and points[i] is not None:
                    pj = (points[i][0], points[i][1], F.one)
                    buckets[d] = pj if buckets[d] is None else self.jac_add(buckets[d], pj, group)
            run = inf
            acc = inf
            for d in range((1 << c) - 1, 0, -1):
                if buckets[d] is not None:
                    run = self.jac_add(run, buckets[d], group)
                acc = self.jac_add(acc, run, group)
            total = self.jac_add(total, acc, group)
        return total


class _FQ:
    def __init__(self, c):
        self.q = c.q
        self.zero = 0
        self.one = 1

    def add(self, a, b):
        return (a + b) % self.q

    def sub(self, a, b):
        return (a - b) % self.q

    def neg(self, a):
        return (-a) % self.q

    def mul(self, a, b):
        return (a * b) % self.q

    def sqr(self, a):
        return (a * a) % self.q

    def inv(self, a):
        return pow(a, self.q - 2, self.q)


class _FQ2:
    def __init__(self, c):
        self.c = c
        self.zero = (0, 0)
        self.one = (1, 0)
        self.add = c.f2_add
        self.sub = c.f2_sub
        self.neg = c.f2_neg
        self.mul = c.f2_mul
        self.sqr = c.f2_sqr
        self.inv = c.f2_inv


# ----------------------------------------------------------------------------
# Curve constants (generators from the standards).
# ----------------------------------------------------------------------------
_BN_Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_BN_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _bn_b2():
    # 3 / (9 + u)
    q = _BN_Q
    n = pow(9 * 9 + 1, q - 2, q)
    inv = ((9 * n) % q, (-1 * n) % q)
    return ((3 * inv[0]) % q, (3 * inv[1]) % q)


BN254 = Curve(
    "bn254", "bn128", _BN_Q, _BN_R, 3, _bn_b2(),
    (1, 2),
    (
        (10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531),
    ),
    28,
)

_BLS_Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_BLS_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

BLS12_381 = Curve(
    "bls12_381", "bls12381", _BLS_Q, _BLS_R, 4, (4, 4),
    (
        0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
        0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
    ),
    (
        (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
         0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
        (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
         0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
    ),
    32,
)

CURVES = {"bn254": BN254, "bls12_381": BLS12_381, "bn128": BN254, "bls12381": BLS12_381}
