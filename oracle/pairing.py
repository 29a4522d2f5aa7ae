"""Ate pairing over BN254 and BLS12-381 with Python ints (oracle, test infrastructure only).

Used only to VERIFY Groth16 proofs: the reference does it through
`ark_groth16::Groth16::verify_proof` (/root/reference/co-circom/co-groth16/src/verifier.rs:23-43);
snarkjs is absent from this image, so the shipped snarkjs `circom.proof` +
`verification_key.json` pairs are the known-answer test for this module.

Fq12 is the flat extension Fq[w]/(w^12 - 2a w^6 + (a^2+1)) with u = w^6 - a, xi = a + u
(BN254: a = 9, BLS12-381: a = 1).  Any bilinear non-degenerate pairing validates the
Groth16 equation, so the sign of the BLS parameter is ignored and the line functions are
scaled by subfield elements (killed by the final exponentiation).
"""
from __future__ import annotations

from .curves import BN254, BLS12_381, Curve


class _F12:
    def __init__(self, curve: Curve, a: int):
        self.q = curve.q
        self.a = a
        # w^12 = c6*w^6 + c0
        self.c6 = (2 * a) % self.q
        self.c0 = (-(a * a + 1)) % self.q
        self.one = [1] + [0] * 11

    def mul(self, x, y):
        q = self.q
        t = [0] * 23
        for i, xi in enumerate(x):
            if xi:
                for j, yj in enumerate(y):
                    if yj:
                        t[i + j] += xi * yj
        for k in range(22, 11, -1):
            v = t[k]
            if v:
                t[k - 6] += v * self.c6
                t[k - 12] += v * self.c0
        return [v % q for v in t[:12]]

    def sqr(self, x):
        return self.mul(x, x)

    def pow(self, x, e):
        res = self.one
        for bit in bin(e)[2:]:
            res = self.sqr(res)
            if bit == "1":
                res = self.mul(res, x)
        return res

    def embed_f2(self, z, shift):
        """Fq2 element z = z0 + z1*u (u = w^6 - a) times w^shift (shift < 6)."""
        out = [0] * 12
        out[shift] = (z[0] - self.a * z[1]) % self.q
        out[shift + 6] = z[1] % self.q
        return out


class PairingEngine:
    def __init__(self, curve: Curve):
        self.c = curve
        if curve is BN254:
            self.f12 = _F12(curve, 9)
            self.xi = (9, 1)
            self.loop = 29793968203157093288  # 6x+2, x = 4965661367192848881
            self.twist_d = True               # D-type: psi(x,y) = (x w^2, y w^3)
        elif curve is BLS12_381:
            self.f12 = _F12(curve, 1)
            self.xi = (1, 1)
            self.loop = 0xD201000000010000   # |x|
            self.twist_d = False              # M-type: psi(x,y) = (x / w^2, y / w^3)
        else:
            raise ValueError("unknown curve")
        q = curve.q
        self.final_exp = (q ** 12 - 1) // curve.r
        if self.twist_d:
            self.gamma2 = curve.f2_pow(self.xi, (q - 1) // 3)
            self.gamma3 = curve.f2_pow(self.xi, (q - 1) // 2)

    # line through R (and Q, or tangent) on the twist, evaluated at P in G1
    def _line(self, R, Q, P):
        c = self.c
        xr, yr = R
        if Q is None or R == Q:
            num = c.f2_muls(c.f2_sqr(xr), 3)
            den = c.f2_add(yr, yr)
        else:
            num = c.f2_sub(Q[1], yr)
            den = c.f2_sub(Q[0], xr)
        lam = c.f2_mul(num, c.f2_inv(den))
        xp, yp = P
        f = self.f12
        c0 = c.f2_sub(yr, c.f2_mul(lam, xr))         # yR - lam*xR
        c1 = c.f2_muls(lam, xp)                      # lam*xP
        if self.twist_d:
            # l = -yP + lam*xP*w + (yR - lam*xR)*w^3
            out = f.embed_f2(c1, 1)
            t = f.embed_f2(c0, 3)
            out = [(x + y) % c.q for x, y in zip(out, t)]
            out[0] = (out[0] - yp) % c.q
        else:
            # l*w^3 = -yP*w^3 + lam*xP*w^2 + (yR - lam*xR)
            out = f.embed_f2(c1, 2)
            t = f.embed_f2(c0, 0)
            out = [(x + y) % c.q for x, y in zip(out, t)]
            out[3] = (out[3] - yp) % c.q
        return out, lam

    def _step(self, R, Q, lam):
        """R+Q (or 2R when Q is None) on the twist given the slope."""
        c = self.c
        xq = R[0] if Q is None else Q[0]
        x3 = c.f2_sub(c.f2_sub(c.f2_sqr(lam), R[0]), xq)
        y3 = c.f2_sub(c.f2_mul(lam, c.f2_sub(R[0], x3)), R[1])
        return (x3, y3)

    def miller(self, P, Q):
        """Miller loop f_{loop,Q}(P); P in G1 affine, Q in G2 affine (neither infinity)."""
        c, f = self.c, self.f12
        acc = f.one
        R = Q
        for bit in bin(self.loop)[3:]:
            l, lam = self._line(R, None, P)
            acc = f.mul(f.sqr(acc), l)
            R = self._step(R, None, lam)
            if bit == "1":
                l, lam = self._line(R, Q, P)
                acc = f.mul(acc, l)
                R = self._step(R, Q, lam)
        if self.twist_d:
            q1 = (c.f2_mul(c.f2_conj(Q[0]), self.gamma2), c.f2_mul(c.f2_conj(Q[1]), self.gamma3))
            q2 = (c.f2_mul(c.f2_conj(q1[0]), self.gamma2), c.f2_mul(c.f2_conj(q1[1]), self.gamma3))
            nq2 = (q2[0], c.f2_neg(q2[1]))
            l, lam = self._line(R, q1, P)
            acc = f.mul(acc, l)
            R = self._step(R, q1, lam)
            l, lam = self._line(R, nq2, P)
            acc = f.mul(acc, l)
        return acc

    def product_is_one(self, pairs):
        """prod e(P_i, Q_i) == 1 ?  Pairs with an infinity member contribute 1."""
        f = self.f12
        acc = f.one
        for P, Q in pairs:
            if P is None or Q is None:
                continue
            acc = f.mul(acc, self.miller(P, Q))
        return f.pow(acc, self.final_exp) == f.one

    def pairing(self, P, Q):
        return self.f12.pow(self.miller(P, Q), self.final_exp)


_ENGINES = {}


def engine(curve: Curve) -> PairingEngine:
    if curve.name not in _ENGINES:
        _ENGINES[curve.name] = PairingEngine(curve)
    return _ENGINES[curve.name]
