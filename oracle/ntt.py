"""Radix-2 NTT over Fr with the snarkjs root convention (oracle, test infrastructure only).

Root table: /root/reference/co-circom/co-circom-snarks/src/lib.rs:208-221 (`roots_of_unity`).
Transform convention: ark-poly 0.4.2 `Radix2EvaluationDomain` (not vendored) with `group_gen`
overridden (/root/reference/co-circom/co-groth16/src/groth16.rs:57-77): forward
out[i] = sum_j in[j] * w^(i*j), natural order in and out; inverse uses w^-1 and scales by n^-1;
inputs shorter than the domain are zero-extended.
"""
from __future__ import annotations

from functools import lru_cache

from .curves import Curve, CURVES


@lru_cache(maxsize=None)
def _roots(name: str):
    c = CURVES[name]
    r = c.r
    q = 1
    while pow(q, (r - 1) // 2, r) != r - 1:   # smallest quadratic non-residue
        q += 1
    s = c.two_adicity
    t = (r - 1) >> s
    assert (r - 1) == t << s and t & 1
    z = pow(q, t, r)
    roots = [z]
    for _ in range(s):
        roots.append((roots[-1] * roots[-1]) % r)
    roots.reverse()          # roots[k] has order 2^k, roots[0] == 1
    assert roots[0] == 1
    return q, roots


def roots_of_unity(curve: Curve):
    return _roots(curve.name)


def groth16_roots(curve: Curve, pow_: int):
    """(domain generator, coset generator) exactly as groth16.rs:57-77."""
    q, roots = roots_of_unity(curve)
    omega = roots[pow_]
    if pow_ == curve.two_adicity:
        g = (q * q) % curve.r
    else:
        g = roots[pow_ + 1]
    return omega, g


def ntt(vals, omega: int, r: int):
    """In-order forward DFT; len(vals) must be a power of two."""
    n = len(vals)
    if n == 1:
        return list(vals)
    logn = n.bit_length() - 1
    assert 1 << logn == n
    a = [0] * n
    for i, v in enumerate(vals):          # bit-reversal permutation
        a[int(format(i, "0%db" % logn)[::-1], 2)] = v % r
    m = 1
    while m < n:
        wm = pow(omega, n // (2 * m), r)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                u = a[k + j]
                t = (a[k + j + m] * w) % r
                a[k + j] = (u + t) % r
                a[k + j + m] = (u - t) % r
                w = (w * wm) % r
        m *= 2
    return a


def intt(vals, omega: int, r: int):
    n = len(vals)
    out = ntt(vals, pow(omega, -1, r), r)
    ninv = pow(n, -1, r)
    return [(v * ninv) % r for v in out]


def dft_naive(vals, omega: int, r: int):
    n = len(vals)
    return [sum(vals[j] * pow(omega, i * j, r) for j in range(n)) % r for i in range(n)]


def distribute_powers(vals, g: int, c: int, r: int):
    """x_i <- x_i * c * g^i  (rep3.rs:681-688 / plain.rs:235-241)."""
    out = []
    p = c % r
    for v in vals:
        out.append((v * p) % r)
        p = (p * g) % r
    return out
