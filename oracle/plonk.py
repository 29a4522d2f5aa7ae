"""Round 1 of the Plonk prover restated on Python ints (oracle, test infrastructure only).

Follows /root/reference/co-circom/co-plonk/src/round1.rs:121-300 with PlainDriver and the reference's own deterministic
blinding hook (`Round1Challenges::deterministic`, b_i = i, round1.rs:101-108) -- the configuration of its bit-exact KATs
(round1.rs:344-427), which pin iNTT (snarkjs roots, co-plonk/src/types.rs:59-99) + MSM results exactly.
"""
from __future__ import annotations

from .formats import PlonkZKey
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


def wire_polynomials(zk: PlonkZKey, values, blinders):
    """compute_wire_polynomials (round1.rs:121-209): coefficients of the blinded a(X), b(X), c(X), length n + 2 each."""
    r = zk.curve.r
    get = witness_with_additions(zk, values)
    _, roots = roots_of_unity(zk.curve)
    omega = roots[zk.pow]
    out = []
    for k, m in enumerate((zk.map_a, zk.map_b, zk.map_c)):
        buf = [get(i) for i in m] + [0] * (zk.domain_size - zk.n_constraints)
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
