"""snarkjs / circom on-disk formats -> oracle types (test infrastructure only).

Follows /root/reference/co-circom/circom-types/src:
  binfile.rs:52-105 (container), groth16/zkey.rs:139-316 (Groth16 zkey),
  traits.rs:47-69,107-155 (Montgomery readers; (0,0) = infinity),
  witness.rs:51-92 (.wtns), groth16/proof.rs:7-29 and traits.rs:186-233 (JSON points),
  groth16/verification_key.rs:15-54.
"""
from __future__ import annotations

import json
import struct

from .curves import CURVES, Curve, BN254, BLS12_381


def _u32(buf, off):
    return struct.unpack_from("<I", buf, off)[0]


def _u64(buf, off):
    return struct.unpack_from("<Q", buf, off)[0]


def read_binfile(data: bytes, magic: bytes):
    """binfile.rs:52-97: magic, u32 version, u32 nSections, then (u32 id, u64 len, bytes)*."""
    if data[:4] != magic:
        raise ValueError("bad magic %r" % data[:4])
    version = _u32(data, 4)
    nsec = _u32(data, 8)
    off = 12
    sections = {}
    for _ in range(nsec):
        sid = _u32(data, off)
        slen = _u64(data, off + 4)
        off += 12
        sections.setdefault(sid, data[off:off + slen])
        off += slen
    return version, sections


def curve_from_q(q: int) -> Curve:
    for c in (BN254, BLS12_381):
        if c.q == q:
            return c
    raise ValueError("unknown base field modulus")


class _Rd:
    def __init__(self, buf):
        self.buf = buf
        self.off = 0

    def u32(self):
        v = _u32(self.buf, self.off)
        self.off += 4
        return v

    def int_le(self, n):
        v = int.from_bytes(self.buf[self.off:self.off + n], "little")
        self.off += n
        return v


def _mont_inv(R, p):
    return pow(R, -1, p)


def read_g1(rd: _Rd, c: Curve, rinv):
    x = rd.int_le(c.n8q)
    y = rd.int_le(c.n8q)
    if x == 0 and y == 0:
        return None
    return ((x * rinv) % c.q, (y * rinv) % c.q)


def read_g2(rd: _Rd, c: Curve, rinv):
    v = [rd.int_le(c.n8q) for _ in range(4)]
    if not any(v):
        return None
    v = [(t * rinv) % c.q for t in v]
    return ((v[0], v[1]), (v[2], v[3]))


class Groth16ZKey:
    """Mirror of circom-types/src/groth16/zkey.rs:47-71 (values canonical)."""

    def __init__(self):
        self.curve = None
        self.n_vars = self.n_public = self.domain_size = self.pow = 0
        self.alpha_g1 = self.beta_g1 = self.beta_g2 = self.gamma_g2 = self.delta_g1 = self.delta_g2 = None
        self.ic = []
        self.a_query = self.b_g1_query = self.b_g2_query = self.l_query = self.h_query = None
        self.a_rows = self.b_rows = None  # list[list[(coeff, signal)]]
        self.num_constraints = 0

    @property
    def num_inputs(self):  # ConstraintMatrices.num_instance_variables (zkey.rs:212)
        return self.n_public + 1


def parse_groth16_zkey(data: bytes, check_points: bool = True) -> Groth16ZKey:
    _, sec = read_binfile(data, b"zkey")
    if _u32(sec[1], 0) != 1:
        raise ValueError("not a groth16 zkey")
    rd = _Rd(sec[2])
    n8q = rd.u32()
    q = rd.int_le(n8q)
    c = curve_from_q(q)
    n8r = rd.u32()
    r = rd.int_le(n8r)
    if r != c.r or n8q != c.n8q or n8r != c.n8r:
        raise ValueError("invalid prime in header")
    zk = Groth16ZKey()
    zk.curve = c
    zk.n_vars = rd.u32()
    zk.n_public = rd.u32()
    zk.domain_size = rd.u32()
    if zk.domain_size == 0 or zk.domain_size & (zk.domain_size - 1):
        raise ValueError("domain size must be a power of two")
    zk.pow = zk.domain_size.bit_length() - 1
    rqi = _mont_inv(c.Rq, c.q)
    zk.alpha_g1 = read_g1(rd, c, rqi)
    zk.beta_g1 = read_g1(rd, c, rqi)
    zk.beta_g2 = read_g2(rd, c, rqi)
    zk.gamma_g2 = read_g2(rd, c, rqi)
    zk.delta_g1 = read_g1(rd, c, rqi)
    zk.delta_g2 = read_g2(rd, c, rqi)

    def g1s(sid, n):
        rd = _Rd(sec[sid])
        return [read_g1(rd, c, rqi) for _ in range(n)]

    def g2s(sid, n):
        rd = _Rd(sec[sid])
        return [read_g2(rd, c, rqi) for _ in range(n)]

    zk.ic = g1s(3, zk.n_public + 1)
    zk.a_query = g1s(5, zk.n_vars)
    zk.b_g1_query = g1s(6, zk.n_vars)
    zk.b_g2_query = g2s(7, zk.n_vars)
    zk.l_query = g1s(8, zk.n_vars - zk.n_public - 1)
    zk.h_query = g1s(9, zk.domain_size)
    if check_points:
        for pts, g in ((zk.ic, 1), (zk.a_query, 1), (zk.b_g1_query, 1), (zk.l_query, 1), (zk.h_query, 1), (zk.b_g2_query, 2)):
            for P in pts:
                if not c.is_on_curve(P, g):
                    raise ValueError("point not on curve")
    # section 4: coefficients stored x R^2 (traits.rs:65-67)
    rd = _Rd(sec[4])
    ncoef = rd.u32()
    rri2 = pow(c.Rr * c.Rr, -1, c.r)
    mats = [[[] for _ in range(zk.domain_size)] for _ in range(2)]
    max_row = 0
    for _ in range(ncoef):
        m = rd.u32()
        row = rd.u32()
        sig = rd.u32()
        val = (rd.int_le(c.n8r) * rri2) % c.r
        max_row = max(max_row, row)
        mats[m][row].append((val, sig))
    zk.num_constraints = max_row - zk.n_public           # zkey.rs:196
    zk.a_rows = mats[0][:zk.num_constraints]             # zkey.rs:198-204
    zk.b_rows = mats[1][:zk.num_constraints]
    return zk


def parse_wtns(data: bytes):
    """witness.rs:51-92 -> (curve, [canonical values])."""
    if data[:4] != b"wtns":
        raise ValueError("bad magic")
    version = _u32(data, 4)
    nsec = _u32(data, 8)
    if version > 2 or nsec > 2:
        raise ValueError("unsupported wtns")
    off = 12 + 12
    n8 = _u32(data, off)
    off += 4
    mod = int.from_bytes(data[off:off + n8], "little")
    off += n8
    curve = None
    for c in (BN254, BLS12_381):
        if c.r == mod:
            curve = c
    if curve is None:
        raise ValueError("wrong scalar field")
    n = _u32(data, off)
    off += 4 + 12
    vals = [int.from_bytes(data[off + i * n8: off + (i + 1) * n8], "little") % mod for i in range(n)]
    return curve, vals


# ---------------- JSON (traits.rs:186-233, proof.rs, verification_key.rs) ----------------
def g1_to_json(P):
    if P is None:
        return ["0", "1", "0"]
    return [str(P[0]), str(P[1]), "1"]


def g2_to_json(P):
    if P is None:
        return [["0", "0"], ["1", "0"], ["0", "0"]]
    return [[str(P[0][0]), str(P[0][1])], [str(P[1][0]), str(P[1][1])], ["1", "0"]]


def g1_from_json(v):
    if v[2] == "0":
        return None
    return (int(v[0]), int(v[1]))


def g2_from_json(v):
    if v[2][0] == "0" and v[2][1] == "0":
        return None
    return ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))


def proof_to_json(curve: Curve, A, B, C):
    return {"pi_a": g1_to_json(A), "pi_b": g2_to_json(B), "pi_c": g1_to_json(C),
            "protocol": "groth16", "curve": curve.circom_name}


def proof_from_json(text: str):
    d = json.loads(text)
    c = CURVES[d["curve"]]
    return c, g1_from_json(d["pi_a"]), g2_from_json(d["pi_b"]), g1_from_json(d["pi_c"])


class VerifyingKey:
    def __init__(self, curve, alpha_g1, beta_g2, gamma_g2, delta_g2, ic):
        self.curve, self.alpha_g1, self.beta_g2 = curve, alpha_g1, beta_g2
        self.gamma_g2, self.delta_g2, self.ic = gamma_g2, delta_g2, ic


def vk_from_json(text: str) -> VerifyingKey:
    d = json.loads(text)
    c = CURVES[d["curve"]]
    return VerifyingKey(c, g1_from_json(d["vk_alpha_1"]), g2_from_json(d["vk_beta_2"]),
                        g2_from_json(d["vk_gamma_2"]), g2_from_json(d["vk_delta_2"]),
                        [g1_from_json(p) for p in d["IC"]])


def vk_from_zkey(zk: Groth16ZKey) -> VerifyingKey:
    return VerifyingKey(zk.curve, zk.alpha_g1, zk.beta_g2, zk.gamma_g2, zk.delta_g2, zk.ic)


# ----------------------------------------------------------------------------
# Plonk zkey (circom-types/src/plonk/zkey.rs:47-90, 160-330): only what round 1 needs
# ----------------------------------------------------------------------------
class PlonkZKey:
    def __init__(self):
        self.curve = None
        self.n_vars = self.n_public = self.domain_size = self.n_additions = self.n_constraints = self.pow = 0
        self.additions = []       # (signal_id1, signal_id2, factor1, factor2), factors canonical
        self.map_a = self.map_b = self.map_c = None
        self.p_tau = None         # domain_size + 6 G1 points (zkey.rs:222-225)
        self.k1 = self.k2 = 0
        self.vk_points = {}       # qm, ql, qr, qo, qc, s1, s2, s3 (G1 affine)
        self.x_2 = None
        self.sigma = None         # [(coefficients, 4n extended evaluations)] x 3 when section 12 is present
        self.selectors = None     # qm, ql, qr, qo, qc -> (coefficients, evaluations)
        self.lagrange = None      # per public input (coefficients, evaluations)


def parse_plonk_zkey(data: bytes) -> PlonkZKey:
    _, sec = read_binfile(data, b"zkey")
    if _u32(sec[1], 0) != 2:
        raise ValueError("not a plonk zkey")
    rd = _Rd(sec[2])                                   # PlonkHeader::read, zkey.rs:375-420
    n8q = rd.u32()
    c = curve_from_q(rd.int_le(n8q))
    n8r = rd.u32()
    if rd.int_le(n8r) != c.r:
        raise ValueError("invalid prime in header")
    zk = PlonkZKey()
    zk.curve = c
    zk.n_vars, zk.n_public, zk.domain_size, zk.n_additions, zk.n_constraints = (rd.u32() for _ in range(5))
    if zk.domain_size == 0 or zk.domain_size & (zk.domain_size - 1):
        raise ValueError("domain size must be a power of two")
    zk.pow = zk.domain_size.bit_length() - 1
    rri = _mont_inv(c.Rr, c.r)
    rqi_h = _mont_inv(c.Rq, c.q)
    # VerifyingKey::new, plonk/zkey.rs:329-355: k1, k2 (Montgomery limbs), Qm Ql Qr Qo Qc S1 S2 S3 in G1, X_2 in G2
    zk.k1, zk.k2 = (rd.int_le(n8r) * rri) % c.r, (rd.int_le(n8r) * rri) % c.r
    zk.vk_points = {name: read_g1(rd, c, rqi_h) for name in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3")}
    zk.x_2 = read_g2(rd, c, rqi_h)
    def _poly_evals(section, count):                   # CircomPolynomial: n coefficients + 4n extended evaluations (zkey.rs:186-208)
        n = zk.domain_size
        rdp = _Rd(section)
        out = []
        for _ in range(count):
            coeffs = [(rdp.int_le(n8r) * rri) % c.r for _ in range(n)]
            evals = [(rdp.int_le(n8r) * rri) % c.r for _ in range(4 * n)]
            out.append((coeffs, evals))
        return out

    if all(k in sec for k in (7, 8, 9, 10, 11, 13)):   # selector polynomials and the Lagrange polynomials of the public inputs
        zk.selectors = {name: _poly_evals(sec[sid], 1)[0] for name, sid in (("qm", 7), ("ql", 8), ("qr", 9), ("qo", 10), ("qc", 11))}
        zk.lagrange = _poly_evals(sec[13], max(zk.n_public, 1) if len(sec[13]) >= 5 * zk.domain_size * n8r * max(zk.n_public, 1) else zk.n_public)
    if 12 in sec:                                      # sigma1..3: n coefficients + 4n extended evaluations each, zkey.rs:186-208, 250-262
        n = zk.domain_size
        rd12 = _Rd(sec[12])
        zk.sigma = []
        for _ in range(3):
            coeffs = [(rd12.int_le(n8r) * rri) % c.r for _ in range(n)]
            evals = [(rd12.int_le(n8r) * rri) % c.r for _ in range(4 * n)]
            zk.sigma.append((coeffs, evals))
    rd = _Rd(sec[3])                                   # additions_indices, zkey.rs:159-177: factors are Montgomery limbs
    for _ in range(zk.n_additions):
        s1, s2 = rd.u32(), rd.u32()
        f1, f2 = rd.int_le(n8r), rd.int_le(n8r)
        zk.additions.append((s1, s2, (f1 * rri) % c.r, (f2 * rri) % c.r))
    maps = []
    for sid in (4, 5, 6):                              # id_map, zkey.rs:179-185
        rd = _Rd(sec[sid])
        maps.append([rd.u32() for _ in range(zk.n_constraints)])
    zk.map_a, zk.map_b, zk.map_c = maps
    rqi = _mont_inv(c.Rq, c.q)
    rd = _Rd(sec[14])                                  # taus
    zk.p_tau = [read_g1(rd, c, rqi) for _ in range(zk.domain_size + 6)]
    return zk
