"""ctypes binding of the C oracle (oracle/c/cocg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays are numpy uint64, little-endian limbs, Montgomery form:
  Fr vector        (n, 4)
  G1 affine        (n, 2*LQ)          LQ = 4 (BN254) / 6 (BLS12-381); (0,0) = infinity
  G2 affine        (n, 4*LQ)          x.c0 | x.c1 | y.c0 | y.c1
  Jacobian         (3*LQ,) / (6*LQ,)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .curves import Curve, BN254, BLS12_381

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "libcocg_oracle.so")
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        L = _LIB
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.orc_num_threads.restype = ci
        L.orc_set_threads.argtypes = [ci]
        L.orc_fr_vec_op.argtypes = [ci, ci, vp, vp, vp, sz]
        L.orc_rep3_mul_local.argtypes = [ci, vp, vp, vp, vp, vp, vp, sz]
        L.orc_distribute_powers.argtypes = [ci, vp, sz, vp, vp]
        L.orc_spmv.argtypes = [ci, vp, vp, vp, vp, vp, sz]
        L.orc_ntt.argtypes = [ci, vp, ctypes.c_uint, vp, ci]
        L.orc_msm.argtypes = [ci, ci, vp, vp, sz, vp]
        L.orc_ec_op.argtypes = [ci, ci, ci, vp, vp, vp]
        L.orc_gen_chain.argtypes = [ci, ci, vp, vp, sz, vp]
    return _LIB


def cid(curve: Curve) -> int:
    return 0 if curve is BN254 else 1


def lq(curve: Curve) -> int:
    return 4 if curve is BN254 else 6


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ---------------- int <-> limb conversions ----------------
def ints_to_limbs(vals, nl: int) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(8 * nl, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(vals), nl).copy()


def limbs_to_ints(arr: np.ndarray):
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    nl = arr.shape[-1]
    flat = arr.reshape(-1, nl)
    raw = flat.tobytes()
    return [int.from_bytes(raw[i * 8 * nl:(i + 1) * 8 * nl], "little") for i in range(flat.shape[0])]


def fr_to_mont(curve: Curve, vals) -> np.ndarray:
    """canonical ints -> (n,4) Montgomery limbs"""
    return ints_to_limbs([(v % curve.r) * curve.Rr % curve.r for v in vals], 4)


def fr_from_mont(curve: Curve, arr: np.ndarray):
    ri = pow(curve.Rr, -1, curve.r)
    return [(v * ri) % curve.r for v in limbs_to_ints(arr)]


def fq_mont(curve: Curve, v: int) -> int:
    return (v % curve.q) * curve.Rq % curve.q


def g_to_mont(curve: Curve, pts, group=1) -> np.ndarray:
    """affine oracle points (or None) -> packed Montgomery array"""
    L = lq(curve)
    rows = []
    for P in pts:
        if P is None:
            rows.append([0] * (2 * group))
        elif group == 1:
            rows.append([fq_mont(curve, P[0]), fq_mont(curve, P[1])])
        else:
            rows.append([fq_mont(curve, P[0][0]), fq_mont(curve, P[0][1]), fq_mont(curve, P[1][0]), fq_mont(curve, P[1][1])])
    flat = [c for row in rows for c in row]
    return ints_to_limbs(flat, L).reshape(len(pts), 2 * group * L)


def g_from_mont(curve: Curve, arr: np.ndarray, group=1):
    """packed Montgomery affine array -> list of oracle affine points"""
    L = lq(curve)
    qi = pow(curve.Rq, -1, curve.q)
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 2 * group * L)
    out = []
    for row in arr:
        cs = [(v * qi) % curve.q for v in limbs_to_ints(row.reshape(2 * group, L))]
        if not any(cs):
            out.append(None)
        elif group == 1:
            out.append((cs[0], cs[1]))
        else:
            out.append(((cs[0], cs[1]), (cs[2], cs[3])))
    return out


def jac_from_mont(curve: Curve, arr: np.ndarray, group=1):
    """Jacobian Montgomery limbs -> oracle AFFINE point (normalised)"""
    L = lq(curve)
    qi = pow(curve.Rq, -1, curve.q)
    cs = [(v * qi) % curve.q for v in limbs_to_ints(np.ascontiguousarray(arr, dtype=np.uint64).reshape(3 * group, L))]
    if group == 1:
        J = (cs[0], cs[1], cs[2])
    else:
        J = ((cs[0], cs[1]), (cs[2], cs[3]), (cs[4], cs[5]))
    return curve.to_affine(J, group)


# ---------------- thin wrappers ----------------
OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_TO_MONT, OP_FROM_MONT = range(6)


def fr_vec_op(curve, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
    lib().orc_fr_vec_op(cid(curve), op, _p(a), _p(b), _p(out), a.shape[0])
    return out


def rep3_mul_local(curve, aa, ab, ba, bb, mask=None):
    out = np.empty_like(aa)
    lib().orc_rep3_mul_local(cid(curve), _p(aa), _p(ab), _p(ba), _p(bb), _p(mask), _p(out), aa.shape[0])
    return out


def distribute_powers(curve, x, g_mont, c_mont):
    x = np.ascontiguousarray(x, dtype=np.uint64).copy()
    lib().orc_distribute_powers(cid(curve), _p(x), x.shape[0], _p(g_mont), _p(c_mont))
    return x


def spmv(curve, rowptr, col, coeff, z):
    rows = rowptr.shape[0] - 1
    out = np.zeros((rows, 4), dtype=np.uint64)
    lib().orc_spmv(cid(curve), _p(rowptr), _p(col), _p(coeff), _p(z), _p(out), rows)
    return out


def ntt(curve, a, omega_mont, inverse=False):
    """in-order DFT with generator `omega_mont` ((1,4) Montgomery); inverse=True expects omega^-1 and scales by 1/n"""
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    n = a.shape[0]
    logn = n.bit_length() - 1
    assert 1 << logn == n
    lib().orc_ntt(cid(curve), _p(a), logn, _p(omega_mont), 1 if inverse else 0)
    return a


def msm(curve, group, pts, scalars_mont):
    n = min(pts.shape[0], scalars_mont.shape[0])
    out = np.zeros(3 * group * lq(curve), dtype=np.uint64)
    pts = np.ascontiguousarray(pts)
    scalars_mont = np.ascontiguousarray(scalars_mont)
    lib().orc_msm(cid(curve), group, _p(pts), _p(scalars_mont), n, _p(out))
    return out


def ec_op(curve, group, op, a, b=None):
    L = lq(curve)
    out = np.zeros((2 if op == 2 else 3) * group * L, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
    lib().orc_ec_op(cid(curve), group, op, _p(a), _p(b), _p(out))
    return out


def gen_chain(curve, group, p0, q, n):
    """n affine points p0 + i*q (packed Montgomery)."""
    out = np.zeros((n, 2 * group * lq(curve)), dtype=np.uint64)
    lib().orc_gen_chain(cid(curve), group, _p(np.ascontiguousarray(p0)), _p(np.ascontiguousarray(q)), n, _p(out))
    return out
