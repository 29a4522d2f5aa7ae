"""Thin numpy-facing wrapper over the C ABI: one `Context` = one `cocg_ctx` (one MPC driver).

Arrays are numpy uint64 limbs exactly as a Rust caller's arkworks values lie in memory:
Fr vectors (n, 4); G1 affine (n, 2*LQ); G2 affine (n, 4*LQ); Jacobian (3*LQ,) / (6*LQ,), LQ = 4 (BN254) or 6
(BLS12-381).  Every method goes to libcocg.so; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import CocgError


class DeviceVec:
    """n Fr elements resident in HBM (a share-vector component)."""

    def __init__(self, ctx: "Context", n: int, ptr: int | None = None, owned: bool = True):
        self.ctx, self.n, self.owned = ctx, n, owned
        if ptr is None:
            p = ctypes.c_void_p()
            ctx._ck(ctx.L.cocg_malloc(ctx.h, max(n, 1) * 32, ctypes.byref(p)))
            ptr = p.value
        self.ptr = ptr

    def free(self):
        if self.owned and self.ptr:
            self.ctx._ck(self.ctx.L.cocg_free(self.ctx.h, self.ptr))
            self.ptr = None

    def to_host(self) -> np.ndarray:
        out = np.empty((self.n, 4), dtype=np.uint64)
        self.ctx._ck(self.ctx.L.cocg_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.n * 32))
        return out

    def slice(self, off: int, n: int) -> "DeviceVec":
        assert 0 <= off and off + n <= self.n
        return DeviceVec(self.ctx, n, self.ptr + off * 32, owned=False)


class Context:
    def __init__(self, curve: int = _lib.BN254, device: int = 0):
        self.L = _lib.load()
        h = ctypes.c_void_p()
        if self.L.cocg_create(ctypes.byref(h), device, curve):
            raise CocgError(self.L.cocg_last_error(None).decode())
        self.h = h
        self.curve = curve
        self.lq = 4 if curve == _lib.BN254 else 6

    def close(self):
        if self.h:
            self.L.cocg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc:
            raise CocgError(self.L.cocg_last_error(self.h).decode())

    # ---------------- memory
    def upload(self, arr: np.ndarray) -> DeviceVec:
        arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
        v = DeviceVec(self, arr.shape[0])
        self._ck(self.L.cocg_h2d(self.h, v.ptr, arr.ctypes.data, arr.shape[0] * 32))
        return v

    def zeros(self, n: int) -> DeviceVec:
        v = DeviceVec(self, n)
        self._ck(self.L.cocg_memset0(self.h, v.ptr, n * 32))
        return v

    def set_stream(self, stream_ptr: int | None):
        self._ck(self.L.cocg_set_stream(self.h, stream_ptr))

    def set_stream_priority(self, high: bool):
        """Re-creates the context's own stream with the device's highest (or the default) scheduling priority; call while idle."""
        self._ck(self.L.cocg_set_stream_priority(self.h, 1 if high else 0))

    def sync(self):
        self._ck(self.L.cocg_sync(self.h))

    def launch_count(self) -> int:
        return int(self.L.cocg_launch_count(self.h))

    # ---------------- element-wise (a5, a6, a7)
    def vec_op(self, op: int, a: DeviceVec, b: DeviceVec | None = None, out: DeviceVec | None = None) -> DeviceVec:
        out = out or DeviceVec(self, a.n)
        self._ck(self.L.cocg_vec_op(self.h, op, a.ptr, b.ptr if b else None, out.ptr, a.n))
        return out

    def vec_axpy(self, a: np.ndarray, x: DeviceVec, y: DeviceVec | None = None, out: DeviceVec | None = None) -> DeviceVec:
        """out = a * x (+ y), a: one Montgomery Fr (4 limbs)."""
        out = out or DeviceVec(self, x.n)
        a = np.ascontiguousarray(a, dtype=np.uint64)
        self._ck(self.L.cocg_vec_axpy(self.h, a.ctypes.data, x.ptr, y.ptr if y else None, out.ptr, x.n))
        return out

    def rep3_mul_local(self, aa, ab, ba, bb, mask=None, out=None) -> DeviceVec:
        out = out or DeviceVec(self, aa.n)
        self._ck(self.L.cocg_rep3_mul_local(self.h, aa.ptr, ab.ptr, ba.ptr, bb.ptr, mask.ptr if mask else None, out.ptr, aa.n))
        return out

    def scale_powers(self, x: DeviceVec, g: np.ndarray, c: np.ndarray):
        g = np.ascontiguousarray(g, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint64)
        self._ck(self.L.cocg_vec_scale_powers(self.h, x.ptr, x.n, g.ctypes.data, c.ctypes.data))

    # ---------------- CoPlonk vector primitives (csrc/poly.cu, csrc/plonk.cu)
    def upload_u32(self, arr: np.ndarray) -> int:
        """uint32 index array -> raw device address (cocg_malloc; free with free_raw)."""
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        p = ctypes.c_void_p()
        self._ck(self.L.cocg_malloc(self.h, max(arr.nbytes, 4), ctypes.byref(p)))
        self._ck(self.L.cocg_h2d(self.h, p, arr.ctypes.data, arr.nbytes))
        return p.value

    def free_raw(self, ptr: int):
        self._ck(self.L.cocg_free(self.h, ptr))

    def vec_gather(self, src: DeviceVec, idx_ptr: int, n: int) -> DeviceVec:
        out = DeviceVec(self, n)
        self._ck(self.L.cocg_vec_gather(self.h, src.ptr, src.n, idx_ptr, out.ptr, n))
        return out

    def vec_scan(self, op: int, x: DeviceVec, out: DeviceVec | None = None) -> DeviceVec:
        out = out or DeviceVec(self, x.n)
        self._ck(self.L.cocg_vec_scan(self.h, op, x.ptr, out.ptr, x.n))
        return out

    def vec_inv(self, x: DeviceVec, out: DeviceVec | None = None):
        """-> (1 / x element-wise with 0 -> 0, number of zero inputs)"""
        out = out or DeviceVec(self, x.n)
        z = ctypes.c_size_t()
        self._ck(self.L.cocg_vec_inv(self.h, x.ptr, out.ptr, x.n, ctypes.byref(z)))
        return out, int(z.value)

    def poly_eval(self, coeffs: DeviceVec, point: np.ndarray, n: int | None = None) -> np.ndarray:
        point = np.ascontiguousarray(point, dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint64)
        self._ck(self.L.cocg_poly_eval(self.h, coeffs.ptr, coeffs.n if n is None else n, point.ctypes.data, out.ctypes.data))
        return out

    def vec_lincomb(self, vecs, factors: np.ndarray, n: int) -> DeviceVec:
        vecs = list(vecs)
        out = DeviceVec(self, n)
        P = (ctypes.c_void_p * len(vecs))(*[v.ptr for v in vecs])
        Ls = (ctypes.c_size_t * len(vecs))(*[v.n for v in vecs])
        f = np.ascontiguousarray(factors, dtype=np.uint64)
        self._ck(self.L.cocg_vec_lincomb(self.h, len(vecs), P, Ls, f.ctypes.data, out.ptr, n))
        return out

    def vec_fill(self, n: int, value: np.ndarray) -> DeviceVec:
        out = DeviceVec(self, n)
        v = np.ascontiguousarray(value, dtype=np.uint64)
        self._ck(self.L.cocg_vec_fill(self.h, out.ptr, n, v.ctypes.data))
        return out

    # ---------------- NTT (a4 + a5)
    def ntt(self, vecs, log_n: int, root: np.ndarray, inverse: bool = False, coset_g: np.ndarray | None = None):
        vecs = list(vecs)
        arr = (ctypes.c_void_p * len(vecs))(*[v.ptr for v in vecs])
        root = np.ascontiguousarray(root, dtype=np.uint64)
        cg = None if coset_g is None else np.ascontiguousarray(coset_g, dtype=np.uint64)
        self._ck(self.L.cocg_ntt(self.h, arr, len(vecs), log_n, root.ctypes.data, 1 if inverse else 0,
                                 None if cg is None else cg.ctypes.data))

    # ---------------- MSM (a1)
    def bases_upload(self, group: int, pts: np.ndarray, mont: bool = True, stride: int | None = None) -> int:
        pts = np.ascontiguousarray(pts)
        pb = 8 * self.lq * 2 * group
        if stride is None:
            stride = pb
        n = pts.nbytes // stride
        hdl = ctypes.c_uint64()
        self._ck(self.L.cocg_bases_upload(self.h, group, pts.ctypes.data, n, stride, 1 if mont else 0, ctypes.byref(hdl)))
        self._groups = getattr(self, "_groups", {})
        self._groups[hdl.value] = group
        return hdl.value

    def bases_generate(self, group: int, n: int, seed: bytes) -> int:
        """n synthetic curve points P0 + i*Q generated in HBM (csrc/gen.cu)."""
        assert len(seed) == 32
        buf = np.frombuffer(seed, dtype=np.uint8).copy()
        hdl = ctypes.c_uint64()
        self._ck(self.L.cocg_bases_generate(self.h, group, n, buf.ctypes.data, ctypes.byref(hdl)))
        self._groups = getattr(self, "_groups", {})
        self._groups[hdl.value] = group
        return hdl.value

    def bases_check(self, handle: int, subgroup: bool = True):
        """-> (number of points off the curve / outside the prime-order subgroup, smallest such index or None)"""
        bad, first = ctypes.c_size_t(), ctypes.c_size_t()
        self._ck(self.L.cocg_bases_check(self.h, handle, 1 if subgroup else 0, ctypes.byref(bad), ctypes.byref(first)))
        return int(bad.value), (int(first.value) if bad.value else None)

    def bases_download(self, handle: int, off: int, n: int) -> np.ndarray:
        group = self._groups[handle]
        out = np.zeros((n, 2 * group * self.lq), dtype=np.uint64)
        self._ck(self.L.cocg_bases_download(self.h, handle, off, n, out.ctypes.data))
        return out

    def profile(self, on: bool = True):
        self._ck(self.L.cocg_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self.L.cocg_profile_reset(self.h))

    def profile_read(self) -> dict:
        names = ["msm_sort", "msm_accumulate", "msm_reduce", "ntt", "vec", "spmv"]
        out = {}
        for i, name in enumerate(names):
            ms, k = ctypes.c_double(), ctypes.c_uint64()
            self._ck(self.L.cocg_profile_read(self.h, i, ctypes.byref(ms), ctypes.byref(k)))
            out[name] = (ms.value, int(k.value))
        return out

    def fp_mul_ceiling(self, base_field: bool = True) -> float:
        """10^9 Montgomery products / s of a pure fp_mul chain on every SM, measured now (Fq of the curve, or Fr)."""
        out = ctypes.c_double()
        self._ck(self.L.cocg_fp_mul_ceiling(self.h, 1 if base_field else 0, ctypes.byref(out)))
        return out.value

    def bases_free(self, handle: int):
        self._ck(self.L.cocg_bases_free(self.h, handle))

    def msm(self, bases: int, scalars, off: int = 0, n: int | None = None, mont: bool = True) -> np.ndarray:
        """scalars: list of DeviceVec (one per share component) -> (k, 3*group*LQ) Jacobian limbs."""
        scalars = list(scalars)
        group = self._groups[bases]
        if n is None:
            n = scalars[0].n
        arr = (ctypes.c_void_p * len(scalars))(*[s.ptr for s in scalars])
        out = np.zeros((len(scalars), 3 * group * self.lq), dtype=np.uint64)
        self._ck(self.L.cocg_msm(self.h, bases, off, n, arr, len(scalars), 1 if mont else 0, out.ctypes.data))
        return out

    def msm_multi(self, bases, offs, scalars, n: int | None = None, mont: bool = True):
        """Several queries times the same scalars (one shared digit sort): list of (k, 3*group*LQ) Jacobian arrays."""
        scalars = list(scalars)
        if n is None:
            n = scalars[0].n
        nq, k = len(bases), len(scalars)
        outs = [np.zeros((k, 3 * self._groups[b] * self.lq), dtype=np.uint64) for b in bases]
        B = (ctypes.c_uint64 * nq)(*bases)
        O = (ctypes.c_size_t * nq)(*offs)
        S = (ctypes.c_void_p * k)(*[s.ptr for s in scalars])
        P = (ctypes.c_void_p * nq)(*[o.ctypes.data for o in outs])
        self._ck(self.L.cocg_msm_multi(self.h, B, O, nq, n, S, k, 1 if mont else 0, P))
        return outs

    def msm_host(self, bases: int, scalars, off: int = 0, n: int | None = None, mont: bool = True) -> np.ndarray:
        """Same with host numpy scalar arrays (the drop-in call of a host-resident caller)."""
        scalars = [np.ascontiguousarray(s, dtype=np.uint64) for s in scalars]
        group = self._groups[bases]
        if n is None:
            n = scalars[0].shape[0]
        arr = (ctypes.c_void_p * len(scalars))(*[s.ctypes.data for s in scalars])
        out = np.zeros((len(scalars), 3 * group * self.lq), dtype=np.uint64)
        self._ck(self.L.cocg_msm_host(self.h, bases, off, n, arr, len(scalars), 1 if mont else 0, out.ctypes.data))
        return out

    # ---------------- SpMV (a8)
    def csr_upload(self, rowptr: np.ndarray, col: np.ndarray, coeff: np.ndarray) -> int:
        rowptr = np.ascontiguousarray(rowptr, dtype=np.uint32)
        col = np.ascontiguousarray(col, dtype=np.uint32)
        coeff = np.ascontiguousarray(coeff, dtype=np.uint64)
        hdl = ctypes.c_uint64()
        self._ck(self.L.cocg_csr_upload(self.h, rowptr.ctypes.data, col.ctypes.data, coeff.ctypes.data,
                                        rowptr.shape[0] - 1, col.shape[0], ctypes.byref(hdl)))
        self._csr_rows = getattr(self, "_csr_rows", {})
        self._csr_rows[hdl.value] = rowptr.shape[0] - 1
        return hdl.value

    def spmv(self, csr: int, z_pub: DeviceVec | None, npub: int, z_wit: DeviceVec, out: DeviceVec | None = None) -> DeviceVec:
        out = out or DeviceVec(self, self._csr_rows[csr])
        self._ck(self.L.cocg_spmv(self.h, csr, z_pub.ptr if z_pub else None, npub, z_wit.ptr, out.ptr))
        return out

    # ---------------- O(1) group ops (K7)
    def ec_op(self, group: int, op: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros((2 if op == _lib.EC_TO_AFFINE else 3) * group * self.lq, dtype=np.uint64)
        self._ck(self.L.cocg_ec_op(self.h, group, op, a.ctypes.data, None if bb is None else bb.ctypes.data, out.ctypes.data))
        return out
