"""GPU: argument / error behaviour of the C ABI and the newer entry points (cocg_vec_axpy, cocg_csr_upload_form, cocg_csr_download,
cocg_bases_generate_range, cocg_msm with sub-ranges of generated bases, profiling counters).  The reference's compute methods are
infallible (mpc-core/src/traits.rs:535-568); the ABI reports misuse with a non-zero return + message instead of aborting."""
import ctypes

import numpy as np
import pytest

from oracle import cref
from oracle.curves import BN254, BLS12_381

pytestmark = pytest.mark.gpu


def rand_fr(n, seed):
    a = np.random.default_rng(seed).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


@pytest.mark.parametrize("curve", [BN254, BLS12_381], ids=["bn254", "bls12_381"])
def test_vec_axpy(cocg, bn, bls, curve):
    ctx = bn if curve is BN254 else bls
    n = 4099
    x, y = rand_fr(n, 1), rand_fr(n, 2)
    a = cref.fr_to_mont(curve, [curve.r - 5])
    dx, dy = ctx.upload(x), ctx.upload(y)
    ax = cref.fr_vec_op(curve, cref.OP_MUL, x, np.repeat(a, n, axis=0))
    assert np.array_equal(ctx.vec_axpy(a, dx).to_host(), ax)
    assert np.array_equal(ctx.vec_axpy(a, dx, dy).to_host(), cref.fr_vec_op(curve, cref.OP_ADD, ax, y))
    ctx.vec_axpy(a, dx, dy, out=dy)  # in place
    assert np.array_equal(dy.to_host(), cref.fr_vec_op(curve, cref.OP_ADD, ax, y))


def test_csr_forms_and_download(cocg, bn):
    c = BN254
    L = cocg.load()
    rowptr = np.array([0, 2, 2, 5], dtype=np.uint32)
    col = np.array([0, 3, 1, 2, 4], dtype=np.uint32)
    vals = [1, c.r - 1, 7, 12345678901234567890, 3]
    z = rand_fr(5, 9)
    want = cref.spmv(c, rowptr, col, cref.fr_to_mont(c, vals), z)
    for form, coeff in ((0, cref.fr_to_mont(c, vals)), (1, cref.ints_to_limbs([v * c.Rr * c.Rr % c.r for v in vals], 4)),
                        (2, cref.ints_to_limbs(vals, 4))):
        h = ctypes.c_uint64()
        coeff = np.ascontiguousarray(coeff)
        assert L.cocg_csr_upload_form(bn.h, rowptr.ctypes.data, col.ctypes.data, coeff.ctypes.data, 3, 5, form, ctypes.byref(h)) == 0
        bn._csr_rows = getattr(bn, "_csr_rows", {})
        bn._csr_rows[h.value] = 3
        got = bn.spmv(h.value, None, 0, bn.upload(z)).to_host()
        assert np.array_equal(got, want), form
        rp, cl, cf, nnz = np.zeros(4, np.uint32), np.zeros(5, np.uint32), np.zeros((5, 4), np.uint64), ctypes.c_size_t()
        assert L.cocg_csr_download(bn.h, h.value, rp.ctypes.data, cl.ctypes.data, cf.ctypes.data, ctypes.byref(nnz)) == 0
        assert nnz.value == 5 and list(rp) == list(rowptr) and list(cl) == list(col) and cref.fr_from_mont(c, cf) == vals
        assert L.cocg_csr_free(bn.h, h.value) == 0
    h = ctypes.c_uint64()
    assert L.cocg_csr_upload_form(bn.h, rowptr.ctypes.data, col.ctypes.data, coeff.ctypes.data, 3, 5, 7, ctypes.byref(h)) != 0
    assert b"form" in L.cocg_last_error(bn.h)
    bad = np.array([0, 2, 2, 4], dtype=np.uint32)  # rowptr[rows] != nnz
    assert L.cocg_csr_upload(bn.h, bad.ctypes.data, col.ctypes.data, coeff.ctypes.data, 3, 5, ctypes.byref(h)) != 0


def test_generated_bases_are_on_curve_distinct_and_range_consistent(cocg, bn):
    c = BN254
    seed = bytes(range(32))
    h = bn.bases_generate(1, 1000, seed)
    pts = cref.g_from_mont(c, bn.bases_download(h, 0, 1000), 1)
    assert all(p is not None and c.is_on_curve(p, 1) for p in pts[:50]) and len(set(pts)) == 1000
    # P_i = P_0 + i Q
    q = c.add(pts[1], c.neg(pts[0], 1), 1)
    assert c.add(pts[7], q, 1) == pts[8]
    # the range variant produces the same sequence
    L = cocg.load()
    h2 = ctypes.c_uint64()
    sd = np.frombuffer(seed, dtype=np.uint8).copy()
    assert L.cocg_bases_generate_range(bn.h, 1, 300, 200, sd.ctypes.data, ctypes.byref(h2)) == 0
    bn._groups[h2.value] = 1
    assert cref.g_from_mont(c, bn.bases_download(h2.value, 0, 200), 1) == pts[300:500]
    # MSM over a slice of the big table == MSM over the separately generated slice
    s = rand_fr(200, 4)
    a = bn.msm(h, [bn.upload(s)], off=300, n=200)[0]
    b = bn.msm(h2.value, [bn.upload(s)])[0]
    assert cref.jac_from_mont(c, a, 1) == cref.jac_from_mont(c, b, 1)
    g2 = bn.bases_generate(2, 64, seed)
    p2 = cref.g_from_mont(c, bn.bases_download(g2, 0, 64), 2)
    assert all(c.is_on_curve(p, 2) for p in p2[:8])
    for hh in (h, h2.value, g2):
        bn.bases_free(hh)


def test_error_paths_do_not_abort(cocg, bn):
    L = cocg.load()
    v = bn.zeros(8)
    out = np.zeros(12, dtype=np.uint64)
    sc = (ctypes.c_void_p * 1)(v.ptr)
    assert L.cocg_msm(bn.h, 9999, 0, 8, sc, 1, 1, out.ctypes.data) != 0 and b"handle" in L.cocg_last_error(bn.h)
    h = bn.bases_generate(1, 8, bytes(32))
    assert L.cocg_msm(bn.h, h, 4, 8, sc, 1, 1, out.ctypes.data) != 0 and b"range" in L.cocg_last_error(bn.h)
    assert L.cocg_msm(bn.h, h, 0, 8, None, 1, 1, out.ctypes.data) != 0
    assert L.cocg_bases_free(bn.h, h) == 0 and L.cocg_bases_free(bn.h, h) != 0  # double free is reported, not fatal
    assert L.cocg_ntt(bn.h, None, 1, 4, None, 0, None) != 0
    assert L.cocg_vec_op(bn.h, 99, v.ptr, v.ptr, v.ptr, 8) != 0
    with pytest.raises(cocg.CocgError):
        cocg.Context(7, 0)  # unknown curve
    with pytest.raises(cocg.CocgError):
        cocg.Context(cocg.BN254, 99)  # no such device
    # the context is still usable afterwards
    assert np.array_equal(bn.vec_op(cocg.OP_ADD, v, v).to_host(), np.zeros((8, 4), dtype=np.uint64))


def test_profile_counters(cocg, bn):
    bn.profile(True)
    bn.profile_reset()
    v = bn.upload(rand_fr(1000, 3))
    bn.vec_op(cocg.OP_MUL, v, v)
    bn.vec_op(cocg.OP_ADD, v, v)
    p = bn.profile_read()
    bn.profile(False)
    assert p["vec"][1] == 2 and p["vec"][0] > 0 and p["ntt"][1] == 0
    n0 = bn.launch_count()
    bn.vec_op(cocg.OP_SUB, v, v)
    assert bn.launch_count() == n0 + 1
