// Quadratic extension Fq2 = Fq[u]/(u^2+1) and short-Weierstrass (a = 0) group law for G1 (over Fq) and
// G2 (over Fq2), BN254 and BLS12-381.
//
// Semantics follow what the reference obtains from arkworks' `short_weierstrass::{Affine, Projective}`
// (ark-ec 0.4.2, not vendored) at its MSM call sites /root/reference/mpc-core/src/protocols/rep3.rs:934-947,
// shamir.rs:1027-1039, plain.rs:408-416: affine inputs, Jacobian (X, Y, Z) outputs, x = X/Z^2, y = Y/Z^3.
// Affine points are packed (x, y) with (0, 0) = point at infinity, the snarkjs zkey layout
// (/root/reference/co-circom/circom-types/src/traits.rs:107-155).
// Bucket accumulators use extended Jacobian "XYZZ" coordinates (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2, which make the mixed addition 8M + 2S; they are converted to Jacobian without an inversion.
#pragma once
#include "fp.cuh"

// Large group-law bodies are kept out of line on the device: they are called from several kernels per curve
// and group, and inlining every copy makes ptxas take tens of minutes.  The hot mixed addition stays inline.
#if defined(__CUDACC__)
#define COCG_EC_OUTLINE __host__ __device__ __noinline__
#else
#define COCG_EC_OUTLINE inline
#endif

namespace cocg {

// ------------------------------------------------------------------ uniform field interface
template <class P> COCG_HD Fp<P> f_add(const Fp<P>& a, const Fp<P>& b) { return fp_add(a, b); }
template <class P> COCG_HD Fp<P> f_sub(const Fp<P>& a, const Fp<P>& b) { return fp_sub(a, b); }
template <class P> COCG_HD Fp<P> f_mul(const Fp<P>& a, const Fp<P>& b) { return fp_mul(a, b); }
template <class P> COCG_HD Fp<P> f_sqr(const Fp<P>& a) { return fp_sqr(a); }
template <class P> COCG_HD Fp<P> f_dbl(const Fp<P>& a) { return fp_add(a, a); }
template <class P> COCG_HD Fp<P> f_neg(const Fp<P>& a) { return fp_neg(a); }
template <class P> COCG_HD Fp<P> f_inv(const Fp<P>& a) { return fp_inv(a); }

template <class P>
struct Fp2 {
  using Base = Fp<P>;
  Fp<P> c0, c1;
  static COCG_HD Fp2 zero() { return Fp2{Fp<P>::zero(), Fp<P>::zero()}; }
  static COCG_HD Fp2 one() { return Fp2{Fp<P>::one(), Fp<P>::zero()}; }
  COCG_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  COCG_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
};
template <class P> COCG_HD Fp2<P> f_add(const Fp2<P>& a, const Fp2<P>& b) { return Fp2<P>{fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
template <class P> COCG_HD Fp2<P> f_sub(const Fp2<P>& a, const Fp2<P>& b) { return Fp2<P>{fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
template <class P> COCG_HD Fp2<P> f_dbl(const Fp2<P>& a) { return Fp2<P>{fp_add(a.c0, a.c0), fp_add(a.c1, a.c1)}; }
template <class P> COCG_HD Fp2<P> f_neg(const Fp2<P>& a) { return Fp2<P>{fp_neg(a.c0), fp_neg(a.c1)}; }
template <class P>
COCG_HD Fp2<P> f2_mul_body(const Fp2<P>& a, const Fp2<P>& b) {  // Karatsuba, u^2 = -1
  Fp<P> v0 = fp_mul(a.c0, b.c0);
  Fp<P> v1 = fp_mul(a.c1, b.c1);
  Fp<P> s = fp_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
  return Fp2<P>{fp_sub(v0, v1), fp_sub(fp_sub(s, v0), v1)};
}
template <class P>
COCG_HD Fp2<P> f2_sqr_body(const Fp2<P>& a) {  // (a0+a1)(a0-a1) + 2 a0 a1 u
  Fp<P> t = fp_mul(a.c0, a.c1);
  Fp<P> r0 = fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
  return Fp2<P>{r0, fp_add(t, t)};
}
#if defined(__CUDA_ARCH__)
// On the device the Fq2 product and square are CALLED, operands and result by value (registers), not inlined: a fully inlined G2
// mixed addition is ~110 KB of SASS, more than the instruction cache holds (ncu: icc hit rate 80 %, `no_instruction` the second
// largest stall), while the three base-field products inside one call still interleave.
template <class P>
__device__ __noinline__ Fp2<P> f2_mul_call(Fp2<P> a, Fp2<P> b) { return f2_mul_body(a, b); }
template <class P>
__device__ __noinline__ Fp2<P> f2_sqr_call(Fp2<P> a) { return f2_sqr_body(a); }
template <class P> COCG_D Fp2<P> f_mul(const Fp2<P>& a, const Fp2<P>& b) { return f2_mul_call<P>(a, b); }
template <class P> COCG_D Fp2<P> f_sqr(const Fp2<P>& a) { return f2_sqr_call<P>(a); }
#else
template <class P> COCG_HD Fp2<P> f_mul(const Fp2<P>& a, const Fp2<P>& b) { return f2_mul_body(a, b); }
template <class P> COCG_HD Fp2<P> f_sqr(const Fp2<P>& a) { return f2_sqr_body(a); }
#endif

template <class P>
COCG_HD Fp2<P> f_inv(const Fp2<P>& a) {  // conj(a) / (a0^2 + a1^2)
  Fp<P> n = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
  return Fp2<P>{fp_mul(a.c0, n), fp_neg(fp_mul(a.c1, n))};
}

// ------------------------------------------------------------------ points
template <class F> struct Affine { F x, y; COCG_HD bool is_inf() const { return x.is_zero() && y.is_zero(); } };
template <class F> struct Jacobian { F x, y, z; COCG_HD bool is_inf() const { return z.is_zero(); } };
template <class F> struct XYZZ { F x, y, zz, zzz; COCG_HD bool is_inf() const { return zz.is_zero(); } };

template <class F> COCG_HD XYZZ<F> xyzz_inf() { return XYZZ<F>{F::zero(), F::zero(), F::zero(), F::zero()}; }
template <class F> COCG_HD Jacobian<F> jac_inf() { return Jacobian<F>{F::one(), F::one(), F::zero()}; }  // arkworks' zero(): (1, 1, 0)

template <class F>
COCG_HD XYZZ<F> xyzz_from_affine(const Affine<F>& p) {
  if (p.is_inf()) return xyzz_inf<F>();
  return XYZZ<F>{p.x, p.y, F::one(), F::one()};
}

// 2*(x, y) for an affine, finite point (mdbl-2008-s-1)
template <class F>
COCG_EC_OUTLINE XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
  F U = f_dbl(p.y);
  F V = f_sqr(U);
  F W = f_mul(U, V);
  F S = f_mul(p.x, V);
  F X2 = f_sqr(p.x);
  F M = f_add(f_dbl(X2), X2);
  F X3 = f_sub(f_sqr(M), f_dbl(S));
  F Y3 = f_sub(f_mul(M, f_sub(S, X3)), f_mul(W, p.y));
  return XYZZ<F>{X3, Y3, V, W};
}

// dbl-2008-s-1
template <class F>
COCG_EC_OUTLINE XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
  if (p.is_inf()) return p;
  F U = f_dbl(p.y);
  F V = f_sqr(U);
  F W = f_mul(U, V);
  F S = f_mul(p.x, V);
  F X2 = f_sqr(p.x);
  F M = f_add(f_dbl(X2), X2);
  F X3 = f_sub(f_sqr(M), f_dbl(S));
  F Y3 = f_sub(f_mul(M, f_sub(S, X3)), f_mul(W, p.y));
  return XYZZ<F>{X3, Y3, f_mul(V, p.zz), f_mul(W, p.zzz)};
}

// acc += q (q affine), madd-2008-s; 8M + 2S on the common path
template <class F>
COCG_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = XYZZ<F>{q.x, q.y, F::one(), F::one()}; return; }
  F U2 = f_mul(q.x, acc.zz);
  F S2 = f_mul(q.y, acc.zzz);
  F Pd = f_sub(U2, acc.x);
  F R = f_sub(S2, acc.y);
  if (Pd.is_zero()) {
    if (R.is_zero()) acc = xyzz_dbl_affine(q);
    else acc = xyzz_inf<F>();
    return;
  }
  F PP = f_sqr(Pd);
  F PPP = f_mul(Pd, PP);
  F Q = f_mul(acc.x, PP);
  F X3 = f_sub(f_sub(f_sqr(R), PPP), f_dbl(Q));
  F Y3 = f_sub(f_mul(R, f_sub(Q, X3)), f_mul(acc.y, PPP));
  acc.x = X3;
  acc.y = Y3;
  acc.zz = f_mul(acc.zz, PP);
  acc.zzz = f_mul(acc.zzz, PPP);
}

// acc += q, add-2008-s; 12M + 2S.  Inline variant for the bucket-reduction loops (the out-of-line one keeps its accumulator
// in local memory across the call).
template <class F>
COCG_HD void xyzz_add_inline(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = f_mul(acc.x, q.zz);
  F U2 = f_mul(q.x, acc.zz);
  F S1 = f_mul(acc.y, q.zzz);
  F S2 = f_mul(q.y, acc.zzz);
  F Pd = f_sub(U2, U1);
  F R = f_sub(S2, S1);
  if (Pd.is_zero()) {
    if (R.is_zero()) acc = xyzz_dbl(acc);
    else acc = xyzz_inf<F>();
    return;
  }
  F PP = f_sqr(Pd);
  F PPP = f_mul(Pd, PP);
  F Q = f_mul(U1, PP);
  F X3 = f_sub(f_sub(f_sqr(R), PPP), f_dbl(Q));
  F Y3 = f_sub(f_mul(R, f_sub(Q, X3)), f_mul(S1, PPP));
  acc.x = X3;
  acc.y = Y3;
  acc.zz = f_mul(f_mul(acc.zz, q.zz), PP);
  acc.zzz = f_mul(f_mul(acc.zzz, q.zzz), PPP);
}

#if defined(__CUDACC__)
// Two independent base-field products per CALL.  The bucket-reduction kernels (msm_impl.cuh) issue the 14 products of an addition over
// Fq as 7 of these: inlined, the addition is ~35 KB of SASS per call site and the marginal-sum kernel ran at a 60 % instruction-cache
// hit rate with `no_instruction` as its first stall reason (profiles/r02_ncu_msm_marginals_multi*); called one product at a time, a
// warp would have a single carry chain in flight.
template <class P>
struct FqPair {
  Fp<P> a, b;
};
template <class P>
__device__ __noinline__ FqPair<P> fp_mul2_call(Fp<P> a0, Fp<P> b0, Fp<P> a1, Fp<P> b1) {
  FqPair<P> r;
  r.a = fp_mul(a0, b0);
  r.b = fp_mul(a1, b1);
  return r;
}
#endif

// acc += q, add-2008-s; 12M + 2S
template <class F>
COCG_EC_OUTLINE void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = f_mul(acc.x, q.zz);
  F U2 = f_mul(q.x, acc.zz);
  F S1 = f_mul(acc.y, q.zzz);
  F S2 = f_mul(q.y, acc.zzz);
  F Pd = f_sub(U2, U1);
  F R = f_sub(S2, S1);
  if (Pd.is_zero()) {
    if (R.is_zero()) acc = xyzz_dbl(acc);
    else acc = xyzz_inf<F>();
    return;
  }
  F PP = f_sqr(Pd);
  F PPP = f_mul(Pd, PP);
  F Q = f_mul(U1, PP);
  F X3 = f_sub(f_sub(f_sqr(R), PPP), f_dbl(Q));
  F Y3 = f_sub(f_mul(R, f_sub(Q, X3)), f_mul(S1, PPP));
  acc.x = X3;
  acc.y = Y3;
  acc.zz = f_mul(f_mul(acc.zz, q.zz), PP);
  acc.zzz = f_mul(f_mul(acc.zzz, q.zzz), PPP);
}

template <class F>
COCG_HD XYZZ<F> xyzz_neg(const XYZZ<F>& p) {
  return XYZZ<F>{p.x, f_neg(p.y), p.zz, p.zzz};
}

// XYZZ -> Jacobian without inversion: Z = ZZZ, X' = X*ZZ^2, Y' = Y*ZZZ^2  (uses ZZ^3 == ZZZ^2)
template <class F>
COCG_HD Jacobian<F> xyzz_to_jacobian(const XYZZ<F>& p) {
  if (p.is_inf()) return jac_inf<F>();
  return Jacobian<F>{f_mul(p.x, f_sqr(p.zz)), f_mul(p.y, f_sqr(p.zzz)), p.zzz};
}
template <class F>
COCG_HD XYZZ<F> xyzz_from_jacobian(const Jacobian<F>& p) {
  if (p.is_inf()) return xyzz_inf<F>();
  F zz = f_sqr(p.z);
  return XYZZ<F>{p.x, p.y, zz, f_mul(zz, p.z)};
}

using Bn254Fq2 = Fp2<Bn254FqP>;
using Bls381Fq2 = Fp2<Bls381FqP>;

}  // namespace cocg
