// Fused point-wise kernels of CoPlonk rounds 2, 3 and 5 (kernel K9), replacing the element loops of
//   compute_z    /root/reference/co-circom/co-plonk/src/round2.rs:146-206  (numerator / denominator factors of the permutation argument)
//   compute_t    /root/reference/co-circom/co-plonk/src/round3.rs:237-471  (gate, permutation and L_1 parts of the quotient on the 4n domain,
//                                                                          mul4vec! / mul4vec_post! :17-71, division by Z_H :438-446)
//   compute_r / compute_wxi  round5.rs:140-330                              (linear combinations of coefficient vectors)
// over share vectors of K components (K = 1: PlainDriver, K = 2: Rep3Protocol, components a | b).
//
// MPC structure.  The reference evaluates compute_t with 52 mul_vec calls, one network round each.  Every product there is one of
//   level 1:  x * y           of wire / permutation evaluations and the blinding polynomials,
//   level 2:  (x * y) * (u * v) of two level-1 results,
// and everything else is linear.  A REP3 product is local up to a re-sharing of ONE additive value per element, and sums of products
// may be re-shared together, so the quotient needs exactly two exchanges: after `plonk_quotient_l1_kernel` (6 vectors of 4n) and after
// `plonk_quotient_l2_kernel` (2 vectors: t and tz).  Products of two BLINDING polynomials (ap*bp, cp*zp, cp*zwp) are polynomials in w
// whose coefficients are products of the shared blinders b_i -- ten scalar multiplications done once on the host driver.
// With Z = Z_H(w^i) in {0, zeta_1, zeta_2, zeta_3} (get_z1..3: z1[m] = zeta_m, z2[m] = zeta_m^2, z3[m] = zeta_m^3) the five outputs of
// mul4vec!, combined by mul4vec_post!, collapse to
//   post = a0 + zeta a1 + zeta^2 a2 + zeta^3 a3 = (X(zeta) Y(zeta) - X0 Y0) / zeta,   X(s) = X0 + s X1 + s^2 X2,  Y likewise,
// with X0 = A B, X1 = Ap B + A Bp, X2 = Ap Bp, Y0 = C D, Y1 = Cp D + C Dp, Y2 = Cp Dp: one product instead of fifteen.
// All of it is exact arithmetic in Fr, so the opened t(X) is the reference's polynomial bit for bit.
#include <string.h>

#include "ctx.cuh"
#include "prf.cuh"

namespace cocg {

// The fused kernels below contain 30-60 field products and up to 20 ChaCha12 blocks per element.  Inlined, that is several hundred KB of
// SASS -- far beyond the instruction cache (measured: 14 ms per 2^20-element launch, ~30x the multiplier bound).  The product and the PRF
// are therefore CALLED (operands and result in registers), as the Fq2 product of the G2 group law is (ec.cuh).
template <class P>
__device__ __noinline__ Fp<P> fmul(Fp<P> a, Fp<P> b) { return fp_mul(a, b); }
template <class P>
__device__ __noinline__ Fp<P> zero_mask(const PrfKey& own, const PrfKey& prev, uint32_t ctr, size_t i) {
  return fp_sub(prf_field<P>(own, ctr, i), prf_field<P>(prev, ctr, i));
}

template <class P, int K>
struct Sh {
  Fp<P> v[K];
};
template <class P, int K>
__device__ __forceinline__ Sh<P, K> sh_load(const void* const* ptr, size_t i) {
  Sh<P, K> r;
#pragma unroll
  for (int k = 0; k < K; k++) r.v[k] = load_fp<P>(ptr[k], i);
  return r;
}
template <class P, int K>
__device__ __forceinline__ Sh<P, K> sh_add(const Sh<P, K>& a, const Sh<P, K>& b) {
  Sh<P, K> r;
#pragma unroll
  for (int k = 0; k < K; k++) r.v[k] = fp_add(a.v[k], b.v[k]);
  return r;
}
template <class P, int K>
__device__ __forceinline__ Sh<P, K> sh_scale(const Sh<P, K>& a, const Fp<P>& f) {  // mul_with_public
  Sh<P, K> r;
#pragma unroll
  for (int k = 0; k < K; k++) r.v[k] = fmul<P>(a.v[k], f);
  return r;
}
// add_with_public (rep3.rs:600-608): the constant enters component `pub_comp` only (party 0: a, party 1: b, party 2: none)
template <class P, int K>
__device__ __forceinline__ Sh<P, K> sh_add_pub(const Sh<P, K>& a, const Fp<P>& f, int pub_comp) {
  Sh<P, K> r = a;
#pragma unroll
  for (int k = 0; k < K; k++)
    if (k == pub_comp) r.v[k] = fp_add(r.v[k], f);
  return r;
}
// The party's additive share of x * y: plain x*y; REP3 x.a*y.a + x.a*y.b + x.b*y.a (rep3.rs:656-660) as two products
template <class P, int K>
__device__ __forceinline__ Fp<P> sh_lmul(const Sh<P, K>& x, const Sh<P, K>& y) {
  if (K == 1) return fmul<P>(x.v[0], y.v[0]);
  return fp_add(fmul<P>(x.v[0], fp_add(y.v[0], y.v[K - 1])), fmul<P>(x.v[K - 1], y.v[0]));
}

// ---------------------------------------------------------------------------------------------- round 2: factors of z
struct ZFactorArgs {
  const void* a; const void* b; const void* c;   // wire buffers, one share component, n elements
  const void* s1; const void* s2; const void* s3; // sigma evaluations on the 4n domain (element 4 i is used)
  const void* wpow;                               // omega^i, i < n
  void* out[6];                                   // n1 n2 n3 d1 d2 d3
};
template <class P>
__global__ void __launch_bounds__(256) plonk_z_factors_kernel(ZFactorArgs g, size_t n, Fp<P> beta, Fp<P> gamma, Fp<P> k1, Fp<P> k2, int add_public) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> a = load_fp<P>(g.a, i), b = load_fp<P>(g.b, i), c = load_fp<P>(g.c, i);
    if (add_public) {
      Fp<P> betaw = fp_mul(beta, load_fp_ro<P>(g.wpow, i));
      store_fp<P>(g.out[0], i, fp_add(fp_add(a, betaw), gamma));
      store_fp<P>(g.out[1], i, fp_add(fp_add(b, fp_mul(k1, betaw)), gamma));
      store_fp<P>(g.out[2], i, fp_add(fp_add(c, fp_mul(k2, betaw)), gamma));
      store_fp<P>(g.out[3], i, fp_add(fp_add(a, fp_mul(beta, load_fp_ro<P>(g.s1, 4 * i))), gamma));
      store_fp<P>(g.out[4], i, fp_add(fp_add(b, fp_mul(beta, load_fp_ro<P>(g.s2, 4 * i))), gamma));
      store_fp<P>(g.out[5], i, fp_add(fp_add(c, fp_mul(beta, load_fp_ro<P>(g.s3, 4 * i))), gamma));
    } else {
      store_fp<P>(g.out[0], i, a); store_fp<P>(g.out[1], i, b); store_fp<P>(g.out[2], i, c);
      store_fp<P>(g.out[3], i, a); store_fp<P>(g.out[4], i, b); store_fp<P>(g.out[5], i, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------- round 3
constexpr int kQuotL1Outs = 6;
struct QuotArgs {  // mirrors cocg_plonk_quotient_args (include/cocg.h), device pointers resolved
  const void* ea[2]; const void* eb[2]; const void* ec[2]; const void* ez[2];
  const void* s1; const void* s2; const void* s3;
  const void* qm; const void* ql; const void* qr; const void* qo; const void* qc;
  const void* lagrange;      // n_lagrange polynomials of 4n evaluations, contiguous
  const void* wpow;          // w^i, i < 4n (w = generator of the 4n domain)
  const void* buf_a[2];      // wire buffer a (first n_lagrange elements are the public-input gates)
  const void* l1[2];         // level-1 results: 6 vectors of 4n, contiguous, per component
  void* out;                 // l1: 6 x 4n; l2: 2 x 4n (t | tz)
};
template <class P>
struct QuotScalars {
  Fp<P> beta, gamma, alpha, alpha2, k1, k2, omega;  // omega = generator of the n domain
  Fp<P> zeta[4], zeta_inv[4];
  Fp<P> b[9][2];    // blinders b0..b8, per component
  Fp<P> sp[10][2];  // b1b3 b0b3 b1b2 b0b2 b5b8 b5b7 b5b6 b4b8 b4b7 b4b6, per component
};

template <class P, int K>
struct Blind {
  Sh<P, K> ap, bp, cp, zp, zwp;
};
template <class P, int K>
__device__ __forceinline__ Sh<P, K> blinder(const QuotScalars<P>& s, int idx) {
  Sh<P, K> r;
#pragma unroll
  for (int k = 0; k < K; k++) r.v[k] = s.b[idx][k];
  return r;
}
template <class P, int K>
__device__ __forceinline__ Sh<P, K> sprod(const QuotScalars<P>& s, int idx) {
  Sh<P, K> r;
#pragma unroll
  for (int k = 0; k < K; k++) r.v[k] = s.sp[idx][k];
  return r;
}
// ap = b1 + b0 w, bp = b3 + b2 w, cp = b5 + b4 w, zp = b8 + b7 w + b6 w^2, zwp = the same at w * omega  (round3.rs:258-262, 318-333)
template <class P, int K>
__device__ __forceinline__ Blind<P, K> blinding_evals(const QuotScalars<P>& s, const Fp<P>& w) {
  Blind<P, K> o;
  o.ap = sh_add(blinder<P, K>(s, 1), sh_scale(blinder<P, K>(s, 0), w));
  o.bp = sh_add(blinder<P, K>(s, 3), sh_scale(blinder<P, K>(s, 2), w));
  o.cp = sh_add(blinder<P, K>(s, 5), sh_scale(blinder<P, K>(s, 4), w));
  const Fp<P> w2 = fmul<P>(w, w);
  o.zp = sh_add(blinder<P, K>(s, 8), sh_add(sh_scale(blinder<P, K>(s, 7), w), sh_scale(blinder<P, K>(s, 6), w2)));
  const Fp<P> ww = fmul<P>(w, s.omega), ww2 = fmul<P>(ww, ww);
  o.zwp = sh_add(blinder<P, K>(s, 8), sh_add(sh_scale(blinder<P, K>(s, 7), ww), sh_scale(blinder<P, K>(s, 6), ww2)));
  return o;
}

template <class P, int K>
__global__ void __launch_bounds__(128) plonk_quotient_l1_kernel(QuotArgs g, QuotScalars<P> s, size_t n4, int pub_comp, PrfKey own, PrfKey prev, uint32_t ctr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const Fp<P> w = load_fp_ro<P>(g.wpow, i);
  const Sh<P, K> a = sh_load<P, K>(g.ea, i), b = sh_load<P, K>(g.eb, i), c = sh_load<P, K>(g.ec, i), z = sh_load<P, K>(g.ez, i);
  const Sh<P, K> zw = sh_load<P, K>(g.ez, (i + 4) & (n4 - 1));
  const Blind<P, K> bl = blinding_evals<P, K>(s, w);
  // The wire factors of the permutation products differ from a, b, c by PUBLIC addends only (A = a + (beta w + gamma), ...), so
  // A*B = a*b + cb*a + ca*b + ca*cb and Ap*B + A*Bp = (ap*b + a*bp) + cb*ap + ca*bp are linear in the two gate products below:
  // six products are re-shared, not ten (level 2 rebuilds X0, X1, Y0, Y1 of both permutation terms from them).
  Fp<P> o[kQuotL1Outs];
  o[0] = sh_lmul(a, b);                                    // a*b
  o[1] = fp_add(sh_lmul(a, bl.bp), sh_lmul(bl.ap, b));     // a*bp + ap*b
  o[2] = sh_lmul(c, z);                                    // c*z
  o[3] = fp_add(sh_lmul(bl.cp, z), sh_lmul(c, bl.zp));     // cp*z + c*zp
  o[4] = sh_lmul(c, zw);                                   // c*zw
  o[5] = fp_add(sh_lmul(bl.cp, zw), sh_lmul(c, bl.zwp));   // cp*zw + c*zwp
#pragma unroll 1
  for (int j = 0; j < kQuotL1Outs; j++) {  // rolled: one copy of the PRF call in the instruction stream
    Fp<P> v = o[j];
    if (K == 2) v = fp_add(v, zero_mask<P>(own, prev, ctr + j, i));
    store_fp<P>(g.out, (size_t)j * n4 + i, v);
  }
}

template <class P, int K>
__global__ void __launch_bounds__(128) plonk_quotient_l2_kernel(QuotArgs g, QuotScalars<P> s, size_t n4, int n_lagrange, int n_public, int pub_comp, PrfKey own,
                                                                 PrfKey prev, uint32_t ctr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int m = (int)(i & 3);
  const bool pub0 = pub_comp == 0;  // the additive output carries public constants for the plain driver and REP3 party 0 only
  const Fp<P> w = load_fp_ro<P>(g.wpow, i), w2 = fmul<P>(w, w), w3 = fmul<P>(w2, w);
  const Blind<P, K> bl = blinding_evals<P, K>(s, w);
  auto L1 = [&](int j) {
    Sh<P, K> r;
#pragma unroll
    for (int k = 0; k < K; k++) r.v[k] = load_fp<P>(g.l1[k], (size_t)j * n4 + i);
    return r;
  };
  // products of two blinding polynomials from the shared scalar products
  const Sh<P, K> X2 = sh_add(sprod<P, K>(s, 0), sh_add(sh_scale(sh_add(sprod<P, K>(s, 1), sprod<P, K>(s, 2)), w), sh_scale(sprod<P, K>(s, 3), w2)));
  const Fp<P> om = s.omega, om2 = fmul<P>(om, om);
  const Sh<P, K> Y2 = sh_add(sh_add(sprod<P, K>(s, 4), sh_scale(sh_add(sprod<P, K>(s, 5), sprod<P, K>(s, 7)), w)),
                             sh_add(sh_scale(sh_add(sprod<P, K>(s, 6), sprod<P, K>(s, 8)), w2), sh_scale(sprod<P, K>(s, 9), w3)));
  const Sh<P, K> Y2w = sh_add(sh_add(sprod<P, K>(s, 4), sh_scale(sh_add(sh_scale(sprod<P, K>(s, 5), om), sprod<P, K>(s, 7)), w)),
                              sh_add(sh_scale(sh_add(sh_scale(sprod<P, K>(s, 6), om2), sh_scale(sprod<P, K>(s, 8), om)), w2),
                                     sh_scale(sprod<P, K>(s, 9), fmul<P>(om2, w3))));
  const Fp<P> qm = load_fp_ro<P>(g.qm, i), ql = load_fp_ro<P>(g.ql, i), qr = load_fp_ro<P>(g.qr, i), qo = load_fp_ro<P>(g.qo, i);
  const Fp<P> l0 = load_fp_ro<P>(g.lagrange, i);
  const Fp<P> a2l0 = fmul<P>(s.alpha2, l0);
  // ---- linear parts, component a only (the party's own additive share)
  Fp<P> t, tz;
  {
    const Fp<P> a = load_fp<P>(g.ea[0], i), b = load_fp<P>(g.eb[0], i), c = load_fp<P>(g.ec[0], i), z = load_fp<P>(g.ez[0], i);
    const Fp<P> p1 = load_fp<P>(g.l1[0], i), p23 = load_fp<P>(g.l1[0], n4 + i);
    t = fp_add(fp_add(fmul<P>(qm, p1), fmul<P>(ql, a)), fp_add(fmul<P>(qr, b), fmul<P>(qo, c)));
    for (int j = 0; j < n_public; j++)  // pi = - sum_j L_j(w^i) * buffer_a[j]   (round3.rs:367-373)
      t = fp_sub(t, fmul<P>(load_fp_ro<P>(g.lagrange, (size_t)j * n4 + i), load_fp<P>(g.buf_a[0], j)));
    t = fp_add(t, fmul<P>(a2l0, z));
    if (pub0) t = fp_add(t, fp_sub(load_fp_ro<P>(g.qc, i), a2l0));  // + qc - alpha^2 L_1
    Fp<P> a0 = p23;
    if (m) a0 = fp_add(a0, fmul<P>(s.zeta[m], X2.v[0]));
    tz = fp_add(fp_add(fmul<P>(qm, a0), fmul<P>(ql, bl.ap.v[0])), fp_add(fmul<P>(qr, bl.bp.v[0]), fmul<P>(qo, bl.cp.v[0])));
    tz = fp_add(tz, fmul<P>(a2l0, bl.zp.v[0]));
  }
  // ---- bilinear parts
  Fp<P> e23 = Fp<P>::zero(), e23z = Fp<P>::zero();
  const Sh<P, K> av = sh_load<P, K>(g.ea, i), bv = sh_load<P, K>(g.eb, i);
  const Sh<P, K> zv = sh_load<P, K>(g.ez, i), zwv = sh_load<P, K>(g.ez, (i + 4) & (n4 - 1));
  const Sh<P, K> P1 = L1(0), P23 = L1(1);
  const Fp<P> betaw = fmul<P>(s.beta, w);
#pragma unroll
  for (int grp = 0; grp < 2; grp++) {
    // public addends of the three wire factors of this permutation term (round3.rs:397-409)
    Fp<P> ca, cb, cc;
    if (grp == 0) {
      ca = fp_add(betaw, s.gamma);
      cb = fp_add(fmul<P>(betaw, s.k1), s.gamma);
      cc = fp_add(fmul<P>(betaw, s.k2), s.gamma);
    } else {
      ca = fp_add(fmul<P>(load_fp_ro<P>(g.s1, i), s.beta), s.gamma);
      cb = fp_add(fmul<P>(load_fp_ro<P>(g.s2, i), s.beta), s.gamma);
      cc = fp_add(fmul<P>(load_fp_ro<P>(g.s3, i), s.beta), s.gamma);
    }
    const Sh<P, K>& D = grp ? zwv : zv;
    const Sh<P, K>& Dp = grp ? bl.zwp : bl.zp;
    const Sh<P, K> X0 = sh_add_pub(sh_add(P1, sh_add(sh_scale(av, cb), sh_scale(bv, ca))), fmul<P>(ca, cb), pub_comp);
    const Sh<P, K> X1 = sh_add(P23, sh_add(sh_scale(bl.ap, cb), sh_scale(bl.bp, ca)));
    const Sh<P, K> Y0 = sh_add(L1(2 + 2 * grp), sh_scale(D, cc));
    const Sh<P, K> Y1 = sh_add(L1(3 + 2 * grp), sh_scale(Dp, cc));
    const Sh<P, K>& Y2g = grp ? Y2w : Y2;
    const Fp<P> r = sh_lmul(X0, Y0);
    Fp<P> post;
    if (m == 0) {
      post = fp_add(sh_lmul(X1, Y0), sh_lmul(X0, Y1));
    } else {
      const Fp<P> ze = s.zeta[m];
      const Sh<P, K> Xz = sh_add(X0, sh_scale(sh_add(X1, sh_scale(X2, ze)), ze));
      const Sh<P, K> Yz = sh_add(Y0, sh_scale(sh_add(Y1, sh_scale(Y2g, ze)), ze));
      post = fmul<P>(fp_sub(sh_lmul(Xz, Yz), r), s.zeta_inv[m]);
    }
    if (grp == 0) { e23 = r; e23z = post; }
    else { e23 = fp_sub(e23, r); e23z = fp_sub(e23z, post); }
  }
  t = fp_add(t, fmul<P>(s.alpha, e23));
  tz = fp_add(tz, fmul<P>(s.alpha, e23z));
  if (K == 2) {
    t = fp_add(t, zero_mask<P>(own, prev, ctr, i));
    tz = fp_add(tz, zero_mask<P>(own, prev, ctr + 1, i));
  }
  store_fp<P>(g.out, i, t);
  store_fp<P>(g.out, n4 + i, tz);
  (void)n_lagrange;
}

// coefficients of T: negate the first n, divide by Z_H = X^n - 1 in coefficient form (c[i] = c[i - n] - c[i], round3.rs:434-446), add the
// blinding part, cut into t1 | t2 | t3 (n, n, n + 6 coefficients; the b9 / b10 patches are applied by the host driver)
struct TFinishArgs {
  const void* ct; const void* ctz;  // 4n each
  void* t1; void* t2; void* t3;     // n + 1, n + 1, n + 6
};
template <class P>
__global__ void __launch_bounds__(256) plonk_t_finish_kernel(TFinishArgs g, size_t n) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Fp<P> c0 = fp_neg(load_fp<P>(g.ct, j));
  Fp<P> c1 = fp_sub(c0, load_fp<P>(g.ct, n + j));
  Fp<P> c2 = fp_sub(c1, load_fp<P>(g.ct, 2 * n + j));
  Fp<P> c3 = fp_sub(c2, load_fp<P>(g.ct, 3 * n + j));
  store_fp<P>(g.t1, j, fp_add(c0, load_fp<P>(g.ctz, j)));
  store_fp<P>(g.t2, j, fp_add(c1, load_fp<P>(g.ctz, n + j)));
  store_fp<P>(g.t3, j, fp_add(c2, load_fp<P>(g.ctz, 2 * n + j)));
  if (j < 6) store_fp<P>(g.t3, n + j, fp_add(c3, load_fp<P>(g.ctz, 3 * n + j)));
}

// ---------------------------------------------------------------------------------------------- linear combinations (round 5)
constexpr int kMaxLin = 8;
template <class P>
struct LinArgs {
  const void* v[kMaxLin];
  size_t len[kMaxLin];
  Fp<P> f[kMaxLin];
  int nv;
};
template <class P>
__global__ void __launch_bounds__(256) vec_lincomb_kernel(LinArgs<P> g, void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> acc = Fp<P>::zero();
    for (int k = 0; k < g.nv; k++)
      if (i < g.len[k]) acc = fp_add(acc, fp_mul(g.f[k], load_fp<P>(g.v[k], i)));
    store_fp<P>(out, i, acc);
  }
}

template <class P>
__global__ void __launch_bounds__(256) vec_fill_kernel(void* __restrict__ out, size_t n, Fp<P> v) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store_fp<P>(out, i, v);
}

template <class P>
static Fp<P> host_fr(const void* p) {
  Fp<P> r;
  memcpy(r.l, p, 32);
  return r;
}

template <class P>
static int z_factors_impl(cocg_ctx* ctx, const cocg_plonk_z_args* a, size_t n, int add_public) {
  using F = Fp<P>;
  if (n == 0) return 0;
  ZFactorArgs g;
  g.a = a->a; g.b = a->b; g.c = a->c; g.s1 = a->sigma1; g.s2 = a->sigma2; g.s3 = a->sigma3;
  for (int k = 0; k < 6; k++) g.out[k] = a->out[k];
  F one = F::one(), om = host_fr<P>(a->omega);
  void* wp = nullptr;
  COCG_TRY(powers_table(ctx, /*kind=*/4, n, om.l, one.l, &wp));
  g.wpow = wp;
  ProfScope prof(ctx, COCG_PROF_VEC);
  plonk_z_factors_kernel<P><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(g, n, host_fr<P>(a->beta), host_fr<P>(a->gamma), host_fr<P>(a->k1), host_fr<P>(a->k2), add_public);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int quotient_impl(cocg_ctx* ctx, int level, const cocg_plonk_quotient_args* a) {
  using F = Fp<P>;
  const size_t n4 = a->n4;
  if (n4 < 4 || (n4 & (n4 - 1))) return fail(ctx, "cocg_plonk_quotient: the extended domain size must be a power of two >= 4");
  const int K = a->components;
  if (K != 1 && K != 2) return fail(ctx, "cocg_plonk_quotient: components must be 1 or 2");
  QuotArgs g;
  memset(&g, 0, sizeof(g));
  for (int k = 0; k < K; k++) {
    g.ea[k] = a->eval_a[k]; g.eb[k] = a->eval_b[k]; g.ec[k] = a->eval_c[k]; g.ez[k] = a->eval_z[k];
    g.buf_a[k] = a->buffer_a[k]; g.l1[k] = a->level1[k];
  }
  g.s1 = a->sigma1; g.s2 = a->sigma2; g.s3 = a->sigma3;
  g.qm = a->qm; g.ql = a->ql; g.qr = a->qr; g.qo = a->qo; g.qc = a->qc;
  g.lagrange = a->lagrange;
  g.out = a->out;
  QuotScalars<P> s;
  s.beta = host_fr<P>(a->beta); s.gamma = host_fr<P>(a->gamma); s.alpha = host_fr<P>(a->alpha); s.alpha2 = fp_sqr(s.alpha);
  s.k1 = host_fr<P>(a->k1); s.k2 = host_fr<P>(a->k2); s.omega = host_fr<P>(a->omega_n);
  // zeta_m = i4^m - 1 where i4 = omega_4n^n is the primitive 4th root (Domains::root_of_unity_2, types.rs:93): Z_H on the 4n domain
  F w4n = host_fr<P>(a->omega_4n);
  F i4 = fp_pow_u64(w4n, (uint64_t)(n4 / 4));
  F pw = F::one();
  for (int m = 0; m < 4; m++) {
    s.zeta[m] = fp_sub(pw, F::one());
    s.zeta_inv[m] = m ? fp_inv(s.zeta[m]) : F::zero();
    pw = fp_mul(pw, i4);
  }
  for (int j = 0; j < 9; j++)
    for (int k = 0; k < 2; k++) s.b[j][k] = k < K ? host_fr<P>((const char*)a->blinders + (size_t)(j * 2 + k) * 32) : F::zero();
  for (int j = 0; j < 10; j++)
    for (int k = 0; k < 2; k++) s.sp[j][k] = (k < K && a->scalar_products) ? host_fr<P>((const char*)a->scalar_products + (size_t)(j * 2 + k) * 32) : F::zero();
  F one = F::one();
  void* wp = nullptr;
  COCG_TRY(powers_table(ctx, /*kind=*/4, n4, w4n.l, one.l, &wp));
  g.wpow = wp;
  PrfKey own, prev;
  memset(&own, 0, sizeof(own));
  memset(&prev, 0, sizeof(prev));
  if (K == 2) {
    if (!a->seed_own || !a->seed_prev) return fail(ctx, "cocg_plonk_quotient: REP3 needs both PRF seeds");
    memcpy(own.k, a->seed_own, 32);
    memcpy(prev.k, a->seed_prev, 32);
  }
  const unsigned grid = (unsigned)((n4 + 127) / 128);
  ProfScope prof(ctx, COCG_PROF_VEC);
  if (level == 1) {
    if (K == 1) plonk_quotient_l1_kernel<P, 1><<<grid, 128, 0, ctx->stream>>>(g, s, n4, a->pub_comp, own, prev, a->ctr);
    else plonk_quotient_l1_kernel<P, 2><<<grid, 128, 0, ctx->stream>>>(g, s, n4, a->pub_comp, own, prev, a->ctr);
  } else {
    if (!a->scalar_products) return fail(ctx, "cocg_plonk_quotient_l2: the blinder products are missing");
    if (K == 1) plonk_quotient_l2_kernel<P, 1><<<grid, 128, 0, ctx->stream>>>(g, s, n4, (int)a->n_lagrange, (int)a->n_public, a->pub_comp, own, prev, a->ctr);
    else plonk_quotient_l2_kernel<P, 2><<<grid, 128, 0, ctx->stream>>>(g, s, n4, (int)a->n_lagrange, (int)a->n_public, a->pub_comp, own, prev, a->ctr);
  }
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int t_finish_impl(cocg_ctx* ctx, const void* ct, const void* ctz, size_t n, void* t1, void* t2, void* t3) {
  if (n < 8) return fail(ctx, "cocg_plonk_t_finish: domain too small (snarkjs keeps n >= 8)");
  TFinishArgs g{ct, ctz, t1, t2, t3};
  ProfScope prof(ctx, COCG_PROF_VEC);
  plonk_t_finish_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(g, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int lincomb_impl(cocg_ctx* ctx, int nv, const void* const* vecs, const size_t* lens, const void* factors, void* out, size_t n) {
  if (n == 0) return 0;
  LinArgs<P> g;
  g.nv = nv;
  for (int k = 0; k < nv; k++) { g.v[k] = vecs[k]; g.len[k] = lens[k]; g.f[k] = host_fr<P>((const char*)factors + (size_t)k * 32); }
  ProfScope prof(ctx, COCG_PROF_VEC);
  vec_lincomb_kernel<P><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(g, out, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int fill_impl(cocg_ctx* ctx, void* out, size_t n, const void* value) {
  if (n == 0) return 0;
  vec_fill_kernel<P><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(out, n, host_fr<P>(value));
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace cocg

using namespace cocg;

extern "C" int cocg_vec_fill(cocg_ctx* ctx, void* out, size_t n, const void* value) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!value || (n && !out)) return fail(ctx, "cocg_vec_fill: null argument");
  return COCG_FR_DISPATCH(ctx, fill_impl, ctx, out, n, value);
}
extern "C" int cocg_plonk_z_factors(cocg_ctx* ctx, const cocg_plonk_z_args* args, size_t n, int add_public) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!args) return fail(ctx, "cocg_plonk_z_factors: null argument");
  return COCG_FR_DISPATCH(ctx, z_factors_impl, ctx, args, n, add_public);
}
extern "C" int cocg_plonk_quotient_l1(cocg_ctx* ctx, const cocg_plonk_quotient_args* args) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!args) return fail(ctx, "cocg_plonk_quotient_l1: null argument");
  return COCG_FR_DISPATCH(ctx, quotient_impl, ctx, 1, args);
}
extern "C" int cocg_plonk_quotient_l2(cocg_ctx* ctx, const cocg_plonk_quotient_args* args) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!args) return fail(ctx, "cocg_plonk_quotient_l2: null argument");
  return COCG_FR_DISPATCH(ctx, quotient_impl, ctx, 2, args);
}
extern "C" int cocg_plonk_t_finish(cocg_ctx* ctx, const void* ct, const void* ctz, size_t n, void* t1, void* t2, void* t3) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ct || !ctz || !t1 || !t2 || !t3) return fail(ctx, "cocg_plonk_t_finish: null argument");
  return COCG_FR_DISPATCH(ctx, t_finish_impl, ctx, ct, ctz, n, t1, t2, t3);
}
extern "C" int cocg_vec_lincomb(cocg_ctx* ctx, int nv, const void* const* vecs, const size_t* lens, const void* factors, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (nv < 0 || nv > kMaxLin) return fail(ctx, "cocg_vec_lincomb: at most 8 vectors");
  if (n && (!out || (nv && (!vecs || !lens || !factors)))) return fail(ctx, "cocg_vec_lincomb: null argument");
  return COCG_FR_DISPATCH(ctx, lincomb_impl, ctx, nv, vecs, lens, factors, out, n);
}
