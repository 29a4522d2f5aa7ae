// Prime-field arithmetic for the cocg kernels: N little-endian 32-bit limbs, Montgomery form, R = 2^(32N).
//
// This is the arithmetic the reference gets from arkworks' `Fp<MontBackend<_, N>>` (ark-ff 0.4.2, pinned by
// /root/reference/Cargo.toml:33-40, not vendored): the in-memory value is the same Montgomery residue, so a
// `[u64; 4]` / `[u64; 6]` handed over by a Rust caller is bit-identical to our `uint32_t[8]` / `[12]`.
//
// Device path: the product is accumulated in two limb arrays -- one aligned on column 0, one on column 1 --
// so that every 32x32 product a_j*s lands on an aligned (lo,hi) register pair and each row is ONE carry chain
// per array (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32.X).  The Montgomery shift by
// one limb per row is done by swapping the roles of the two arrays instead of moving registers.
// Host path (used by the C-ABI for O(1) bookkeeping such as twiddle seeds): portable 64-bit CIOS.
#pragma once
#include <stdint.h>

#include "params_gen.h"

#if defined(__CUDACC__)
#define COCG_HD __host__ __device__ __forceinline__
#define COCG_D __device__ __forceinline__
#else
#define COCG_HD inline
#define COCG_D inline
#endif

namespace cocg {

// ---------------------------------------------------------------------------------------------------------
// Field parameter packs.  Constants come from tools/gen_params.py (params_gen.h); they are exposed through
// constexpr accessors so that fully unrolled device code sees them as immediates.
// ---------------------------------------------------------------------------------------------------------
#define COCG_DEFINE_FIELD(NAME, PFX)                                                                       \
  struct NAME {                                                                                            \
    static constexpr int N = PFX##_LIMBS;                                                                  \
    static constexpr int BITS = PFX##_BITS;                                                                \
    static constexpr uint32_t INV = PFX##_INV;                                                             \
    static COCG_HD constexpr uint32_t mod(int i) {                                                         \
      constexpr uint32_t v[PFX##_LIMBS] = PFX##_MOD;                                                       \
      return v[i];                                                                                         \
    }                                                                                                      \
    static COCG_HD constexpr uint32_t r1(int i) {                                                          \
      constexpr uint32_t v[PFX##_LIMBS] = PFX##_R1;                                                        \
      return v[i];                                                                                         \
    }                                                                                                      \
    static COCG_HD constexpr uint32_t r2(int i) {                                                          \
      constexpr uint32_t v[PFX##_LIMBS] = PFX##_R2;                                                        \
      return v[i];                                                                                         \
    }                                                                                                      \
  };

COCG_DEFINE_FIELD(Bn254FrP, BN254_FR)
COCG_DEFINE_FIELD(Bn254FqP, BN254_FQ)
COCG_DEFINE_FIELD(Bls381FrP, BLS381_FR)
COCG_DEFINE_FIELD(Bls381FqP, BLS381_FQ)

// ---------------------------------------------------------------------------------------------------------
// PTX carry-chain primitives (device only).
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
namespace ptx {
COCG_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
COCG_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {  // one IMAD.WIDE.U32 instead of a mul.lo / mul.hi pair
  uint64_t r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  lo = (uint32_t)r;
  hi = (uint32_t)(r >> 32);
}
COCG_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
COCG_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
COCG_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
}  // namespace ptx
#endif

template <class P>
struct Fp {
  static constexpr int N = P::N;
  using Params = P;
  alignas(16) uint32_t l[N];  // device structs are moved with 128-bit accesses; host callers' buffers go through memcpy

  static COCG_HD Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = 0;
    return r;
  }
  static COCG_HD Fp one() {  // Montgomery form of 1
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = P::r1(i);
    return r;
  }
  static COCG_HD Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
    return r;
  }
  static COCG_HD Fp modulus() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = P::mod(i);
    return r;
  }
  COCG_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i];
    return o == 0;
  }
  COCG_HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
    return o == 0;
  }
  COCG_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// ------------------------------------------------------------------ reduce x in [0, 2p) to [0, p)
template <class P>
COCG_HD void fp_cond_sub(uint32_t* x) {
  constexpr int N = P::N;
  uint32_t t[N];
#if defined(__CUDA_ARCH__)
  t[0] = ptx::sub_cc(x[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(x[i], P::mod(i));
  uint32_t borrow = ptx::subc(0, 0);  // 0 if x >= p, 0xffffffff otherwise
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = borrow ? x[i] : t[i];
#else
  uint64_t br = 0;
  for (int i = 0; i < N; i++) {
    uint64_t d = (uint64_t)x[i] - P::mod(i) - br;
    t[i] = (uint32_t)d;
    br = (d >> 63) & 1;
  }
  if (!br)
    for (int i = 0; i < N; i++) x[i] = t[i];
#endif
}

template <class P>
COCG_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
#if defined(__CUDA_ARCH__)
  r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
  r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);  // p < 2^(32N-1): a+b < 2^(32N), no carry out
#else
  uint64_t c = 0;
  for (int i = 0; i < N; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    r.l[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  fp_cond_sub<P>(r.l);
  return r;
}

template <class P>
COCG_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
#if defined(__CUDA_ARCH__)
  r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
  uint32_t mask = ptx::subc(0, 0);  // all-ones when a < b
  r.l[0] = ptx::add_cc(r.l[0], P::mod(0) & mask);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], P::mod(i) & mask);
  r.l[N - 1] = ptx::addc(r.l[N - 1], P::mod(N - 1) & mask);
#else
  uint64_t br = 0;
  for (int i = 0; i < N; i++) {
    uint64_t d = (uint64_t)a.l[i] - b.l[i] - br;
    r.l[i] = (uint32_t)d;
    br = (d >> 63) & 1;
  }
  if (br) {
    uint64_t c = 0;
    for (int i = 0; i < N; i++) {
      c += (uint64_t)r.l[i] + P::mod(i);
      r.l[i] = (uint32_t)c;
      c >>= 32;
    }
  }
#endif
  return r;
}

template <class P>
COCG_HD Fp<P> fp_neg(const Fp<P>& a) {
  return fp_sub(Fp<P>::zero(), a);
}
template <class P>
COCG_HD Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add(a, a);
}

// ------------------------------------------------------------------ Montgomery product a*b*R^-1 mod p
#if defined(__CUDA_ARCH__)
namespace detail {
// c0: limb array aligned on column 0 (c0[0..N-1], c0[N] = spill-over limb, c0[N+1] == 0)
// c1: limb array aligned on column 1 (c1[0..N-1] hold columns 1..N)
// Adds v*s to the pair: even limbs of v go to c0, odd limbs to c1, one carry chain each.
template <class P, bool MODULUS>
COCG_D void row_add(uint32_t* c0, uint32_t* c1, const uint32_t* v, uint32_t s) {
  constexpr int N = P::N;
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    uint32_t vj = MODULUS ? P::mod(j) : v[j];
    c0[j] = (j == 0) ? ptx::mad_lo_cc(vj, s, c0[j]) : ptx::madc_lo_cc(vj, s, c0[j]);
    c0[j + 1] = ptx::madc_hi_cc(vj, s, c0[j + 1]);
  }
  c0[N] = ptx::addc(c0[N], 0);
#pragma unroll
  for (int j = 1; j < N; j += 2) {
    uint32_t vj = MODULUS ? P::mod(j) : v[j];
    c1[j - 1] = (j == 1) ? ptx::mad_lo_cc(vj, s, c1[j - 1]) : ptx::madc_lo_cc(vj, s, c1[j - 1]);
    c1[j] = ptx::madc_hi_cc(vj, s, c1[j]);
  }
  // no carry out: c1 * 2^32 <= total < 2^(32(N+1))
}
// One CIOS iteration after the first.  On entry c0[0] == 0 (previous reduction); the value is
// (c0 >> 32) + c1.  On exit the roles are swapped: c1 is the column-0 array, c0 the column-1 array.
template <class P>
COCG_D void mul_step(uint32_t* c0, uint32_t* c1, const uint32_t* a, uint32_t bi) {
  constexpr int N = P::N;
  // fold c0[1] (new column 0) into c1[0]; its carry enters the new column-1 chain below
  c1[0] = ptx::add_cc(c1[0], c0[1]);
#pragma unroll
  for (int j = 1; j < N; j += 2) {  // new column-1 array = old c0 shifted down two limbs, plus odd products
    c0[j - 1] = ptx::madc_lo_cc(a[j], bi, c0[j + 1]);
    c0[j] = ptx::madc_hi_cc(a[j], bi, c0[j + 2]);
  }
  c0[N] = 0;
#pragma unroll
  for (int j = 0; j < N; j += 2) {  // new column-0 array = old c1 plus even products
    c1[j] = (j == 0) ? ptx::mad_lo_cc(a[j], bi, c1[j]) : ptx::madc_lo_cc(a[j], bi, c1[j]);
    c1[j + 1] = ptx::madc_hi_cc(a[j], bi, c1[j + 1]);
  }
  c1[N] = ptx::addc(0, 0);
  uint32_t m = c1[0] * P::INV;
  row_add<P, true>(c1, c0, nullptr, m);
}
}  // namespace detail
#endif

template <class P>
COCG_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  static_assert(N % 2 == 0, "limb count must be even");
  Fp<P> r;
#if defined(__CUDA_ARCH__)
  uint32_t e[N + 2], o[N + 2];
  // row 0: disjoint (lo,hi) pairs, no chain needed
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    ptx::mul_wide(e[j], e[j + 1], a.l[j], b.l[0]);
    ptx::mul_wide(o[j], o[j + 1], a.l[j + 1], b.l[0]);
  }
  e[N] = 0; e[N + 1] = 0; o[N] = 0; o[N + 1] = 0;
  {
    uint32_t m = e[0] * P::INV;
    detail::row_add<P, true>(e, o, nullptr, m);
  }
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    detail::mul_step<P>(e, o, a.l, b.l[i]);               // column-0 array is now o
    if (i + 1 < N) detail::mul_step<P>(o, e, a.l, b.l[i + 1]);  // and e again
  }
  // N-1 swaps (odd): column-0 array is o (o[0] == 0), column-1 array is e.  result = (o >> 32) + e
  r.l[0] = ptx::add_cc(e[0], o[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(e[k], o[k + 1]);
  r.l[N - 1] = ptx::addc(e[N - 1], o[N]);
#else
  // host: 64-bit limbs (the uint32 array is little-endian, so limb pairs are u64 limbs on x86/aarch64-LE)
  constexpr int M = N / 2;
  uint64_t A[M], B[M], Q[M], t[M + 2];
  for (int i = 0; i < M; i++) {
    A[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
    B[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
    Q[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
  }
  // -p^-1 mod 2^64 from the 32-bit constant by one Newton step
  uint64_t inv64 = (uint64_t)P::INV;
  inv64 = inv64 * (2 + Q[0] * inv64);  // x' = x(2 + q x) for x = -q^-1
  for (int i = 0; i < M + 2; i++) t[i] = 0;
  for (int i = 0; i < M; i++) {
    unsigned __int128 c = 0;
    for (int j = 0; j < M; j++) {
      c += (unsigned __int128)A[j] * B[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[M];
    t[M] = (uint64_t)c;
    t[M + 1] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * inv64;
    c = ((unsigned __int128)m * Q[0] + t[0]) >> 64;
    for (int j = 1; j < M; j++) {
      c += (unsigned __int128)m * Q[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[M];
    t[M - 1] = (uint64_t)c;
    t[M] = t[M + 1] + (uint64_t)(c >> 64);
  }
  for (int i = 0; i < M; i++) {
    r.l[2 * i] = (uint32_t)t[i];
    r.l[2 * i + 1] = (uint32_t)(t[i] >> 32);
  }
#endif
  fp_cond_sub<P>(r.l);
  return r;
}

template <class P>
COCG_HD Fp<P> fp_sqr(const Fp<P>& a) {
  return fp_mul(a, a);
}

template <class P>
COCG_HD Fp<P> fp_to_mont(const Fp<P>& a) {
  return fp_mul(a, Fp<P>::r2());
}
template <class P>
COCG_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = Fp<P>::zero();
  o.l[0] = 1;
  return fp_mul(a, o);
}

// a^e for a small public exponent (twiddle seeds, chunk starts)
template <class P>
COCG_HD Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
  Fp<P> r = Fp<P>::one(), b = a;
  while (e) {
    if (e & 1) r = fp_mul(r, b);
    b = fp_sqr(b);
    e >>= 1;
  }
  return r;
}

// a^(p-2): Fermat inverse (0 -> 0).  Only used O(1) times per call on the path.
template <class P>
COCG_HD Fp<P> fp_inv(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t e[N];
  {
    uint64_t br = 2;  // p - 2
    for (int i = 0; i < N; i++) {
      uint64_t d = (uint64_t)P::mod(i) - br;
      e[i] = (uint32_t)d;
      br = (d >> 63) & 1;
    }
  }
  Fp<P> r = Fp<P>::one();
  for (int i = 32 * N - 1; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
  }
  return r;
}

using Bn254Fr = Fp<Bn254FrP>;
using Bn254Fq = Fp<Bn254FqP>;
using Bls381Fr = Fp<Bls381FrP>;
using Bls381Fq = Fp<Bls381FqP>;

}  // namespace cocg
