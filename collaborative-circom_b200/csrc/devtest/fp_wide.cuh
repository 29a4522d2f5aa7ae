// Separated multiply / reduce for the prime fields of fp.cuh (device only): a 2N-limb product accumulated in two column-aligned
// limb arrays, then a word-serial Montgomery reduction over the same arrays.  Same IMAD count as the interleaved CIOS of fp_mul
// (N^2 + N^2 + N), but the split admits
//   * squaring with N(N+1)/2 products (cross products once, doubled),
//   * Karatsuba on the product (3 half-size products; the extra additions run on the ALU pipe, which idles while the fmaheavy
//     pipe -- IMAD.WIDE occupies it 4 cycles per warp -- is the bottleneck of every kernel on this path),
//   * lazy reduction in Fq2: three unreduced products, two reductions instead of three.
// Representation: a value V = sum_k e[k] 2^(32k) + sum_k o[k] 2^(32(k+1)); a 32x32 product whose low half belongs to column c goes
// to (e[c], e[c+1]) when c is even and to (o[c-1], o[c]) when c is odd, so every product lands on an aligned (lo, hi) register
// pair and a row of products of one parity is a single carry chain (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X).
#pragma once
#include "../fp.cuh"

#if defined(__CUDA_ARCH__)
namespace cocg {
namespace wide {

// acc arrays have 2N + 2 entries; all indices are compile-time constants after unrolling
template <int N>
struct Acc {
  uint32_t e[2 * N + 2], o[2 * N + 2];
  COCG_D void clear() {
#pragma unroll
    for (int k = 0; k < 2 * N + 2; k++) { e[k] = 0; o[k] = 0; }
  }
};

// acc += (v[0..M) * s) << (32 * col0), M even
template <int N, int M, class V>
COCG_D void row_mad(Acc<N>& A, const V& v, uint32_t s, int col0) {
  // parity of the first column decides which array takes the even-indexed limbs of v
  if ((col0 & 1) == 0) {
#pragma unroll
    for (int j = 0; j < M; j += 2) {
      A.e[col0 + j] = (j == 0) ? ptx::mad_lo_cc(v(j), s, A.e[col0 + j]) : ptx::madc_lo_cc(v(j), s, A.e[col0 + j]);
      A.e[col0 + j + 1] = ptx::madc_hi_cc(v(j), s, A.e[col0 + j + 1]);
    }
    A.e[col0 + M] = ptx::addc(A.e[col0 + M], 0);
#pragma unroll
    for (int j = 1; j < M; j += 2) {
      A.o[col0 + j - 1] = (j == 1) ? ptx::mad_lo_cc(v(j), s, A.o[col0 + j - 1]) : ptx::madc_lo_cc(v(j), s, A.o[col0 + j - 1]);
      A.o[col0 + j] = ptx::madc_hi_cc(v(j), s, A.o[col0 + j]);
    }
    A.o[col0 + M] = ptx::addc(A.o[col0 + M], 0);
  } else {
#pragma unroll
    for (int j = 0; j < M; j += 2) {
      A.o[col0 + j - 1] = (j == 0) ? ptx::mad_lo_cc(v(j), s, A.o[col0 + j - 1]) : ptx::madc_lo_cc(v(j), s, A.o[col0 + j - 1]);
      A.o[col0 + j] = ptx::madc_hi_cc(v(j), s, A.o[col0 + j]);
    }
    A.o[col0 + M - 1] = ptx::addc(A.o[col0 + M - 1], 0);
#pragma unroll
    for (int j = 1; j < M; j += 2) {
      A.e[col0 + j] = (j == 1) ? ptx::mad_lo_cc(v(j), s, A.e[col0 + j]) : ptx::madc_lo_cc(v(j), s, A.e[col0 + j]);
      A.e[col0 + j + 1] = ptx::madc_hi_cc(v(j), s, A.e[col0 + j + 1]);
    }
    A.e[col0 + M + 1] = ptx::addc(A.e[col0 + M + 1], 0);
  }
}

template <class P>
struct ModLimbs {
  COCG_D uint32_t operator()(int j) const { return P::mod(j); }
};
struct PtrLimbs {
  const uint32_t* p;
  COCG_D uint32_t operator()(int j) const { return p[j]; }
};

// collapse the pair of arrays into plain limbs t[0..2N)
template <int N>
COCG_D void collapse(const Acc<N>& A, uint32_t* t) {
  t[0] = A.e[0];
  t[1] = ptx::add_cc(A.e[1], A.o[0]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) t[k] = ptx::addc_cc(A.e[k], A.o[k - 1]);
  t[2 * N - 1] = ptx::addc(A.e[2 * N - 1], A.o[2 * N - 2]);
}

// A <- a * b (schoolbook, N^2 products)
template <class P>
COCG_D void mul_wide(Acc<P::N>& A, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  A.clear();
  PtrLimbs va{a};
#pragma unroll
  for (int i = 0; i < N; i++) row_mad<N, N>(A, va, b[i], i);
}

// A <- a * a: cross products a_i a_j (i < j) once, doubled, plus the squares on the diagonal
template <class P>
COCG_D void sqr_wide(Acc<P::N>& A, const uint32_t* a) {
  constexpr int N = P::N;
  A.clear();
  // rows i: a[i+1..N) * a[i] at column 2i + 1; row lengths N-1-i (made even by taking pairs; an odd leftover is a single product)
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    const int len = N - 1 - i;
    const int col0 = 2 * i + 1;
    // even part of the row
    if (len >= 2) {
      PtrLimbs v{a + i + 1};
      if ((len & 1) == 0) {
        // full even length
        switch (len) {
          case 2: row_mad<N, 2>(A, v, a[i], col0); break;
          case 4: row_mad<N, 4>(A, v, a[i], col0); break;
          case 6: row_mad<N, 6>(A, v, a[i], col0); break;
          case 8: row_mad<N, 8>(A, v, a[i], col0); break;
          case 10: row_mad<N, 10>(A, v, a[i], col0); break;
        }
      } else {
        switch (len - 1) {
          case 2: row_mad<N, 2>(A, v, a[i], col0); break;
          case 4: row_mad<N, 4>(A, v, a[i], col0); break;
          case 6: row_mad<N, 6>(A, v, a[i], col0); break;
          case 8: row_mad<N, 8>(A, v, a[i], col0); break;
          case 10: row_mad<N, 10>(A, v, a[i], col0); break;
        }
      }
    }
    if (len & 1) {  // the last product of the row: a[N-1] * a[i] at column i + N - 1
      const int c = i + N - 1;
      if ((c & 1) == 0) {
        A.e[c] = ptx::mad_lo_cc(a[N - 1], a[i], A.e[c]);
        A.e[c + 1] = ptx::madc_hi_cc(a[N - 1], a[i], A.e[c + 1]);
        A.e[c + 2] = ptx::addc(A.e[c + 2], 0);
      } else {
        A.o[c - 1] = ptx::mad_lo_cc(a[N - 1], a[i], A.o[c - 1]);
        A.o[c] = ptx::madc_hi_cc(a[N - 1], a[i], A.o[c]);
        A.o[c + 1] = ptx::addc(A.o[c + 1], 0);
      }
    }
  }
  // double: both arrays shift left by one bit (their sum doubles)
#pragma unroll
  for (int k = 2 * N; k >= 1; k--) {
    A.e[k] = __funnelshift_l(A.e[k - 1], A.e[k], 1);
    A.o[k] = __funnelshift_l(A.o[k - 1], A.o[k], 1);
  }
  A.e[0] <<= 1;
  A.o[0] <<= 1;
  // diagonal a_i^2 at column 2i (even): one chain through e
#pragma unroll
  for (int i = 0; i < N; i++) {
    A.e[2 * i] = (i == 0) ? ptx::mad_lo_cc(a[i], a[i], A.e[2 * i]) : ptx::madc_lo_cc(a[i], a[i], A.e[2 * i]);
    A.e[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], A.e[2 * i + 1]);
  }
  A.e[2 * N] = ptx::addc(A.e[2 * N], 0);
}

// Montgomery reduction of t[0..2N) (< p * 2^(32N)): r = t / 2^(32N) mod p, canonical.  The rows m_i * p go to a FRESH pair of
// arrays (so that the slot receiving a chain's final carry never holds more than earlier carries); column i is resolved exactly
// -- t[i] + e[i] + o[i-1] + carry from below -- before m_i is chosen.
template <class P>
COCG_D void redc(const uint32_t* t, uint32_t* r) {
  constexpr int N = P::N;
  ModLimbs<P> vp;
  Acc<N> B;
  B.clear();
  uint32_t carry = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint32_t lo = ptx::add_cc(t[i], carry);
    uint32_t hi = ptx::addc(0, 0);
    lo = ptx::add_cc(lo, B.e[i]);
    hi = ptx::addc(hi, 0);
    if (i > 0) {
      lo = ptx::add_cc(lo, B.o[i - 1]);
      hi = ptx::addc(hi, 0);
    }
    const uint32_t m = lo * P::INV;
    const uint32_t plo = m * P::mod(0);
    (void)ptx::add_cc(lo, plo);  // == 0 mod 2^32: only its carry matters
    carry = ptx::addc(hi, 0);
    B.e[i] = 0;                  // column i is consumed (accounted for in `carry`)
    if (i > 0) B.o[i - 1] = 0;
    row_mad<N, N>(B, vp, m, i);
    if ((i & 1) == 0) B.e[i] = 0; else B.o[i - 1] = 0;  // the slot that took lo(m * p_0): consumed as well
  }
  // r = columns N .. 2N-1 of t + B + carry
  uint32_t u[N];
  u[0] = ptx::add_cc(t[N], carry);
#pragma unroll
  for (int k = 1; k < N - 1; k++) u[k] = ptx::addc_cc(t[N + k], 0);
  u[N - 1] = ptx::addc(t[2 * N - 1], 0);
  u[0] = ptx::add_cc(u[0], B.e[N]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) u[k] = ptx::addc_cc(u[k], B.e[N + k]);
  u[N - 1] = ptx::addc(u[N - 1], B.e[2 * N - 1]);
  r[0] = ptx::add_cc(u[0], B.o[N - 1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r[k] = ptx::addc_cc(u[k], B.o[N + k - 1]);
  r[N - 1] = ptx::addc(u[N - 1], B.o[2 * N - 2]);
  fp_cond_sub<P>(r);
}

}  // namespace wide

template <class P>
COCG_D Fp<P> fp_sqr_wide(const Fp<P>& a) {
  wide::Acc<P::N> A;
  wide::sqr_wide<P>(A, a.l);
  uint32_t t[2 * P::N];
  wide::collapse<P::N>(A, t);
  Fp<P> r;
  wide::redc<P>(t, r.l);
  return r;
}
template <class P>
COCG_D Fp<P> fp_mul_wide(const Fp<P>& a, const Fp<P>& b) {
  wide::Acc<P::N> A;
  wide::mul_wide<P>(A, a.l, b.l);
  uint32_t t[2 * P::N];
  wide::collapse<P::N>(A, t);
  Fp<P> r;
  wide::redc<P>(t, r.l);
  return r;
}

}  // namespace cocg
#endif
