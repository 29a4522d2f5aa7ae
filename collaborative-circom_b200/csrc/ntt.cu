// In-order radix-2 NTT / iNTT over Fr (SURVEY rows a4 + a5), replacing the reference's
// `domain.fft_in_place(&mut data.a)` / `ifft_in_place` per share component
// (/root/reference/mpc-core/src/protocols/rep3.rs:880-921, shamir.rs:826-863, plain.rs:369-400 -> ark-poly
// Radix2EvaluationDomain with the snarkjs generator, co-circom/co-groth16/src/groth16.rs:57-77) and the
// `distribute_powers_and_mul_by_const` coset scaling that brackets it (groth16.rs:177-199).
//
// Convention (must match ark-poly): forward out[i] = sum_j in[j] * w^(i*j), natural order in and out;
// inverse uses w^-1 and multiplies by n^-1.
//
// Structure: the decimation-in-frequency butterfly network (natural in, bit-reversed out) is cut into
// ceil(log n / 8) passes.  A pass runs `s` consecutive stages on tiles of 2^s "group" elements x C adjacent
// columns held in shared memory (<= 1024 elements = 32 KiB, split into two 16-byte planes so LDS.128/STS.128
// are conflict-free), one 256-thread CTA per tile:
//   * every pass but the last: group elements are Q = 2^(log n - k0 - s) apart in memory, the C columns are
//     adjacent, so global traffic is in C*32-byte runs;
//   * the last pass works on C contiguous chunks of 2^s elements chosen so that their bit-reversed
//     destinations are adjacent; it applies the bit reversal (and the optional n^-1 / coset post-scale) in its
//     store, which makes the transform in-order without a separate permutation pass.
// Twiddles come from one table w^k, k < n/2, cached per (n, w) in HBM and hot in L2.
// Data moves data -> scratch (pass 0) -> ... -> data (last pass): 32*n bytes read + written per pass.
#include <stdlib.h>
#include <string.h>

#include "ctx.cuh"

namespace cocg {

#ifndef COCG_NTT_CALL
#define COCG_NTT_CALL __forceinline__
#endif
#ifndef COCG_NTT_MIN_BLOCKS
#define COCG_NTT_MIN_BLOCKS 4
#endif
constexpr int kNttMaxStages = 8;    // stages per pass (COCG_NTT_MAX_STAGES overrides, <= 10)
constexpr int kNttTile = 1024;      // elements per CTA tile
constexpr int kNttThreads = 128;    // one thread per 8 tile elements: a radix-8 round keeps every thread busy
constexpr int kNttMaxLogC = 4;      // columns per tile: global runs of up to 16 x 32 B
constexpr int kNttMaxVecs = 8;

struct NttVecs {
  const void* in[kNttMaxVecs];
  void* out[kNttMaxVecs];
};

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

// The product is CALLED (operands and result in registers): a radix-8 round holds 12 of them, inlined they are ~50 KB of SASS per
// round shape and the instruction cache, not the multiplier, sets the pace (the same effect as in plonk.cu).
template <class P>
__device__ COCG_NTT_CALL Fp<P> ntt_fmul(Fp<P> a, Fp<P> b) { return fp_mul(a, b); }
// two independent products per call: the two carry chains interleave inside the callee (one chain per warp leaves the multiplier
// waiting on its own latency at the 12-16 warps per SM this kernel's register budget allows)
template <class P>
struct FpPair {
  Fp<P> a, b;
};
template <class P>
__device__ COCG_NTT_CALL FpPair<P> ntt_fmul2(Fp<P> a, Fp<P> wa, Fp<P> b, Fp<P> wb) {
  FpPair<P> r;
  r.a = fp_mul(a, wa);
  r.b = fp_mul(b, wb);
  return r;
}

template <class P>
__device__ __forceinline__ Fp<P> lds_fp(const uint4* plane_lo, const uint4* plane_hi, int m) {
  const uint4 lo = plane_lo[m], hi = plane_hi[m];
  Fp<P> x;
  x.l[0] = lo.x; x.l[1] = lo.y; x.l[2] = lo.z; x.l[3] = lo.w; x.l[4] = hi.x; x.l[5] = hi.y; x.l[6] = hi.z; x.l[7] = hi.w;
  return x;
}
template <class P>
__device__ __forceinline__ void sts_fp(uint4* plane_lo, uint4* plane_hi, int m, const Fp<P>& x) {
  plane_lo[m] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
  plane_hi[m] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
}

// One round = G consecutive butterfly stages (local stages kk .. kk+G-1 of the pass) on 2^G elements held in REGISTERS: the tile is
// read and written once per round instead of once per stage, one barrier per round, 2^G - 1 twiddle loads per 2^G elements.
// DIF stage: (x, y) -> (x + y, (x - y) * w^((e mod 2^logh) * Q' << k)).  The 2^G elements of a work item differ in bits
// d .. d+G-1 of the group index e (d = s - kk - G); stage t pairs elements whose j differ in bit G-1-t.
template <class P, int G, bool LAST>
__device__ __forceinline__ void ntt_round(uint4* plane_lo, uint4* plane_hi, int tile_elems, int s, int logC, int kk, int k0, int L, int logQ,
                                          uint32_t col0, const void* __restrict__ tw) {
  constexpr int R = 1 << G;
  const int C = 1 << logC;
  const int d = s - kk - G;
  const int items = tile_elems >> G;
  for (int it = threadIdx.x; it < items; it += kNttThreads) {
    const int c = it & (C - 1);
    const int p = it >> logC;
    const int low = p & ((1 << d) - 1);
    const int e_base = ((p >> d) << (d + G)) + low;
    Fp<P> x[R];
#pragma unroll
    for (int j = 0; j < R; j++) x[j] = lds_fp<P>(plane_lo, plane_hi, ((e_base + (j << d)) << logC) + c);
#pragma unroll
    for (int t = 0; t < G; t++) {
      const int k = k0 + kk + t;
      constexpr int kPairs = R / 2;
      const int half = 1 << (G - 1 - t);
      const bool unit = (k == L - 1);  // last global stage: every twiddle is 1
      // butterfly b of this stage: q = b mod half (position below the pairing bit -> its twiddle), elements j = (b / half) * 2 half + q, j + half
      Fp<P> wa, wb;           // the (at most two) twiddles in use; reloaded only when the pair of butterflies needs other ones
      int held_a = -1, held_b = -1;  // compile-time after unrolling
      auto twiddle = [&](int q) {
        const int lo_i = (q << d) + low;
        const size_t ex = LAST ? (size_t)lo_i : (((size_t)lo_i << logQ) + col0 + c);
        return load_fp_ro<P>(tw, ex << k);
      };
#pragma unroll
      for (int b = 0; b < kPairs; b += 2) {
        const int q0 = b & (half - 1), j0 = ((b & ~(half - 1)) << 1) + q0;
        if (!unit && held_a != q0) { wa = twiddle(q0); held_a = q0; }
        if (kPairs == 1) {
          const Fp<P> u = fp_add(x[j0], x[j0 + half]);
          Fp<P> v = fp_sub(x[j0], x[j0 + half]);
          if (!unit) v = ntt_fmul<P>(v, wa);
          x[j0] = u;
          x[j0 + half] = v;
        } else {
          const int b1 = b + 1, q1 = b1 & (half - 1), j1 = ((b1 & ~(half - 1)) << 1) + q1;
          if (!unit && q1 != q0 && held_b != q1) { wb = twiddle(q1); held_b = q1; }
          const Fp<P> u0 = fp_add(x[j0], x[j0 + half]), u1 = fp_add(x[j1], x[j1 + half]);
          Fp<P> v0 = fp_sub(x[j0], x[j0 + half]), v1 = fp_sub(x[j1], x[j1 + half]);
          if (!unit) {
            const FpPair<P> pr = ntt_fmul2<P>(v0, wa, v1, q1 != q0 ? wb : wa);
            v0 = pr.a;
            v1 = pr.b;
          }
          x[j0] = u0; x[j0 + half] = v0;
          x[j1] = u1; x[j1 + half] = v1;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; j++) sts_fp<P>(plane_lo, plane_hi, ((e_base + (j << d)) << logC) + c, x[j]);
  }
  __syncthreads();
}

template <class P, bool LAST>
__global__ void __launch_bounds__(kNttThreads, COCG_NTT_MIN_BLOCKS) ntt_pass_kernel(NttVecs vecs, int L, int k0, int s, int logC,
                                                                const void* __restrict__ tw, const void* __restrict__ pre,
                                                                const void* __restrict__ post, Fp<P> post_const, int use_post_const) {
  __shared__ uint4 plane_lo[kNttTile];
  __shared__ uint4 plane_hi[kNttTile];
  const int E = 1 << s, C = 1 << logC;
  const int tile_elems = E << logC;
  const uint32_t tile = blockIdx.x;
  const uint4* __restrict__ in = reinterpret_cast<const uint4*>(vecs.in[blockIdx.y]);
  uint4* __restrict__ out = reinterpret_cast<uint4*>(vecs.out[blockIdx.y]);

  // ---- tile geometry
  size_t base;          // global index of (e = 0, c = 0)
  size_t stride_e;      // global stride between group elements
  size_t stride_c;      // global stride between columns
  uint32_t col0 = 0;    // first column (non-last passes): enters the twiddle exponent
  const int logQ = L - k0 - s;
  if (!LAST) {
    uint32_t tiles_per_block = 1u << (logQ - logC);
    uint32_t hb = tile / tiles_per_block;
    col0 = (tile % tiles_per_block) << logC;
    base = ((size_t)hb << (L - k0)) + col0;
    stride_e = (size_t)1 << logQ;
    stride_c = 1;
  } else {
    // chunks tile + cidx * NT, NT = 2^(L - s - logC)
    base = (size_t)tile << s;
    stride_e = 1;
    stride_c = ((size_t)1 << (L - s - logC)) << s;
  }

  // ---- load (optionally pre-scaled by pre[global index])
  for (int t = threadIdx.x; t < tile_elems; t += kNttThreads) {
    int e, c;
    if (LAST) { e = t & (E - 1); c = t >> s; } else { c = t & (C - 1); e = t >> logC; }
    size_t g = base + (size_t)e * stride_e + (size_t)c * stride_c;
    uint4 lo = in[2 * g], hi = in[2 * g + 1];
    if (pre) {
      Fp<P> x;
      x.l[0] = lo.x; x.l[1] = lo.y; x.l[2] = lo.z; x.l[3] = lo.w; x.l[4] = hi.x; x.l[5] = hi.y; x.l[6] = hi.z; x.l[7] = hi.w;
      x = ntt_fmul<P>(x, load_fp_ro<P>(pre, g));
      lo = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
      hi = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
    }
    int m = (e << logC) + c;
    plane_lo[m] = lo;
    plane_hi[m] = hi;
  }
  __syncthreads();

  // ---- s butterfly stages in register rounds of 3 (2, 2 instead of 3, 1, so that no round is a lone stage unless s == 1).
  // Requesting a round's twiddle lines ahead of its barrier (prefetch.global.L1) was measured and made the pass slower
  // (0.42 -> 0.46 ms for two 2^20 vectors): the extra address arithmetic and spills cost more than the L2 latency they hide.
  int kk = 0;
  while (s - kk > 4 || s - kk == 3) {
    ntt_round<P, 3, LAST>(plane_lo, plane_hi, tile_elems, s, logC, kk, k0, L, logQ, col0, tw);
    kk += 3;
  }
  while (s - kk >= 2) {
    ntt_round<P, 2, LAST>(plane_lo, plane_hi, tile_elems, s, logC, kk, k0, L, logQ, col0, tw);
    kk += 2;
  }
  if (s - kk == 1) ntt_round<P, 1, LAST>(plane_lo, plane_hi, tile_elems, s, logC, kk, k0, L, logQ, col0, tw);

  // ---- store
  if (!LAST) {
    for (int t = threadIdx.x; t < tile_elems; t += kNttThreads) {
      int c = t & (C - 1), e = t >> logC;
      size_t g = base + (size_t)e * stride_e + c;
      out[2 * g] = plane_lo[t];
      out[2 * g + 1] = plane_hi[t];
    }
  } else {
    // element (e, cidx) sits at bit-reversed position j = chunk*E + e; its natural index is
    // i = bitrev_s(e) * 2^(L-s) + bitrev_{L-s-logC}(tile) * C + bitrev_logC(cidx).  Iterate with the
    // destination column fastest so that stores are C*32-byte runs.
    const size_t row = (size_t)1 << (L - s);
    const size_t colbase = (size_t)bitrev(tile, L - s - logC) << logC;
    for (int t = threadIdx.x; t < tile_elems; t += kNttThreads) {
      int q = t & (C - 1), r = t >> logC;
      int cidx = (int)bitrev((uint32_t)q, logC);
      int e = (int)bitrev((uint32_t)r, s);
      int m = (e << logC) + cidx;
      size_t i = (size_t)r * row + colbase + q;
      uint4 lo = plane_lo[m], hi = plane_hi[m];
      if (post || use_post_const) {
        Fp<P> x;
        x.l[0] = lo.x; x.l[1] = lo.y; x.l[2] = lo.z; x.l[3] = lo.w; x.l[4] = hi.x; x.l[5] = hi.y; x.l[6] = hi.z; x.l[7] = hi.w;
        x = ntt_fmul<P>(x, post ? load_fp_ro<P>(post, i) : post_const);
        lo = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
        hi = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
      }
      out[2 * i] = lo;
      out[2 * i + 1] = hi;
    }
  }
}

template <class P>
static int ntt_impl(cocg_ctx* ctx, void* const* vecs, int k, unsigned log_n, const void* root, int inverse, const void* coset_g) {
  using F = Fp<P>;
  const int L = (int)log_n;
  const size_t n = (size_t)1 << L;
  F w, g;
  memcpy(w.l, root, 32);
  if (coset_g) memcpy(g.l, coset_g, 32);
  F one = F::one();
  if (L == 0) {  // length-1 transform is the identity (n^-1 = 1, g^0 = 1)
    return 0;
  }
  if (inverse) w = fp_inv(w);
  // n^-1 in Montgomery form
  F ninv = one;
  if (inverse) {
    F nn = F::zero();
    nn.l[0] = (uint32_t)n;
    nn.l[1] = (uint32_t)((uint64_t)n >> 32);
    ninv = fp_inv(fp_to_mont(nn));
  }
  void* tw = nullptr;
  COCG_TRY(powers_table(ctx, /*kind=*/2, n / 2, w.l, one.l, &tw));
  void* pre = nullptr;
  void* post = nullptr;
  if (coset_g && !inverse) COCG_TRY(powers_table(ctx, 3, n, g.l, one.l, &pre));
  if (coset_g && inverse) COCG_TRY(powers_table(ctx, 3, n, g.l, ninv.l, &post));
  int use_post_const = (inverse && !coset_g) ? 1 : 0;

  // pass plan
  static const int max_stages = [] {
    const char* e = getenv("COCG_NTT_MAX_STAGES");
    int v = e ? atoi(e) : kNttMaxStages;
    return v < 1 ? 1 : v > 10 ? 10 : v;
  }();
  int np = (L + max_stages - 1) / max_stages;
  int sbase = L / np, extra = L % np;
  void* scratch = nullptr;
  if (np > 1) COCG_TRY(scratch_get(ctx, 0, (size_t)k * n * sizeof(F), &scratch));
  int k0 = 0;
  ProfScope prof(ctx, COCG_PROF_NTT);
  for (int p = 0; p < np; p++) {
    int s = sbase + (p < extra ? 1 : 0);
    bool last = (p == np - 1);
    int logC = 0;
    if (!last) { logC = L - k0 - s; if (logC > kNttMaxLogC) logC = kNttMaxLogC; }
    else { logC = L - s; if (logC > kNttMaxLogC) logC = kNttMaxLogC; }
    while ((1 << (s + logC)) > kNttTile) logC--;
    for (int v0 = 0; v0 < k; v0 += kNttMaxVecs) {
      int kv = k - v0 < kNttMaxVecs ? k - v0 : kNttMaxVecs;
      NttVecs nv;
      for (int v = 0; v < kv; v++) {
        char* data = (char*)vecs[v0 + v];
        char* scr = scratch ? (char*)scratch + (size_t)(v0 + v) * n * sizeof(F) : data;
        nv.in[v] = (p == 0) ? data : scr;
        nv.out[v] = last ? data : scr;
      }
      dim3 grid((unsigned)(n >> (s + logC)), (unsigned)kv);
      if (last)
        ntt_pass_kernel<P, true><<<grid, kNttThreads, 0, ctx->stream>>>(nv, L, k0, s, logC, tw, p == 0 ? pre : nullptr, post, ninv, use_post_const);
      else
        ntt_pass_kernel<P, false><<<grid, kNttThreads, 0, ctx->stream>>>(nv, L, k0, s, logC, tw, p == 0 ? pre : nullptr, nullptr, ninv, 0);
      COCG_LAUNCH_CHECK(ctx);
    }
    k0 += s;
  }
  return 0;
}

}  // namespace cocg

using namespace cocg;

extern "C" int cocg_ntt(cocg_ctx* ctx, void* const* vecs, int k, unsigned log_n, const void* root, int inverse, const void* coset_g) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (k <= 0) return 0;
  if (!vecs || !root) return fail(ctx, "cocg_ntt: null argument");
  if (log_n > 30) return fail(ctx, "cocg_ntt: log_n too large");
  for (int i = 0; i < k; i++)
    if (!vecs[i]) return fail(ctx, "cocg_ntt: null vector");
  return COCG_FR_DISPATCH(ctx, ntt_impl, ctx, vecs, k, log_n, root, inverse, coset_g);
}
