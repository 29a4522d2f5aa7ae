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
#include <string.h>

#include "ctx.cuh"

namespace cocg {

constexpr int kNttMaxStages = 8;    // stages per pass
constexpr int kNttTile = 1024;      // elements per CTA tile
constexpr int kNttThreads = 256;
constexpr int kNttMaxVecs = 8;

struct NttVecs {
  const void* in[kNttMaxVecs];
  void* out[kNttMaxVecs];
};

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

template <class P, bool LAST>
__global__ void __launch_bounds__(kNttThreads) ntt_pass_kernel(NttVecs vecs, int L, int k0, int s, int logC,
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
      x = fp_mul(x, load_fp_ro<P>(pre, g));
      lo = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
      hi = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
    }
    int m = (e << logC) + c;
    plane_lo[m] = lo;
    plane_hi[m] = hi;
  }
  __syncthreads();

  // ---- s butterfly stages: (x, y) -> (x + y, (x - y) * w^((j mod 2^(L-1-k)) * 2^k))
  const int nbf = tile_elems >> 1;
  for (int kk = 0; kk < s; kk++) {
    const int k = k0 + kk;
    const int logh = s - 1 - kk;
    const bool unit = (k == L - 1);  // last global stage: twiddle is 1
    for (int b = threadIdx.x; b < nbf; b += kNttThreads) {
      int c = b & (C - 1);
      int p = b >> logC;
      int lo_i = p & ((1 << logh) - 1);
      int hi_i = p >> logh;
      int e0 = (hi_i << (logh + 1)) + lo_i;
      int m0 = (e0 << logC) + c;
      int m1 = m0 + ((1 << logh) << logC);
      uint4 xl = plane_lo[m0], xh = plane_hi[m0], yl = plane_lo[m1], yh = plane_hi[m1];
      Fp<P> x, y;
      x.l[0] = xl.x; x.l[1] = xl.y; x.l[2] = xl.z; x.l[3] = xl.w; x.l[4] = xh.x; x.l[5] = xh.y; x.l[6] = xh.z; x.l[7] = xh.w;
      y.l[0] = yl.x; y.l[1] = yl.y; y.l[2] = yl.z; y.l[3] = yl.w; y.l[4] = yh.x; y.l[5] = yh.y; y.l[6] = yh.z; y.l[7] = yh.w;
      Fp<P> u = fp_add(x, y);
      Fp<P> v = fp_sub(x, y);
      if (!unit) {
        size_t ex = LAST ? (size_t)lo_i : (((size_t)lo_i << logQ) + col0 + c);
        v = fp_mul(v, load_fp_ro<P>(tw, ex << k));
      }
      plane_lo[m0] = make_uint4(u.l[0], u.l[1], u.l[2], u.l[3]);
      plane_hi[m0] = make_uint4(u.l[4], u.l[5], u.l[6], u.l[7]);
      plane_lo[m1] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
      plane_hi[m1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
    __syncthreads();
  }

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
        x = fp_mul(x, post ? load_fp_ro<P>(post, i) : post_const);
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
  int np = (L + kNttMaxStages - 1) / kNttMaxStages;
  int sbase = L / np, extra = L % np;
  void* scratch = nullptr;
  if (np > 1) COCG_TRY(scratch_get(ctx, 0, (size_t)k * n * sizeof(F), &scratch));
  int k0 = 0;
  ProfScope prof(ctx, COCG_PROF_NTT);
  for (int p = 0; p < np; p++) {
    int s = sbase + (p < extra ? 1 : 0);
    bool last = (p == np - 1);
    int logC = 0;
    if (!last) { logC = L - k0 - s; if (logC > 2) logC = 2; }
    else { logC = L - s; if (logC > 2) logC = 2; }
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
