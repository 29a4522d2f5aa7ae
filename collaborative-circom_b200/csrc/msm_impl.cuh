// Variable-base multi-scalar multiplication over resident public bases, replacing
// `C::msm_unchecked(points, &scalars.{a,b})` behind MSMProvider::msm_public_points
// (/root/reference/mpc-core/src/traits.rs:561-568; impls rep3.rs:934-947, shamir.rs:1027-1039,
// plain.rs:408-416; the algorithm itself is ark-ec 0.4.2's Pippenger, not vendored).  The result is the same
// group element; only its Jacobian representative may differ.
//
// Pipeline per share component (all on the context's stream):
//   1. digits      scalars leave Montgomery form and are recoded into signed c-bit digits (|d| <= 2^(c-1));
//                  a histogram of (window, |d|) is taken with global atomics
//   2. scan        exclusive prefix sum of the histogram -> bucket offsets
//   3. scatter     counting sort: point indices (sign in bit 31) grouped by (window, bucket)
//   4. accumulate  one thread per bucket walks its index run, gathers affine points with 128-bit loads and
//                  adds them into an XYZZ accumulator (8M + 2S per point); runs longer than kHeavy (skewed
//                  scalars) are deferred to a warp-per-bucket kernel that tree-combines with warp shuffles
//   5. reduce      sum_b (b+1) * B_b per window: threads take 16-bucket segments (running-sum trick) and
//                  lift them by a short double-and-add; a CTA per window sums the lifted pieces
//   6. fold        the <= 64 window sums go to the host, which does the 2^c Horner fold (sequential doublings
//                  are latency-bound on a GPU and O(1) for the caller)
// Shares are uniformly random, so buckets are balanced (N / 2^(c-1) ~ 32 points each at c = log2 N - 4).
#pragma once
#include <string.h>

#include "ctx.cuh"

namespace cocg {

constexpr int kMaxWindows = 96;
constexpr int kHeavy = 256;     // runs longer than this go to the warp-per-bucket kernel
constexpr int kSeg = 16;        // buckets per thread in the reduce kernel

static int msm_window_bits(size_t n) {
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= n) lg++;
  int c = lg - 4;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  return c;
}

// ------------------------------------------------------------------ 1. digits + histogram
template <class FrP>
__global__ void __launch_bounds__(256) msm_digits_kernel(const void* __restrict__ scalars, size_t n, int c, int nwin, int mont,
                                                          uint32_t* __restrict__ dig, uint32_t* __restrict__ counts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<FrP> s = load_fp<FrP>(scalars, i);
  if (mont) s = fp_from_mont(s);
  const uint32_t nb = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1;
  uint32_t carry = 0;
  for (int w = 0; w < nwin; w++) {
    int bit = w * c;
    uint32_t raw = 0;
    if (bit < 32 * FrP::N) {
      int word = bit >> 5, sh = bit & 31;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int q = 0; q < FrP::N; q++) {  // select without dynamic register indexing
        lo = (q == word) ? s.l[q] : lo;
        hi = (q == word + 1) ? s.l[q] : hi;
      }
      raw = __funnelshift_r(lo, hi, sh) & mask;
    }
    uint32_t d = raw + carry;
    carry = 0;
    uint32_t enc = 0;  // 0 = skip; else (bucket+1) | sign << 31
    if (d > nb) {
      enc = ((1u << c) - d) | 0x80000000u;
      carry = 1;
      if (((1u << c) - d) == 0) enc = 0;  // d == 2^c: digit 0 with carry
    } else {
      enc = d;
    }
    dig[(size_t)w * n + i] = enc;
    if (enc & 0x7fffffffu) atomicAdd(&counts[(size_t)w * nb + ((enc & 0x7fffffffu) - 1)], 1u);
  }
}

// ------------------------------------------------------------------ 2. exclusive scan (3 small kernels)
constexpr int kScanThreads = 256, kScanPer = 4, kScanBlock = kScanThreads * kScanPer;
static __global__ void __launch_bounds__(kScanThreads) scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t m,
                                                                   uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sh[kScanThreads];
  size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPer;
  uint32_t v[kScanPer], sum = 0;
#pragma unroll
  for (int q = 0; q < kScanPer; q++) { v[q] = base + q < m ? in[base + q] : 0; sum += v[q]; }
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {
    uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t excl = sh[threadIdx.x] - sum;
#pragma unroll
  for (int q = 0; q < kScanPer; q++) { if (base + q < m) out[base + q] = excl; excl += v[q]; }
  if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}
static __global__ void scan_top_kernel(uint32_t* block_sums, size_t nblocks, uint32_t* total) {
  // single thread block; nblocks is at most a few thousand
  __shared__ uint32_t sh[kScanThreads];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += kScanThreads) {
    size_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? block_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kScanThreads; off <<= 1) {
      uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_sums[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
static __global__ void __launch_bounds__(kScanThreads) scan_add_kernel(uint32_t* __restrict__ out, size_t m, const uint32_t* __restrict__ block_sums,
                                                                 const uint32_t* __restrict__ total) {
  size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPer;
  uint32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int q = 0; q < kScanPer; q++)
    if (base + q < m) out[base + q] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[m] = *total;
}

// ------------------------------------------------------------------ 3. scatter (counting sort by bucket)
static __global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t* __restrict__ dig, size_t n, int nwin, uint32_t nb,
                                                           const uint32_t* __restrict__ start, uint32_t* __restrict__ counts,
                                                           uint32_t* __restrict__ sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int w = 0; w < nwin; w++) {
    uint32_t enc = dig[(size_t)w * n + i];
    uint32_t b = enc & 0x7fffffffu;
    if (!b) continue;
    size_t gb = (size_t)w * nb + (b - 1);
    uint32_t slot = atomicSub(&counts[gb], 1u) - 1u;  // counts[] drains back to zero
    sorted[start[gb] + slot] = (uint32_t)i | (enc & 0x80000000u);
  }
}

// ------------------------------------------------------------------ point load helpers
template <class F>
__device__ __forceinline__ Affine<F> load_affine(const void* bases, size_t idx) {
  Affine<F> p;
  constexpr int W = sizeof(Affine<F>) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(bases) + idx * W;
  uint32_t* dst = reinterpret_cast<uint32_t*>(&p);
#pragma unroll
  for (int k = 0; k < W; k++) {
    uint4 v = __ldg(src + k);
    dst[4 * k] = v.x; dst[4 * k + 1] = v.y; dst[4 * k + 2] = v.z; dst[4 * k + 3] = v.w;
  }
  return p;
}
template <class T>
__device__ __forceinline__ T warp_shfl_down(const T& v, int delta) {
  T r;
  constexpr int W = sizeof(T) / 4;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < W; k++) d[k] = __shfl_down_sync(0xffffffffu, s[k], delta);
  return r;
}
template <class F>
__device__ __forceinline__ void accumulate_run(XYZZ<F>& acc, const void* bases, const uint32_t* sorted, uint32_t beg, uint32_t end, uint32_t step) {
  for (uint32_t e = beg; e < end; e += step) {
    uint32_t v = sorted[e];
    Affine<F> p = load_affine<F>(bases, v & 0x7fffffffu);
    if (v >> 31) p.y = f_neg(p.y);
    xyzz_madd(acc, p);
  }
}

// ------------------------------------------------------------------ 4. bucket accumulation
template <class F>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const void* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                              const uint32_t* __restrict__ start, size_t nbuckets, XYZZ<F>* __restrict__ buckets,
                                                              uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ heavy_count) {
  size_t gb = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= nbuckets) return;
  uint32_t beg = start[gb], end = start[gb + 1];
  if (end - beg > (uint32_t)kHeavy) {
    heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)gb;
    return;
  }
  XYZZ<F> acc = xyzz_inf<F>();
  accumulate_run<F>(acc, bases, sorted, beg, end, 1);
  buckets[gb] = acc;
}
// one warp per heavy bucket: lanes stride through the run, then a shuffle tree combines the 32 partial sums
template <class F>
__global__ void __launch_bounds__(128) msm_heavy_kernel(const void* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                         const uint32_t* __restrict__ start, XYZZ<F>* __restrict__ buckets,
                                                         const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t total = *heavy_count;
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < total; h += nwarps) {
    uint32_t gb = heavy_list[h];
    XYZZ<F> acc = xyzz_inf<F>();
    accumulate_run<F>(acc, bases, sorted, start[gb] + lane, start[gb + 1], 32);
    for (int delta = 16; delta >= 1; delta >>= 1) {
      XYZZ<F> other = warp_shfl_down(acc, delta);
      if (lane < (uint32_t)delta) xyzz_add(acc, other);
    }
    if (lane == 0) buckets[gb] = acc;
  }
}

// ------------------------------------------------------------------ 5. per-window bucket reduction
template <class F>
__device__ __forceinline__ XYZZ<F> xyzz_mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> acc = xyzz_inf<F>();
  for (int b = 31 - __clz(k | 1); b >= 0; b--) {
    acc = xyzz_dbl(acc);
    if ((k >> b) & 1) xyzz_add(acc, p);
  }
  return acc;
}
// thread (w, g): buckets [g*kSeg, g*kSeg + len) of window w hold weights g*kSeg+1 .. ; emits
//   piece = sum_t (t+1) * B[g*kSeg+t] + (g*kSeg) * sum_t B[g*kSeg+t]
template <class F>
__global__ void __launch_bounds__(128) msm_reduce_segments_kernel(const XYZZ<F>* __restrict__ buckets, uint32_t nb, uint32_t segs_per_window,
                                                                   size_t total_segs, XYZZ<F>* __restrict__ pieces) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_segs) return;
  uint32_t w = (uint32_t)(t / segs_per_window), g = (uint32_t)(t % segs_per_window);
  uint32_t lo = g * kSeg, hi = lo + kSeg < nb ? lo + kSeg : nb;
  const XYZZ<F>* B = buckets + (size_t)w * nb;
  XYZZ<F> run = xyzz_inf<F>(), acc = xyzz_inf<F>();
  for (uint32_t b = hi; b-- > lo;) {
    xyzz_add(run, B[b]);
    xyzz_add(acc, run);
  }
  if (lo) xyzz_add(acc, xyzz_mul_small(run, lo));
  pieces[t] = acc;
}
// one CTA per window: sum the window's pieces
template <class F>
__global__ void __launch_bounds__(128) msm_sum_pieces_kernel(const XYZZ<F>* __restrict__ pieces, uint32_t segs_per_window, XYZZ<F>* __restrict__ window_sums) {
  __shared__ XYZZ<F> sh[4];
  const XYZZ<F>* Pw = pieces + (size_t)blockIdx.x * segs_per_window;
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t i = threadIdx.x; i < segs_per_window; i += blockDim.x) xyzz_add(acc, Pw[i]);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int delta = 16; delta >= 1; delta >>= 1) {
    XYZZ<F> other = warp_shfl_down(acc, delta);
    if (lane < (uint32_t)delta) xyzz_add(acc, other);
  }
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 4; k++) xyzz_add(acc, sh[k]);
    window_sums[blockIdx.x] = acc;
  }
}

// ------------------------------------------------------------------ driver
template <class F, class FrP>
int msm_impl(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac) {
  using X = XYZZ<F>;
  char* out = reinterpret_cast<char*>(out_jac);  // caller memory: no alignment assumed
  if (n == 0) {
    Jacobian<F> inf = jac_inf<F>();
    for (int j = 0; j < k; j++) memcpy(out + (size_t)j * sizeof(inf), &inf, sizeof(inf));
    return 0;
  }
  if (n >= ((size_t)1 << 31)) return fail(ctx, "cocg_msm: n must be < 2^31");
  const int c = msm_window_bits(n);
  const int nwin = (FrP::BITS + c) / c;  // ceil((BITS+1)/c): room for the final carry
  if (nwin > kMaxWindows) return fail(ctx, "cocg_msm: too many windows");
  const uint32_t nb = 1u << (c - 1);
  const size_t nbuckets = (size_t)nwin * nb;
  const uint32_t segs = (nb + kSeg - 1) / kSeg;
  const size_t total_segs = (size_t)nwin * segs;
  const size_t scan_blocks = (nbuckets + kScanBlock - 1) / kScanBlock;

  uint32_t *dig, *sorted, *counts, *start, *bsums, *heavy;
  X *buckets, *pieces, *wsums;
  void* p;
  COCG_TRY(scratch_get(ctx, 1, (size_t)nwin * n * 4, &p)); dig = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 2, (size_t)nwin * n * 4, &p)); sorted = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 3, nbuckets * 4, &p)); counts = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 4, (nbuckets + 1) * 4, &p)); start = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 5, (scan_blocks + 2) * 4, &p)); bsums = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 6, (nbuckets + 1) * 4, &p)); heavy = (uint32_t*)p;  // [0] = count, [1..] = list
  COCG_TRY(scratch_get(ctx, 7, nbuckets * sizeof(X), &p)); buckets = (X*)p;
  COCG_TRY(scratch_get(ctx, 8, total_segs * sizeof(X), &p)); pieces = (X*)p;
  COCG_TRY(scratch_get(ctx, 9, (size_t)k * nwin * sizeof(X), &p)); wsums = (X*)p;
  const char* base_ptr = (const char*)be.d + off * be.point_bytes;
  cudaStream_t st = ctx->stream;

  COCG_CUDA(ctx, cudaMemsetAsync(counts, 0, nbuckets * 4, st));
  for (int j = 0; j < k; j++) {
    COCG_CUDA(ctx, cudaMemsetAsync(heavy, 0, 4, st));
    {
    ProfScope prof(ctx, COCG_PROF_MSM_SORT);
    msm_digits_kernel<FrP><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scalars[j], n, c, nwin, mont, dig, counts);
    COCG_LAUNCH_CHECK(ctx);
    scan_block_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, start, nbuckets, bsums);
    COCG_LAUNCH_CHECK(ctx);
    scan_top_kernel<<<1, kScanThreads, 0, st>>>(bsums, scan_blocks, bsums + scan_blocks);
    COCG_LAUNCH_CHECK(ctx);
    scan_add_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(start, nbuckets, bsums, bsums + scan_blocks);
    COCG_LAUNCH_CHECK(ctx);
    msm_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dig, n, nwin, nb, start, counts, sorted);
    COCG_LAUNCH_CHECK(ctx);
    }
    {
    ProfScope prof(ctx, COCG_PROF_MSM_ACCUMULATE);
    msm_accumulate_kernel<F><<<(unsigned)((nbuckets + 127) / 128), 128, 0, st>>>(base_ptr, sorted, start, nbuckets, buckets, heavy + 1, heavy);
    COCG_LAUNCH_CHECK(ctx);
    msm_heavy_kernel<F><<<kNumSMs, 128, 0, st>>>(base_ptr, sorted, start, buckets, heavy + 1, heavy);
    COCG_LAUNCH_CHECK(ctx);
    }
    ProfScope prof(ctx, COCG_PROF_MSM_REDUCE);
    msm_reduce_segments_kernel<F><<<(unsigned)((total_segs + 127) / 128), 128, 0, st>>>(buckets, nb, segs, total_segs, pieces);
    COCG_LAUNCH_CHECK(ctx);
    msm_sum_pieces_kernel<F><<<nwin, 128, 0, st>>>(pieces, segs, wsums + (size_t)j * nwin);
    COCG_LAUNCH_CHECK(ctx);
  }
  // 6. window sums -> host, Horner fold sum_w 2^(c*w) W_w
  void* hp;
  COCG_TRY(pinned_get(ctx, (size_t)k * nwin * sizeof(X), &hp));
  COCG_CUDA(ctx, cudaMemcpyAsync(hp, wsums, (size_t)k * nwin * sizeof(X), cudaMemcpyDeviceToHost, st));
  COCG_CUDA(ctx, cudaStreamSynchronize(st));
  const X* hw = reinterpret_cast<const X*>(hp);
  for (int j = 0; j < k; j++) {
    X acc = hw[(size_t)j * nwin + nwin - 1];
    for (int w = nwin - 2; w >= 0; w--) {
      for (int q = 0; q < c; q++) acc = xyzz_dbl(acc);
      xyzz_add(acc, hw[(size_t)j * nwin + w]);
    }
    Jacobian<F> jr = xyzz_to_jacobian(acc);
    memcpy(out + (size_t)j * sizeof(jr), &jr, sizeof(jr));
  }
  return 0;
}


// per-(curve, group) entry points, one translation unit each (msm_<curve>_<group>.cu)
int msm_bn254_g1(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac);
int msm_bn254_g2(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac);
int msm_bls381_g1(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac);
int msm_bls381_g2(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac);

}  // namespace cocg
