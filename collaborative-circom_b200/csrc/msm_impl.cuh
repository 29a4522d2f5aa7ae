// Variable-base multi-scalar multiplication over resident public bases, replacing
// `C::msm_unchecked(points, &scalars.{a,b})` behind MSMProvider::msm_public_points
// (/root/reference/mpc-core/src/traits.rs:561-568; impls rep3.rs:934-947, shamir.rs:1027-1039,
// plain.rs:408-416; the algorithm itself is ark-ec 0.4.2's Pippenger, not vendored).  The result is the same
// group element; only its Jacobian representative may differ.
//
// B200-first design: the bases of a zkey query are public and reused by every proof (the reference borrows &ZKey for
// the whole prove, groth16.rs:113-117), and HBM is the abundant resource (180 GB), so the window structure of Pippenger
// is moved into memory.  At upload time the table T[j][i] = 2^(c*j) * P_i, j < nwin = ceil((bits+1)/c), is built once
// (affine, 13 x the query at c = 20).  A signed digit d of window j of scalar s_i then contributes
// sign(d) * T[j][i] to bucket |d| of ONE shared bucket set, so that
//     sum_i s_i P_i = sum_b b * B_b,      B_b = sum of the table points whose digit is +-b,
// with a single bucket reduction per MSM instead of one per window and no doublings at all on the path.
// Pipeline per share component (all on the context's stream):
//   1. digits      scalars leave Montgomery form and are recoded into signed c-bit digits (|d| <= 2^(c-1));
//                  histogram of |d| with global atomics
//   2. scan        exclusive prefix sum of the histogram -> bucket offsets
//   3. scatter     counting sort: (window, point index, sign) entries grouped by bucket
//   4. accumulate  one thread per bucket walks its run, gathers affine table points with 128-bit loads and adds them
//                  into an XYZZ accumulator (8M + 2S per point); runs longer than kHeavy (skewed scalars: plain-driver
//                  witnesses) are deferred to a warp-per-bucket kernel that tree-combines with warp shuffles
//   5. reduce      sum_b (b+1) B_b with b = hi*L + lo:  sum_hi (L*hi + 1) R_hi + sum_lo lo * C_lo  for the row sums
//                  R_hi = sum_lo B and column sums C_lo = sum_hi B -- two parallel marginal reductions (a warp per
//                  row / column) instead of a sequential running sum, then one small tree sum
// Shares are uniformly random, so buckets are balanced (N * nwin / 2^(c-1) ~ 26 points each at N = 2^20, c = 20).
#pragma once
#include <string.h>

#include "ctx.cuh"

namespace cocg {

constexpr int kHeavy = 1024;         // runs longer than this leave the thread-per-bucket path ...
constexpr int kHeavyChunk = 256;     // ... and are cut into chunks of this many entries, one warp each.  A top window only a few bits wide
                                     // (c = 18: bits 252..253) puts n / 4 entries into each of 3 buckets; with 1024-entry chunks those were
                                     // 256 warps of 37 dependent additions = 207 us per 2^18-term MSM (ncu, round 2), a third of the
                                     // accumulate time.  256-entry chunks give 4x the warps and a quarter of the chain.
constexpr int kIdxBits = 25;         // entry = sign << 31 | window << 25 | point index

template <class FrP>
static int msm_num_windows(int c) {
  return (FrP::BITS + c) / c;  // ceil((BITS+1)/c): room for the final carry of the signed recoding
}

// ------------------------------------------------------------------ 0. table precompute (upload time)
// thread i: T[j][i] = 2^(c*j) P_i for j = 1..nwin-1, normalised to affine with one shared inversion per group of
// kPreGroup windows
constexpr int kPreGroup = 16;
template <class F>
__global__ void __launch_bounds__(128) msm_precompute_kernel(Affine<F>* __restrict__ table, size_t n, int c, int nwin) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = table[i];
  if (p.is_inf()) {
    for (int j = 1; j < nwin; j++) table[(size_t)j * n + i] = p;
    return;
  }
  XYZZ<F> pts[kPreGroup];
  F pref[kPreGroup];
  XYZZ<F> acc = xyzz_from_affine(p);
  for (int j0 = 1; j0 < nwin; j0 += kPreGroup) {
    const int cnt = nwin - j0 < kPreGroup ? nwin - j0 : kPreGroup;
    F run = F::one();
    for (int t = 0; t < cnt; t++) {
      for (int q = 0; q < c; q++) acc = xyzz_dbl(acc);
      pts[t] = acc;
      pref[t] = run;
      if (!acc.is_inf()) run = f_mul(run, acc.zzz);
    }
    F inv = f_inv(run);
    for (int t = cnt - 1; t >= 0; t--) {
      Affine<F> a;
      if (pts[t].is_inf()) {
        a.x = F::zero();
        a.y = F::zero();
      } else {
        F w = f_mul(inv, pref[t]);   // 1 / zzz_t
        inv = f_mul(inv, pts[t].zzz);
        F zi = f_mul(pts[t].zz, w);  // 1 / z
        a.x = f_mul(pts[t].x, f_sqr(zi));
        a.y = f_mul(pts[t].y, w);
      }
      table[(size_t)(j0 + t) * n + i] = a;
    }
  }
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
    uint32_t enc = 0;  // 0 = skip; else |digit| | sign << 31
    if (d > nb) {
      enc = ((1u << c) - d) | 0x80000000u;
      carry = 1;
      if (((1u << c) - d) == 0) enc = 0;  // d == 2^c: digit 0 with carry
    } else {
      enc = d;
    }
    dig[(size_t)w * n + i] = enc;
    if (enc & 0x7fffffffu) atomicAdd(&counts[(enc & 0x7fffffffu) - 1], 1u);
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

// ------------------------------------------------------------------ 2b. bucket order by size
// Threads of a warp walk their buckets in lock-step, so a warp costs as much as its largest bucket.  Buckets are handed
// out in descending size order (counting sort on the size, sizes > kHeavy share the first bin).
constexpr int kSizeBins = kHeavy + 2;
static __global__ void __launch_bounds__(256) bucket_size_hist_kernel(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kSizeBins];
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
    uint32_t s = counts[b];
    atomicAdd(&sh[s > (uint32_t)kHeavy ? kHeavy + 1 : s], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// hist[s] <- number of buckets larger than s (start of size class s in the descending order).  One block: the bins are staged in
// shared memory and scanned by thread 0 there (a serial walk over global memory was 35 us of load latency per sort).
static __global__ void __launch_bounds__(256) bucket_size_scan_kernel(uint32_t* hist) {
  __shared__ uint32_t sh[kSizeBins];
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x) sh[i] = hist[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int s = kSizeBins - 1; s >= 0; s--) {
      uint32_t v = sh[s];
      sh[s] = run;
      run += v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x) hist[i] = sh[i];
}
// Bucket sizes concentrate on a few dozen values (mean n * nwin / nb), so one global atomic per bucket was ~2^19 atomics on ~40
// addresses (100 us per sort).  A block counts its buckets per size class in shared memory, reserves one range per class with ONE
// global atomic, and its threads take consecutive slots of that range.  The order inside a size class is irrelevant (it only decides
// which thread accumulates which bucket).
constexpr int kOrderThreads = 1024;
static __global__ void __launch_bounds__(kOrderThreads) bucket_order_kernel(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t* __restrict__ pos,
                                                                     uint32_t* __restrict__ order) {
  __shared__ uint32_t sh_cnt[kSizeBins];
  __shared__ uint32_t sh_base[kSizeBins];
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x) sh_cnt[i] = 0;
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cls = 0, mine = 0;
  if (b < nb) {
    const uint32_t s = counts[b];
    cls = s > (uint32_t)kHeavy ? kHeavy + 1 : s;
    mine = atomicAdd(&sh_cnt[cls], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSizeBins; i += blockDim.x)
    if (sh_cnt[i]) sh_base[i] = atomicAdd(&pos[i], sh_cnt[i]);
  __syncthreads();
  if (b < nb) order[sh_base[cls] + mine] = b;
}

// ------------------------------------------------------------------ 3. scatter (counting sort by bucket)
static __global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t* __restrict__ dig, size_t n, int nwin,
                                                           const uint32_t* __restrict__ start, uint32_t* __restrict__ counts,
                                                           uint32_t* __restrict__ sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int w = 0; w < nwin; w++) {
    uint32_t enc = dig[(size_t)w * n + i];
    uint32_t b = enc & 0x7fffffffu;
    if (!b) continue;
    uint32_t slot = atomicSub(&counts[b - 1], 1u) - 1u;  // counts[] drains back to zero
    sorted[start[b - 1] + slot] = (uint32_t)i | ((uint32_t)w << kIdxBits) | (enc & 0x80000000u);
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
// table: T[w][off + i], window stride `tstride` points
template <class F>
__device__ __forceinline__ void accumulate_run(XYZZ<F>& acc, const void* table, size_t tstride, const uint32_t* sorted, uint32_t beg, uint32_t end,
                                               uint32_t step) {
  // The gather address of entry e + step is known one addition ahead: prefetching its cache line(s) into L1 hides the DRAM latency
  // of the random table access behind the current addition without holding a second point in registers.
  auto entry_index = [&](uint32_t v) { return (size_t)((v >> kIdxBits) & 63u) * tstride + (v & ((1u << kIdxBits) - 1)); };
  if (beg >= end) return;
  uint32_t v = sorted[beg];
  for (uint32_t e = beg; e < end; e += step) {
    const size_t idx = entry_index(v);
    uint32_t vn = 0;
    if (e + step < end) {
      vn = sorted[e + step];
      const char* nxt = reinterpret_cast<const char*>(table) + entry_index(vn) * sizeof(Affine<F>);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
      if (128 % sizeof(Affine<F>) != 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + sizeof(Affine<F>) - 16));  // may straddle two lines
    }
    Affine<F> p = load_affine<F>(table, idx);
    if (v >> 31) p.y = f_neg(p.y);
    xyzz_madd(acc, p);
    v = vn;
  }
}

// ------------------------------------------------------------------ 4. bucket accumulation
static inline int msm_bucket_slot(int group) { return group == COCG_G1 ? 7 : 14; }  // scratch arena of the group's bucket sets
// Blocks of 128 threads, registers left to ptxas (104 for Fq, 188 for Fq2 = 2 blocks per SM).  Measured alternatives on B200
// (BN254 G2, 2^20 terms, ms): 7.6 as is; 64-thread blocks (5 per SM) 8.08; capped at 168 registers (3 blocks per SM, 156 B of
// spills) 8.23.  G1 capped at 96 registers (5 blocks per SM): 2.12, unchanged.
constexpr int kAccumulateThreads = 128;
// After the Fq2 product became a call (ec.cuh) the BN254 G2 instantiation is occupancy-limited rather than fetch-limited: capping it at
// 128 registers (4 blocks per SM, 456 B of L1-resident spills) measured 7.01 -> 6.52 ms at 2^20 terms; 3 blocks 6.67, 5 blocks 6.82.
template <class F> constexpr int kAccumulateMinBlocks = sizeof(F) == 64 ? 4 : 1;
// Fq2 (BN254 G2): the bucket accumulator lives in SHARED memory, 16 B words interleaved over the block's threads (conflict-free),
// and each coordinate is read where it is used and written back as soon as its new value exists.  In registers the accumulator is
// 64 of the 128 the 4-blocks-per-SM cap allows, the Fq2 product is a call that values cannot be moved across, and ptxas answered
// with 456 B of spills per thread (ncu: ~0.95 GB of spill traffic reaching DRAM per 2^20-term launch).
template <class F> constexpr int kAccSmemWords = sizeof(F) == 64 ? (int)(sizeof(XYZZ<F>) / 16) * kAccumulateThreads : 1;
template <class F>
__device__ __forceinline__ F sm_get(const uint4* sm, int comp) {  // coordinate `comp` (0 X, 1 Y, 2 ZZ, 3 ZZZ) of this thread's accumulator
  constexpr int W = sizeof(F) / 16;
  F r;
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < W; k++) {
    const uint4 v = sm[(comp * W + k) * kAccumulateThreads + threadIdx.x];
    d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
  }
  return r;
}
template <class F>
__device__ __forceinline__ void sm_put(uint4* sm, int comp, const F& x) {
  constexpr int W = sizeof(F) / 16;
  const uint32_t* d = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
  for (int k = 0; k < W; k++) sm[(comp * W + k) * kAccumulateThreads + threadIdx.x] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
}
// acc += q with the accumulator in shared memory (the formulas and the special cases of xyzz_madd, ec.cuh); inf = accumulator is O
template <class F>
__device__ __forceinline__ void xyzz_madd_sm(uint4* sm, bool& inf, const Affine<F>& q) {
  if (q.is_inf()) return;
  if (inf) {
    sm_put<F>(sm, 0, q.x); sm_put<F>(sm, 1, q.y); sm_put<F>(sm, 2, F::one()); sm_put<F>(sm, 3, F::one());
    inf = false;
    return;
  }
  const F Pd = f_sub(f_mul(q.x, sm_get<F>(sm, 2)), sm_get<F>(sm, 0));
  const F R = f_sub(f_mul(q.y, sm_get<F>(sm, 3)), sm_get<F>(sm, 1));
  if (Pd.is_zero()) {
    if (R.is_zero()) {
      const XYZZ<F> d = xyzz_dbl_affine(q);
      sm_put<F>(sm, 0, d.x); sm_put<F>(sm, 1, d.y); sm_put<F>(sm, 2, d.zz); sm_put<F>(sm, 3, d.zzz);
    } else {
      inf = true;
    }
    return;
  }
  const F PP = f_sqr(Pd);
  sm_put<F>(sm, 2, f_mul(sm_get<F>(sm, 2), PP));
  const F Q = f_mul(sm_get<F>(sm, 0), PP);
  const F PPP = f_mul(Pd, PP);
  sm_put<F>(sm, 3, f_mul(sm_get<F>(sm, 3), PPP));
  const F T = f_mul(sm_get<F>(sm, 1), PPP);
  const F X3 = f_sub(f_sub(f_sqr(R), PPP), f_dbl(Q));
  sm_put<F>(sm, 0, X3);
  sm_put<F>(sm, 1, f_sub(f_mul(R, f_sub(Q, X3)), T));
}
template <class F>
__device__ __forceinline__ void accumulate_run_sm(uint4* sm, bool& inf, const void* table, size_t tstride, const uint32_t* sorted, uint32_t beg,
                                                  uint32_t end) {
  auto entry_index = [&](uint32_t v) { return (size_t)((v >> kIdxBits) & 63u) * tstride + (v & ((1u << kIdxBits) - 1)); };
  if (beg >= end) return;
  uint32_t v = sorted[beg];
  for (uint32_t e = beg; e < end; e++) {
    const size_t idx = entry_index(v);
    uint32_t vn = 0;
    if (e + 1 < end) {
      vn = sorted[e + 1];
      const char* nxt = reinterpret_cast<const char*>(table) + entry_index(vn) * sizeof(Affine<F>);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
      if (128 % sizeof(Affine<F>) != 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + sizeof(Affine<F>) - 16));
    }
    Affine<F> p = load_affine<F>(table, idx);
    if (v >> 31) p.y = f_neg(p.y);
    xyzz_madd_sm<F>(sm, inf, p);
    v = vn;
  }
}
// One thread per bucket, buckets taken in descending size order so that the lanes of a warp finish together (ncu: 31.7 of 32
// lanes active, fmaheavy pipe 90 % busy at 2^19 buckets of ~26 points).
struct HeavyRec {
  uint32_t bucket, first_chunk, nchunks;
};
template <class F>
__global__ void __launch_bounds__(kAccumulateThreads, kAccumulateMinBlocks<F>) msm_accumulate_kernel(const void* __restrict__ table, size_t tstride, const uint32_t* __restrict__ sorted,
                                                              const uint32_t* __restrict__ start, const uint32_t* __restrict__ order, uint32_t nbuckets,
                                                              XYZZ<F>* __restrict__ buckets, HeavyRec* __restrict__ heavy_list,
                                                              uint32_t* __restrict__ chunk_owner,
                                                              uint32_t* __restrict__ heavy_count /* [0] buckets, [1] chunks */) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbuckets) return;
  const uint32_t gb = order[t];
  const uint32_t beg = start[gb], end = start[gb + 1];
  if (end - beg > (uint32_t)kHeavy) {
    uint32_t nch = (end - beg + kHeavyChunk - 1) / kHeavyChunk;
    uint32_t h = atomicAdd(&heavy_count[0], 1u);
    uint32_t first = atomicAdd(&heavy_count[1], nch);
    heavy_list[h] = HeavyRec{gb, first, nch};
    for (uint32_t q = 0; q < nch; q++) chunk_owner[first + q] = h;
    return;
  }
  if constexpr (sizeof(F) == 64) {
    __shared__ uint4 acc_sm[kAccSmemWords<F>];
    bool inf = true;
    accumulate_run_sm<F>(acc_sm, inf, table, tstride, sorted, beg, end);
    XYZZ<F> acc = xyzz_inf<F>();
    if (!inf) acc = XYZZ<F>{sm_get<F>(acc_sm, 0), sm_get<F>(acc_sm, 1), sm_get<F>(acc_sm, 2), sm_get<F>(acc_sm, 3)};
    buckets[gb] = acc;
  } else {
    XYZZ<F> acc = xyzz_inf<F>();
    accumulate_run<F>(acc, table, tstride, sorted, beg, end, 1);
    buckets[gb] = acc;
  }
}
// skewed scalars (plain-driver witnesses, or a top window narrower than c bits): a warp per kHeavy-entry chunk ...
template <class F>
__global__ void __launch_bounds__(128) msm_heavy_chunks_kernel(const void* __restrict__ table, size_t tstride, const uint32_t* __restrict__ sorted,
                                                                const uint32_t* __restrict__ start, const HeavyRec* __restrict__ heavy_list,
                                                                const uint32_t* __restrict__ chunk_owner, const uint32_t* __restrict__ heavy_count,
                                                                XYZZ<F>* __restrict__ partial) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nchunks = heavy_count[1];
  for (uint32_t ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks; ch += nwarps) {
    const HeavyRec rec = heavy_list[chunk_owner[ch]];
    uint32_t beg = start[rec.bucket] + (ch - rec.first_chunk) * kHeavyChunk;
    uint32_t end = min(beg + (uint32_t)kHeavyChunk, start[rec.bucket + 1]);
    XYZZ<F> acc = xyzz_inf<F>();
    accumulate_run<F>(acc, table, tstride, sorted, beg + lane, end, 32);
    for (int delta = 16; delta >= 1; delta >>= 1) {
      XYZZ<F> other = warp_shfl_down(acc, delta);
      if (lane < (uint32_t)delta) xyzz_add(acc, other);
    }
    if (lane == 0) partial[ch] = acc;
  }
}
// ... then a warp per heavy bucket folds its chunk sums
template <class F>
__global__ void __launch_bounds__(128) msm_heavy_fold_kernel(const HeavyRec* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count,
                                                              const XYZZ<F>* __restrict__ partial, XYZZ<F>* __restrict__ buckets) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nheavy = heavy_count[0];
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < nheavy; h += nwarps) {
    const HeavyRec rec = heavy_list[h];
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t q = lane; q < rec.nchunks; q += 32) xyzz_add(acc, partial[rec.first_chunk + q]);
    for (int delta = 16; delta >= 1; delta >>= 1) {
      XYZZ<F> other = warp_shfl_down(acc, delta);
      if (lane < (uint32_t)delta) xyzz_add(acc, other);
    }
    if (lane == 0) buckets[rec.bucket] = acc;
  }
}

// ------------------------------------------------------------------ 5. bucket reduction
template <class F>
__device__ __forceinline__ XYZZ<F> xyzz_mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> acc = xyzz_inf<F>();
  if (k == 0) return acc;
  for (int b = 31 - __clz(k); b >= 0; b--) {
    acc = xyzz_dbl(acc);
    if ((k >> b) & 1) xyzz_add(acc, p);
  }
  return acc;
}
// Buckets form an H x L matrix (b = hi * L + lo); sum_b (b + 1) B_b = sum_hi (L*hi + 1) R_hi + sum_lo lo * C_lo with the
// row sums R_hi = sum_lo B[hi][lo] and column sums C_lo = sum_hi B[hi][lo].  One warp per row; columns (twice as long when
// H = 2L) are cut into kColSeg segments of one warp each, so that all warps run the same number of additions.
// out[row], then out[H + col * kColSeg + seg].
// The reductions of up to kMaxSets bucket sets (the queries and share components of one cocg_msm_multi call) run as ONE launch
// of each kernel, set index = blockIdx.y: a single 2^19-bucket set gives only ~3.5 warps per scheduler and the weigh / final
// kernels are pure latency chains, so batching sets is free parallelism.
constexpr uint32_t kColSeg = 2;
constexpr int kMaxSets = 8;
constexpr size_t kResultSlot = 512;  // bytes reserved per XYZZ result (G2 over BLS12-381 needs 384)
struct ReduceSets {
  uint32_t n;
  uint32_t slot[kMaxSets];  // result slot (kResultSlot bytes each) of every set
};
static inline uint32_t marg_stride(uint32_t nmarg) { return nmarg + (nmarg + 31) / 32 + 2; }  // partial marginals | weighted warp sums
// Marginal sums with the accumulators in SHARED memory.  A block of 128 threads owns 4 marginals; thread T sums part T / 4 (of 32)
// of marginal T mod 4 into its own accumulator slot (16-byte words interleaved over the block: conflict-free), reading each
// coordinate of the accumulator and of the addend where the formula uses it -- no XYZZ value is ever held in registers as a whole,
// which is what lets the Fq2 instantiation run 4 blocks per SM instead of 2 (250 registers before).  The 32 partial sums of a
// marginal are then folded in place: T < 64 adds slot T + 64, T < 32 adds slot T + 32, ... -- whole warps drop out level after level
// (2, 1, 1, 1, 1 warp-additions per block) where a shuffle tree kept 16, 8, 4, 2, 1 lanes of EVERY warp busy.
// Over Fq the 14 products of an addition are 7 calls of two independent products (instruction cache, see fp_mul2_call in ec.cuh).
template <class F>
struct Mul2 {
  F a, b;
};
template <class P>
__device__ __forceinline__ Mul2<Fp<P>> f_mul2(const Fp<P>& a0, const Fp<P>& b0, const Fp<P>& a1, const Fp<P>& b1) {
  const FqPair<P> r = fp_mul2_call<P>(a0, b0, a1, b1);
  return Mul2<Fp<P>>{r.a, r.b};
}
template <class P>
__device__ __forceinline__ Mul2<Fp2<P>> f_mul2(const Fp2<P>& a0, const Fp2<P>& b0, const Fp2<P>& a1, const Fp2<P>& b1) {
  return Mul2<Fp2<P>>{f_mul(a0, b0), f_mul(a1, b1)};
}
template <class F>
struct GlobalXyzz {  // coordinate getter of a bucket in global memory
  const uint4* p;
  __device__ __forceinline__ F operator()(int comp) const {
    constexpr int W = sizeof(F) / 16;
    F r;
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint4 v = __ldg(p + comp * W + k);
      d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
    }
    return r;
  }
};
template <class F>
struct SlotXyzz {  // coordinate getter of another thread's accumulator slot
  const uint4* sm;
  uint32_t slot;
  __device__ __forceinline__ F operator()(int comp) const {
    constexpr int W = sizeof(F) / 16;
    F r;
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < W; k++) {
      const uint4 v = sm[(comp * W + k) * 128 + slot];
      d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
    }
    return r;
  }
};
template <class F>
__device__ __forceinline__ F slot_get(const uint4* sm, int comp) {
  return SlotXyzz<F>{sm, threadIdx.x}(comp);
}
template <class F>
__device__ __forceinline__ void slot_put(uint4* sm, int comp, const F& x) {
  constexpr int W = sizeof(F) / 16;
  const uint32_t* d = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
  for (int k = 0; k < W; k++) sm[(comp * W + k) * 128 + threadIdx.x] = make_uint4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
}
// own slot += q (add-2008-s, the special cases of xyzz_add).  The accumulator is overwritten before P == Q is known; in that case
// it equals q, which is intact, and is doubled from there.
template <class F, class Q>
__device__ __forceinline__ void xyzz_add_slot(uint4* sm, bool& inf, const Q& q) {
  const F zz2 = q(2);
  if (zz2.is_zero()) return;
  if (inf) {
    slot_put<F>(sm, 0, q(0)); slot_put<F>(sm, 1, q(1)); slot_put<F>(sm, 2, zz2); slot_put<F>(sm, 3, q(3));
    inf = false;
    return;
  }
  F U1, U2, S1, Pd, R;
  {
    const F zz1 = slot_get<F>(sm, 2);
    const Mul2<F> m = f_mul2(zz1, zz2, q(0), zz1);  // ZZ1 ZZ2, U2
    slot_put<F>(sm, 2, m.a);
    U2 = m.b;
  }
  {
    const F zzz1 = slot_get<F>(sm, 3), zzz2 = q(3);
    const Mul2<F> m = f_mul2(slot_get<F>(sm, 0), zz2, zzz1, zzz2);  // U1, ZZZ1 ZZZ2
    U1 = m.a;
    slot_put<F>(sm, 3, m.b);
    const Mul2<F> s = f_mul2(q(1), zzz1, slot_get<F>(sm, 1), zzz2);  // S2, S1
    S1 = s.b;
    R = f_sub(s.a, s.b);
  }
  Pd = f_sub(U2, U1);
  if (Pd.is_zero()) {
    if (R.is_zero()) {
      const XYZZ<F> d = xyzz_dbl(XYZZ<F>{q(0), q(1), q(2), q(3)});
      slot_put<F>(sm, 0, d.x); slot_put<F>(sm, 1, d.y); slot_put<F>(sm, 2, d.zz); slot_put<F>(sm, 3, d.zzz);
    } else {
      inf = true;
    }
    return;
  }
  const Mul2<F> sq = f_mul2(Pd, Pd, R, R);      // PP, R^2
  const Mul2<F> pq = f_mul2(Pd, sq.a, U1, sq.a);  // PPP, Q
  const F X3 = f_sub(f_sub(sq.b, pq.a), f_dbl(pq.b));
  slot_put<F>(sm, 0, X3);
  {
    const Mul2<F> z = f_mul2(slot_get<F>(sm, 2), sq.a, slot_get<F>(sm, 3), pq.a);
    slot_put<F>(sm, 2, z.a);
    slot_put<F>(sm, 3, z.b);
  }
  const Mul2<F> t = f_mul2(R, f_sub(pq.b, X3), S1, pq.a);  // R (Q - X3), S1 PPP
  slot_put<F>(sm, 1, f_sub(t.a, t.b));
}
template <class F> constexpr int kMarginalsMinBlocks = sizeof(F) == 32 ? 4 : sizeof(F) == 64 ? 3 : sizeof(F) == 48 ? 3 : 1;  // blocks of 128 threads per SM
template <class F>
__global__ void __launch_bounds__(128, kMarginalsMinBlocks<F>) msm_marginals_kernel(const XYZZ<F>* __restrict__ buckets, uint32_t logH, uint32_t logL,
                                                             XYZZ<F>* __restrict__ out, uint32_t mstride) {
  __shared__ uint4 acc_sm[sizeof(XYZZ<F>) / 16 * 128];
  const uint32_t H = 1u << logH, L = 1u << logL;
  const uint32_t part = threadIdx.x >> 2;
  const uint32_t marg = blockIdx.x * 4 + (threadIdx.x & 3);
  const uint32_t nmarg = H + L * kColSeg;
  buckets += (size_t)blockIdx.y << (logH + logL);
  out += (size_t)blockIdx.y * mstride;
  bool inf = true;
  constexpr int kLines = (sizeof(XYZZ<F>) + 127) / 128;
  auto add_bucket = [&](const XYZZ<F>* b, const XYZZ<F>* next) {
    if (next) {
#pragma unroll
      for (int l = 0; l < kLines; l++) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(next) + 128 * l));
    }
    xyzz_add_slot<F>(acc_sm, inf, GlobalXyzz<F>{reinterpret_cast<const uint4*>(b)});
  };
  if (marg < H) {
    const XYZZ<F>* r = buckets + ((size_t)marg << logL);
    for (uint32_t lo = part; lo < L; lo += 32) add_bucket(r + lo, lo + 32 < L ? r + lo + 32 : nullptr);
  } else if (marg < nmarg) {
    const uint32_t w = marg - H;
    const uint32_t col = w / kColSeg, seg = w % kColSeg;
    const uint32_t hi0 = (uint32_t)((uint64_t)seg * H / kColSeg), hi1 = (uint32_t)((uint64_t)(seg + 1) * H / kColSeg);
    for (uint32_t hi = hi0 + part; hi < hi1; hi += 32)
      add_bucket(buckets + ((size_t)hi << logL) + col, hi + 32 < hi1 ? buckets + ((size_t)(hi + 32) << logL) + col : nullptr);
  }
  if (inf) slot_put<F>(acc_sm, 2, F::zero());  // the fold reads other threads' slots: mark O the way XYZZ does (ZZ = 0)
  __syncthreads();
  for (uint32_t active = 64; active >= 4; active >>= 1) {
    if (threadIdx.x < active) {
      xyzz_add_slot<F>(acc_sm, inf, SlotXyzz<F>{acc_sm, threadIdx.x + active});
      if (inf) slot_put<F>(acc_sm, 2, F::zero());
    }
    __syncthreads();
  }
  if (threadIdx.x < 4 && marg < nmarg) {
    XYZZ<F> r = xyzz_inf<F>();
    if (!inf) r = XYZZ<F>{slot_get<F>(acc_sm, 0), slot_get<F>(acc_sm, 1), slot_get<F>(acc_sm, 2), slot_get<F>(acc_sm, 3)};
    out[marg] = r;
  }
}
// weights: one THREAD per partial marginal (all lanes busy, unlike a multiply on the reducing lane), then a warp tree
template <class F>
__global__ void __launch_bounds__(128) msm_weigh_kernel(const XYZZ<F>* __restrict__ marg, uint32_t logH, uint32_t logL, XYZZ<F>* __restrict__ partial,
                                                         uint32_t mstride) {
  const uint32_t H = 1u << logH, nmarg = H + (1u << logL) * kColSeg;
  marg += (size_t)blockIdx.y * mstride;
  partial += (size_t)blockIdx.y * mstride;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  XYZZ<F> acc = xyzz_inf<F>();
  if (i < nmarg) acc = xyzz_mul_small(marg[i], i < H ? (i << logL) + 1 : (i - H) / kColSeg);
  for (int delta = 16; delta >= 1; delta >>= 1) {
    XYZZ<F> other = warp_shfl_down(acc, delta);
    if (lane < (uint32_t)delta) xyzz_add(acc, other);
  }
  if (lane == 0 && i < ((nmarg + 31) & ~31u)) partial[i >> 5] = acc;  // warps past the last marginal own no slot (the next set's region follows)
}
// one warp per set: plain sum of m points into the set's result slot
template <class F>
__global__ void __launch_bounds__(32) msm_final_kernel(const XYZZ<F>* __restrict__ in, uint32_t m, uint32_t mstride, char* __restrict__ results,
                                                        ReduceSets sets) {
  const uint32_t lane = threadIdx.x;
  in += (size_t)blockIdx.x * mstride;
  XYZZ<F>* out = reinterpret_cast<XYZZ<F>*>(results + (size_t)sets.slot[blockIdx.x] * kResultSlot);
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t i = lane; i < m; i += 32) xyzz_add(acc, in[i]);
  for (int delta = 16; delta >= 1; delta >>= 1) {
    XYZZ<F> other = warp_shfl_down(acc, delta);
    if (lane < (uint32_t)delta) xyzz_add(acc, other);
  }
  if (lane == 0) *out = acc;
}

// ------------------------------------------------------------------ drivers
template <class F, class FrP>
int msm_precompute_impl(cocg_ctx* ctx, BasesEntry& be) {
  be.c = msm_plan_window_bits(be.n, FrP::BITS);
  be.nwin = msm_num_windows<FrP>(be.c);
  if (be.n == 0) return 0;
  msm_precompute_kernel<F><<<(unsigned)((be.n + 127) / 128), 128, 0, ctx->stream>>>(reinterpret_cast<Affine<F>*>(be.d), be.n, be.c, be.nwin);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

// State left in the context's scratch by the digit sort of one scalar vector; consumed by any number of accumulate +
// reduce passes over tables built with the same window width (the four Groth16 queries that take aux_assignment,
// /root/reference/co-circom/co-groth16/src/groth16.rs:221-225, 251-255, share it).
struct MsmSorted {
  size_t n = 0;
  int c = 0, nwin = 0;
  uint32_t nb = 0;
  uint32_t *sorted = nullptr, *start = nullptr, *order = nullptr, *heavy = nullptr, *chunk_owner = nullptr;
  HeavyRec* heavy_list = nullptr;
  size_t max_heavy = 0, max_chunks = 0;
};

template <class FrP>
int msm_sort_impl(cocg_ctx* ctx, const void* scalars, size_t n, int c, int mont, MsmSorted& S) {
  const int nwin = msm_num_windows<FrP>(c);
  const uint32_t nb = 1u << (c - 1);
  const size_t scan_blocks = ((size_t)nb + kScanBlock - 1) / kScanBlock;
  uint32_t *dig, *counts, *bsums, *shist;
  S.n = n; S.c = c; S.nwin = nwin; S.nb = nb;
  S.max_heavy = (size_t)nwin * n / kHeavy + 1;  // buckets with more than kHeavy entries
  S.max_chunks = (size_t)nwin * n / kHeavyChunk + S.max_heavy + 1;  // sum over heavy buckets of ceil(len / kHeavyChunk)
  void* p;
  COCG_TRY(scratch_get(ctx, 1, (size_t)nwin * n * 4, &p)); dig = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 2, (size_t)nwin * n * 4, &p)); S.sorted = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 3, (size_t)nb * 4, &p)); counts = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 4, ((size_t)nb + 1) * 4, &p)); S.start = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 5, (scan_blocks + 2) * 4, &p)); bsums = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 6, 16 + S.max_heavy * sizeof(HeavyRec), &p)); S.heavy = (uint32_t*)p;  // [0], [1] = counters, records from +16 B
  S.heavy_list = reinterpret_cast<HeavyRec*>(S.heavy + 4);
  COCG_TRY(scratch_get(ctx, 13, S.max_chunks * 4, &p)); S.chunk_owner = (uint32_t*)p;
  COCG_TRY(scratch_get(ctx, 10, ((size_t)nb + kSizeBins) * 4, &p)); S.order = (uint32_t*)p; shist = S.order + nb;
  cudaStream_t st = ctx->stream;
  ProfScope prof(ctx, COCG_PROF_MSM_SORT);
  COCG_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)nb * 4, st));
  msm_digits_kernel<FrP><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scalars, n, c, nwin, mont, dig, counts);
  COCG_LAUNCH_CHECK(ctx);
  scan_block_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, S.start, nb, bsums);
  COCG_LAUNCH_CHECK(ctx);
  scan_top_kernel<<<1, kScanThreads, 0, st>>>(bsums, scan_blocks, bsums + scan_blocks);
  COCG_LAUNCH_CHECK(ctx);
  scan_add_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(S.start, nb, bsums, bsums + scan_blocks);
  COCG_LAUNCH_CHECK(ctx);
  COCG_CUDA(ctx, cudaMemsetAsync(shist, 0, kSizeBins * 4, st));
  bucket_size_hist_kernel<<<grid_for(nb, 256, 2), 256, 0, st>>>(counts, nb, shist);
  COCG_LAUNCH_CHECK(ctx);
  bucket_size_scan_kernel<<<1, 256, 0, st>>>(shist);
  COCG_LAUNCH_CHECK(ctx);
  bucket_order_kernel<<<(nb + kOrderThreads - 1) / kOrderThreads, kOrderThreads, 0, st>>>(counts, nb, shist, S.order);
  COCG_LAUNCH_CHECK(ctx);
  msm_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dig, n, nwin, S.start, counts, S.sorted);  // drains counts[] to zero
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

// One query against a finished sort, in two halves that live in separate translation units (the group law over Fq2 makes each
// of them minutes of ptxas time): bucket accumulation into the context's bucket scratch, then the bucket reduction whose XYZZ
// result is written to d_result (device).
template <class F>
int msm_buckets_impl(cocg_ctx* ctx, const BasesEntry& be, size_t off, const MsmSorted& S, int set) {
  using X = XYZZ<F>;
  if (be.c != S.c) return fail(ctx, "cocg_msm: the table's window width differs from the sort's");
  const uint32_t nb = S.nb;
  X *buckets, *hpartial;
  void* p;
  // bucket sets of one group live side by side (slot sized by cocg_msm_multi before the first set is written)
  COCG_TRY(scratch_get(ctx, msm_bucket_slot(be.group), (size_t)(set + 1) * nb * sizeof(X), &p)); buckets = (X*)p + (size_t)set * nb;
  COCG_TRY(scratch_get(ctx, 12, S.max_chunks * sizeof(X), &p)); hpartial = (X*)p;
  const char* table = (const char*)be.d + off * be.point_bytes;  // T[w][off + i] = table[w * be.n + i]
  cudaStream_t st = ctx->stream;
  ProfScope prof(ctx, COCG_PROF_MSM_ACCUMULATE);
  COCG_CUDA(ctx, cudaMemsetAsync(S.heavy, 0, 16, st));
  constexpr int kThreads = kAccumulateThreads;
  if (sizeof(F) == 64) {  // 4 blocks x 32 KB of accumulators per SM: ask for the carve-out that holds them
    static const cudaError_t carve = cudaFuncSetAttribute(msm_accumulate_kernel<F>, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
    (void)carve;
  }
  msm_accumulate_kernel<F><<<(nb + kThreads - 1) / kThreads, kThreads, 0, st>>>(table, be.n, S.sorted, S.start, S.order, nb, buckets, S.heavy_list,
                                                                                  S.chunk_owner, S.heavy);
  COCG_LAUNCH_CHECK(ctx);
  msm_heavy_chunks_kernel<F><<<kNumSMs * 4, 128, 0, st>>>(table, be.n, S.sorted, S.start, S.heavy_list, S.chunk_owner, S.heavy, hpartial);
  COCG_LAUNCH_CHECK(ctx);
  msm_heavy_fold_kernel<F><<<kNumSMs, 128, 0, st>>>(S.heavy_list, S.heavy, hpartial, buckets);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}
// Reduces bucket sets 0 .. sets.n-1 of `group` (all built with window width c); set s goes to d_results + sets.slot[s] * kResultSlot.
template <class F>
int msm_reduce_impl(cocg_ctx* ctx, int group, int c, const ReduceSets& sets, void* d_results) {
  using X = XYZZ<F>;
  if (sets.n == 0) return 0;
  const uint32_t nb = 1u << (c - 1);
  const uint32_t logL = (uint32_t)(c - 1) / 2, logH = (uint32_t)(c - 1) - logL;
  const uint32_t nmarg = (1u << logH) + (1u << logL) * kColSeg;
  const uint32_t mstride = marg_stride(nmarg);
  X *buckets, *marg;
  void* p;
  COCG_TRY(scratch_get(ctx, msm_bucket_slot(group), (size_t)sets.n * nb * sizeof(X), &p)); buckets = (X*)p;  // filled by msm_buckets_impl
  COCG_TRY(scratch_get(ctx, 8, (size_t)sets.n * mstride * sizeof(X), &p)); marg = (X*)p;
  cudaStream_t st = ctx->stream;
  ProfScope prof(ctx, COCG_PROF_MSM_REDUCE);
  msm_marginals_kernel<F><<<dim3((nmarg + 3) / 4, sets.n), 128, 0, st>>>(buckets, logH, logL, marg, mstride);
  COCG_LAUNCH_CHECK(ctx);
  const uint32_t nwarps = (nmarg + 31) / 32;
  msm_weigh_kernel<F><<<dim3((nmarg + 127) / 128, sets.n), 128, 0, st>>>(marg, logH, logL, marg + nmarg, mstride);
  COCG_LAUNCH_CHECK(ctx);
  msm_final_kernel<F><<<sets.n, 32, 0, st>>>(marg + nmarg, nwarps, mstride, (char*)d_results, sets);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}
// host: XYZZ result (as copied back) -> Jacobian in caller memory (no alignment assumed)
template <class F>
void msm_finish_impl(const void* h_xyzz, void* out_jac) {
  XYZZ<F> x;
  memcpy(&x, h_xyzz, sizeof(x));
  Jacobian<F> jr = xyzz_to_jacobian(x);
  memcpy(out_jac, &jr, sizeof(jr));
}

// per-(curve, group) entry points, one translation unit each (msm_<curve>_<group>.cu)
#define COCG_MSM_DECL(NAME)                                                                                     \
  int msm_buckets_##NAME(cocg_ctx* ctx, const BasesEntry& be, size_t off, const MsmSorted& S, int set);          \
  int msm_reduce_##NAME(cocg_ctx* ctx, int c, const ReduceSets& sets, void* d_results);                          \
  void msm_finish_##NAME(const void* h_xyzz, void* out_jac);                                                      \
  int msm_precompute_##NAME(cocg_ctx* ctx, BasesEntry& be);
COCG_MSM_DECL(bn254_g1)
COCG_MSM_DECL(bn254_g2)
COCG_MSM_DECL(bls381_g1)
COCG_MSM_DECL(bls381_g2)

}  // namespace cocg
