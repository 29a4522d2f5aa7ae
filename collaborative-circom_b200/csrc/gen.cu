// Synthetic MSM bases generated in HBM: bases[i] = P0 + i*Q, affine, for two public points P0, Q derived from a seed.
//
// The reference has no counterpart (its query arrays come from a snarkjs zkey, circom-types/src/groth16/zkey.rs:139-251);
// BASELINE.json's 2^20-constraint configurations have no shipped zkey and there is no circom/snarkjs here to make one, so
// bench.py and the full-size tests need 2^20..2^22 valid, distinct curve points per query without a CPU-side generator.
// Each thread walks kGenRun consecutive points with mixed additions and normalises them with one shared inversion
// (Montgomery's trick on ZZZ; 1/ZZ = (ZZ/ZZZ)^2).
#include <string.h>

#include "ctx.cuh"
#include "prf.cuh"

namespace cocg {
int msm_table_windows(int curve, size_t n);
int msm_precompute(cocg_ctx* ctx, BasesEntry& be);

constexpr int kGenRun = 16;

template <class F>
__global__ void __launch_bounds__(128) bases_generate_kernel(Affine<F> p0, Affine<F> q, size_t first, size_t n, Affine<F>* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kGenRun;
  if (lo >= n) return;
  // start = p0 + (first + lo)*q
  XYZZ<F> step = xyzz_from_affine(q), acc = xyzz_from_affine(p0);
  {
    XYZZ<F> m = xyzz_inf<F>();
    uint64_t k = first + lo;
    for (int b = 63 - __clzll(k | 1); b >= 0; b--) {
      m = xyzz_dbl(m);
      if ((k >> b) & 1) xyzz_add(m, step);
    }
    xyzz_add(acc, m);
  }
  XYZZ<F> pts[kGenRun];
  F pref[kGenRun];
  F run = F::one();
  int cnt = (int)((n - lo < (size_t)kGenRun) ? n - lo : kGenRun);
  for (int j = 0; j < cnt; j++) {
    pts[j] = acc;
    pref[j] = run;
    if (!acc.is_inf()) run = f_mul(run, acc.zzz);
    xyzz_madd(acc, q);
  }
  F inv = f_inv(run);
  for (int j = cnt - 1; j >= 0; j--) {
    Affine<F> a;
    if (pts[j].is_inf()) {
      a.x = F::zero();
      a.y = F::zero();
    } else {
      F w = f_mul(inv, pref[j]);          // 1 / zzz_j
      inv = f_mul(inv, pts[j].zzz);
      F zi = f_mul(pts[j].zz, w);         // 1 / z
      a.x = f_mul(pts[j].x, f_sqr(zi));
      a.y = f_mul(pts[j].y, w);
    }
    out[lo + j] = a;
  }
}

template <class F>
static int generate_impl(cocg_ctx* ctx, int group, size_t first, size_t n, const uint8_t seed[32], void* d) {
  // P0 = k0*G, Q = k1*G with k0, k1 from the field PRF; computed on the host with the O(1) group operations
  uint64_t gen[36], p0j[36], qj[36], aff[24];
  uint32_t k0[8], k1[8];
  COCG_TRY(cocg_prf_field_host(ctx->curve, seed, 0x67656e30u, 0, k0));
  COCG_TRY(cocg_prf_field_host(ctx->curve, seed, 0x67656e31u, 0, k1));
  COCG_TRY(cocg_ec_op(ctx, group, 6, nullptr, nullptr, gen));
  COCG_TRY(cocg_ec_op(ctx, group, 1, gen, k0, p0j));
  COCG_TRY(cocg_ec_op(ctx, group, 1, gen, k1, qj));
  Affine<F> p0, q;
  COCG_TRY(cocg_ec_op(ctx, group, 2, p0j, nullptr, aff));
  memcpy(&p0, aff, sizeof(p0));
  COCG_TRY(cocg_ec_op(ctx, group, 2, qj, nullptr, aff));
  memcpy(&q, aff, sizeof(q));
  size_t threads = (n + kGenRun - 1) / kGenRun;
  bases_generate_kernel<F><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(p0, q, first, n, reinterpret_cast<Affine<F>*>(d));
  COCG_LAUNCH_CHECK(ctx);
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // namespace cocg

using namespace cocg;

extern "C" int cocg_bases_generate(cocg_ctx* ctx, int group, size_t n, const void* seed, uint64_t* handle) {
  return cocg_bases_generate_range(ctx, group, 0, n, seed, handle);
}

extern "C" int cocg_bases_generate_range(cocg_ctx* ctx, int group, size_t first, size_t n, const void* seed, uint64_t* handle) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (group != COCG_G1 && group != COCG_G2) return fail(ctx, "cocg_bases_generate: group must be 1 or 2");
  if (!handle || !seed) return fail(ctx, "cocg_bases_generate: null argument");
  BasesEntry be;
  be.n = n;
  be.group = group;
  be.point_bytes = (ctx->curve == COCG_BN254 ? 32 : 48) * 2 * (size_t)group;
  COCG_CUDA(ctx, cudaMalloc(&be.d, n ? (size_t)msm_table_windows(ctx->curve, n) * n * be.point_bytes : 16));
  if (n) {
    int rc;
    const uint8_t* sd = (const uint8_t*)seed;
    if (ctx->curve == COCG_BN254) rc = group == COCG_G1 ? generate_impl<Bn254Fq>(ctx, group, first, n, sd, be.d) : generate_impl<Bn254Fq2>(ctx, group, first, n, sd, be.d);
    else rc = group == COCG_G1 ? generate_impl<Bls381Fq>(ctx, group, first, n, sd, be.d) : generate_impl<Bls381Fq2>(ctx, group, first, n, sd, be.d);
    if (rc) { cudaFree(be.d); return rc; }
  }
  COCG_TRY(msm_precompute(ctx, be));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < ctx->bases.size(); i++)
    if (!ctx->bases[i].d) { ctx->bases[i] = be; *handle = i + 1; return 0; }
  ctx->bases.push_back(be);
  *handle = ctx->bases.size();
  return 0;
}

extern "C" int cocg_bases_download(cocg_ctx* ctx, uint64_t handle, size_t off, size_t n, void* out) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (handle == 0 || handle > ctx->bases.size() || !ctx->bases[handle - 1].d) return fail(ctx, "cocg_bases_download: bad handle");
  const BasesEntry& be = ctx->bases[handle - 1];
  if (off > be.n || n > be.n - off) return fail(ctx, "cocg_bases_download: range exceeds the bases");
  if (n == 0) return 0;
  if (!out) return fail(ctx, "cocg_bases_download: null output");
  COCG_CUDA(ctx, cudaMemcpyAsync(out, (const char*)be.d + off * be.point_bytes, n * be.point_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
