// C-ABI entry points of the MSM path (bases residency + dispatch); the pipeline itself is msm_impl.cuh.
#include <string.h>

#include "ctx.cuh"

namespace cocg {
#define COCG_MSM_DECL(NAME)                                                                                                           \
  int msm_##NAME(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac); \
  int msm_precompute_##NAME(cocg_ctx* ctx, BasesEntry& be);
COCG_MSM_DECL(bn254_g1)
COCG_MSM_DECL(bn254_g2)
COCG_MSM_DECL(bls381_g1)
COCG_MSM_DECL(bls381_g2)

// window bits / count for a query of n points: must match msm_impl.cuh (the table is allocated before it is built)
int msm_table_windows(int curve, size_t n) {
  int c = msm_plan_window_bits(n);
  int bits = curve == COCG_BN254 ? 254 : 255;
  return (bits + c) / c;
}
int msm_precompute(cocg_ctx* ctx, BasesEntry& be) {
  if (ctx->curve == COCG_BN254) return be.group == COCG_G1 ? msm_precompute_bn254_g1(ctx, be) : msm_precompute_bn254_g2(ctx, be);
  return be.group == COCG_G1 ? msm_precompute_bls381_g1(ctx, be) : msm_precompute_bls381_g2(ctx, be);
}

template <class F>
__global__ void __launch_bounds__(256) bases_to_mont_kernel(void* pts, size_t ncoords) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncoords) return;
  F* p = reinterpret_cast<F*>(pts);
  p[i] = fp_to_mont(p[i]);
}

}  // namespace cocg


using namespace cocg;

static size_t coord_bytes(int curve) { return curve == COCG_BN254 ? 32 : 48; }

extern "C" int cocg_bases_upload(cocg_ctx* ctx, int group, const void* pts, size_t n, size_t stride, int mont, uint64_t* handle) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (group != COCG_G1 && group != COCG_G2) return fail(ctx, "cocg_bases_upload: group must be 1 or 2");
  if (!handle || (n && !pts)) return fail(ctx, "cocg_bases_upload: null argument");
  size_t pb = coord_bytes(ctx->curve) * 2 * group;
  if (stride < pb) return fail(ctx, "cocg_bases_upload: stride smaller than a packed point");
  BasesEntry be;
  be.n = n; be.group = group; be.point_bytes = pb;
  COCG_CUDA(ctx, cudaMalloc(&be.d, n ? (size_t)msm_table_windows(ctx->curve, n) * n * pb : 16));
  if (n) {
    COCG_CUDA(ctx, cudaMemcpy2DAsync(be.d, pb, pts, stride, pb, n, cudaMemcpyHostToDevice, ctx->stream));
    if (!mont) {
      size_t ncoords = n * 2 * group;
      if (ctx->curve == COCG_BN254) bases_to_mont_kernel<Bn254Fq><<<(unsigned)((ncoords + 255) / 256), 256, 0, ctx->stream>>>(be.d, ncoords);
      else bases_to_mont_kernel<Bls381Fq><<<(unsigned)((ncoords + 255) / 256), 256, 0, ctx->stream>>>(be.d, ncoords);
      COCG_LAUNCH_CHECK(ctx);
    }
  }
  COCG_TRY(msm_precompute(ctx, be));  // T[j] = 2^(c*j) * bases, j < nwin
  if (be.nwin != msm_table_windows(ctx->curve, n)) return fail(ctx, "cocg_bases_upload: window plan mismatch");
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // reuse a free slot if any
  for (size_t i = 0; i < ctx->bases.size(); i++)
    if (!ctx->bases[i].d) { ctx->bases[i] = be; *handle = i + 1; return 0; }
  ctx->bases.push_back(be);
  *handle = ctx->bases.size();
  return 0;
}

extern "C" int cocg_bases_free(cocg_ctx* ctx, uint64_t handle) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (handle == 0 || handle > ctx->bases.size() || !ctx->bases[handle - 1].d) return fail(ctx, "cocg_bases_free: bad handle");
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->bases[handle - 1].owned) COCG_CUDA(ctx, cudaFree(ctx->bases[handle - 1].d));
  ctx->bases[handle - 1] = BasesEntry();
  return 0;
}

extern "C" int cocg_bases_share(cocg_ctx* ctx, cocg_ctx* owner, uint64_t owner_handle, uint64_t* handle) {
  if (!ctx) return 1;
  if (!owner || !handle) return fail(ctx, "cocg_bases_share: null argument");
  if (owner->device != ctx->device || owner->curve != ctx->curve) return fail(ctx, "cocg_bases_share: contexts differ in device or curve");
  if (owner_handle == 0 || owner_handle > owner->bases.size() || !owner->bases[owner_handle - 1].d) return fail(ctx, "cocg_bases_share: bad handle");
  BasesEntry be = owner->bases[owner_handle - 1];
  be.owned = false;
  for (size_t i = 0; i < ctx->bases.size(); i++)
    if (!ctx->bases[i].d) { ctx->bases[i] = be; *handle = i + 1; return 0; }
  ctx->bases.push_back(be);
  *handle = ctx->bases.size();
  return 0;
}

extern "C" int cocg_msm(cocg_ctx* ctx, uint64_t bases, size_t off, size_t n, const void* const* scalars, int k, int scalars_mont,
                        void* out_jacobian) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bases == 0 || bases > ctx->bases.size() || !ctx->bases[bases - 1].d) return fail(ctx, "cocg_msm: bad bases handle");
  if (k <= 0) return 0;
  if (!scalars || !out_jacobian) return fail(ctx, "cocg_msm: null argument");
  const BasesEntry& be = ctx->bases[bases - 1];
  if (off > be.n || n > be.n - off) return fail(ctx, "cocg_msm: range exceeds the uploaded bases");
  for (int j = 0; j < k; j++)
    if (n && !scalars[j]) return fail(ctx, "cocg_msm: null scalar vector");
  if (ctx->curve == COCG_BN254) {
    if (be.group == COCG_G1) return msm_bn254_g1(ctx, be, off, n, scalars, k, scalars_mont, out_jacobian);
    return msm_bn254_g2(ctx, be, off, n, scalars, k, scalars_mont, out_jacobian);
  }
  if (be.group == COCG_G1) return msm_bls381_g1(ctx, be, off, n, scalars, k, scalars_mont, out_jacobian);
  return msm_bls381_g2(ctx, be, off, n, scalars, k, scalars_mont, out_jacobian);
}

extern "C" int cocg_msm_host(cocg_ctx* ctx, uint64_t bases, size_t off, size_t n, const void* const* scalars, int k, int scalars_mont,
                             void* out_jacobian) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (k <= 0) return 0;
  if (k > 8) return fail(ctx, "cocg_msm_host: at most 8 components");
  if (!scalars) return fail(ctx, "cocg_msm_host: null argument");
  void* dev = nullptr;
  COCG_TRY(scratch_get(ctx, 11, (size_t)k * n * 32 + 32, &dev));
  const void* dptr[8];
  for (int j = 0; j < k; j++) {
    if (n && !scalars[j]) return fail(ctx, "cocg_msm_host: null scalar vector");
    dptr[j] = (char*)dev + (size_t)j * n * 32;
    if (n) COCG_CUDA(ctx, cudaMemcpyAsync((void*)dptr[j], scalars[j], n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  return cocg_msm(ctx, bases, off, n, dptr, k, scalars_mont, out_jacobian);
}
