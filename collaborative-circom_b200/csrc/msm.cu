// C-ABI entry points of the MSM path (bases residency + dispatch); the pipeline itself is msm_impl.cuh.
#include <string.h>

#include <vector>

#include "msm_impl.cuh"

namespace cocg {
// window bits / count for a query of n points: must match msm_impl.cuh (the table is allocated before it is built)
int msm_table_windows(int curve, size_t n) {
  int bits = curve == COCG_BN254 ? 254 : 255;
  int c = msm_plan_window_bits(n, bits);
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

extern "C" int cocg_msm_plan(int curve, size_t n, int* window_bits, int* windows) {
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return 1;
  const int bits = curve == COCG_BN254 ? 254 : 255;
  const int c = msm_plan_window_bits(n, bits);
  if (window_bits) *window_bits = c;
  if (windows) *windows = (bits + c) / c;
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

namespace {
struct MsmOps {
  int (*buckets)(cocg_ctx*, const BasesEntry&, size_t, const MsmSorted&, int);
  int (*reduce)(cocg_ctx*, int, const ReduceSets&, void*);
  void (*finish)(const void*, void*);
  size_t xyzz_bytes, jac_bytes;
};
MsmOps msm_ops(int curve, int group) {
  const size_t cb = curve == COCG_BN254 ? 32 : 48;
  MsmOps o;
  if (curve == COCG_BN254) {
    o.buckets = group == COCG_G1 ? msm_buckets_bn254_g1 : msm_buckets_bn254_g2;
    o.reduce = group == COCG_G1 ? msm_reduce_bn254_g1 : msm_reduce_bn254_g2;
    o.finish = group == COCG_G1 ? msm_finish_bn254_g1 : msm_finish_bn254_g2;
  } else {
    o.buckets = group == COCG_G1 ? msm_buckets_bls381_g1 : msm_buckets_bls381_g2;
    o.reduce = group == COCG_G1 ? msm_reduce_bls381_g1 : msm_reduce_bls381_g2;
    o.finish = group == COCG_G1 ? msm_finish_bls381_g1 : msm_finish_bls381_g2;
  }
  o.xyzz_bytes = 4 * cb * group;
  o.jac_bytes = 3 * cb * group;
  return o;
}
}  // namespace

extern "C" int cocg_msm_multi(cocg_ctx* ctx, const uint64_t* bases, const size_t* offs, int nq, size_t n, const void* const* scalars, int k,
                              int scalars_mont, void* const* out_jacobian) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (nq <= 0 || k <= 0) return 0;
  if (nq > 16 || k > 8) return fail(ctx, "cocg_msm_multi: at most 16 queries and 8 components");
  if (!bases || !offs || !scalars || !out_jacobian) return fail(ctx, "cocg_msm_multi: null argument");
  const BasesEntry* be[16];
  for (int q = 0; q < nq; q++) {
    if (bases[q] == 0 || bases[q] > ctx->bases.size() || !ctx->bases[bases[q] - 1].d) return fail(ctx, "cocg_msm: bad bases handle");
    be[q] = &ctx->bases[bases[q] - 1];
    if (offs[q] > be[q]->n || n > be[q]->n - offs[q]) return fail(ctx, "cocg_msm: range exceeds the uploaded bases");
    if (!out_jacobian[q]) return fail(ctx, "cocg_msm: null output");
    if (be[q]->n >= ((size_t)1 << kIdxBits)) return fail(ctx, "cocg_msm: at most 2^25 - 1 bases per query");
  }
  for (int j = 0; j < k; j++)
    if (n && !scalars[j]) return fail(ctx, "cocg_msm: null scalar vector");
  if (n == 0) {  // empty sum: arkworks' zero()
    for (int q = 0; q < nq; q++) {
      MsmOps o = msm_ops(ctx->curve, be[q]->group);
      std::vector<uint8_t> inf(o.xyzz_bytes, 0);
      for (int j = 0; j < k; j++) o.finish(inf.data(), (char*)out_jacobian[q] + (size_t)j * o.jac_bytes);
    }
    return 0;
  }
  void* d_res;
  COCG_TRY(scratch_get(ctx, 9, (size_t)nq * k * kResultSlot, &d_res));
  // Queries are grouped by window width so that each group shares one digit sort per component.  Every (query, component)
  // accumulates into its own bucket set; the sets of one curve group are reduced together, kMaxSets per launch (msm_impl.cuh).
  bool done[16] = {};
  for (int q0 = 0; q0 < nq; q0++) {
    if (done[q0]) continue;
    const int c = be[q0]->c;
    const size_t nb = (size_t)1 << (c - 1);
    ReduceSets pend[3] = {};  // indexed by group (COCG_G1 = 1, COCG_G2 = 2)
    for (int g = COCG_G1; g <= COCG_G2; g++) {  // size the bucket arenas before the first set is written (growth does not preserve contents)
      size_t cnt = 0;
      for (int q = q0; q < nq; q++) cnt += (be[q]->c == c && be[q]->group == g) ? (size_t)k : 0;
      if (cnt > (size_t)kMaxSets) cnt = kMaxSets;
      void* p;
      if (cnt) COCG_TRY(scratch_get(ctx, msm_bucket_slot(g), cnt * nb * msm_ops(ctx->curve, g).xyzz_bytes, &p));
    }
    for (int j = 0; j < k; j++) {
      MsmSorted S;
      if (ctx->curve == COCG_BN254) COCG_TRY(msm_sort_impl<Bn254FrP>(ctx, scalars[j], n, c, scalars_mont, S));
      else COCG_TRY(msm_sort_impl<Bls381FrP>(ctx, scalars[j], n, c, scalars_mont, S));
      for (int q = q0; q < nq; q++) {
        if (be[q]->c != c) continue;
        const int g = be[q]->group;
        MsmOps o = msm_ops(ctx->curve, g);
        ReduceSets& P = pend[g];
        COCG_TRY(o.buckets(ctx, *be[q], offs[q], S, (int)P.n));
        P.slot[P.n++] = (uint32_t)(q * k + j);
        if (P.n == (uint32_t)kMaxSets) {
          COCG_TRY(o.reduce(ctx, c, P, d_res));
          P.n = 0;
        }
      }
    }
    for (int g = COCG_G1; g <= COCG_G2; g++)
      if (pend[g].n) COCG_TRY(msm_ops(ctx->curve, g).reduce(ctx, c, pend[g], d_res));
    for (int q = q0; q < nq; q++)
      if (be[q]->c == c) done[q] = true;
  }
  void* hp;
  COCG_TRY(pinned_get(ctx, (size_t)nq * k * kResultSlot, &hp));
  COCG_CUDA(ctx, cudaMemcpyAsync(hp, d_res, (size_t)nq * k * kResultSlot, cudaMemcpyDeviceToHost, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < nq; q++) {
    MsmOps o = msm_ops(ctx->curve, be[q]->group);
    for (int j = 0; j < k; j++) o.finish((const char*)hp + ((size_t)q * k + j) * kResultSlot, (char*)out_jacobian[q] + (size_t)j * o.jac_bytes);
  }
  return 0;
}

extern "C" int cocg_msm(cocg_ctx* ctx, uint64_t bases, size_t off, size_t n, const void* const* scalars, int k, int scalars_mont,
                        void* out_jacobian) {
  if (!ctx) return 1;
  if (k <= 0) return 0;
  if (!out_jacobian) return fail(ctx, "cocg_msm: null argument");
  void* outs[1] = {out_jacobian};
  return cocg_msm_multi(ctx, &bases, &off, 1, n, scalars, k, scalars_mont, outs);
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
