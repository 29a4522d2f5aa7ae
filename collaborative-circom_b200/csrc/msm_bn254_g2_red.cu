// Bucket reduction, result conversion and table precompute of one (curve, group).
#include "msm_impl.cuh"
namespace cocg {
int msm_reduce_bn254_g2(cocg_ctx* ctx, int c, const ReduceSets& sets, void* d_results) { return msm_reduce_impl<Bn254Fq2>(ctx, COCG_G2, c, sets, d_results); }
void msm_finish_bn254_g2(const void* h_xyzz, void* out_jac) { msm_finish_impl<Bn254Fq2>(h_xyzz, out_jac); }
int msm_precompute_bn254_g2(cocg_ctx* ctx, BasesEntry& be) { return msm_precompute_impl<Bn254Fq2, Bn254FrP>(ctx, be); }
}  // namespace cocg
