// Bucket reduction, result conversion and table precompute of one (curve, group).
#include "msm_impl.cuh"
namespace cocg {
int msm_reduce_bls381_g1(cocg_ctx* ctx, int c, const ReduceSets& sets, void* d_results) { return msm_reduce_impl<Bls381Fq>(ctx, COCG_G1, c, sets, d_results); }
void msm_finish_bls381_g1(const void* h_xyzz, void* out_jac) { msm_finish_impl<Bls381Fq>(h_xyzz, out_jac); }
int msm_precompute_bls381_g1(cocg_ctx* ctx, BasesEntry& be) { return msm_precompute_impl<Bls381Fq, Bls381FrP>(ctx, be); }
}  // namespace cocg
