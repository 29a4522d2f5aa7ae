// Bucket reduction, result conversion and table precompute of one (curve, group).
#include "msm_impl.cuh"
namespace cocg {
int msm_reduce_bls381_g2(cocg_ctx* ctx, const MsmSorted& S, void* d_result) { return msm_reduce_impl<Bls381Fq2>(ctx, S, d_result); }
void msm_finish_bls381_g2(const void* h_xyzz, void* out_jac) { msm_finish_impl<Bls381Fq2>(h_xyzz, out_jac); }
int msm_precompute_bls381_g2(cocg_ctx* ctx, BasesEntry& be) { return msm_precompute_impl<Bls381Fq2, Bls381FrP>(ctx, be); }
}  // namespace cocg
