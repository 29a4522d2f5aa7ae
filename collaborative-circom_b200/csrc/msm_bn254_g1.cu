// One (curve, group) instantiation of the MSM pipeline per translation unit, so that they compile in parallel.
#include "msm_impl.cuh"
namespace cocg {
int msm_bn254_g1(cocg_ctx* ctx, const BasesEntry& be, size_t off, size_t n, const void* const* scalars, int k, int mont, void* out_jac) {
  return msm_impl<Bn254Fq, Bn254FrP>(ctx, be, off, n, scalars, k, mont, out_jac);
}
int msm_precompute_bn254_g1(cocg_ctx* ctx, BasesEntry& be) { return msm_precompute_impl<Bn254Fq, Bn254FrP>(ctx, be); }
}  // namespace cocg
