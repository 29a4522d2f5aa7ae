// Bucket accumulation of one (curve, group): one translation unit per instantiation and per half, so that they compile in parallel.
#include "msm_impl.cuh"
namespace cocg {
int msm_buckets_bn254_g2(cocg_ctx* ctx, const BasesEntry& be, size_t off, const MsmSorted& S, int set) { return msm_buckets_impl<Bn254Fq2>(ctx, be, off, S, set); }
}  // namespace cocg
