// On-curve and prime-order-subgroup validation of resident bases, replacing the per-point checks the reference runs
// inside its rayon loop while parsing a zkey: `g1_vec_from_reader` / `g2_vec_from_reader`
// (/root/reference/co-circom/circom-types/src/traits.rs:555-570) -> `g1_from_reader` (traits.rs:107-155): `p.is_on_curve()` and
// `p.is_in_correct_subgroup_assuming_on_curve()`, failing with SerializationError::InvalidData.  (0, 0) is the point at infinity.
//
// One thread per point: y^2 == x^3 + b, then [r]P == O by double-and-add over the constant bits of the group order (uniform control
// flow: every lane takes the same branch at every bit).  BN254 G1 has cofactor 1, so every curve point is in the group and arkworks'
// check is vacuous there; G2 of both curves and BLS12-381 G1 have cofactors != 1.  The answer (member / not a member) does not depend
// on how membership is decided, so the plain multiplication by r agrees with arkworks' endomorphism-based tests.
// Cost: ~2,600 (G1) / ~8,000 (G2) base-field products per point, once per zkey load.
#include <string.h>

#include "ctx.cuh"

namespace cocg {

template <class F, class FrP>
__global__ void __launch_bounds__(128) bases_check_kernel(const Affine<F>* __restrict__ pts, size_t n, F b, int subgroup, unsigned long long* __restrict__ bad /* [0] count, [1] first index */) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = pts[i];
  if (p.is_inf()) return;
  bool ok = f_sqr(p.y) == f_add(f_mul(f_sqr(p.x), p.x), b);
  if (ok && subgroup) {
    XYZZ<F> acc = xyzz_inf<F>();
    for (int bit = FrP::BITS - 1; bit >= 0; bit--) {
      acc = xyzz_dbl(acc);
      if ((FrP::mod(bit >> 5) >> (bit & 31)) & 1) xyzz_madd(acc, p);
    }
    ok = acc.is_inf();
  }
  if (!ok) {
    atomicAdd(&bad[0], 1ull);
    atomicMin(&bad[1], (unsigned long long)i);
  }
}

template <class F, class FrP>
static int bases_check_impl(cocg_ctx* ctx, const BasesEntry& be, const F& b, int subgroup, size_t* n_bad, size_t* first_bad) {
  void* d;
  COCG_TRY(scratch_get(ctx, 15, 64, &d));
  unsigned long long init[2] = {0ull, ~0ull}, res[2];
  COCG_CUDA(ctx, cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, ctx->stream));
  if (be.n) {
    bases_check_kernel<F, FrP><<<(unsigned)((be.n + 127) / 128), 128, 0, ctx->stream>>>(reinterpret_cast<const Affine<F>*>(be.d), be.n, b, subgroup, (unsigned long long*)d);
    COCG_LAUNCH_CHECK(ctx);
  }
  COCG_CUDA(ctx, cudaMemcpyAsync(res, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_bad = (size_t)res[0];
  if (first_bad) *first_bad = (size_t)res[1];
  return 0;
}

template <class P>
static Fp<P> fq_const(const uint32_t* v) {
  Fp<P> r;
  for (int i = 0; i < P::N; i++) r.l[i] = v[i];
  return r;
}

}  // namespace cocg

using namespace cocg;

extern "C" int cocg_bases_check(cocg_ctx* ctx, uint64_t handle, int check_subgroup, size_t* n_bad, size_t* first_bad) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!n_bad) return fail(ctx, "cocg_bases_check: null argument");
  if (handle == 0 || handle > ctx->bases.size() || !ctx->bases[handle - 1].d) return fail(ctx, "cocg_bases_check: bad handle");
  const BasesEntry& be = ctx->bases[handle - 1];
  if (ctx->curve == COCG_BN254) {
    static const uint32_t b1[8] = BN254_G1_B, c0[8] = BN254_G2_B_C0, c1[8] = BN254_G2_B_C1;
    // BN254 G1: cofactor 1 -- membership follows from the curve equation
    if (be.group == COCG_G1) return bases_check_impl<Bn254Fq, Bn254FrP>(ctx, be, fq_const<Bn254FqP>(b1), 0, n_bad, first_bad);
    return bases_check_impl<Bn254Fq2, Bn254FrP>(ctx, be, Bn254Fq2{fq_const<Bn254FqP>(c0), fq_const<Bn254FqP>(c1)}, check_subgroup, n_bad, first_bad);
  }
  static const uint32_t b1[12] = BLS381_G1_B, c0[12] = BLS381_G2_B_C0, c1[12] = BLS381_G2_B_C1;
  if (be.group == COCG_G1) return bases_check_impl<Bls381Fq, Bls381FrP>(ctx, be, fq_const<Bls381FqP>(b1), check_subgroup, n_bad, first_bad);
  return bases_check_impl<Bls381Fq2, Bls381FrP>(ctx, be, Bls381Fq2{fq_const<Bls381FqP>(c0), fq_const<Bls381FqP>(c1)}, check_subgroup, n_bad, first_bad);
}
