// Element-wise share-vector kernels and the power-table builder.
//
// Reference call sites replaced (paths under /root/reference):
//   add_vec / sub_assign_vec / neg_vec_in_place      mpc-core/src/protocols/rep3.rs:581-593, 634-648, 672-679
//   mul_vec local step                               mpc-core/src/protocols/rep3.rs:656-660 (shamir.rs:618-621, plain.rs:224)
//   distribute_powers_and_mul_by_const               mpc-core/src/protocols/rep3.rs:681-688
// The reference runs these as serial iterators on one core; here one thread owns one 32-byte element
// (two 128-bit loads per operand, a warp covers 1 KiB contiguous per operand), grids are a multiple of the
// 148 SMs and grid-stride over the vector.  add/sub/neg are HBM-bound (96 B/element); the multiplications
// are bound by the integer-multiply pipe.
#include <string.h>

#include "ctx.cuh"
#include "prf.cuh"

namespace cocg {

template <class P, int OP>
__global__ void __launch_bounds__(256) vec_op_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                      void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> x = load_fp<P>(a, i);
    Fp<P> r;
    if (OP == COCG_OP_MUL) r = fp_mul(x, load_fp<P>(b, i));
    else if (OP == COCG_OP_ADD) r = fp_add(x, load_fp<P>(b, i));
    else if (OP == COCG_OP_SUB) r = fp_sub(x, load_fp<P>(b, i));
    else if (OP == COCG_OP_NEG) r = fp_neg(x);
    else if (OP == COCG_OP_TO_MONT) r = fp_to_mont(x);
    else r = fp_from_mont(x);
    store_fp<P>(out, i, r);
  }
}

// out[i] = aa*ba + aa*bb + ab*ba (+ mask)  ==  aa*(ba+bb) + ab*ba (+ mask): two products instead of three,
// same value mod r, hence bit-identical canonical output.
template <class P, bool MASK>
__global__ void __launch_bounds__(256) rep3_mul_local_kernel(const void* __restrict__ aa, const void* __restrict__ ab,
                                                              const void* __restrict__ ba, const void* __restrict__ bb,
                                                              const void* __restrict__ mask, void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> xa = load_fp<P>(aa, i), xb = load_fp<P>(ab, i), ya = load_fp<P>(ba, i), yb = load_fp<P>(bb, i);
    Fp<P> r = fp_add(fp_mul(xa, fp_add(ya, yb)), fp_mul(xb, ya));
    if (MASK) r = fp_add(r, load_fp<P>(mask, i));
    store_fp<P>(out, i, r);
  }
}

// Same with the zero-mask generated in the kernel: mask[i] = F(k_own, ctr, i) - F(k_prev, ctr, i), the counter-addressed
// form of masking_field_element (rep3/rngs.rs:37-46; see prf.cuh).  No mask vector ever exists in HBM.
template <class P>
__global__ void __launch_bounds__(256) rep3_mul_local_prf_kernel(const void* __restrict__ aa, const void* __restrict__ ab,
                                                                  const void* __restrict__ ba, const void* __restrict__ bb,
                                                                  PrfKey k_own, PrfKey k_prev, uint32_t ctr, void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> xa = load_fp<P>(aa, i), xb = load_fp<P>(ab, i), ya = load_fp<P>(ba, i), yb = load_fp<P>(bb, i);
    Fp<P> r = fp_add(fp_mul(xa, fp_add(ya, yb)), fp_mul(xb, ya));
    r = fp_add(r, fp_sub(prf_field<P>(k_own, ctr, i), prf_field<P>(k_prev, ctr, i)));
    store_fp<P>(out, i, r);
  }
}
template <class P>
__global__ void __launch_bounds__(256) prf_fill_kernel(PrfKey key, uint32_t ctr, void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store_fp<P>(out, i, prf_field<P>(key, ctr, i));
}

// out[i] = a * x[i] (+ y[i]): the linear combinations of the Shamir driver (Lagrange interpolation at the king, polynomial
// re-sharing, Vandermonde extraction: shamir.rs:302-384, 904-1010, shamir/shamir_core.rs:8-33)
template <class P, bool ADD>
__global__ void __launch_bounds__(256) vec_axpy_kernel(Fp<P> a, const void* __restrict__ x, const void* __restrict__ y, void* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Fp<P> r = fp_mul(a, load_fp<P>(x, i));
    if (ADD) r = fp_add(r, load_fp<P>(y, i));
    store_fp<P>(out, i, r);
  }
}

// out[i] = first * base^i.  Thread t owns a run of RUN consecutive exponents: one square-and-multiply to
// reach base^(t*RUN), then RUN-1 dependent products.
constexpr int kPowRun = 32;
template <class P>
__global__ void __launch_bounds__(128) powers_kernel(void* __restrict__ out, size_t n, Fp<P> base, Fp<P> first) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kPowRun;
  if (lo >= n) return;
  Fp<P> pw = fp_mul(first, fp_pow_u64(base, (uint64_t)lo));
  size_t hi = lo + kPowRun < n ? lo + kPowRun : n;
  for (size_t i = lo; i < hi; i++) {
    store_fp<P>(out, i, pw);
    pw = fp_mul(pw, base);
  }
}

// x[i] *= first * base^i without a table: (base, first) are per-proof challenges in the CoPlonk rounds (division by X - xi), and a
// cached table per challenge would never be reused.
template <class P>
__global__ void __launch_bounds__(128) scale_powers_kernel(void* __restrict__ x, size_t n, Fp<P> base, Fp<P> first) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * kPowRun;
  if (lo >= n) return;
  Fp<P> pw = fp_mul(first, fp_pow_u64(base, (uint64_t)lo));
  size_t hi = lo + kPowRun < n ? lo + kPowRun : n;
  for (size_t i = lo; i < hi; i++) {
    store_fp<P>(x, i, fp_mul(load_fp<P>(x, i), pw));
    pw = fp_mul(pw, base);
  }
}
template <class P>
static int scale_powers_impl(cocg_ctx* ctx, void* x, size_t n, const void* g, const void* c) {
  Fp<P> b, f;
  memcpy(b.l, g, 32);
  memcpy(f.l, c, 32);
  size_t threads = (n + kPowRun - 1) / kPowRun;
  ProfScope prof(ctx, COCG_PROF_VEC);
  scale_powers_kernel<P><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(x, n, b, f);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int powers_table_impl(cocg_ctx* ctx, uint32_t kind, size_t n, const uint32_t* base, const uint32_t* first, void** out) {
  std::array<uint32_t, 18> key;
  key[0] = kind;
  key[1] = (uint32_t)n;
  for (int i = 0; i < 8; i++) { key[2 + i] = base[i]; key[10 + i] = first[i]; }
  auto it = ctx->tables.find(key);
  if (it != ctx->tables.end()) { *out = it->second; return 0; }
  void* d = nullptr;
  COCG_CUDA(ctx, cudaMalloc(&d, (n ? n : 1) * sizeof(Fp<P>)));
  Fp<P> b, f;
  for (int i = 0; i < 8; i++) { b.l[i] = base[i]; f.l[i] = first[i]; }
  size_t threads = (n + kPowRun - 1) / kPowRun;
  if (n) {
    powers_kernel<P><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(d, n, b, f);
    COCG_LAUNCH_CHECK(ctx);
  }
  ctx->tables[key] = d;
  *out = d;
  return 0;
}
int powers_table(cocg_ctx* ctx, uint32_t kind, size_t n, const uint32_t* base, const uint32_t* first, void** out) {
  return ctx->curve == COCG_BN254 ? powers_table_impl<Bn254FrP>(ctx, kind, n, base, first, out)
                                  : powers_table_impl<Bls381FrP>(ctx, kind, n, base, first, out);
}

template <class P>
static int vec_op_impl(cocg_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n) {
  if (n == 0) return 0;
  int grid = grid_for(n, 256, 8);
  ProfScope prof(ctx, COCG_PROF_VEC);
  switch (op) {
    case COCG_OP_MUL: vec_op_kernel<P, COCG_OP_MUL><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case COCG_OP_ADD: vec_op_kernel<P, COCG_OP_ADD><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case COCG_OP_SUB: vec_op_kernel<P, COCG_OP_SUB><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case COCG_OP_NEG: vec_op_kernel<P, COCG_OP_NEG><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case COCG_OP_TO_MONT: vec_op_kernel<P, COCG_OP_TO_MONT><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    case COCG_OP_FROM_MONT: vec_op_kernel<P, COCG_OP_FROM_MONT><<<grid, 256, 0, ctx->stream>>>(a, b, out, n); break;
    default: return fail(ctx, "cocg_vec_op: unknown op");
  }
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int rep3_mul_local_impl(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                               const void* mask, void* out, size_t n) {
  if (n == 0) return 0;
  int grid = grid_for(n, 256, 8);
  ProfScope prof(ctx, COCG_PROF_VEC);
  if (mask) rep3_mul_local_kernel<P, true><<<grid, 256, 0, ctx->stream>>>(aa, ab, ba, bb, mask, out, n);
  else rep3_mul_local_kernel<P, false><<<grid, 256, 0, ctx->stream>>>(aa, ab, ba, bb, mask, out, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

template <class P>
static int rep3_mul_local_prf_impl(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                                   const void* seed_own, const void* seed_prev, uint32_t ctr, void* out, size_t n) {
  if (n == 0) return 0;
  PrfKey k1, k2;
  memcpy(k1.k, seed_own, 32);
  memcpy(k2.k, seed_prev, 32);
  ProfScope prof(ctx, COCG_PROF_VEC);
  rep3_mul_local_prf_kernel<P><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(aa, ab, ba, bb, k1, k2, ctr, out, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}
template <class P>
static int vec_axpy_impl(cocg_ctx* ctx, const void* a, const void* x, const void* y, void* out, size_t n) {
  if (n == 0) return 0;
  Fp<P> av;
  memcpy(av.l, a, 32);
  ProfScope prof(ctx, COCG_PROF_VEC);
  if (y) vec_axpy_kernel<P, true><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(av, x, y, out, n);
  else vec_axpy_kernel<P, false><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(av, x, y, out, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}
template <class P>
static int prf_fill_impl(cocg_ctx* ctx, const void* seed, uint32_t ctr, void* out, size_t n) {
  if (n == 0) return 0;
  PrfKey k;
  memcpy(k.k, seed, 32);
  prf_fill_kernel<P><<<grid_for(n, 256, 8), 256, 0, ctx->stream>>>(k, ctr, out, n);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace cocg

using namespace cocg;

namespace cocg {
// A pure chain of Montgomery products on every SM: the multiplier ceiling the MSM / NTT kernels are measured against
// (bench.py's roofline.issue.peak), so that the bound is measured in the same run as the kernels instead of being read from a file.
template <class P>
__global__ void __launch_bounds__(256) fp_mul_chain_kernel(uint32_t* out, int iters) {
  Fp<P> a, b;
#pragma unroll
  for (int k = 0; k < P::N; k++) { a.l[k] = threadIdx.x * 2654435761u + k; b.l[k] = blockIdx.x * 40503u + 7 * k + 1; }
  a.l[P::N - 1] &= 0x0fffffffu;
  b.l[P::N - 1] &= 0x0fffffffu;
  for (int it = 0; it < iters; it++) a = fp_mul(a, b);
  if (a.l[0] == 0x12345678u && a.l[1] == 0x9abcdef0u) out[0] = a.l[2];
}
template <class P>
static int fp_mul_ceiling_impl(cocg_ctx* ctx, double* out) {
  void* d;
  COCG_TRY(scratch_get(ctx, 15, 64, &d));
  const int blocks = kNumSMs * 8, iters = 1500;
  cudaEvent_t e0, e1;
  COCG_CUDA(ctx, cudaEventCreate(&e0));
  COCG_CUDA(ctx, cudaEventCreate(&e1));
  fp_mul_chain_kernel<P><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)d, 50);  // warm-up
  COCG_LAUNCH_CHECK(ctx);
  COCG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  fp_mul_chain_kernel<P><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)d, iters);
  COCG_LAUNCH_CHECK(ctx);
  COCG_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  COCG_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0;
  COCG_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *out = (double)blocks * 256 * iters / (ms * 1e-3) / 1e9;
  return 0;
}
}  // namespace cocg

extern "C" int cocg_fp_mul_ceiling(cocg_ctx* ctx, int base_field, double* gmul_per_s) {
  if (!ctx) return 1;
  if (!gmul_per_s) return fail(ctx, "cocg_fp_mul_ceiling: null argument");
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->curve == COCG_BN254) return base_field ? fp_mul_ceiling_impl<Bn254FqP>(ctx, gmul_per_s) : fp_mul_ceiling_impl<Bn254FrP>(ctx, gmul_per_s);
  return base_field ? fp_mul_ceiling_impl<Bls381FqP>(ctx, gmul_per_s) : fp_mul_ceiling_impl<Bls381FrP>(ctx, gmul_per_s);
}

extern "C" int cocg_rep3_mul_local_prf(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                                       const void* seed_own, const void* seed_prev, uint32_t ctr, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!seed_own || !seed_prev || (n && (!aa || !ab || !ba || !bb || !out))) return fail(ctx, "cocg_rep3_mul_local_prf: null operand");
  return COCG_FR_DISPATCH(ctx, rep3_mul_local_prf_impl, ctx, aa, ab, ba, bb, seed_own, seed_prev, ctr, out, n);
}
extern "C" int cocg_vec_axpy(cocg_ctx* ctx, const void* a, const void* x, const void* y, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!a || (n && (!x || !out))) return fail(ctx, "cocg_vec_axpy: null operand");
  return COCG_FR_DISPATCH(ctx, vec_axpy_impl, ctx, a, x, y, out, n);
}
extern "C" int cocg_prf_fill(cocg_ctx* ctx, const void* seed, uint32_t ctr, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!seed || (n && !out)) return fail(ctx, "cocg_prf_fill: null operand");
  return COCG_FR_DISPATCH(ctx, prf_fill_impl, ctx, seed, ctr, out, n);
}
extern "C" int cocg_prf_field_host(int curve, const void* seed, uint32_t ctr, uint64_t idx, void* out) {
  if (!seed || !out) return 1;
  PrfKey k;
  memcpy(k.k, seed, 32);
  if (curve == COCG_BN254) { auto v = prf_field<Bn254FrP>(k, ctr, idx); memcpy(out, v.l, 32); }
  else if (curve == COCG_BLS12_381) { auto v = prf_field<Bls381FrP>(k, ctr, idx); memcpy(out, v.l, 32); }
  else return 1;
  return 0;
}

extern "C" int cocg_vec_op(cocg_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!a || !out || ((op == COCG_OP_MUL || op == COCG_OP_ADD || op == COCG_OP_SUB) && !b))
    return n == 0 ? 0 : fail(ctx, "cocg_vec_op: null operand");
  return COCG_FR_DISPATCH(ctx, vec_op_impl, ctx, op, a, b, out, n);
}

extern "C" int cocg_rep3_mul_local(cocg_ctx* ctx, const void* aa, const void* ab, const void* ba, const void* bb,
                                   const void* mask, void* out, size_t n) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n && (!aa || !ab || !ba || !bb || !out)) return fail(ctx, "cocg_rep3_mul_local: null operand");
  return COCG_FR_DISPATCH(ctx, rep3_mul_local_impl, ctx, aa, ab, ba, bb, mask, out, n);
}

extern "C" int cocg_vec_scale_powers(cocg_ctx* ctx, void* x, size_t n, const void* g, const void* c) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  if (!x || !g || !c) return fail(ctx, "cocg_vec_scale_powers: null operand");
  return COCG_FR_DISPATCH(ctx, scale_powers_impl, ctx, x, n, g, c);
}
