// evaluate_constraint as a CSR sparse matrix-vector product.
//
// Replaces the double loop /root/reference/co-circom/co-groth16/src/groth16.rs:159-166 over
// `evaluate_constraint` (/root/reference/mpc-core/src/protocols/rep3.rs:690-708, plain.rs:243-251): for each
// constraint row, sum coeff * z[index] with z = (public inputs | witness share).  In REP3 the public part only
// enters party 0's `a` and party 1's `b` component (rep3.rs:600-608): the caller passes z_pub = NULL for the
// components that must not see it.  One thread per row (rows have 2-4 non-zeros); coefficients stream, the
// z operands are 32-byte gathers.
#include "ctx.cuh"

namespace cocg {

template <class P>
__global__ void __launch_bounds__(256) spmv_kernel(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col,
                                                    const void* __restrict__ coeff, const void* __restrict__ z_pub, uint32_t npub,
                                                    const void* __restrict__ z_wit, void* __restrict__ out, size_t rows) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  Fp<P> acc = Fp<P>::zero();
  uint32_t end = rowptr[r + 1];
  for (uint32_t k = rowptr[r]; k < end; k++) {
    uint32_t c = __ldg(col + k);
    if (c < npub) {
      if (z_pub) acc = fp_add(acc, fp_mul(load_fp_ro<P>(coeff, k), load_fp_ro<P>(z_pub, c)));
    } else {
      acc = fp_add(acc, fp_mul(load_fp_ro<P>(coeff, k), load_fp_ro<P>(z_wit, c - npub)));
    }
  }
  store_fp<P>(out, r, acc);
}

template <class P>
static int spmv_impl(cocg_ctx* ctx, const CsrEntry& m, const void* z_pub, size_t npub, const void* z_wit, void* out) {
  if (m.rows == 0) return 0;
  ProfScope prof(ctx, COCG_PROF_SPMV);
  spmv_kernel<P><<<(unsigned)((m.rows + 255) / 256), 256, 0, ctx->stream>>>(m.rowptr, m.col, m.coeff, z_pub, (uint32_t)npub, z_wit, out, m.rows);
  COCG_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace cocg

using namespace cocg;

extern "C" int cocg_csr_upload(cocg_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, const void* coeff, size_t rows, size_t nnz,
                               uint64_t* handle) {
  return cocg_csr_upload_form(ctx, rowptr, col, coeff, rows, nnz, COCG_FORM_MONT, handle);
}

extern "C" int cocg_csr_upload_form(cocg_ctx* ctx, const uint32_t* rowptr, const uint32_t* col, const void* coeff, size_t rows, size_t nnz,
                                    int coeff_form, uint64_t* handle) {
  if (!ctx) return 1;
  if (coeff_form < COCG_FORM_MONT || coeff_form > COCG_FORM_CANONICAL) return fail(ctx, "cocg_csr_upload: unknown coefficient form");
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!handle || !rowptr || (nnz && (!col || !coeff))) return fail(ctx, "cocg_csr_upload: null argument");
  if (rowptr[rows] != nnz) return fail(ctx, "cocg_csr_upload: rowptr[rows] != nnz");
  for (size_t r = 0; r < rows; r++)  // a malformed row pointer would become an out-of-bounds read in spmv_kernel
    if (rowptr[r] > rowptr[r + 1]) return fail(ctx, "cocg_csr_upload: rowptr is not non-decreasing");
  CsrEntry m;
  m.rows = rows; m.nnz = nnz;
  COCG_CUDA(ctx, cudaMalloc(&m.rowptr, (rows + 1) * 4));
  COCG_CUDA(ctx, cudaMalloc(&m.col, nnz ? nnz * 4 : 16));
  COCG_CUDA(ctx, cudaMalloc(&m.coeff, nnz ? nnz * 32 : 32));
  COCG_CUDA(ctx, cudaMemcpyAsync(m.rowptr, rowptr, (rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    COCG_CUDA(ctx, cudaMemcpyAsync(m.col, col, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    COCG_CUDA(ctx, cudaMemcpyAsync(m.coeff, coeff, nnz * 32, cudaMemcpyHostToDevice, ctx->stream));
    // zkey section 4 stores value * R^2: one Montgomery reduction gives the Montgomery form value * R (circom-types traits.rs:65-67)
    if (coeff_form == COCG_FORM_R2) COCG_TRY(cocg_vec_op(ctx, COCG_OP_FROM_MONT, m.coeff, nullptr, m.coeff, nnz));
    if (coeff_form == COCG_FORM_CANONICAL) COCG_TRY(cocg_vec_op(ctx, COCG_OP_TO_MONT, m.coeff, nullptr, m.coeff, nnz));
  }
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < ctx->csrs.size(); i++)
    if (!ctx->csrs[i].rowptr) { ctx->csrs[i] = m; *handle = i + 1; return 0; }
  ctx->csrs.push_back(m);
  *handle = ctx->csrs.size();
  return 0;
}

extern "C" int cocg_csr_free(cocg_ctx* ctx, uint64_t handle) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (handle == 0 || handle > ctx->csrs.size() || !ctx->csrs[handle - 1].rowptr) return fail(ctx, "cocg_csr_free: bad handle");
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CsrEntry& m = ctx->csrs[handle - 1];
  if (m.owned) { cudaFree(m.rowptr); cudaFree(m.col); cudaFree(m.coeff); }
  m = CsrEntry();
  return 0;
}

extern "C" int cocg_csr_download(cocg_ctx* ctx, uint64_t handle, uint32_t* rowptr, uint32_t* col, void* coeff, size_t* nnz) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (handle == 0 || handle > ctx->csrs.size() || !ctx->csrs[handle - 1].rowptr) return fail(ctx, "cocg_csr_download: bad handle");
  const CsrEntry& m = ctx->csrs[handle - 1];
  if (nnz) *nnz = m.nnz;
  if (rowptr) COCG_CUDA(ctx, cudaMemcpyAsync(rowptr, m.rowptr, (m.rows + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (col && m.nnz) COCG_CUDA(ctx, cudaMemcpyAsync(col, m.col, m.nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (coeff && m.nnz) COCG_CUDA(ctx, cudaMemcpyAsync(coeff, m.coeff, m.nnz * 32, cudaMemcpyDeviceToHost, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int cocg_csr_share(cocg_ctx* ctx, cocg_ctx* owner, uint64_t owner_handle, uint64_t* handle) {
  if (!ctx) return 1;
  if (!owner || !handle) return fail(ctx, "cocg_csr_share: null argument");
  if (owner->device != ctx->device || owner->curve != ctx->curve) return fail(ctx, "cocg_csr_share: contexts differ in device or curve");
  if (owner_handle == 0 || owner_handle > owner->csrs.size() || !owner->csrs[owner_handle - 1].rowptr) return fail(ctx, "cocg_csr_share: bad handle");
  CsrEntry m = owner->csrs[owner_handle - 1];
  m.owned = false;
  for (size_t i = 0; i < ctx->csrs.size(); i++)
    if (!ctx->csrs[i].rowptr) { ctx->csrs[i] = m; *handle = i + 1; return 0; }
  ctx->csrs.push_back(m);
  *handle = ctx->csrs.size();
  return 0;
}

extern "C" int cocg_spmv(cocg_ctx* ctx, uint64_t csr, const void* z_pub, size_t npub, const void* z_wit, void* out) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (csr == 0 || csr > ctx->csrs.size() || !ctx->csrs[csr - 1].rowptr) return fail(ctx, "cocg_spmv: bad handle");
  const CsrEntry& m = ctx->csrs[csr - 1];
  if (m.rows && !out) return fail(ctx, "cocg_spmv: null output");
  return COCG_FR_DISPATCH(ctx, spmv_impl, ctx, m, z_pub, npub, z_wit, out);
}
