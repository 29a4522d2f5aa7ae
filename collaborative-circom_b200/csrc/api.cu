// Lifecycle, device-memory and O(1) group-operation entry points of the C ABI (include/cocg.h).
#include <string.h>

#include <mutex>

#include "ctx.cuh"

using namespace cocg;

static std::mutex g_err_mu;
static std::string g_create_err;

extern "C" int cocg_version(void) { return 100; }

extern "C" int cocg_create(cocg_ctx** out, int device, int curve) {
  auto seterr = [](const std::string& m) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_create_err = m;
    return 1;
  };
  if (!out) return seterr("cocg_create: null out pointer");
  *out = nullptr;
  if (curve != COCG_BN254 && curve != COCG_BLS12_381) return seterr("cocg_create: unknown curve");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return seterr(std::string("cocg_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
  if (device < 0 || device >= ndev) return seterr("cocg_create: device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return seterr(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return seterr(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10) return seterr("cocg_create: kernels are built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  cocg_ctx* ctx = new cocg_ctx();
  ctx->device = device;
  ctx->curve = curve;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return seterr(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return 0;
}

extern "C" void cocg_destroy(cocg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < cocg_ctx::kScratchSlots; i++)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (auto& pr : ctx->prof_pending) { cudaEventDestroy(pr.a); cudaEventDestroy(pr.b); }
  for (auto& ev : ctx->prof_free) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& kv : ctx->tables) cudaFree(kv.second);
  for (auto& b : ctx->bases)
    if (b.d && b.owned) cudaFree(b.d);
  for (auto& m : ctx->csrs)
    if (m.rowptr && m.owned) { cudaFree(m.rowptr); cudaFree(m.col); cudaFree(m.coeff); }
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char* cocg_last_error(cocg_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

extern "C" int cocg_set_stream(cocg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return 1;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return 0;
}
// Re-creates the context's own stream with the device's highest (high != 0) or default priority.  Thread blocks of a high-priority
// stream are scheduled ahead of the pending blocks of other streams' running kernels: a party whose short witness-map kernels share
// the GPU with another party's MSM launches is no longer queued behind each of them (multi-GPU block mode).
extern "C" int cocg_set_stream_priority(cocg_ctx* ctx, int high) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->stream != ctx->own_stream) return fail(ctx, "cocg_set_stream_priority: the context runs on a caller's stream");
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
  int lo = 0, hi = 0;
  COCG_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));  // numerically lower = higher priority
  cudaStream_t st;
  COCG_CUDA(ctx, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, high ? hi : 0));
  cudaStreamDestroy(ctx->own_stream);
  ctx->own_stream = ctx->stream = st;
  return 0;
}
extern "C" int cocg_sync(cocg_ctx* ctx) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" uint64_t cocg_launch_count(cocg_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cocg_malloc(cocg_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx) return 1;
  if (!dptr) return fail(ctx, "cocg_malloc: null out pointer");
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 16));
  return 0;
}
extern "C" int cocg_free(cocg_ctx* ctx, void* dptr) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  COCG_CUDA(ctx, cudaFree(dptr));
  return 0;
}
extern "C" int cocg_h2d(cocg_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) return 0;
  COCG_CUDA(ctx, cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int cocg_d2h(cocg_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) return 0;
  COCG_CUDA(ctx, cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int cocg_memset0(cocg_ctx* ctx, void* dptr, size_t bytes) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) return 0;
  COCG_CUDA(ctx, cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
  return 0;
}

static int prof_drain(cocg_ctx* ctx) {
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& pr : ctx->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, pr.a, pr.b) == cudaSuccess) { ctx->prof_ms[pr.cls] += ms; ctx->prof_count[pr.cls]++; }
    ctx->prof_free.push_back({pr.a, pr.b});
  }
  ctx->prof_pending.clear();
  return 0;
}
extern "C" int cocg_profile_enable(cocg_ctx* ctx, int on) {
  if (!ctx) return 1;
  ctx->profile = on != 0;
  return 0;
}
extern "C" int cocg_profile_read(cocg_ctx* ctx, int cls, double* total_ms, uint64_t* scopes) {
  if (!ctx) return 1;
  if (cls < 0 || cls >= COCG_PROF_CLASSES) return fail(ctx, "cocg_profile_read: unknown class");
  COCG_TRY(prof_drain(ctx));
  if (total_ms) *total_ms = ctx->prof_ms[cls];
  if (scopes) *scopes = ctx->prof_count[cls];
  return 0;
}
extern "C" int cocg_profile_reset(cocg_ctx* ctx) {
  if (!ctx) return 1;
  COCG_TRY(prof_drain(ctx));
  for (int i = 0; i < COCG_PROF_CLASSES; i++) { ctx->prof_ms[i] = 0; ctx->prof_count[i] = 0; }
  return 0;
}

extern "C" int cocg_d2d(cocg_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) return 0;
  COCG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
extern "C" int cocg_host_alloc(cocg_ctx* ctx, size_t bytes, void** hptr) {
  if (!ctx) return 1;
  if (!hptr) return fail(ctx, "cocg_host_alloc: null out pointer");
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 16));
  return 0;
}
extern "C" int cocg_host_free(cocg_ctx* ctx, void* hptr) {
  if (!ctx) return 1;
  COCG_CUDA(ctx, cudaSetDevice(ctx->device));
  COCG_CUDA(ctx, cudaFreeHost(hptr));
  return 0;
}

// ------------------------------------------------------------------ O(1) group operations on the host
// The handful of single-point operations of proof assembly
// (/root/reference/co-circom/co-groth16/src/groth16.rs:257-312: scalar_mul_public_point, add_assign_points, ...)
// are latency-, not throughput-bound; they run on the calling host thread with the same field code as the
// kernels (host branch of fp.cuh) instead of paying a launch + two copies each.
namespace {
template <class F>
Jacobian<F> host_scalar_mul(const Jacobian<F>& p, const uint8_t* k32) {
  // fixed 4-bit windows: 15-entry table, 256 doublings + 64 additions
  XYZZ<F> tab[16];
  tab[0] = xyzz_inf<F>();
  tab[1] = xyzz_from_jacobian(p);
  for (int i = 2; i < 16; i++) {
    tab[i] = tab[i - 1];
    xyzz_add(tab[i], tab[1]);
  }
  XYZZ<F> acc = xyzz_inf<F>();
  for (int nib = 63; nib >= 0; nib--) {
    for (int q = 0; q < 4; q++) acc = xyzz_dbl(acc);
    int d = (k32[nib >> 1] >> ((nib & 1) * 4)) & 15;
    if (d) xyzz_add(acc, tab[d]);
  }
  return xyzz_to_jacobian(acc);
}

template <class F>
int ec_op_impl(cocg_ctx* ctx, int op, const void* a, const void* b, void* out) {
  Jacobian<F> A, R;
  Affine<F> Af;
  switch (op) {
    case 0: {
      Jacobian<F> B;
      memcpy(&A, a, sizeof(A));
      memcpy(&B, b, sizeof(B));
      XYZZ<F> acc = xyzz_from_jacobian(A);
      xyzz_add(acc, xyzz_from_jacobian(B));
      R = xyzz_to_jacobian(acc);
      break;
    }
    case 1:
      memcpy(&A, a, sizeof(A));
      R = host_scalar_mul(A, (const uint8_t*)b);
      break;
    case 2: {
      memcpy(&A, a, sizeof(A));
      if (A.is_inf()) {
        memset(out, 0, sizeof(Af));
        return 0;
      }
      // z^-1 over Fq or Fq2
      F zi = f_inv(A.z);
      F zi2 = f_sqr(zi);
      Af.x = f_mul(A.x, zi2);
      Af.y = f_mul(A.y, f_mul(zi2, zi));
      memcpy(out, &Af, sizeof(Af));
      return 0;
    }
    case 3:
      memcpy(&Af, a, sizeof(Af));
      R = Af.is_inf() ? jac_inf<F>() : Jacobian<F>{Af.x, Af.y, F::one()};
      break;
    case 4:
      memcpy(&A, a, sizeof(A));
      R = Jacobian<F>{A.x, f_neg(A.y), A.z};
      break;
    case 5: {
      memcpy(&A, a, sizeof(A));
      R = xyzz_to_jacobian(xyzz_dbl(xyzz_from_jacobian(A)));
      break;
    }
    default:
      return fail(ctx, "cocg_ec_op: unknown op");
  }
  memcpy(out, &R, sizeof(R));
  return 0;
}
}  // namespace

// op 6: the group generator as a Jacobian point (x, y, 1)
static int ec_generator(cocg_ctx* ctx, int group, void* out) {
  static const uint32_t bn_g1[2][8] = BN254_G1_GEN;
  static const uint32_t bn_g2[4][8] = BN254_G2_GEN;
  static const uint32_t bls_g1[2][12] = BLS381_G1_GEN;
  static const uint32_t bls_g2[4][12] = BLS381_G2_GEN;
  char* o = (char*)out;
  if (ctx->curve == COCG_BN254) {
    Bn254Fq one = Bn254Fq::one();
    size_t nb = group == COCG_G1 ? sizeof(bn_g1) : sizeof(bn_g2);
    memcpy(o, group == COCG_G1 ? (const void*)bn_g1 : (const void*)bn_g2, nb);
    memset(o + nb, 0, nb / 2);
    memcpy(o + nb, one.l, 32);
  } else {
    Bls381Fq one = Bls381Fq::one();
    size_t nb = group == COCG_G1 ? sizeof(bls_g1) : sizeof(bls_g2);
    memcpy(o, group == COCG_G1 ? (const void*)bls_g1 : (const void*)bls_g2, nb);
    memset(o + nb, 0, nb / 2);
    memcpy(o + nb, one.l, 48);
  }
  return 0;
}

extern "C" int cocg_ec_op(cocg_ctx* ctx, int group, int op, const void* a, const void* b, void* out) {
  if (!ctx) return 1;
  if (group != COCG_G1 && group != COCG_G2) return fail(ctx, "cocg_ec_op: group must be 1 or 2");
  if (op == 6) return out ? ec_generator(ctx, group, out) : fail(ctx, "cocg_ec_op: null argument");
  if (!a || !out || ((op == 0 || op == 1) && !b)) return fail(ctx, "cocg_ec_op: null argument");
  if (group != COCG_G1 && group != COCG_G2) return fail(ctx, "cocg_ec_op: group must be 1 or 2");
  if (ctx->curve == COCG_BN254) return group == COCG_G1 ? ec_op_impl<Bn254Fq>(ctx, op, a, b, out) : ec_op_impl<Bn254Fq2>(ctx, op, a, b, out);
  return group == COCG_G1 ? ec_op_impl<Bls381Fq>(ctx, op, a, b, out) : ec_op_impl<Bls381Fq2>(ctx, op, a, b, out);
}
