// Context object behind the C ABI (include/cocg.h): device, stream, grow-only scratch arenas, cached NTT
// tables, resident bases / CSR matrices.  One context per MPC driver (the reference's drivers are `&mut self`,
// /root/reference/mpc-core/src/traits.rs:43), so nothing in here is shared between threads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <array>
#include <deque>
#include <map>
#include <string>
#include <vector>

#include "../../include/cocg.h"
#include "ec.cuh"

namespace cocg {
constexpr int kNumSMs = 148;  // B200; grids are sized in multiples of this

struct BasesEntry {
  void* d = nullptr;  // table T[j][i] = 2^(c*j) P_i of packed Montgomery affine points, j < nwin (T[0] = the bases)
  size_t n = 0;
  int c = 0, nwin = 1;  // window bits the table was built for (msm_impl.cuh)
  int group = 0;
  size_t point_bytes = 0;
  bool owned = true;  // false: alias of another context's allocation (cocg_bases_share)
};
struct CsrEntry {
  uint32_t* rowptr = nullptr;
  uint32_t* col = nullptr;
  void* coeff = nullptr;
  size_t rows = 0, nnz = 0;
  bool owned = true;
};
}  // namespace cocg

struct cocg_ctx {
  int device = 0;
  int curve = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // scratch arenas (grow-only, reused across calls)
  enum { kScratchSlots = 16 };
  void* scratch[kScratchSlots] = {};
  size_t scratch_bytes[kScratchSlots] = {};
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // cached power tables: key = (kind, n, base limbs, first limbs)
  std::map<std::array<uint32_t, 18>, void*> tables;
  std::vector<cocg::BasesEntry> bases;
  std::vector<cocg::CsrEntry> csrs;
  // per-kernel-class device timing (cocg_profile_*): event pairs recorded on the launching stream
  bool profile = false;
  struct ProfPair { cudaEvent_t a, b; int cls; };
  std::deque<ProfPair> prof_pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_free;
  double prof_ms[COCG_PROF_CLASSES] = {};
  uint64_t prof_count[COCG_PROF_CLASSES] = {};
};

namespace cocg {

inline int fail(cocg_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return 1;
}

#define COCG_CUDA(ctx, call)                                                                          \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return ::cocg::fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")");                                   \
  } while (0)

#define COCG_LAUNCH_CHECK(ctx)                 \
  do {                                         \
    (ctx)->launches++;                         \
    COCG_CUDA(ctx, cudaGetLastError());        \
  } while (0)

// Times everything enqueued on ctx->stream during the scope (only when cocg_profile_enable is on).
struct ProfScope {
  cocg_ctx* c;
  cudaEvent_t stop = nullptr;
  ProfScope(cocg_ctx* ctx, int cls) : c(ctx) {
    if (!c->profile) return;
    // recycle finished pairs instead of creating events on the hot path (cudaEventCreate costs ~10 us)
    while (c->prof_free.empty() && !c->prof_pending.empty()) {
      if (cudaEventQuery(c->prof_pending.front().b) != cudaSuccess) {
        (void)cudaGetLastError();  // cudaErrorNotReady is not an error; do not leave it for the next launch check
        break;
      }
      auto pr = c->prof_pending.front();
      float ms = 0;
      if (cudaEventElapsedTime(&ms, pr.a, pr.b) == cudaSuccess) { c->prof_ms[pr.cls] += ms; c->prof_count[pr.cls]++; }
      c->prof_free.push_back({pr.a, pr.b});
      c->prof_pending.pop_front();
    }
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!c->prof_free.empty()) { ev = c->prof_free.back(); c->prof_free.pop_back(); }
    else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) return;
    cudaEventRecord(ev.first, c->stream);
    c->prof_pending.push_back({ev.first, ev.second, cls});
    stop = ev.second;
  }
  ~ProfScope() { if (stop) cudaEventRecord(stop, c->stream); }
};

#define COCG_TRY(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__) return rc__;    \
  } while (0)

// grow-only scratch; contents are not preserved across growth
inline int scratch_get(cocg_ctx* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) {
      COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      COCG_CUDA(ctx, cudaFree(ctx->scratch[slot]));
      ctx->scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
    }
    size_t cap = bytes + bytes / 8 + 256;
    COCG_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], cap));
    ctx->scratch_bytes[slot] = cap;
  }
  *out = ctx->scratch[slot];
  return 0;
}
inline int pinned_get(cocg_ctx* ctx, size_t bytes, void** out) {
  if (ctx->pinned_bytes < bytes) {
    if (ctx->pinned) {
      COCG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      COCG_CUDA(ctx, cudaFreeHost(ctx->pinned));
      ctx->pinned = nullptr;
      ctx->pinned_bytes = 0;
    }
    COCG_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
  }
  *out = ctx->pinned;
  return 0;
}

// Window width c of the MSM table built for a query of n points (msm_impl.cuh): about log2(n), capped at 20 bits
// (13 windows for 254/255-bit scalars, 2^19 buckets of ~26 points at n = 2^20).  Measured alternatives at n = 2^20 (G1, ms):
// c = 17 with 4 lanes per bucket 3.70, c = 19 3.92, c = 20 3.36.  If one bit less needs no extra window it is preferred: half the
// buckets to reduce for the same number of additions, and a top window that is not degenerate (254 bits: c = 18 -> 15 windows with a
// 2-bit top window whose 3 buckets collect n / 4 entries each; c = 17 -> 15 windows with a 16-bit top window).
inline int msm_plan_window_bits(size_t n, int scalar_bits) {
  int lg = 0;
  size_t v = n + n / 2;  // round to the nearest power of two
  while (((size_t)1 << (lg + 1)) <= v) lg++;
  int c = lg;
  if (c < 4) c = 4;
  int cap = 20;
  if (const char* e = getenv("COCG_MSM_MAX_WINDOW")) {  // experiment knob: cap of the window width, 8..22
    int v2 = atoi(e);
    if (v2 >= 8 && v2 <= 22) cap = v2;
  }
  if (c > cap) c = cap;
  while (c > 4 && (scalar_bits + c - 1) / (c - 1) == (scalar_bits + c) / c) c--;
  return c;
}

inline int grid_for(size_t work_items, int threads, int max_waves = 16) {
  size_t b = (work_items + threads - 1) / threads;
  size_t cap = (size_t)kNumSMs * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- 16-byte vector access to 32-byte Fr elements ----
template <class P>
__device__ __forceinline__ Fp<P> load_fp(const void* base, size_t idx) {
  static_assert(P::N % 4 == 0, "limbs must pack into uint4");
  Fp<P> r;
  const uint4* p = reinterpret_cast<const uint4*>(base) + idx * (P::N / 4);
#pragma unroll
  for (int k = 0; k < P::N / 4; k++) {
    uint4 v = p[k];
    r.l[4 * k] = v.x; r.l[4 * k + 1] = v.y; r.l[4 * k + 2] = v.z; r.l[4 * k + 3] = v.w;
  }
  return r;
}
template <class P>
__device__ __forceinline__ Fp<P> load_fp_ro(const void* base, size_t idx) {  // read-only path
  Fp<P> r;
  const uint4* p = reinterpret_cast<const uint4*>(base) + idx * (P::N / 4);
#pragma unroll
  for (int k = 0; k < P::N / 4; k++) {
    uint4 v = __ldg(p + k);
    r.l[4 * k] = v.x; r.l[4 * k + 1] = v.y; r.l[4 * k + 2] = v.z; r.l[4 * k + 3] = v.w;
  }
  return r;
}
template <class P>
__device__ __forceinline__ void store_fp(void* base, size_t idx, const Fp<P>& v) {
  uint4* p = reinterpret_cast<uint4*>(base) + idx * (P::N / 4);
#pragma unroll
  for (int k = 0; k < P::N / 4; k++) p[k] = make_uint4(v.l[4 * k], v.l[4 * k + 1], v.l[4 * k + 2], v.l[4 * k + 3]);
}

// Power tables out[i] = first * base^i (device, cached per context).  kind separates users of the same base.
int powers_table(cocg_ctx* ctx, uint32_t kind, size_t n, const uint32_t* base, const uint32_t* first, void** out);

// per-curve dispatch helper
#define COCG_FR_DISPATCH(ctx, FN, ...)                                              \
  ((ctx)->curve == COCG_BN254 ? FN<::cocg::Bn254FrP>(__VA_ARGS__) : FN<::cocg::Bls381FrP>(__VA_ARGS__))

}  // namespace cocg
