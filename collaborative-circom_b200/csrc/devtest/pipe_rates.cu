// Measures the per-SM issue rates that decide how a 256/384-bit Montgomery product should be built on B200 (sm_100a):
// IMAD.WIDE.U32 (the pipe fp.cuh uses today), DFMA (FP64 pipe, 52-bit-limb products), IADD3 (ALU pipe), and mixes of them.
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChains = 8;  // independent dependency chains per thread (hides the pipe latency)

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(uint64_t* out, int iters, double seed) {
  uint64_t acc[kChains];
  double d[kChains];
  uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
  uint32_t s[kChains];
  for (int k = 0; k < kChains; k++) { acc[k] = a + k; d[k] = seed + k + threadIdx.x; s[k] = a ^ k; }
  const double m = seed * 0.999, c = seed * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < kChains; k++) {
      if (MODE == 0 || MODE == 3) {  // IMAD.WIDE.U32
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[k]), "r"(b));
      }
      if (MODE == 1 || MODE == 3 || MODE == 4) {  // DFMA
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(m), "d"(c));
      }
      if (MODE == 2 || MODE == 4) {  // IADD3 (two per DFMA in mode 4: a 64-bit add)
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(s[k]) : "r"(a));
        asm volatile("addc.u32 %0, %0, %1;" : "+r"(s[(k + 1) % kChains]) : "r"(b));
      }
      if (MODE == 5) {  // IMAD lo + IMAD.HI pair (instead of one IMAD.WIDE)
        uint32_t lo, hi;
        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(s[k]), "r"(b), "r"(a));
        asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(s[k]), "r"(b), "r"(lo));
        s[k] = hi;
      }
    }
  }
  uint64_t r = 0;
  for (int k = 0; k < kChains; k++) r += acc[k] + (uint64_t)__double_as_longlong(d[k]) + s[k];
  if (r == 0x1234567812345678ull) out[0] = r;
}

template <int MODE>
static void run(const char* name, double ops_per_iter_per_thread, int sm_count, int clock_khz) {
  uint64_t* d;
  cudaMalloc(&d, 8);
  const int blocks = sm_count * 8, threads = 256, iters = 4000;
  rate_kernel<MODE><<<blocks, threads>>>(d, 10, 1.5);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  rate_kernel<MODE><<<blocks, threads>>>(d, iters, 1.5);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * iters * kChains * ops_per_iter_per_thread;
  const double per_clk_sm = ops / (ms * 1e-3) / ((double)clock_khz * 1e3) / sm_count;
  printf("%-34s %8.3f ms  %7.1f G op/s  %6.1f op/clk/SM (at %d MHz)\n", name, ms, ops / ms / 1e6, per_clk_sm, clock_khz / 1000);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, khz / 1000);
  run<0>("IMAD.WIDE.U32", 1, p.multiProcessorCount, khz);
  run<1>("DFMA", 1, p.multiProcessorCount, khz);
  run<2>("IADD3 pair (add.cc + addc)", 2, p.multiProcessorCount, khz);
  run<3>("IMAD.WIDE + DFMA (counted: both)", 2, p.multiProcessorCount, khz);
  run<4>("DFMA + 2 IADD3 (counted: DFMA)", 1, p.multiProcessorCount, khz);
  run<5>("IMAD.lo + IMAD.HI pair (counted: 1)", 1, p.multiProcessorCount, khz);
  return 0;
}
