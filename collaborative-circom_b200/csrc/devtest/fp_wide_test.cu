// Standalone device check of fp_wide.cuh against the interleaved fp_mul of fp.cuh (itself checked bit-exactly against the oracle by
// tests/test_gpu_parity.py): random and edge operands, all four fields.  nvcc -arch ... -o fp_wide_test fp_wide_test.cu && ./fp_wide_test
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fp_wide.cuh"

using namespace cocg;

template <class P>
__global__ void check_kernel(const uint32_t* in, size_t n, unsigned* bad, double* cycles) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<P> a, b;
  for (int k = 0; k < P::N; k++) { a.l[k] = in[(2 * i) * P::N + k]; b.l[k] = in[(2 * i + 1) * P::N + k]; }
  Fp<P> want = fp_mul(a, b), got = fp_mul_wide(a, b);
  Fp<P> wants = fp_mul(a, a), gots = fp_sqr_wide(a);
  if (want != got) atomicAdd(&bad[0], 1u);
  if (wants != gots) atomicAdd(&bad[1], 1u);
}

template <class P, int MODE>
__global__ void speed_kernel(uint32_t* out, int iters) {
  Fp<P> a, b;
  for (int k = 0; k < P::N; k++) { a.l[k] = threadIdx.x * 2654435761u + k; b.l[k] = blockIdx.x * 40503u + 7 * k + 1; }
  a.l[P::N - 1] &= 0x0fffffffu; b.l[P::N - 1] &= 0x0fffffffu;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) a = fp_mul(a, b);
    else if (MODE == 1) a = fp_mul_wide(a, b);
    else if (MODE == 2) a = fp_mul(a, a);
    else a = fp_sqr_wide(a);
  }
  if (a.l[0] == 0x12345678u) out[0] = a.l[1];
}

template <class P>
int run(const char* name) {
  const size_t n = 1 << 18;
  std::vector<uint32_t> h(2 * n * P::N);
  srand(12345);
  for (auto& v : h) v = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
  // reduce operands below the modulus: clear top bits, then force a few edge cases
  const uint32_t top = (P::BITS % 32) ? ((1u << (P::BITS % 32 - 1)) - 1u) : 0x7fffffffu;
  for (size_t i = 0; i < 2 * n; i++) h[i * P::N + P::N - 1] &= top;
  for (int k = 0; k < P::N; k++) { h[k] = 0; h[P::N + k] = P::mod(k); }            // 0 * (p) -- p itself is not canonical but must not crash
  for (int k = 0; k < P::N; k++) { h[2 * P::N + k] = P::mod(k) - (k == 0); h[3 * P::N + k] = P::mod(k) - (k == 0); }  // (p-1)^2
  for (int k = 0; k < P::N; k++) { h[4 * P::N + k] = 0xffffffffu & (k == P::N - 1 ? top : 0xffffffffu); h[5 * P::N + k] = h[4 * P::N + k]; }
  uint32_t* d; unsigned* bad; double* cyc;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&bad, 8); cudaMalloc(&cyc, 8);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(bad, 0, 8);
  check_kernel<P><<<(unsigned)((n + 127) / 128), 128>>>(d + 2 * P::N, n - 1, bad, cyc);  // skip the non-canonical pair 0
  unsigned hb[2];
  cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  float ms[4];
  for (int mode = 0; mode < 4; mode++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    if (mode == 0) speed_kernel<P, 0><<<148 * 8, 256>>>(d, 2000);
    if (mode == 1) speed_kernel<P, 1><<<148 * 8, 256>>>(d, 2000);
    if (mode == 2) speed_kernel<P, 2><<<148 * 8, 256>>>(d, 2000);
    if (mode == 3) speed_kernel<P, 3><<<148 * 8, 256>>>(d, 2000);
    cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms[mode], a, b);
  }
  const double ops = 148.0 * 8 * 256 * 2000;
  printf("%-10s mul mismatches %u, sqr mismatches %u of %zu (%s) | G op/s: fp_mul %.1f  mul_wide %.1f  fp_mul(a,a) %.1f  sqr_wide %.1f\n", name, hb[0], hb[1], n - 1,
         cudaGetErrorString(e), ops / ms[0] / 1e6, ops / ms[1] / 1e6, ops / ms[2] / 1e6, ops / ms[3] / 1e6);
  return hb[0] + hb[1];
}

int main() {
  int bad = 0;
  bad += run<Bn254FqP>("bn254 Fq");
  bad += run<Bn254FrP>("bn254 Fr");
  bad += run<Bls381FqP>("bls381 Fq");
  bad += run<Bls381FrP>("bls381 Fr");
  printf(bad ? "FAILED\n" : "OK\n");
  return bad != 0;
}
