// mix.cu — does a non-FP64 instruction issue in the shadow of a DFMA?  Per thread: 4 independent DFMA chains and K independent
// 32-bit integer (or FP64-select) instructions per DFMA; reports SM cycles per DFMA warp-instruction for several warp counts.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int KIND>
__global__ void k_mix(double* out, long long* cyc, int iters, double a, double b, int ia) {
  double x[4]; int y[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) x[q] = threadIdx.x * 1e-3 + q;
#pragma unroll
  for (int q = 0; q < 8; ++q) y[q] = threadIdx.x + q;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        x[q] = fma(x[q], a, b);
#pragma unroll
        for (int s = 0; s < K; ++s) {
          const int j = (q * K + s) & 7;
          if (KIND == 0) y[j] = y[j] * ia + 12345;                      // IMAD
          else if (KIND == 1) y[j] = (y[j] ^ ia) + (y[j] >> 3);          // LOP3 / shift / IADD
          else y[j] = y[j] > ia ? y[j] - 7 : y[j] + 5;
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0; int si = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += x[q];
#pragma unroll
  for (int q = 0; q < 8; ++q) si += y[q];
  if (s == 1.2345e-300 || si == 123456789) *out = s + si;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K, int KIND> void run(int warps, double* out, long long* cyc) {
  const int iters = 1000;
  k_mix<K, KIND><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9, 3);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double ndfma = (double)warps * iters * 32;
  printf("kind %d K=%d warps/SM %2d: SM cycles per DFMA %.3f (DFMA/clk/SM %.2f; all instr/clk/SM %.2f)\n", KIND, K, warps,
         h / ndfma, ndfma / h, ndfma * (1 + K) / h);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 12, 16}) { run<0, 0>(w, out, cyc); run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<3, 0>(w, out, cyc); }
  for (int w : {8, 12}) { run<1, 1>(w, out, cyc); run<2, 1>(w, out, cyc); }
  return 0;
}
