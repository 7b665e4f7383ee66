// peaks.cu — measures the two ceilings the step kernels are judged against on this GPU:
// HBM copy bandwidth (STREAM-style copy, read+write bytes) and FP64 DFMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
__global__ void k_read(const double4* __restrict__ a, double* out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += st) { double4 v = a[i]; s += v.x + v.y + v.z + v.w; }
  if (s == 1.2345e-300) *out = s;
}
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x * 1e-3 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  if (s == 1.2345e-300) *out = s;
}
__global__ void k_mufu(double* out, int iters, double a) {
  double x[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = threadIdx.x * 1e-3 + q + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[q])); x[q] = r + a; }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += x[q];
  if (s == 1.2345e-300) *out = s;
}
__global__ void k_ddiv(double* out, int iters, double a) {
  double x[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) x[q] = threadIdx.x * 1e-3 + q + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = a / x[q] + 1.0;
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += x[q];
  if (s == 1.2345e-300) *out = s;
}
__global__ void k_dsqrt(double* out, int iters, double a) {
  double x[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) x[q] = threadIdx.x * 1e-3 + q + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = sqrt(x[q]) + a;
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += x[q];
  if (s == 1.2345e-300) *out = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const size_t bytes = 4ull << 30, n = bytes / sizeof(double4);
  double4 *a, *b; double* out;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 8);
  cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
  double best_copy = 0, best_read = 0;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0); k_copy<<<sms * 16, 512>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); double g = 2.0 * bytes / (ms * 1e-3) / 1e9; if (rep > 1 && g > best_copy) best_copy = g;
    cudaEventRecord(e0); k_read<<<sms * 16, 512>>>(a, out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); g = 1.0 * bytes / (ms * 1e-3) / 1e9; if (rep > 1 && g > best_read) best_read = g;
  }
  double dfma = 0, mufu = 0, ddiv = 0, dsq = 0;
  const int iters = 20000;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); k_dfma<8><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); double g = (double)sms * 8 * 256 * iters * 8 / (ms * 1e-3); if (g > dfma) dfma = g;
    cudaEventRecord(e0); k_mufu<<<sms * 8, 256>>>(out, iters / 4, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); g = (double)sms * 8 * 256 * (iters / 4) * 8 / (ms * 1e-3); if (g > mufu) mufu = g;
    cudaEventRecord(e0); k_ddiv<<<sms * 8, 256>>>(out, iters / 8, 3.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); g = (double)sms * 8 * 256 * (iters / 8) * 4 / (ms * 1e-3); if (g > ddiv) ddiv = g;
    cudaEventRecord(e0); k_dsqrt<<<sms * 8, 256>>>(out, iters / 8, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); g = (double)sms * 8 * 256 * (iters / 8) * 4 / (ms * 1e-3); if (g > dsq) dsq = g;
  }
  printf("{\"device\": \"%s\", \"sms\": %d, \"hbm_copy_gbs\": %.1f, \"hbm_read_gbs\": %.1f, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.2f, "
         "\"rcp_add_pairs_per_s\": %.4e, \"ieee_div_add_per_s\": %.4e, \"ieee_sqrt_add_per_s\": %.4e}\n",
         p.name, sms, best_copy, best_read, dfma, 2 * dfma / 1e12, mufu, ddiv, dsq);
  return 0;
}
