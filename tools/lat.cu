// lat.cu — FP64 dependent-issue latency and per-SMSP issue interval on this GPU
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x * 1e-3 + q;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  if (s == 1.2345e-300) *out = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm_q, double* out, long long* cyc) {
  const int iters = 2000;
  k_chain<ILP><<<1, 32 * warps_per_sm_q>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("ILP %d, warps/SM %2d: %.2f cycles per DFMA per warp; SM rate %.2f DFMA-warp-instr/cycle\n", ILP, warps_per_sm_q,
         (double)h / (iters * 8.0 * ILP), (double)warps_per_sm_q * iters * 8 * ILP / h);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 12, 16, 32}) run<1>(w, out, cyc);
  for (int w : {1, 4, 8, 12, 16}) run<2>(w, out, cyc);
  for (int w : {4, 8, 12, 16}) run<4>(w, out, cyc);
  return 0;
}
