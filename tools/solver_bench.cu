// solver_bench.cu — the production HLLD solve alone, in registers (no memory, no barriers): cycles per solve per SM sub-partition
// for NJ interfaces solved together per thread (the compiler may interleave them) at several warp counts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGX_FLAVOUR_FAST -fmad=true -o tools/solver_bench tools/solver_bench.cu
#define GX_SOLVE_MASK 0xffffffffu
#include <cstdio>
#include "../guacho_b200/csrc/gx_physics.cuh"

template <int NJ, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_solve(gxp::Phys P, double* out, long long* cyc, int iters) {
  double wl[NJ][8], wr[NJ][8];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const double t = 1e-3 * (threadIdx.x + 37 * j);
    wl[j][0] = 1.0 + t; wl[j][1] = 0.3 - t; wl[j][2] = 0.1 + t; wl[j][3] = -0.2 + t; wl[j][4] = 0.6 + t; wl[j][5] = 0.4; wl[j][6] = 0.3 + t; wl[j][7] = -0.1 - t;
    wr[j][0] = 0.9 - t; wr[j][1] = 0.2 + t; wr[j][2] = -0.1 + t; wr[j][3] = 0.1 - t; wr[j][4] = 0.5 + t; wr[j][5] = 0.45; wr[j][6] = 0.2 - t; wr[j][7] = 0.15 + t;
  }
  int err = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    double ff[NJ][8];
    gxp::PasInfo I;
#pragma unroll
    for (int j = 0; j < NJ; ++j) err |= gxp::riemann<GX_SOLVER_HLLD>(P, wl[j], wr[j], ff[j], I);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {           // feed the flux back so iterations depend on each other (like a new interface)
        if (q == 0 || q == 4) { wl[j][q] = fma(1e-9, fabs(ff[j][q]), wl[j][q]); wr[j][q] = fma(1e-9, fabs(ff[j][q]), wr[j][q]); }
        else { wl[j][q] = fma(1e-9, ff[j][q], wl[j][q]); wr[j][q] = fma(-1e-9, ff[j][q], wr[j][q]); }
      }
    }
  }
  const long long t1 = clock64();
  double s = err;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) s += wl[j][q] + wr[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NJ, int MAXT> void run(int warps, const gxp::Phys& P, double* out, long long* cyc) {
  const int iters = 2000;
  k_solve<NJ, MAXT><<<1, 32 * warps>>>(P, out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double v; cudaMemcpy(&v, out, 8, cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_solve<NJ, MAXT>);
  const double solves_per_smsp = (double)iters * NJ * warps / 4.0;
  printf("NJ=%d maxthreads=%3d regs=%3d spill=%zuB warps/SM=%2d: %.0f cycles per solve per SMSP (%.0f cycles per warp-iteration) %s out=%g\n", NJ, MAXT,
         fa.numRegs, (size_t)fa.localSizeBytes, warps, h / solves_per_smsp, (double)h / iters, e == cudaSuccess ? "" : cudaGetErrorString(e), v);
}
int main() {
  gxp::Phys P; P.cv = 1.5; P.gamma = 5.0 / 3.0; P.Tempsc = 1.0; P.inv_cv = 1.0 / 1.5; P.m4gamma = -4.0 * P.gamma; P.eos = 1; P.neqdyn = 8; P.npas = 0;
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  for (int w : {4, 8}) run<1, 256>(w, P, out, cyc);
  for (int w : {4, 8, 12}) run<1, 384>(w, P, out, cyc);
  for (int w : {4, 8, 12, 16}) run<1, 512>(w, P, out, cyc);
  for (int w : {4, 8}) run<2, 256>(w, P, out, cyc);
  for (int w : {4, 8, 12}) run<2, 384>(w, P, out, cyc);
  for (int w : {4, 8}) run<3, 256>(w, P, out, cyc);
  return 0;
}
