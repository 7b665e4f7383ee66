// spec_bench.cu — does warp specialisation pay for a second-order interface job?  A "job" = gather 4 cells x 8 variables from shared
// memory + minmod reconstruction + the production HLLD solve + 8 flux stores.
//   mode A: every warp does whole jobs (what k_stage does today);
//   mode B: warp pairs — a producer warp gathers + reconstructs and hands the 16 states to its consumer warp through a 4 KB
//           shared-memory buffer (full / empty mbarriers); the consumer warp solves and stores.
// Prints SM cycles per job per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGX_FLAVOUR_FAST -fmad=true
#define GX_SOLVE_MASK 0xffffffffu
#include <cstdio>
#include "../guacho_b200/csrc/gx_physics.cuh"
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, int parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
constexpr int CX = 36, RY = 15, PC = CX * RY, NPL = 4;          // a few planes of primitives
__device__ __forceinline__ void gather(const double* c, int st, double (&wl)[8], double (&wr)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double pl = c[q * PC - st], pr = c[q * PC];
    gxp::reconstruct<GX_LIMITER_MINMOD>(c[q * PC - 2 * st], pl, pr, c[q * PC + st]);
    wl[q] = pl; wr[q] = pr;
  }
}
template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k(gxp::Phys P, double* out, long long* cyc, int iters) {
  extern __shared__ double sm[];
  double* ring = sm;                                  // [NPL][8][PC]
  double* fx = ring + NPL * 8 * PC;                   // [warps][8][32] flux stores
  double* buf = fx + 16 * 8 * 32;                     // [pairs][16][32] state hand-over
  __shared__ unsigned long long bars[32];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < NPL * 8 * PC; t += blockDim.x) {
    const int q = (t / PC) % 8;
    const double x = 1e-3 * (t % 977);
    ring[t] = q == 0 ? 1.0 + x : q == 4 ? 0.6 + x : q == 5 ? 0.4 : 0.3 * sin(0.01 * t);
  }
  if (tid == 0) for (int b = 0; b < 32; ++b) mbar_init(smem_u32(&bars[b]), 32);
  __syncthreads();
  int err = 0;
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      const int row = 2 + (wrp + it) % 11, pl = it % NPL, st = (it % 3 == 1) ? CX : 1;
      const double* c = ring + pl * 8 * PC + row * CX + 2 + lane;
      double wl[8], wr[8], ff[8];
      gather(c, st, wl, wr);
      gxp::PasInfo I;
      err |= gxp::riemann<GX_SOLVER_HLLD>(P, wl, wr, ff, I);
#pragma unroll
      for (int q = 0; q < 8; ++q) fx[(wrp * 8 + q) * 32 + lane] = ff[q];
    }
  } else {
    const int pair = wrp >> 1;
    const unsigned full = smem_u32(&bars[2 * pair]), empty = smem_u32(&bars[2 * pair + 1]);
    double* b = buf + pair * 16 * 32 + lane;
    if ((wrp & 1) == 0) {                             // producer
      for (int it = 0; it < iters; ++it) {
        const int row = 2 + (pair + it) % 11, pl = it % NPL, st = (it % 3 == 1) ? CX : 1;
        const double* c = ring + pl * 8 * PC + row * CX + 2 + lane;
        double wl[8], wr[8];
        gather(c, st, wl, wr);
        if (it > 0) mbar_wait(empty, (it - 1) & 1);
#pragma unroll
        for (int q = 0; q < 8; ++q) { b[q * 32] = wl[q]; b[(8 + q) * 32] = wr[q]; }
        mbar_arrive(full);
      }
    } else {                                          // consumer
      for (int it = 0; it < iters; ++it) {
        double wl[8], wr[8], ff[8];
        mbar_wait(full, it & 1);
#pragma unroll
        for (int q = 0; q < 8; ++q) { wl[q] = b[q * 32]; wr[q] = b[(8 + q) * 32]; }
        mbar_arrive(empty);
        gxp::PasInfo I;
        err |= gxp::riemann<GX_SOLVER_HLLD>(P, wl, wr, ff, I);
#pragma unroll
        for (int q = 0; q < 8; ++q) fx[(pair * 8 + q) * 32 + lane] = ff[q];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[tid] = fx[tid] + err;
  if (tid == 0) *cyc = t1 - t0;
}
template <int MODE, int MAXT> void run(int warps, const gxp::Phys& P, double* out, long long* cyc) {
  const int iters = 3000;
  const int smem = (NPL * 8 * PC + 16 * 8 * 32 + 8 * 16 * 32) * 8;
  cudaFuncSetAttribute(k<MODE, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE, MAXT><<<1, 32 * warps, smem>>>(P, out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<MODE, MAXT>);
  const double jobs = (double)iters * (MODE == 0 ? warps : warps / 2);
  printf("mode %c warps %2d (regs %3d, spill %zu B): %.0f cycles per job per scheduler  %s\n", MODE ? 'B' : 'A', warps, fa.numRegs, (size_t)fa.localSizeBytes,
         h / (jobs / 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  gxp::Phys P; P.cv = 1.5; P.gamma = 5.0 / 3.0; P.Tempsc = 1.0; P.inv_cv = 1.0 / 1.5; P.m4gamma = -4.0 * P.gamma; P.inv_Tempsc = 1.0; P.eos = 1; P.neqdyn = 8; P.npas = 0;
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  run<0, 256>(8, P, out, cyc);
  run<0, 384>(12, P, out, cyc);
  run<0, 512>(16, P, out, cyc);
  run<1, 256>(8, P, out, cyc);
  run<1, 384>(12, P, out, cyc);
  run<1, 512>(16, P, out, cyc);
  return 0;
}
