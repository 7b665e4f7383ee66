// tma_test.cu — minimal check of the TMA tile load used by the stage kernel: 4-D FLOAT64 tensor (x, y, plane, variable), box (CX, RY, 1, NU)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int CX, int RY, int NU>
__global__ void k(const __grid_constant__ CUtensorMap tm, double* out, int c0, int c1, int c2, int mode) {
  extern __shared__ double sm[];
  double* ring = sm + (((smem_u32(sm) + 127u) & ~127u) - smem_u32(sm)) / 8u;
  __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
    if (mode & 2) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long tma = reinterpret_cast<unsigned long long>(&tm);
  if (threadIdx.x == (mode == 1 ? 32 : 0)) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(CX * RY * NU * 8)) : "memory");
    if (mode & 4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(ring)), "l"(tma), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(b) : "memory");
    else asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(ring)), "l"(tma), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(b), "r"(0) : "memory");
  for (int t = threadIdx.x; t < CX * RY * NU; t += blockDim.x) out[t] = ring[t];
}
template <int CX, int RY, int NU> int run(EncodeTiledFn enc, int px, int py, int pz, int mode) {
  const long long vs = (long long)px * py * pz;
  std::vector<double> h((size_t)vs * NU);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (double)t;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, (size_t)CX * RY * NU * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)pz, (cuuint64_t)NU};
  const cuuint64_t strides[3] = {(cuuint64_t)px * 8, (cuuint64_t)px * py * 8, (cuuint64_t)vs * 8};
  const cuuint32_t box[4] = {CX, RY, 1, NU}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   (mode & 8) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%dx1x%d mode %d: encode -> %d; ", CX, RY, NU, mode, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  const int smem = CX * RY * NU * 8 + 128;
  cudaFuncSetAttribute(k<CX, RY, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int c0 = (mode & 16) ? 15 : 14, c1 = 0, c2 = 3;
  k<CX, RY, NU><<<1, 384, smem>>>(tm, o, c0, c1, c2, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel -> %s; ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> g((size_t)CX * RY * NU);
    cudaMemcpy(g.data(), o, g.size() * 8, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int v = 0; v < NU; ++v) for (int y = 0; y < RY; ++y) for (int x = 0; x < CX; ++x) {
      const double want = (double)(v * vs + ((long long)c2 * py + (c1 + y)) * px + (c0 + x));
      if (g[((size_t)v * RY + y) * CX + x] != want) ++bad;
    }
    printf("mismatches %lld", bad);
  }
  printf("\n");
  cudaFree(d); cudaFree(o);
  return e != cudaSuccess;
}
int main() {
  cudaDriverEntryPointQueryResult qr; void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { printf("no encode entry point\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  if (run<32, 8, 4>(enc, 96, 48, 12, 0)) return 0;     // 256-byte inner box
  if (run<34, 13, 8>(enc, 96, 48, 12, 0)) return 0;    // the first-order stage box
  if (run<36, 15, 8>(enc, 96, 48, 12, 0)) return 0;    // the second-order stage box
  if (run<36, 15, 8>(enc, 96, 48, 12, 1)) return 0;    // issued from another warp
  if (run<36, 15, 8>(enc, 96, 48, 12, 8)) return 0;    // + L2 promotion 256 B
  if (run<36, 15, 8>(enc, 96, 48, 12, 4)) return 0;    // + .tile qualifier
  if (run<36, 15, 8>(enc, 96, 48, 12, 2)) return 0;    // + fence.mbarrier_init
  if (run<36, 15, 8>(enc, 96, 48, 12, 16)) return 0;   // odd start column: box start not 16-byte aligned
  return 0;
}
