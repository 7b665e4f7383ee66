// tma_test2.cu — which async-copy forms work here: (A) cp.async.bulk 1-D, (B) 2-D FLOAT32 tensor, (C) 4-D FLOAT64 tensor with the descriptor in global memory
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(unsigned b) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(b), "r"(0) : "memory");
}
__global__ void kA(const double* src, double* out, int n) {       // n doubles, multiple of 2
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(n * 8)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(src), "r"((unsigned)(n * 8)), "r"(b) : "memory");
  }
  wait0(b);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = sm[t];
}
__global__ void kB(const __grid_constant__ CUtensorMap tm, float* out, int n) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(n * 4)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(0), "r"(0), "r"(b) : "memory");
  }
  wait0(b);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = ((float*)sm)[t];
}
__global__ void kC(const CUtensorMap* tm, double* out, int n) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(n * 8)) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(0), "r"(0), "r"(0), "r"(0), "r"(b) : "memory");
  }
  wait0(b);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = sm[t];
}
int main() {
  cudaDriverEntryPointQueryResult qr; void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<double> h(1 << 16);
  for (size_t t = 0; t < h.size(); ++t) h[t] = (double)t;
  double *d, *o; cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 1 << 16);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  {   // A
    kA<<<1, 128, 8192>>>(d + 14, o, 36);
    cudaError_t e = cudaDeviceSynchronize();
    double g[36]; if (e == cudaSuccess) cudaMemcpy(g, o, sizeof g, cudaMemcpyDeviceToHost);
    printf("A cp.async.bulk 288 B: %s %s\n", cudaGetErrorString(e), (e == cudaSuccess && g[0] == 14.0 && g[35] == 49.0) ? "data ok" : "data ?");
    if (e != cudaSuccess) return 0;
  }
  {   // B: 2-D float32 64 x 8 box of a 256 x 64 tensor
    std::vector<float> hf(256 * 64); for (size_t t = 0; t < hf.size(); ++t) hf[t] = (float)t;
    float *df, *of; cudaMalloc(&df, hf.size() * 4); cudaMalloc(&of, 64 * 8 * 4);
    cudaMemcpy(df, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm; const cuuint64_t dims[2] = {256, 64}, strides[1] = {256 * 4}; const cuuint32_t box[2] = {64, 8}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, df, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    kB<<<1, 128, 8192>>>(tm, of, 64 * 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("B 2-D f32 tensor (encode %d): %s\n", (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
  }
  {   // C: 4-D f64, descriptor in global memory
    CUtensorMap tm; const cuuint64_t dims[4] = {64, 16, 8, 4}, strides[3] = {64 * 8, 64 * 16 * 8, 64 * 16 * 8 * 8}; const cuuint32_t box[4] = {32, 8, 1, 4}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof tm); cudaMemcpy(dtm, &tm, sizeof tm, cudaMemcpyHostToDevice);
    kC<<<1, 128, 16384>>>(dtm, o, 32 * 8 * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("C 4-D f64 tensor, descriptor in global memory (encode %d): %s\n", (int)r, cudaGetErrorString(e));
  }
  return 0;
}
