// Standalone probe: which TMA tiled-box configurations execute on this GPU?
//   tma_probe <bx> <by> <bz> <swizzle 0|1|2|3> <elem 2|4> <smem_align>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool GLOBAL>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int bytes, int align, int c0, int c1, int c2, unsigned* out) {
  const CUtensorMap* mp = GLOBAL ? gmap : &pmap;
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  uint32_t dst = (smem_u32(raw) + align - 1) & ~(uint32_t)(align - 1);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)mp), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  unsigned sum = 0;
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(dst + 4 * i)); sum += v; }
  atomicAdd(out, sum);
}
int main(int argc, char** argv) {
  int bx = atoi(argv[1]), by = atoi(argv[2]), bz = atoi(argv[3]), sw = atoi(argv[4]), es = atoi(argv[5]), align = atoi(argv[6]);
  const int X = 512, Y = 300, Z = 40;
  void* d; cudaMalloc(&d, (size_t)X * Y * Z * es); cudaMemset(d, 1, (size_t)X * Y * Z * es);
  unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t gdim[3] = {X, Y, Z}; cuuint64_t gs[2] = {(cuuint64_t)X * es, (cuuint64_t)X * Y * es};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}; cuuint32_t est[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)sym)(&map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gs, box, est,
      CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  int bytes = bx * by * bz * es; int smem = bytes + align;
  int useg = argc > 7 ? atoi(argv[7]) : 0; int c0 = argc > 8 ? atoi(argv[8]) : 0, c1 = argc > 9 ? atoi(argv[9]) : 0, c2 = argc > 10 ? atoi(argv[10]) : 0;
  CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (useg) { cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); probe<true><<<1, 128, smem>>>(map, gmap, bytes, align, c0, c1, c2, out); }
  else { cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); probe<false><<<1, 128, smem>>>(map, gmap, bytes, align, c0, c1, c2, out); }
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("box=(%d,%d,%d) sw=%d es=%d align=%d bytes=%d global=%d c=(%d,%d,%d) -> %s sum=%u\n", bx, by, bz, sw, es, align, bytes, useg, c0, c1, c2, cudaGetErrorString(e), h);
  return e == cudaSuccess ? 0 : 1;
}
