// Exhaustive check of div_by_const (csrc/b2_deskew.cu): for ALL 2^32 bit patterns of the dividend
// a (finite, |a| < 2^26 — the coordinate range of the deskew kernels) and a set of divisors
// b = Zi - 1, the two-correction sequence with rb = RN(1/b) must equal __fdiv_rn(a, b) bit for bit.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/deskew_div_check scripts/deskew_div_check.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_by_const(float a, float b, float rb) {
  float q = __fmul_rn(a, rb);
  q = __fmaf_rn(__fmaf_rn(-b, q, a), rb, q);
  return __fmaf_rn(__fmaf_rn(-b, q, a), rb, q);
}
__global__ void check(float b, unsigned long long* bad, unsigned long long* n) {
  const float rb = __frcp_rn(b);
  unsigned long long lb = 0, ln = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32);
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float a = __uint_as_float((unsigned)i);
    if (!(fabsf(a) < 67108864.0f)) continue;  // also skips NaN / inf
    ++ln;
    const float q = div_by_const(a, b, rb), r = __fdiv_rn(a, b);
    if (__float_as_uint(q) != __float_as_uint(r) && !(q == 0.0f && r == 0.0f)) ++lb;
  }
  if (lb) atomicAdd(bad, lb);
  atomicAdd(n, ln);
}
int main() {
  unsigned long long *bad, *n;
  cudaMallocManaged(&bad, 8);
  cudaMallocManaged(&n, 8);
  const float bs[] = {1, 2, 3, 7, 63, 99, 127, 255, 256, 399, 511, 799, 999, 1023, 1399, 2047, 4095, 16383,
                      65535, 1048575, 8388607, 16777215};
  int rc = 0;
  for (float b : bs) {
    *bad = 0; *n = 0;
    check<<<148 * 16, 256>>>(b, bad, n);
    cudaError_t e = cudaDeviceSynchronize();
    printf("b = %9.0f: %llu dividends checked, mismatches %llu (%s)\n", b, *n, *bad, cudaGetErrorString(e));
    if (*bad || e != cudaSuccess) rc = 1;
  }
  return rc;
}
