// Exhaustive check of the division shortcut in csrc/b2_flatfield.cu: for every uint16 sample v and
// every possible median p (half-integers 0.5 .. 65535), q = fma(fma(-p, v*r, v), r, v*r) with
// r = RN(1/p) must equal the IEEE quotient __ddiv_rn(v, p).  8.6e9 pairs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/flatfield_div_check scripts/flatfield_div_check.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void check(unsigned long long* bad) {
  const unsigned s2 = blockIdx.x * blockDim.x + threadIdx.x + 1;  // 2 * p
  if (s2 > 131070u) return;
  const double p = 0.5 * (double)s2;
  const double r = __drcp_rn(p);
  unsigned long long local = 0;
  for (unsigned v = 0; v < 65536u; ++v) {
    const double dv = (double)v;
    const double q0 = __dmul_rn(dv, r);
    const double rem = __fma_rn(-p, q0, dv);
    const double q = __fma_rn(rem, r, q0);
    if (q != __ddiv_rn(dv, p)) ++local;
  }
  if (local) atomicAdd(bad, local);
}
int main() {
  unsigned long long* bad;
  cudaMallocManaged(&bad, 8);
  *bad = 0;
  check<<<(131070 + 255) / 256, 256>>>(bad);
  cudaError_t e = cudaDeviceSynchronize();
  printf("cuda: %s; pairs checked: %llu; mismatches: %llu\n", cudaGetErrorString(e),
         131070ull * 65536ull, *bad);
  return (*bad == 0 && e == cudaSuccess) ? 0 : 1;
}
