// Overhang fill for the deskewed volume (reference `_fill_overhang_torch`,
// biahub/deskew.py:339-368): mask = (vol == 0); dilate `it` times with a 3x3x3 cube (implicit
// -inf padding, so the array border never seeds); fill = fp32 mean of un-masked voxels or a
// constant; vol = mask ? fill : vol.
//
// `it` cube dilations compose into one (2*it+1)^3 cube = a separable running maximum, so the
// mask is built in three byte-volume passes (x, y, z); the z pass also reduces sum/count of the
// un-masked voxels (warp shuffle -> one double atomic per warp), and a last pass applies the fill.
#include "b2_common.cuh"

namespace b2 {

struct FillScratch {
  double sum;
  unsigned long long count;
};

__global__ void __launch_bounds__(256)
    fill_mask_x_kernel(const float* __restrict__ vol, uint8_t* __restrict__ m1, int64_t total, int X,
                       int r) {
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % X);
    const int64_t row = idx - x;
    const int lo = max(0, x - r), hi = min(X - 1, x + r);
    uint8_t m = 0;
    for (int xx = lo; xx <= hi; ++xx) m |= (vol[row + xx] == 0.0f) ? 1 : 0;
    m1[idx] = m;
  }
}

__global__ void __launch_bounds__(256)
    fill_mask_y_kernel(const uint8_t* __restrict__ m1, uint8_t* __restrict__ m2, int64_t total, int Y,
                       int X, int r) {
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>((idx / X) % Y);
    const int lo = max(0, y - r), hi = min(Y - 1, y + r);
    uint8_t m = 0;
    for (int yy = lo; yy <= hi; ++yy) m |= m1[idx + static_cast<int64_t>(yy - y) * X];
    m2[idx] = m;
  }
}

__global__ void __launch_bounds__(256)
    fill_mask_z_reduce_kernel(const float* __restrict__ vol, const uint8_t* __restrict__ m2,
                              uint8_t* __restrict__ m3, int64_t total, int Z, int64_t plane, int r,
                              FillScratch* scratch) {
  double sum = 0.0;
  unsigned long long cnt = 0;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(idx / plane);
    const int lo = max(0, z - r), hi = min(Z - 1, z + r);
    uint8_t m = 0;
    for (int zz = lo; zz <= hi; ++zz) m |= m2[idx + static_cast<int64_t>(zz - z) * plane];
    m3[idx] = m;
    if (!m) {
      sum += static_cast<double>(vol[idx]);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(&scratch->sum, sum);
    atomicAdd(&scratch->count, cnt);
  }
}

__global__ void __launch_bounds__(256)
    fill_apply_kernel(float* __restrict__ vol, const uint8_t* __restrict__ m3, int64_t total,
                      int use_mean, float fill_value, const FillScratch* scratch) {
  float fill = fill_value;
  if (use_mean) {
    // mean of an empty selection is NaN in the reference (torch mean of an empty tensor)
    fill = scratch->count ? static_cast<float>(scratch->sum / static_cast<double>(scratch->count))
                          : __int_as_float(0x7fc00000);
  }
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (m3[idx]) vol[idx] = fill;
  }
}

size_t fill_workspace_bytes(int64_t z, int64_t y, int64_t x) {
  const size_t vox = static_cast<size_t>(z) * y * x;
  const size_t vol_bytes = (vox + 255) / 256 * 256;
  return 2 * vol_bytes + 256;
}

int fill_device(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                int iterations, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!vol || !ws) {
    set_error("overhang_fill: null pointer");
    return B2_ERR_INVALID;
  }
  if (z < 1 || y < 1 || x < 1 || iterations < 0 || x > 2147483647LL || y > 2147483647LL ||
      z > 2147483647LL) {
    set_error("overhang_fill: invalid shape");
    return B2_ERR_INVALID;
  }
  if (ws_bytes < fill_workspace_bytes(z, y, x)) {
    set_error("overhang_fill: workspace too small (%zu < %zu)", ws_bytes,
              fill_workspace_bytes(z, y, x));
    return B2_ERR_INVALID;
  }
  const int64_t total = z * y * x;
  const size_t vol_bytes = (static_cast<size_t>(total) + 255) / 256 * 256;
  uint8_t* a = static_cast<uint8_t*>(ws);
  uint8_t* b = a + vol_bytes;
  FillScratch* scratch = reinterpret_cast<FillScratch*>(b + vol_bytes);
  int sms = 148;
  sm_count(&sms);
  const int64_t want = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms) * 32;
  const int grid = static_cast<int>(want < cap ? want : cap);
  B2_CUDA(cudaMemsetAsync(scratch, 0, sizeof(FillScratch), stream));
  fill_mask_x_kernel<<<grid, 256, 0, stream>>>(vol, a, total, (int)x, iterations);
  fill_mask_y_kernel<<<grid, 256, 0, stream>>>(a, b, total, (int)y, (int)x, iterations);
  fill_mask_z_reduce_kernel<<<grid, 256, 0, stream>>>(vol, b, a, total, (int)z, y * x, iterations,
                                                      scratch);
  fill_apply_kernel<<<grid, 256, 0, stream>>>(vol, a, total, use_mean, fill_value, scratch);
  B2_CUDA(cudaGetLastError());
  count_launch(4);
  return B2_OK;
}

}  // namespace b2
