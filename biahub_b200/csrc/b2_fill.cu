// Overhang fill for the deskewed volume (reference `_fill_overhang_torch`,
// biahub/deskew.py:339-368): mask = (vol == 0); dilate `it` times with a 3x3x3 cube (implicit
// -inf padding, so the array border never seeds); fill = fp32 mean of un-masked voxels or a
// constant; vol = mask ? fill : vol.
//
// `it` cube dilations compose into one (2*it+1)^3 cube = a separable running maximum.  The mask
// is kept as BITS (one uint32 word per 32 voxels along x, 1/32 byte per voxel):
//   A  zero test: a warp reads 32 consecutive voxels (one 128-byte line), __ballot_sync -> word
//   B  x and y dilation on the words (shift/or across word boundaries, or over 2r+1 rows)
//   C  z dilation (or over 2r+1 planes) + fp64 sum / count of the un-masked voxels (warp shuffle,
//      one atomic per warp) — or, for a constant fill, the fill itself
//   D  (mean only) write the fill where the final mask is set; untouched words skip their row
// HBM traffic: the volume is read in A and C and written only where masked; the words are 3 % of
// it.  (The first version made four byte-per-voxel passes with 7-tap loops: 13.8 ms on the
// mantis keep_overhang volume; this one: see DESIGN.md.)
#include <algorithm>
#include <utility>

#include "b2_common.cuh"

namespace b2 {

struct FillScratch {
  double sum;
  unsigned long long count;
};

constexpr int kFillThreads = 256;

// All kernels: one CTA per (z, y) row (blockIdx.x = z * Y + y), its 8 warps stride over the row's
// W words; no per-word index division, four independent words per warp iteration.
constexpr int kFillWarps = kFillThreads / 32;
constexpr int kFillUnroll = 4;
constexpr int kFillRowsPerCta = 8;  // rows per CTA: a 9 KB row alone is too little work per CTA

// A: bits[word] = ballot(vol == 0) for the 32 voxels of the word
__global__ void __launch_bounds__(kFillThreads)
    fill_bits_kernel(const float* __restrict__ vol, uint32_t* __restrict__ bits, int W, int X,
                     int64_t rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kFillRowsPerCta;
       row < rows && row < (static_cast<int64_t>(blockIdx.x) + 1) * kFillRowsPerCta; ++row) {
  const float* __restrict__ rowp = vol + row * X;
  uint32_t* __restrict__ wp = bits + row * W;
  for (int w0 = warp; w0 < W; w0 += kFillWarps * kFillUnroll) {
    float v[kFillUnroll];
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      const int x = (w0 + u * kFillWarps) * 32 + lane;
      v[u] = x < X ? __ldcs(rowp + x) : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      const uint32_t b = __ballot_sync(0xffffffffu, v[u] == 0.0f);
      const int w = w0 + u * kFillWarps;
      if (lane == 0 && w < W) wp[w] = b;
    }
  }
  }
}

// B: dilation by r along x (inside the row) and y (inside the plane)
__global__ void __launch_bounds__(kFillThreads)
    fill_dilate_xy_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int W, int Y,
                          int r, int64_t rows) {
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kFillRowsPerCta;
       row < rows && row < (static_cast<int64_t>(blockIdx.x) + 1) * kFillRowsPerCta; ++row) {
  const int y = static_cast<int>(row % Y);
  const int lo = max(0, y - r), hi = min(Y - 1, y + r);
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    // separable: OR the (left, centre, right) words over the 2r+1 rows first, then ONE x
    // dilation with funnel shifts (2 SHF + 2 LOP per step instead of 8 operations per row)
    uint32_t c = 0, l = 0, g = 0;
    for (int yy = lo; yy <= hi; ++yy) {
      const uint32_t* p = in + (row + (yy - y)) * W + w;
      c |= __ldg(p);
      if (w > 0) l |= __ldg(p - 1);
      if (w + 1 < W) g |= __ldg(p + 1);
    }
    uint32_t acc = c;
    for (int s = 1; s <= r; ++s) acc |= __funnelshift_l(l, c, s) | __funnelshift_r(c, g, s);
    out[row * W + w] = acc;
  }
  }
}

// C: dilation by r along z; REDUCE: sum / count of the un-masked voxels, final words stored;
//    !REDUCE (constant fill): the masked voxels are overwritten right away
template <bool REDUCE>
__global__ void __launch_bounds__(kFillThreads)
    fill_dilate_z_kernel(float* __restrict__ vol, const uint32_t* __restrict__ in,
                         uint32_t* __restrict__ out, int W, int X, int Y, int Z, int r,
                         float fill_value, FillScratch* scratch, int64_t rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t plane_words = static_cast<int64_t>(Y) * W;
  double sum = 0.0;
  unsigned long long cnt = 0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kFillRowsPerCta;
       row < rows && row < (static_cast<int64_t>(blockIdx.x) + 1) * kFillRowsPerCta; ++row) {
  const int z = static_cast<int>(row / Y);
  const int lo = max(0, z - r), hi = min(Z - 1, z + r);
  const uint32_t* __restrict__ inrow = in + row * W;
  float* __restrict__ rowp = vol + row * X;
  for (int w0 = warp; w0 < W; w0 += kFillWarps * kFillUnroll) {
    // lane k fetches plane lo + k of the window (2r+1 <= 31 planes), the warp ORs them;
    // kFillUnroll independent words per iteration keep several loads in flight
    uint32_t m[kFillUnroll];
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      const int w = w0 + u * kFillWarps;
      m[u] = (w < W && lo + lane <= hi)
                 ? __ldg(inrow + w + static_cast<int64_t>(lo + lane - z) * plane_words) : 0u;
    }
    float v[kFillUnroll];
    bool take[kFillUnroll];
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      m[u] = __reduce_or_sync(0xffffffffu, m[u]);
      const int w = w0 + u * kFillWarps;
      const int x = w * 32 + lane;
      const bool masked = (m[u] >> lane) & 1u;
      take[u] = false;
      v[u] = 0.0f;
      if (REDUCE) {
        if (lane == 0 && w < W) out[row * W + w] = m[u];
        take[u] = !masked && w < W && x < X;
        if (take[u]) v[u] = __ldcs(rowp + x);
      } else if (masked && w < W && x < X) {
        rowp[x] = fill_value;
      }
    }
    if (REDUCE) {
#pragma unroll
      for (int u = 0; u < kFillUnroll; ++u) {
        if (take[u]) {
          sum += static_cast<double>(v[u]);
          ++cnt;
        }
      }
    }
  }
  }
  if (REDUCE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_sum[kFillWarps];
    __shared__ unsigned long long s_cnt[kFillWarps];
    if (lane == 0) {
      s_sum[warp] = sum;
      s_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      unsigned long long c = 0;
#pragma unroll
      for (int i = 0; i < kFillWarps; ++i) {
        t += s_sum[i];
        c += s_cnt[i];
      }
      if (c) {
        atomicAdd(&scratch->sum, t);
        atomicAdd(&scratch->count, c);
      }
    }
  }
}

// D: vol = mean where the final mask is set
__global__ void __launch_bounds__(kFillThreads)
    fill_apply_mean_kernel(float* __restrict__ vol, const uint32_t* __restrict__ bits, int W, int X,
                           const FillScratch* scratch, int64_t rows) {
  // mean of an empty selection is NaN in the reference (torch mean of an empty tensor)
  const float fill = scratch->count
                         ? static_cast<float>(scratch->sum / static_cast<double>(scratch->count))
                         : __int_as_float(0x7fc00000);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kFillRowsPerCta;
       row < rows && row < (static_cast<int64_t>(blockIdx.x) + 1) * kFillRowsPerCta; ++row) {
  const uint32_t* __restrict__ wp = bits + row * W;
  float* __restrict__ rowp = vol + row * X;
  for (int w0 = warp; w0 < W; w0 += kFillWarps * kFillUnroll) {
    uint32_t m[kFillUnroll];
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      const int w = w0 + u * kFillWarps;
      m[u] = w < W ? __ldg(wp + w) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kFillUnroll; ++u) {
      const int x = (w0 + u * kFillWarps) * 32 + lane;
      if (((m[u] >> lane) & 1u) && x < X) rowp[x] = fill;
    }
  }
  }
}

// One iteration of scipy.ndimage.binary_dilation with its default structuring element (the 3-D
// cross, 6-connectivity; border_value 0) on the bit words — the numpy variant of the fill used by
// the legacy `deskew_zyx` (reference biahub/deskew.py:277-336, `_fill_overhang_with_mean`).
__global__ void __launch_bounds__(kFillThreads)
    fill_dilate_cross_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int W, int X,
                             int Y, int Z, int64_t rows) {
  const int64_t plane_words = static_cast<int64_t>(Y) * W;
  const uint32_t tail = (X & 31) ? ((1u << (X & 31)) - 1u) : 0xffffffffu;  // valid bits of the last word
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kFillRowsPerCta;
       row < rows && row < (static_cast<int64_t>(blockIdx.x) + 1) * kFillRowsPerCta; ++row) {
    const int y = static_cast<int>(row % Y), z = static_cast<int>(row / Y);
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
      const uint32_t* p = in + row * W + w;
      const uint32_t c = __ldg(p);
      const uint32_t l = w > 0 ? __ldg(p - 1) : 0u;
      const uint32_t g = w + 1 < W ? __ldg(p + 1) : 0u;
      uint32_t acc = c | __funnelshift_l(l, c, 1) | __funnelshift_r(c, g, 1);
      if (y > 0) acc |= __ldg(p - W);
      if (y + 1 < Y) acc |= __ldg(p + W);
      if (z > 0) acc |= __ldg(p - plane_words);
      if (z + 1 < Z) acc |= __ldg(p + plane_words);
      if (w == W - 1) acc &= tail;
      out[row * W + w] = acc;
    }
  }
}

// Legacy averaging (reference `_average_n_slices_torch`, biahub/deskew.py:71-96): mean over
// groups of n slices of the DESKEWED stack, the last slice repeated to fill the last group.
// fp32 sequential sum, true division (torch.mean over a short dimension).
__global__ void __launch_bounds__(256)
    average_slices_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t plane,
                          int Z, int n, int64_t total) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t a = i / plane, r = i - a * plane;
    float s = 0.0f;
    for (int k = 0; k < n; ++k) {
      const int64_t z = min(a * n + k, static_cast<int64_t>(Z - 1));
      s = __fadd_rn(s, __ldcs(src + z * plane + r));
    }
    dst[i] = __fdiv_rn(s, static_cast<float>(n));
  }
}

int average_slices_device(const float* src, int64_t z, int64_t plane, int n, float* dst,
                          cudaStream_t stream) {
  if (!src || !dst || z < 1 || plane < 1 || n < 1 || z > 2147483647LL) {
    set_error("average_slices: invalid argument");
    return B2_ERR_INVALID;
  }
  const int64_t za = (z + n - 1) / n;
  const int64_t total = za * plane;
  int sms = 148;
  sm_count(&sms);
  const int64_t want = (total + 255) / 256;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(want, static_cast<int64_t>(sms) * 32));
  average_slices_kernel<<<grid, 256, 0, stream>>>(src, dst, plane, static_cast<int>(z), n, total);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

size_t fill_workspace_bytes(int64_t z, int64_t y, int64_t x) {
  const size_t words = static_cast<size_t>(z) * y * ((x + 31) / 32);
  const size_t arr = (words * 4 + 255) / 256 * 256;
  return 2 * arr + 256;
}

int fill_device_ex(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                   int iterations, int connectivity, void* ws, size_t ws_bytes, cudaStream_t stream);

int fill_device(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                int iterations, void* ws, size_t ws_bytes, cudaStream_t stream) {
  return fill_device_ex(vol, z, y, x, use_mean, fill_value, iterations, 26, ws, ws_bytes, stream);
}

// connectivity 26: 3x3x3 cube per iteration (torch variant, max_pool3d); 6: 3-D cross per
// iteration (numpy variant, scipy binary_dilation default)
int fill_device_ex(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                   int iterations, int connectivity, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!vol || !ws) {
    set_error("overhang_fill: null pointer");
    return B2_ERR_INVALID;
  }
  if (z < 1 || y < 1 || x < 1 || iterations < 0 || x > 2147483647LL || y > 2147483647LL ||
      z > 2147483647LL) {
    set_error("overhang_fill: invalid shape");
    return B2_ERR_INVALID;
  }
  if (connectivity != 26 && connectivity != 6) {
    set_error("overhang_fill: connectivity must be 26 (cube) or 6 (cross)");
    return B2_ERR_INVALID;
  }
  if (iterations > 15) {
    set_error("overhang_fill: at most 15 dilation iterations (the reference uses 3)");
    return B2_ERR_UNSUPPORTED;
  }
  if (ws_bytes < fill_workspace_bytes(z, y, x)) {
    set_error("overhang_fill: workspace too small (%zu < %zu)", ws_bytes,
              fill_workspace_bytes(z, y, x));
    return B2_ERR_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(ws) % 8 != 0) {
    set_error("overhang_fill: workspace must be 8-byte aligned");
    return B2_ERR_INVALID;
  }
  const int W = static_cast<int>((x + 31) / 32);
  const size_t arr = (static_cast<size_t>(z) * y * W * 4 + 255) / 256 * 256;
  uint32_t* a = static_cast<uint32_t*>(ws);
  uint32_t* b = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + arr);
  FillScratch* scratch = reinterpret_cast<FillScratch*>(static_cast<char*>(ws) + 2 * arr);
  const int64_t rows = z * y;
  if (rows > 2147483647LL) {
    set_error("overhang_fill: too many rows");
    return B2_ERR_INVALID;
  }
  const unsigned grid = static_cast<unsigned>((rows + kFillRowsPerCta - 1) / kFillRowsPerCta);
  int r = iterations;
  B2_CUDA(cudaMemsetAsync(scratch, 0, sizeof(FillScratch), stream));
  fill_bits_kernel<<<grid, kFillThreads, 0, stream>>>(vol, a, W, (int)x, rows);
  // one thread per word of the row: a block only as wide as the row has words
  const int bt = W >= kFillThreads ? kFillThreads : (W + 31) / 32 * 32;
  if (connectivity == 6) {
    // `iterations` cross dilations ping-pong between the two word arrays; the final pass below
    // reads b, so an even count takes one more plain copy pass (xy dilation with r = 0)
    uint32_t* cur = a;
    uint32_t* nxt = b;
    for (int it = 0; it < iterations; ++it) {
      fill_dilate_cross_kernel<<<grid, bt, 0, stream>>>(cur, nxt, W, (int)x, (int)y, (int)z, rows);
      count_launch();
      std::swap(cur, nxt);
    }
    if (cur != b) {
      fill_dilate_xy_kernel<<<grid, bt, 0, stream>>>(cur, b, W, (int)y, 0, rows);  // copy
      count_launch();
    }
    r = 0;  // the final pass only reduces / fills
  } else {
    fill_dilate_xy_kernel<<<grid, bt, 0, stream>>>(a, b, W, (int)y, r, rows);
  }
  if (use_mean) {
    fill_dilate_z_kernel<true><<<grid, kFillThreads, 0, stream>>>(vol, b, a, W, (int)x, (int)y, (int)z,
                                                                  r, 0.0f, scratch, rows);
    fill_apply_mean_kernel<<<grid, kFillThreads, 0, stream>>>(vol, a, W, (int)x, scratch, rows);
    count_launch(4);
  } else {
    fill_dilate_z_kernel<false><<<grid, kFillThreads, 0, stream>>>(vol, b, a, W, (int)x, (int)y,
                                                                   (int)z, r, fill_value, scratch, rows);
    count_launch(3);
  }
  B2_CUDA(cudaGetLastError());
  return B2_OK;
}

}  // namespace b2
