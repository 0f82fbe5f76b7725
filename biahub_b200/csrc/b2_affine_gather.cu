// Generic affine pull warp: one thread per output voxel, float64 coordinates in the oracle's op
// order, taps fetched with LDG through L1/L2.  Correct for ANY 3x4 matrix, shape and alignment;
// it is the fallback of the TMA-staged kernels (b2_affine_zsep.cu), not a CPU fallback.
#include <cstdlib>

#include "b2_affine.cuh"

namespace b2 {

constexpr int kGatherXPerThread = 4;

// grid.x = output rows (z*oy + y), grid.y = chunks of 256*kGatherXPerThread voxels along x
template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
__global__ void __launch_bounds__(256) affine_gather_kernel(const AffineParams p) {
  const int row = blockIdx.x;
  const int z = row / p.oy;
  const int y = row - z * p.oy;
  double rp[3];
  affine_row_part(p, z, y, rp);
  float* __restrict__ out = p.dst + static_cast<int64_t>(row) * p.dpitch;
  const int x0 = blockIdx.y * (256 * kGatherXPerThread) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kGatherXPerThread; ++i) {
    const int x = x0 + i * 256;
    if (x < p.ox) out[x] = affine_sample_row<T, ORDER, BOUNDARY, SCRUB>(p, rp, x);
  }
}

template <typename T, int ORDER, int BOUNDARY>
static int launch_gather_scrub(const AffineParams& p, cudaStream_t stream) {
  const int64_t rows = static_cast<int64_t>(p.oz) * p.oy;
  if (rows == 0 || p.ox == 0) return B2_OK;
  const int xchunks = (p.ox + 256 * kGatherXPerThread - 1) / (256 * kGatherXPerThread);
  if (rows > 2147483647LL || xchunks > 65535) {
    set_error("affine3d: output too large for the gather kernel grid");
    return B2_ERR_INVALID;
  }
  const dim3 grid(static_cast<unsigned>(rows), static_cast<unsigned>(xchunks), 1);
  if (p.scrub && sizeof(T) == 4)
    affine_gather_kernel<T, ORDER, BOUNDARY, true><<<grid, 256, 0, stream>>>(p);
  else
    affine_gather_kernel<T, ORDER, BOUNDARY, false><<<grid, 256, 0, stream>>>(p);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

template <typename T>
static int launch_gather_typed(const AffineParams& p, cudaStream_t stream) {
  if (p.order == 0) {
    if (p.boundary == B2_BOUNDARY_CONSTANT)
      return launch_gather_scrub<T, 0, B2_BOUNDARY_CONSTANT>(p, stream);
    return launch_gather_scrub<T, 0, B2_BOUNDARY_ITK>(p, stream);
  }
  if (p.boundary == B2_BOUNDARY_CONSTANT)
    return launch_gather_scrub<T, 1, B2_BOUNDARY_CONSTANT>(p, stream);
  return launch_gather_scrub<T, 1, B2_BOUNDARY_ITK>(p, stream);
}

int affine_gather_launch(const AffineParams& p, int src_dtype, cudaStream_t stream) {
  if (src_dtype == B2_DTYPE_U16) return launch_gather_typed<uint16_t>(p, stream);
  return launch_gather_typed<float>(p, stream);
}

// ---------------------------------------------------------------------------------------------
// entry point shared by both affine kernels
// ---------------------------------------------------------------------------------------------
int affine_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* dst,
                  int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                  int order, int boundary, int scrub, int path, cudaStream_t stream,
                  int64_t src_row_pitch, int64_t dst_row_pitch) {
  if (!src || !dst || !M12) {
    set_error("affine3d: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("affine3d: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (order != 0 && order != 1) {
    set_error("affine3d: order must be 0 (nearest) or 1 (linear), got %d", order);
    return B2_ERR_INVALID;
  }
  if (boundary != B2_BOUNDARY_CONSTANT && boundary != B2_BOUNDARY_ITK) {
    set_error("affine3d: unknown boundary %d", boundary);
    return B2_ERR_INVALID;
  }
  const int64_t lim = 2147483647LL;
  if (sz < 1 || sy < 1 || sx < 1 || oz < 0 || oy < 0 || ox < 0 || sz > lim || sy > lim ||
      sx > lim || oz > lim || oy > lim || ox > lim) {
    set_error("affine3d: invalid shape");
    return B2_ERR_INVALID;
  }
  for (int i = 0; i < 12; ++i) {
    if (!(M12[i] == M12[i]) || M12[i] > 1e300 || M12[i] < -1e300) {
      set_error("affine3d: matrix entry %d is not finite", i);
      return B2_ERR_INVALID;
    }
  }
  AffineParams p;
  p.src = src;
  p.dst = dst;
  p.sz = (int)sz; p.sy = (int)sy; p.sx = (int)sx;
  p.oz = (int)oz; p.oy = (int)oy; p.ox = (int)ox;
  p.cz = crop_start ? (int)crop_start[0] : 0;
  p.cy = crop_start ? (int)crop_start[1] : 0;
  p.cx = crop_start ? (int)crop_start[2] : 0;
  for (int i = 0; i < 12; ++i) p.m[i] = M12[i];
  p.order = order;
  p.boundary = boundary;
  p.scrub = scrub ? 1 : 0;
  if ((src_row_pitch != 0 && (src_row_pitch < sx || src_row_pitch > lim)) ||
      (dst_row_pitch != 0 && (dst_row_pitch < ox || dst_row_pitch > lim))) {
    set_error("affine3d: row pitch smaller than the row length");
    return B2_ERR_INVALID;
  }
  p.spitch = src_row_pitch ? (int)src_row_pitch : (int)sx;
  p.dpitch = dst_row_pitch ? (int)dst_row_pitch : (int)ox;
  if (oz == 0 || oy == 0 || ox == 0) return B2_OK;

  if (path != B2_PATH_GATHER) {
    bool eligible = false;
    int rc = affine_zsep_launch(p, src_dtype, stream, &eligible);
    if (eligible) return rc;
    rc = affine_brick_launch(p, src_dtype, stream, &eligible);  // generic matrices
    if (eligible) return rc;
    if (path == B2_PATH_TMA) {
      set_error("affine3d: TMA paths not eligible (need 16-byte aligned source rows and a "
                "back-projected tile brick that fits shared memory)");
      return B2_ERR_UNSUPPORTED;
    }
  }
  return affine_gather_launch(p, src_dtype, stream);
}

}  // namespace b2
