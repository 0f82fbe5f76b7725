// Shared definitions of the affine pull-warp kernels (b2_affine_gather.cu, b2_affine_zsep.cu).
//
// Semantics (reference biahub/register.py:202-281):
//   c = M[:, :3] @ o + M[:, 3]   with o in the UNCROPPED output frame, evaluated in float64 in
//   scipy's op order  c_d = ((t_d + z*m_d0) + y*m_d1) + x*m_d2  (accumulator starts at the shift,
//   separate multiply/add; probed against scipy 1.18.1, see oracle/affine_oracle.py).
//   boundary CONSTANT: outside [0, n-1] on any axis -> 0;  ITK: outside [-0.5, n-0.5) -> 0 and
//   clamp-to-edge inside the half-voxel band.  order 0: index floor(c + 0.5); order 1: trilinear.
//   Source scrub (np.nan_to_num(nan=0), biahub/register.py:254): NaN -> 0, +-inf -> +-FLT_MAX.
#pragma once

#include <float.h>

#include "b2_common.cuh"

namespace b2 {

struct AffineParams {
  const void* src;
  float* dst;
  int sz, sy, sx;     // source shape
  int oz, oy, ox;     // (cropped) output shape
  int cz, cy, cx;     // crop start in the uncropped output frame
  double m[12];       // row-major 3x4
  int order;          // 0 | 1
  int boundary;       // B2_BOUNDARY_*
  int scrub;          // scrub NaN/inf on load (float32 sources only)
  int spitch, dpitch; // source / output row pitch in elements (planes are rows*pitch apart)
};

#ifdef __CUDACC__

struct AxisTap {
  int i0, i1;  // tap indices (always valid indices when inside)
  float w;     // weight of i1 (order 1); 0 for order 0
  bool inside;
};

// Resolve one axis of the pull coordinate into tap indices + weight, in float64.
template <int ORDER, int BOUNDARY>
__device__ __forceinline__ AxisTap resolve_axis(double c, int n) {
  AxisTap t;
  const double last = static_cast<double>(n - 1);
  if (BOUNDARY == B2_BOUNDARY_CONSTANT) {
    t.inside = (c >= 0.0) && (c <= last);
  } else {
    t.inside = (c >= -0.5) && (c < last + 0.5);
  }
  // keep the conversion to int well-defined for far-away / non-finite coordinates
  const double cc = t.inside ? c : 0.0;
  if (ORDER == 0) {
    int i = __double2int_rd(__dadd_rn(cc, 0.5));
    i = max(0, min(i, n - 1));
    t.i0 = i;
    t.i1 = i;
    t.w = 0.0f;
  } else {
    int b = __double2int_rd(cc);
    b = max(0, min(b, n - 1));  // ITK: base clamped to the start index
    double d = __dsub_rn(cc, static_cast<double>(b));
    d = d < 0.0 ? 0.0 : d;  // ITK: non-positive distance -> no blend
    const bool has_next = (b + 1) <= (n - 1);
    t.i0 = b;
    t.i1 = has_next ? b + 1 : b;
    t.w = has_next ? static_cast<float>(d) : 0.0f;  // neighbour beyond the edge is dropped
  }
  return t;
}

__device__ __forceinline__ float scrub_value(float v) {
  if (v != v) return 0.0f;
  return fminf(fmaxf(v, -FLT_MAX), FLT_MAX);
}

template <typename T, bool SCRUB>
__device__ __forceinline__ float load_tap(const T* __restrict__ p) {
  const float v = to_f32<T>(__ldg(p));
  if (SCRUB && sizeof(T) == 4) return scrub_value(v);
  return v;
}

__device__ __forceinline__ float lerp_w(float v0, float v1, float w) {
  // (1-w)*v0 + w*v1 : exact v0 when w == 0 (integer shifts), never forms v1 - v0
  return __fmaf_rn(w, v1, __fmul_rn(__fsub_rn(1.0f, w), v0));
}

// One output voxel of the generic warp: float64 coordinates in the oracle's op order, taps via LDG.
// (t_d + z*m_d0) + y*m_d1 for the three axes: the part of the coordinate shared by a whole row
__device__ __forceinline__ void affine_row_part(const AffineParams& p, int z, int y, double (&rp)[3]) {
  const double zf = static_cast<double>(z + p.cz);
  const double yf = static_cast<double>(y + p.cy);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double* m = p.m + 4 * d;
    rp[d] = __dadd_rn(__dadd_rn(m[3], __dmul_rn(zf, m[0])), __dmul_rn(yf, m[1]));
  }
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
__device__ __forceinline__ float affine_sample_row(const AffineParams& p, const double (&rp)[3],
                                                   int x) {
  const T* __restrict__ src = static_cast<const T*>(p.src);
  const int64_t sxy = static_cast<int64_t>(p.sy) * p.spitch;
  const double xf = static_cast<double>(x + p.cx);
  double c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) c[d] = __dadd_rn(rp[d], __dmul_rn(xf, p.m[4 * d + 2]));
  const AxisTap tz = resolve_axis<ORDER, BOUNDARY>(c[0], p.sz);
  const AxisTap ty = resolve_axis<ORDER, BOUNDARY>(c[1], p.sy);
  const AxisTap tx = resolve_axis<ORDER, BOUNDARY>(c[2], p.sx);
  if (!(tz.inside && ty.inside && tx.inside)) return 0.0f;
  const T* b0 = src + tz.i0 * sxy;
  if (ORDER == 0) return load_tap<T, SCRUB>(b0 + static_cast<int64_t>(ty.i0) * p.spitch + tx.i0);
  const T* b1 = src + tz.i1 * sxy;
  const int64_t r0 = static_cast<int64_t>(ty.i0) * p.spitch;
  const int64_t r1 = static_cast<int64_t>(ty.i1) * p.spitch;
  const float v000 = load_tap<T, SCRUB>(b0 + r0 + tx.i0);
  const float v001 = load_tap<T, SCRUB>(b0 + r0 + tx.i1);
  const float v010 = load_tap<T, SCRUB>(b0 + r1 + tx.i0);
  const float v011 = load_tap<T, SCRUB>(b0 + r1 + tx.i1);
  const float v100 = load_tap<T, SCRUB>(b1 + r0 + tx.i0);
  const float v101 = load_tap<T, SCRUB>(b1 + r0 + tx.i1);
  const float v110 = load_tap<T, SCRUB>(b1 + r1 + tx.i0);
  const float v111 = load_tap<T, SCRUB>(b1 + r1 + tx.i1);
  const float p0 = lerp_w(lerp_w(v000, v001, tx.w), lerp_w(v010, v011, tx.w), ty.w);
  const float p1 = lerp_w(lerp_w(v100, v101, tx.w), lerp_w(v110, v111, tx.w), ty.w);
  return lerp_w(p0, p1, tz.w);
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
__device__ __forceinline__ float affine_sample_generic(const AffineParams& p, int z, int y, int x) {
  double rp[3];
  affine_row_part(p, z, y, rp);
  return affine_sample_row<T, ORDER, BOUNDARY, SCRUB>(p, rp, x);
}

#endif  // __CUDACC__

int affine_gather_launch(const AffineParams& p, int src_dtype, cudaStream_t stream);
// returns B2_ERR_UNSUPPORTED (without setting an error) when the matrix/shape is not eligible
int affine_zsep_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible);
int affine_brick_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible);

}  // namespace b2
