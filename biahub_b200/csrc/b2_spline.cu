// Cubic B-spline affine pull warp = the reference's `method="scipy"` branch
// (biahub/register.py:271-272: scipy.ndimage.affine_transform(zyx, M, shape) with scipy's
// defaults order=3, mode="constant", cval=0, prefilter=True; the third positional argument
// lands in scipy's `offset` slot and is ignored for a homogeneous matrix, so the output grid is
// the INPUT's shape and the output dtype the input's).
//
// Two stages, both float64 like scipy:
//  1. prefilter (scipy.ndimage.spline_filter, order 3): per axis the recursive filter
//       c+[i] = 6 s[i] + z c+[i-1],  c[i] = z (c[i+1] - c+[i]),  z = sqrt(3) - 2,
//     with the mirror initialisation scipy uses for mode="constant".  On the mirror-extended
//     signal this cascade equals the symmetric exponential filter
//       c[i] = g * sum_k z^|k| s[i+k] = g * (C+[i] + C-[i] - s[i]),   g = -6z / (1 - z^2),
//     with C+[i] = s[i] + z C+[i-1] and C-[i] = s[i] + z C-[i+1] both running on the INPUT (the
//     parallel form of the same transfer function).  z^20 = 4e-12, so a recursion started from
//     zero 20 samples before the chunk is exact far below float32 resolution: every thread owns a
//     chunk of 32 samples of one line and needs no carry from its neighbours — any axis, no scan.
//  2. evaluation: c = M o (float64, scipy's op order), outside [0, n-1] -> 0, else the 4x4x4
//     taps floor(c)-1 .. floor(c)+2, mirror-extended, weighted with scipy's cubic weights.
//     Integer outputs use scipy's conversion (t > 0 ? t + 0.5 : 0, clamped, truncated).
//
// Not a roofline kernel: the reference calls this branch "10x slower than ANTs"
// (biahub/register.py:256-257); here it is gather-bound on L1/L2 at about 25 ms per 0.5 Gvoxel.
#include "b2_affine.cuh"

namespace b2 {

namespace {

constexpr double kPole = -0.26794919243112270647;  // sqrt(3) - 2
constexpr double kGain = 1.7320508075688772935;    // -6 z / (1 - z^2) = sqrt(3)
// z^20 = 3.6e-12: the zero-started recursions are exact to 4e-12 of the signal after 20 warm-up
// samples (float32 output: 6e-8); 32-sample chunks read each sample 3.25 times (2 x (20 + 32) / 32)
constexpr int kWarm = 20;
constexpr int kChunk = 32;

__device__ __forceinline__ int mirror_index(int j, int n) {
  if (j >= 0 && j < n) return j;
  const int p = 2 * n - 2;  // n >= 2
  int m = j % p;
  if (m < 0) m += p;
  return m < n ? m : p - m;
}

template <typename T, bool SCRUB>
__device__ __forceinline__ double spline_load(const T* p);
template <>
__device__ __forceinline__ double spline_load<uint16_t, false>(const uint16_t* p) {
  return static_cast<double>(__ldg(p));
}
template <>
__device__ __forceinline__ double spline_load<uint16_t, true>(const uint16_t* p) {
  return static_cast<double>(__ldg(p));
}
template <>
__device__ __forceinline__ double spline_load<float, false>(const float* p) {
  return static_cast<double>(__ldg(p));
}
template <>
__device__ __forceinline__ double spline_load<float, true>(const float* p) {
  return static_cast<double>(scrub_value(__ldg(p)));
}
template <>
__device__ __forceinline__ double spline_load<double, false>(const double* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ double spline_load<double, true>(const double* p) {
  return __ldg(p);
}

// One axis of the prefilter.  Lines are indexed by (a, b): base = a*strideA + b*strideB, sample i
// of a line sits at base + i*stride.  CHUNK_FAST: consecutive threads own consecutive chunks of
// one line (contiguous axis), otherwise consecutive lines (coalesced along b).
template <typename T, bool SCRUB, bool CHUNK_FAST>
__global__ void __launch_bounds__(256)
    spline3_filter_kernel(const T* __restrict__ in, double* __restrict__ out, int n, int64_t stride,
                          int64_t lines, int64_t B, int64_t strideA, int64_t strideB, int chunks) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= lines * chunks) return;
  const int64_t line = CHUNK_FAST ? tid / chunks : tid % lines;
  const int c = static_cast<int>(CHUNK_FAST ? tid % chunks : tid / lines);
  const int64_t base = (line / B) * strideA + (line % B) * strideB;
  const int i0 = c * kChunk;
  const T* __restrict__ src = in + base;

  double cm[kChunk];
  double s = 0.0;
  for (int j = i0 + kChunk - 1 + kWarm; j >= i0 + kChunk; --j)
    s = fma(kPole, s, spline_load<T, SCRUB>(src + mirror_index(j, n) * stride));
#pragma unroll
  for (int k = kChunk - 1; k >= 0; --k) {
    s = fma(kPole, s, spline_load<T, SCRUB>(src + mirror_index(i0 + k, n) * stride));
    cm[k] = s;
  }
  s = 0.0;
  for (int j = i0 - kWarm; j < i0; ++j)
    s = fma(kPole, s, spline_load<T, SCRUB>(src + mirror_index(j, n) * stride));
  double* __restrict__ dst = out + base;
#pragma unroll
  for (int k = 0; k < kChunk; ++k) {
    const int j = i0 + k;
    const double v = spline_load<T, SCRUB>(src + mirror_index(j, n) * stride);
    s = fma(kPole, s, v);
    if (j < n) dst[j * stride] = kGain * ((s + cm[k]) - v);
  }
}

template <typename T, bool SCRUB>
__global__ void spline3_convert_kernel(const T* __restrict__ in, double* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = spline_load<T, SCRUB>(in + i);
}

__device__ __forceinline__ void cubic_weights(double t, double (&w)[4]) {
  const double u = 1.0 - t;
  w[1] = (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0;
  w[2] = (u * u * (u - 2.0) * 3.0 + 4.0) / 6.0;
  w[0] = u * u * u / 6.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

template <typename TOut>
__device__ __forceinline__ TOut spline_store_value(double v);
template <>
__device__ __forceinline__ float spline_store_value<float>(double v) {
  return static_cast<float>(v);
}
template <>
__device__ __forceinline__ uint16_t spline_store_value<uint16_t>(double v) {
  double t = v > 0.0 ? v + 0.5 : 0.0;
  t = t > 65535.0 ? 65535.0 : t;
  return static_cast<uint16_t>(static_cast<int>(t));
}

template <typename TOut>
__global__ void __launch_bounds__(128)
    spline3_eval_kernel(const double* __restrict__ coef, const __grid_constant__ AffineParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, z = blockIdx.z;
  if (x >= p.ox) return;
  const double zf = static_cast<double>(z + p.cz), yf = static_cast<double>(y + p.cy),
               xf = static_cast<double>(x + p.cx);
  const int n[3] = {p.sz, p.sy, p.sx};
  double w[3][4];
  int idx[3][4];
  bool inside = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double* m = p.m + 4 * d;
    const double c = __dadd_rn(__dadd_rn(__dadd_rn(m[3], __dmul_rn(zf, m[0])), __dmul_rn(yf, m[1])),
                               __dmul_rn(xf, m[2]));
    inside = inside && (c >= 0.0) && (c <= static_cast<double>(n[d] - 1));
    const double cc = inside ? c : 0.0;
    const double f = floor(cc);
    cubic_weights(cc - f, w[d]);
    const int b = static_cast<int>(f) - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[d][k] = n[d] == 1 ? 0 : mirror_index(b + k, n[d]);
  }
  TOut* __restrict__ dst = reinterpret_cast<TOut*>(p.dst);
  const int64_t o = (static_cast<int64_t>(z) * p.oy + y) * p.dpitch + x;
  if (!inside) {
    dst[o] = spline_store_value<TOut>(0.0);
    return;
  }
  const int64_t plane = static_cast<int64_t>(p.sy) * p.sx;
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double* __restrict__ pz = coef + idx[0][a] * plane;
    double accy = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const double* __restrict__ row = pz + static_cast<int64_t>(idx[1][b]) * p.sx;
      double accx = __ldg(row + idx[2][0]) * w[2][0];
      accx = fma(__ldg(row + idx[2][1]), w[2][1], accx);
      accx = fma(__ldg(row + idx[2][2]), w[2][2], accx);
      accx = fma(__ldg(row + idx[2][3]), w[2][3], accx);
      accy = fma(accx, w[1][b], accy);
    }
    acc = fma(accy, w[0][a], acc);
  }
  dst[o] = spline_store_value<TOut>(acc);
}

template <typename T, bool SCRUB>
int launch_filter_axis(const T* in, double* out, const int64_t (&n)[3], int axis, cudaStream_t st) {
  const int64_t Z = n[0], Y = n[1], X = n[2];
  const int len = static_cast<int>(n[axis]);
  const int chunks = (len + kChunk - 1) / kChunk;
  int64_t lines, B, strideA, strideB, stride;
  if (axis == 0) {
    lines = Y * X; B = lines; strideA = 0; strideB = 1; stride = Y * X;
  } else if (axis == 1) {
    lines = Z * X; B = X; strideA = Y * X; strideB = 1; stride = X;
  } else {
    lines = Z * Y; B = lines; strideA = 0; strideB = X; stride = 1;
  }
  const int64_t threads = lines * chunks;
  const int64_t blocks = (threads + 255) / 256;
  if (blocks > 2147483647LL) {
    set_error("spline prefilter: volume too large");
    return B2_ERR_UNSUPPORTED;
  }
  if (axis == 2)
    spline3_filter_kernel<T, SCRUB, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        in, out, len, stride, lines, B, strideA, strideB, chunks);
  else
    spline3_filter_kernel<T, SCRUB, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        in, out, len, stride, lines, B, strideA, strideB, chunks);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

template <typename T, bool SCRUB>
int prefilter_typed(const T* src, const int64_t (&n)[3], double* buf_a, double* buf_b, double** coef,
                    cudaStream_t st) {
  const int64_t vol = n[0] * n[1] * n[2];
  bool first = true;
  double* cur = nullptr;
  int rc;
  for (int axis = 2; axis >= 0; --axis) {
    if (n[axis] < 2) continue;  // scipy skips axes of length 1
    double* nxt = (cur == buf_a) ? buf_b : buf_a;
    if (first)
      rc = launch_filter_axis<T, SCRUB>(src, nxt, n, axis, st);
    else
      rc = launch_filter_axis<double, false>(cur, nxt, n, axis, st);
    if (rc) return rc;
    cur = nxt;
    first = false;
  }
  if (first) {  // 1x1x1: the coefficients are the samples
    const int64_t blocks = (vol + 255) / 256;
    spline3_convert_kernel<T, SCRUB><<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, buf_a, vol);
    B2_CUDA(cudaGetLastError());
    count_launch();
    cur = buf_a;
  }
  *coef = cur;
  return B2_OK;
}

}  // namespace

size_t spline_workspace_bytes(int64_t sz, int64_t sy, int64_t sx) {
  if (sz < 1 || sy < 1 || sx < 1) return 0;
  const size_t vol = (static_cast<size_t>(sz) * sy * sx * sizeof(double) + 255) / 256 * 256;
  return 2 * vol;
}

// prefilter `src` into the workspace; *coef points at the finished float64 coefficients
int spline_prefilter_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                            int scrub, void* ws, size_t ws_bytes, const double** coef,
                            cudaStream_t st) {
  if (!src || !ws || !coef) {
    set_error("spline prefilter: null pointer");
    return B2_ERR_INVALID;
  }
  if (sz < 1 || sy < 1 || sx < 1 || sz > 2147483647LL / 4 || sy > 2147483647LL / 4 ||
      sx > 2147483647LL / 4) {
    set_error("spline prefilter: invalid shape");
    return B2_ERR_INVALID;
  }
  if (ws_bytes < spline_workspace_bytes(sz, sy, sx)) {
    set_error("spline prefilter: workspace too small (%zu < %zu)", ws_bytes,
              spline_workspace_bytes(sz, sy, sx));
    return B2_ERR_INVALID;
  }
  const int64_t n[3] = {sz, sy, sx};
  double* a = static_cast<double*>(ws);
  double* b = reinterpret_cast<double*>(static_cast<char*>(ws) + spline_workspace_bytes(sz, sy, sx) / 2);
  double* out = nullptr;
  int rc;
  if (src_dtype == B2_DTYPE_U16)
    rc = prefilter_typed<uint16_t, false>(static_cast<const uint16_t*>(src), n, a, b, &out, st);
  else if (src_dtype == B2_DTYPE_F32)
    rc = scrub ? prefilter_typed<float, true>(static_cast<const float*>(src), n, a, b, &out, st)
               : prefilter_typed<float, false>(static_cast<const float*>(src), n, a, b, &out, st);
  else {
    set_error("spline prefilter: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  *coef = out;
  return rc;
}

// evaluate output planes of the (cropped) box; dst dtype uint16 (scipy's integer conversion) or float32
int spline_eval_device(const double* coef, int64_t sz, int64_t sy, int64_t sx, void* dst,
                       int dst_dtype, int64_t oz, int64_t oy, int64_t ox, const double* M12,
                       const int64_t* crop_start, cudaStream_t st) {
  if (!coef || !dst || !M12) {
    set_error("spline eval: null pointer");
    return B2_ERR_INVALID;
  }
  if (dst_dtype != B2_DTYPE_U16 && dst_dtype != B2_DTYPE_F32) {
    set_error("spline eval: dst_dtype must be uint16 or float32");
    return B2_ERR_INVALID;
  }
  if (oz < 0 || oy < 0 || ox < 0 || oy > 65535 || ox > 2147483647LL / 2) {
    set_error("spline eval: invalid output shape");
    return B2_ERR_INVALID;
  }
  if (oz == 0 || oy == 0 || ox == 0) return B2_OK;
  AffineParams p{};
  p.src = coef;
  p.dst = static_cast<float*>(dst);
  p.sz = static_cast<int>(sz); p.sy = static_cast<int>(sy); p.sx = static_cast<int>(sx);
  p.oy = static_cast<int>(oy); p.ox = static_cast<int>(ox);
  p.cy = static_cast<int>(crop_start ? crop_start[1] : 0);
  p.cx = static_cast<int>(crop_start ? crop_start[2] : 0);
  for (int i = 0; i < 12; ++i) p.m[i] = M12[i];
  p.order = 3;
  p.dpitch = static_cast<int>(ox);
  const size_t esz = dst_dtype == B2_DTYPE_U16 ? 2 : 4;
  for (int64_t z0 = 0; z0 < oz; z0 += 65535) {  // gridDim.z limit
    const int64_t cnt = oz - z0 < 65535 ? oz - z0 : 65535;
    p.oz = static_cast<int>(cnt);
    p.cz = static_cast<int>((crop_start ? crop_start[0] : 0) + z0);
    p.dst = reinterpret_cast<float*>(static_cast<char*>(dst) + static_cast<size_t>(z0) * oy * ox * esz);
    const dim3 grid(static_cast<unsigned>((ox + 127) / 128), static_cast<unsigned>(oy),
                    static_cast<unsigned>(cnt));
    if (dst_dtype == B2_DTYPE_U16)
      spline3_eval_kernel<uint16_t><<<grid, 128, 0, st>>>(coef, p);
    else
      spline3_eval_kernel<float><<<grid, 128, 0, st>>>(coef, p);
    B2_CUDA(cudaGetLastError());
    count_launch();
  }
  return B2_OK;
}

int affine_spline_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                         void* dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                         const double* M12, const int64_t* crop_start, int scrub, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  const double* coef = nullptr;
  int rc = spline_prefilter_device(src, src_dtype, sz, sy, sx, scrub, ws, ws_bytes, &coef, st);
  if (rc) return rc;
  return spline_eval_device(coef, sz, sy, sx, dst, dst_dtype, oz, oy, ox, M12, crop_start, st);
}

}  // namespace b2
