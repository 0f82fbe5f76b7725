// TMA-staged affine pull warp for Z-SEPARABLE matrices (m01 = m02 = m10 = m20 = 0, m00 > 0):
// in-plane rotation / scale / shear / translation in YX combined with scale + translation in Z.
// This covers `biahub register` matrices built by the reference's helpers
// (biahub/register.py:32-145: YX rotation, YX/Z scaling, flips) and every `biahub stabilize`
// translation matrix (biahub/estimate_stabilization.py:981-990, 1186-1198).
//
// One CTA owns an output tile (TY x TX in YX) and MARCHES along output z.  For that tile the
// in-plane source footprint is the same for every z: the back-projected bounding box of the
// tile, a (BY x BX) brick per source plane.  A producer warp streams the needed source planes
// through a ring of shared-memory stages with 3-D TMA box loads (out-of-bounds zero fill);
// consumer threads hold their in-plane tap offsets/weights in registers (computed once, in
// float64, in the oracle's op order), reduce each arriving plane to one in-plane-interpolated
// value per output point, and blend consecutive planes along z.  Every source plane brick is
// read from global memory exactly once per CTA and every output voxel is written once,
// coalesced along x.
#include "b2_affine.cuh"

namespace b2 {

constexpr int kZsTY = 16;
constexpr int kZsTX = 64;
constexpr int kZsConsumers = 256;
constexpr int kZsThreads = kZsConsumers + 32;  // + one producer warp
constexpr int kZsStages = 4;
constexpr int kZsPPT = (kZsTY * kZsTX) / kZsConsumers;  // output points per consumer thread
constexpr int kZsRowsPerPass = kZsConsumers / kZsTX;

struct ZsepGeom {
  int BY, BX;          // brick extent per plane (elements)
  int stage_bytes;     // BY*BX*sizeof(T) rounded up to 128
  int zchunk;          // output planes per CTA
  int debug;           // fault bisection switches (env B2_ZSEP_DEBUG), 0 in production
};

template <typename T>
__device__ __forceinline__ float lds_elem(uint32_t base, int off);
template <>
__device__ __forceinline__ float lds_elem<float>(uint32_t base, int off) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + static_cast<uint32_t>(off) * 4u));
  return v;
}
template <>
__device__ __forceinline__ float lds_elem<uint16_t>(uint32_t base, int off) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + static_cast<uint32_t>(off) * 2u));
  return static_cast<float>(v);
}

__device__ __forceinline__ double coord_yx(double yf, double xf, const double* m) {
  // ((t_d + z*m_d0) + y*m_d1) + x*m_d2 with m_d0 == 0 (t_d + z*0 == t_d exactly)
  return __dadd_rn(__dadd_rn(m[3], __dmul_rn(yf, m[1])), __dmul_rn(xf, m[2]));
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
__global__ void __launch_bounds__(kZsThreads)
    affine_zsep_kernel(const __grid_constant__ CUtensorMap src_map, const AffineParams p,
                       const ZsepGeom g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kZsStages];
  __shared__ uint64_t empty_bar[kZsStages];

  const uint32_t stage0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x;
  const int y0 = blockIdx.y * kZsTY;
  const int x0 = blockIdx.x * kZsTX;
  const int zb = blockIdx.z * g.zchunk;
  const int ze = min(zb + g.zchunk, p.oz);

  // ---- brick origin: back-project the tile's four corners (uniform across the CTA)
  const int yl = min(y0 + kZsTY - 1, p.oy - 1);
  const int xl = min(x0 + kZsTX - 1, p.ox - 1);
  double cy_min = 1e300, cy_max = -1e300, cx_min = 1e300, cx_max = -1e300;
#pragma unroll
  for (int corner = 0; corner < 4; ++corner) {
    const double yf = static_cast<double>(((corner & 1) ? yl : y0) + p.cy);
    const double xf = static_cast<double>(((corner & 2) ? xl : x0) + p.cx);
    const double cy = coord_yx(yf, xf, p.m + 4);
    const double cx = coord_yx(yf, xf, p.m + 8);
    cy_min = fmin(cy_min, cy);
    cy_max = fmax(cy_max, cy);
    cx_min = fmin(cx_min, cx);
    cx_max = fmax(cx_max, cx);
  }
  // clamp far-away tiles so the int conversion is defined; such tiles are entirely outside
  const double big = 1.0e9;
  // The innermost TMA coordinate must be 16-byte aligned (measured on B200 / driver 580: an
  // unaligned inner coordinate raises "illegal instruction"), so the brick starts at the
  // 16-byte boundary at or below the back-projected minimum.
  constexpr int kVec = 16 / static_cast<int>(sizeof(T));
  const int by0 = __double2int_rd(fmax(-big, fmin(big, cy_min)));
  const int bx0 = __double2int_rd(fmax(-big, fmin(big, cx_min))) & ~(kVec - 1);
  const int by_hi = __double2int_rd(fmax(-big, fmin(big, cy_max))) + 1;
  const int bx_hi = __double2int_rd(fmax(-big, fmin(big, cx_max))) + 1;
  const bool brick_ok = (by_hi - by0) < g.BY && (bx_hi - bx0) < g.BX;  // CTA-uniform

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kZsStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kZsConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const double m00 = p.m[0], t0 = p.m[3];

  if (tid >= kZsConsumers) {
    // ================= producer warp: all lanes walk the plane sequence, lane 0 issues =================
    if (brick_ok) {
      const bool issuer = (tid == kZsConsumers);
      int s_last = INT_MIN;
      uint32_t seq = 0;
      for (int z = zb; z < ze; ++z) {
        const double cz = __dadd_rn(t0, __dmul_rn(static_cast<double>(z + p.cz), m00));
        const AxisTap tz = resolve_axis<ORDER, BOUNDARY>(cz, p.sz);
        if (!tz.inside) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = h ? tz.i1 : tz.i0;
          if (s > s_last) {
            const uint32_t stage = seq % kZsStages;
            if (!(g.debug & 1)) {
              if (seq >= kZsStages && !(g.debug & 2))
                mbar_wait(&empty_bar[stage], ((seq / kZsStages) - 1) & 1);
              if (issuer) {
                if (g.debug & 4) {
                  mbar_arrive(&full_bar[stage]);
                } else {
                  mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(g.BY) * g.BX * sizeof(T));
                  tma_load_3d(stage0 + stage * g.stage_bytes, &src_map, &full_bar[stage], bx0, by0, s);
                }
              }
              __syncwarp();
            }
            s_last = s;
            ++seq;
          }
        }
      }
    }
    return;
  }

  // ================================= consumer threads =================================
  const T* __restrict__ src = static_cast<const T*>(p.src);
  const int lx = tid % kZsTX;
  const int ly = tid / kZsTX;
  const int x = x0 + lx;

  int off00[kZsPPT], off01[kZsPPT], off10[kZsPPT], off11[kZsPPT];
  float wy[kZsPPT], wx[kZsPPT];
  bool live[kZsPPT];   // point is a real output voxel
  bool inyx[kZsPPT];   // ... and its in-plane coordinate is inside the source
#pragma unroll
  for (int i = 0; i < kZsPPT; ++i) {
    const int y = y0 + ly + kZsRowsPerPass * i;
    live[i] = (y < p.oy) && (x < p.ox);
    const double yf = static_cast<double>(y + p.cy);
    const double xf = static_cast<double>(x + p.cx);
    const AxisTap ty = resolve_axis<ORDER, BOUNDARY>(coord_yx(yf, xf, p.m + 4), p.sy);
    const AxisTap tx = resolve_axis<ORDER, BOUNDARY>(coord_yx(yf, xf, p.m + 8), p.sx);
    inyx[i] = live[i] && ty.inside && tx.inside;
    wy[i] = ty.w;
    wx[i] = tx.w;
    if (brick_ok) {
      off00[i] = (ty.i0 - by0) * g.BX + (tx.i0 - bx0);
      off01[i] = (ty.i0 - by0) * g.BX + (tx.i1 - bx0);
      off10[i] = (ty.i1 - by0) * g.BX + (tx.i0 - bx0);
      off11[i] = (ty.i1 - by0) * g.BX + (tx.i1 - bx0);
    } else {  // element offsets into a global source plane instead
      off00[i] = ty.i0 * p.sx + tx.i0;
      off01[i] = ty.i0 * p.sx + tx.i1;
      off10[i] = ty.i1 * p.sx + tx.i0;
      off11[i] = ty.i1 * p.sx + tx.i1;
    }
    if (!inyx[i]) off00[i] = off01[i] = off10[i] = off11[i] = 0;
  }

  float p_prev[kZsPPT], p_last[kZsPPT];
#pragma unroll
  for (int i = 0; i < kZsPPT; ++i) p_prev[i] = p_last[i] = 0.0f;
  int s_last = INT_MIN;
  uint32_t seq = 0;
  const int64_t sxy = static_cast<int64_t>(p.sy) * p.sx;

  for (int z = zb; z < ze; ++z) {
    const double cz = __dadd_rn(t0, __dmul_rn(static_cast<double>(z + p.cz), m00));
    const AxisTap tz = resolve_axis<ORDER, BOUNDARY>(cz, p.sz);
    float* __restrict__ out_plane = p.dst + static_cast<int64_t>(z) * p.oy * p.ox;
    if (!tz.inside) {
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i) {
        const int y = y0 + ly + kZsRowsPerPass * i;
        if (live[i]) st_global_cs(out_plane + static_cast<int64_t>(y) * p.ox + x, 0.0f);
      }
      continue;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = h ? tz.i1 : tz.i0;
      if (s > s_last) {  // CTA-uniform: the next plane of the producer's sequence
        uint32_t base = 0;
        const T* gplane = nullptr;
        const uint32_t stage = seq % kZsStages;
        if (brick_ok) {
          if (!(g.debug & 1)) mbar_wait(&full_bar[stage], (seq / kZsStages) & 1);
          base = stage0 + stage * g.stage_bytes;
        } else {
          gplane = src + s * sxy;
        }
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i) {
          p_prev[i] = p_last[i];
          float v = 0.0f;
          if (inyx[i]) {
            float v00, v01, v10, v11;
            if (brick_ok) {
              v00 = lds_elem<T>(base, off00[i]);
              if (ORDER == 1) {
                v01 = lds_elem<T>(base, off01[i]);
                v10 = lds_elem<T>(base, off10[i]);
                v11 = lds_elem<T>(base, off11[i]);
              }
            } else {
              v00 = to_f32<T>(__ldg(gplane + off00[i]));
              if (ORDER == 1) {
                v01 = to_f32<T>(__ldg(gplane + off01[i]));
                v10 = to_f32<T>(__ldg(gplane + off10[i]));
                v11 = to_f32<T>(__ldg(gplane + off11[i]));
              }
            }
            if (ORDER == 0) {
              v = (SCRUB && sizeof(T) == 4) ? scrub_value(v00) : v00;
            } else {
              v = lerp_w(lerp_w(v00, v01, wx[i]), lerp_w(v10, v11, wx[i]), wy[i]);
              if (SCRUB && sizeof(T) == 4 && !(fabsf(v) <= FLT_MAX)) {
                // a NaN/inf tap was involved (possibly with zero weight): redo with the scrub
                v = lerp_w(lerp_w(scrub_value(v00), scrub_value(v01), wx[i]),
                           lerp_w(scrub_value(v10), scrub_value(v11), wx[i]), wy[i]);
              }
            }
          }
          p_last[i] = v;
        }
        if (brick_ok && !(g.debug & 3)) {
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
        }
        s_last = s;
        ++seq;
      }
    }
    const bool first_is_last = (tz.i0 == s_last);
#pragma unroll
    for (int i = 0; i < kZsPPT; ++i) {
      const int y = y0 + ly + kZsRowsPerPass * i;
      if (live[i]) {
        const float v0 = first_is_last ? p_last[i] : p_prev[i];
        const float v = (ORDER == 0) ? v0 : lerp_w(v0, p_last[i], tz.w);
        st_global_cs(out_plane + static_cast<int64_t>(y) * p.ox + x, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool zsep_geometry(const AffineParams& p, ZsepGeom* g, size_t* smem_bytes) {
  const double* m = p.m;
  if (m[1] != 0.0 || m[2] != 0.0 || m[4] != 0.0 || m[8] != 0.0) return false;
  if (!(m[0] > 0.0)) return false;
  if (reinterpret_cast<uintptr_t>(p.src) % 16 != 0) return false;
  if ((static_cast<int64_t>(p.sx) * sizeof(T)) % 16 != 0) return false;
  const double ey = fabs(m[5]) * (kZsTY - 1) + fabs(m[6]) * (kZsTX - 1);
  const double ex = fabs(m[9]) * (kZsTY - 1) + fabs(m[10]) * (kZsTX - 1);
  if (!(ey < 250.0) || !(ex < 250.0)) return false;
  const int vec = 16 / sizeof(T);
  int BY = static_cast<int>(ey) + 4;
  int BX = static_cast<int>(ex) + 4 + (vec - 1);  // + alignment slack of the brick origin
  BX = (BX + vec - 1) / vec * vec;
  if (BY > 256 || BX > 256) return false;
  const int stage = (BY * BX * static_cast<int>(sizeof(T)) + 127) / 128 * 128;
  if (stage * kZsStages > 96 * 1024) return false;
  if (static_cast<int64_t>(p.sy) * p.sx >= (1LL << 31)) return false;  // int32 fallback offsets
  g->BY = BY;
  g->BX = BX;
  g->stage_bytes = stage;
  *smem_bytes = static_cast<size_t>(stage) * kZsStages + 128;
  return true;
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
static int launch_zsep(const AffineParams& p, ZsepGeom g, size_t smem_bytes, cudaStream_t stream) {
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return B2_ERR_NO_DEVICE;
  }
  CUtensorMap map;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p.sx), static_cast<cuuint64_t>(p.sy),
                              static_cast<cuuint64_t>(p.sz)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p.sx) * sizeof(T),
                                 static_cast<cuuint64_t>(p.sx) * p.sy * sizeof(T)};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(g.BX), static_cast<cuuint32_t>(g.BY), 1u};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUtensorMapDataType dt =
      sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = encode(&map, dt, 3, const_cast<void*>(p.src), gdim, gstride, box, estride,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for affine source (%d,%d,%d)", (int)r,
              p.sz, p.sy, p.sx);
    return B2_ERR_UNSUPPORTED;
  }
  int sms = 148;
  sm_count(&sms);
  const int tiles_x = (p.ox + kZsTX - 1) / kZsTX;
  const int tiles_y = (p.oy + kZsTY - 1) / kZsTY;
  const int64_t tiles = static_cast<int64_t>(tiles_x) * tiles_y;
  const int64_t target = static_cast<int64_t>(sms) * 16;  // enough CTAs for a few waves
  int zsplit = 1;
  if (tiles < target) zsplit = static_cast<int>((target + tiles - 1) / tiles);
  int zchunk = (p.oz + zsplit - 1) / zsplit;
  if (zchunk < 8) zchunk = 8;
  if (zchunk > p.oz) zchunk = p.oz;
  g.zchunk = zchunk;
  {
    const char* dbg = getenv("B2_ZSEP_DEBUG");
    g.debug = dbg ? atoi(dbg) : 0;
  }
  const int grid_z = (p.oz + zchunk - 1) / zchunk;
  if (tiles_y > 65535 || grid_z > 65535) return affine_gather_launch(p, sizeof(T) == 2 ? 0 : 1, stream);

  auto kern = affine_zsep_kernel<T, ORDER, BOUNDARY, SCRUB>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_bytes)));
  const dim3 grid(tiles_x, tiles_y, grid_z);
  kern<<<grid, kZsThreads, smem_bytes, stream>>>(map, p, g);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

template <typename T>
static int zsep_typed(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  ZsepGeom g{};
  size_t smem = 0;
  *eligible = zsep_geometry<T>(p, &g, &smem);
  if (!*eligible) return B2_ERR_UNSUPPORTED;
  const bool scrub = p.scrub && sizeof(T) == 4;
#define B2_ZS(ORD, BND)                                                          \
  (scrub ? launch_zsep<T, ORD, BND, true>(p, g, smem, stream)                    \
         : launch_zsep<T, ORD, BND, false>(p, g, smem, stream))
  if (p.order == 0)
    return p.boundary == B2_BOUNDARY_CONSTANT ? B2_ZS(0, B2_BOUNDARY_CONSTANT)
                                              : B2_ZS(0, B2_BOUNDARY_ITK);
  return p.boundary == B2_BOUNDARY_CONSTANT ? B2_ZS(1, B2_BOUNDARY_CONSTANT)
                                            : B2_ZS(1, B2_BOUNDARY_ITK);
#undef B2_ZS
}

int affine_zsep_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible) {
  if (src_dtype == B2_DTYPE_U16) return zsep_typed<uint16_t>(p, stream, eligible);
  return zsep_typed<float>(p, stream, eligible);
}

}  // namespace b2
