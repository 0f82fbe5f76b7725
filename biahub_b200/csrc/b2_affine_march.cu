// Z-marching affine pull warp for GENERIC 3x4 matrices, order 1 (trilinear): the production case of
// `biahub register` (estimated transforms have all twelve terms non-zero).  Same arithmetic as
// `affine_brick_kernel` (b2_affine_brick.cu), different data movement:
//
//   * one CTA owns an output tile TY x TX in YX and MARCHES along output z in chunks of kMaCZ
//     planes (m00 > 0: the source plane demand is monotone in z);
//   * the source planes a chunk needs are the z extent of its back-projected hull.  They live in a
//     RING of R plane slots in shared memory; every plane brick (BY x BX, the in-plane hull of the
//     CTA's whole z range, origin rounded down to 16 bytes) is loaded exactly ONCE per CTA by a
//     3-D TMA box of depth 1 (out-of-bounds planes / rows / columns are zero-filled = cval 0);
//   * loads run two chunks ahead: while chunk c is interpolated, the planes of chunk c+1 have
//     landed or are landing and those of chunk c+2 are issued as soon as chunk c is finished
//     (one __syncthreads per chunk; two mbarriers, one per chunk parity).
//
// Compared with one 3-D brick per tile this removes the z halo (12 planes loaded per 8 written ->
// 1 per 1) and the exposed TMA latency at the start of every tile (27 % of the warp stall samples
// of the brick kernel).  Ineligible shapes (m00 <= 0, ring or hull too large, unaligned rows)
// fall through to the brick / gather kernels.
#include "b2_affine.cuh"

namespace b2 {

constexpr int kMaCZ = 4;  // output planes per chunk
constexpr int kMaTY = 16;
constexpr int kMaTX = 32;
constexpr int kMaThreads = 256;
constexpr int kMaCols = (kMaTY * kMaTX) / kMaThreads;  // (y, x) columns per thread
constexpr int kMaRowStep = kMaThreads / kMaTX;
constexpr float kMaEdge = 2.0e-3f;
constexpr float kMaMagic = 12582912.0f;  // 1.5 * 2^23
constexpr uint32_t kMaMagicBits = 0x4B400000u;

struct MarchGeom {
  int R;           // ring depth in planes
  int BY, BX;      // plane brick extent (elements)
  int slot_bytes;  // BY*BX*sizeof(T) rounded up to 128
  int zchunks;     // chunks per CTA
  // hull of ONE chunk (kMaCZ x TY x TX voxels) relative to its origin voxel, per source axis:
  // sums of the negative / positive parts of m_dj * (extent_j - 1)  (host, float64)
  double negc[3], posc[3];
  // the same for the CTA's whole z range (zchunks * kMaCZ planes): in-plane brick origin / extent
  double negt[3], post[3];
  float mcol[9];  // float32 copy of the 3x3 linear part (row-major)
};

template <typename T>
__device__ __forceinline__ float ring_elem(uint32_t addr);
template <>
__device__ __forceinline__ float ring_elem<float>(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <>
__device__ __forceinline__ float ring_elem<uint16_t>(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return static_cast<float>(v);
}

// four taps of one source plane (rows a0 / a1, columns +0 / +1), loaded only when `need`
template <typename T>
__device__ __forceinline__ void ring_quad_if(bool need, uint32_t a0, uint32_t a1, float& v00,
                                             float& v01, float& v10, float& v11);
template <>
__device__ __forceinline__ void ring_quad_if<float>(bool need, uint32_t a0, uint32_t a1, float& v00,
                                                    float& v01, float& v10, float& v11) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %4, 0;\n"
      "@p ld.shared.f32 %0, [%5];\n"
      "@p ld.shared.f32 %1, [%5+4];\n"
      "@p ld.shared.f32 %2, [%6];\n"
      "@p ld.shared.f32 %3, [%6+4];\n"
      "}\n"
      : "+f"(v00), "+f"(v01), "+f"(v10), "+f"(v11)
      : "r"(static_cast<uint32_t>(need)), "r"(a0), "r"(a1));
}
template <>
__device__ __forceinline__ void ring_quad_if<uint16_t>(bool need, uint32_t a0, uint32_t a1,
                                                       float& v00, float& v01, float& v10,
                                                       float& v11) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b16 h0, h1, h2, h3;\n"
      "setp.ne.u32 p, %4, 0;\n"
      "@p ld.shared.u16 h0, [%5];\n"
      "@p ld.shared.u16 h1, [%5+2];\n"
      "@p ld.shared.u16 h2, [%6];\n"
      "@p ld.shared.u16 h3, [%6+2];\n"
      "@p cvt.rn.f32.u16 %0, h0;\n"
      "@p cvt.rn.f32.u16 %1, h1;\n"
      "@p cvt.rn.f32.u16 %2, h2;\n"
      "@p cvt.rn.f32.u16 %3, h3;\n"
      "}\n"
      : "+f"(v00), "+f"(v01), "+f"(v10), "+f"(v11)
      : "r"(static_cast<uint32_t>(need)), "r"(a0), "r"(a1));
}

__device__ __forceinline__ double march_coord(const double* m, double zf, double yf, double xf) {
  return __dadd_rn(__dadd_rn(__dadd_rn(m[3], __dmul_rn(zf, m[0])), __dmul_rn(yf, m[1])),
                   __dmul_rn(xf, m[2]));
}

struct MarchRing {
  uint32_t ring;        // shared-window address of slot 0
  uint32_t slot_b;      // bytes per plane slot
  uint32_t row_b;       // bytes per brick row
  int R;                // slots
  int b0z, b0y, b0x;    // source index of ring plane 0 / brick row 0 / brick column 0
  int BX;
};

// exact (float64) evaluation of one voxel with taps read from the ring
template <typename T, int BOUNDARY, bool SCRUB>
__device__ __noinline__ float march_sample_exact(const AffineParams& p, const MarchRing& rg, int z,
                                                 int y, int x) {
  const double zf = static_cast<double>(z + p.cz), yf = static_cast<double>(y + p.cy),
               xf = static_cast<double>(x + p.cx);
  const AxisTap tz = resolve_axis<1, BOUNDARY>(march_coord(p.m, zf, yf, xf), p.sz);
  const AxisTap ty = resolve_axis<1, BOUNDARY>(march_coord(p.m + 4, zf, yf, xf), p.sy);
  const AxisTap tx = resolve_axis<1, BOUNDARY>(march_coord(p.m + 8, zf, yf, xf), p.sx);
  if (!(tz.inside && ty.inside && tx.inside)) return 0.0f;
  auto tap = [&](int iz, int iy, int ix) {
    const uint32_t slot = static_cast<uint32_t>(iz - rg.b0z) % static_cast<uint32_t>(rg.R);
    const uint32_t off = static_cast<uint32_t>((iy - rg.b0y) * rg.BX + (ix - rg.b0x));
    float v = ring_elem<T>(rg.ring + slot * rg.slot_b + off * static_cast<uint32_t>(sizeof(T)));
    if (SCRUB && sizeof(T) == 4) v = scrub_value(v);
    return v;
  };
  const float p0 = lerp_w(lerp_w(tap(tz.i0, ty.i0, tx.i0), tap(tz.i0, ty.i0, tx.i1), tx.w),
                          lerp_w(tap(tz.i0, ty.i1, tx.i0), tap(tz.i0, ty.i1, tx.i1), tx.w), ty.w);
  const float p1 = lerp_w(lerp_w(tap(tz.i1, ty.i0, tx.i0), tap(tz.i1, ty.i0, tx.i1), tx.w),
                          lerp_w(tap(tz.i1, ty.i1, tx.i0), tap(tz.i1, ty.i1, tx.i1), tx.w), ty.w);
  return lerp_w(p0, p1, tz.w);
}

struct MarchCol {
  uint32_t abase;       // ring - (magic + wrap base) * slot_b - magic * (row_b + es)
  uint32_t thr;         // magic + wrap base + R: float bits of t_z at which the slot wraps
  uint32_t ring_b;      // R * slot_b
  uint32_t slot_b, row_b;
  int64_t out_plane;
  float u0z, u0y, u0x;  // brick-local coordinate of the column's first voxel of the chunk
  float mz, my, mx;     // coordinate step per output plane
};

// One (y, x) column of one chunk: kMaCZ voxels along output z.  Strictly interior voxels are
// interpolated in fp32 from the ring and stored; returns the bit mask of the voxels NOT written
// (not strictly interior, or a non-finite tap): the caller finishes those on the exact path.
// CHECK = false: the chunk's whole hull is strictly inside the source (decided once per chunk).
template <typename T, bool SCRUB, bool CHECK>
__device__ __forceinline__ uint32_t march_column(const MarchCol& c, const float (&mid)[3],
                                                 const float (&half)[3], float* __restrict__ out) {
  constexpr uint32_t es = static_cast<uint32_t>(sizeof(T));
  uint32_t rest = 0;
  // upper-plane taps of the previous voxel: its successor usually sits in the same (y, x) cell
  // one plane further and takes them as its lower-plane taps without reading them again
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
  uint32_t a_up = 0xffffffffu;
  float* __restrict__ o = out;
#pragma unroll
  for (int k = 0; k < kMaCZ; ++k, o += c.out_plane) {
    const float kf = static_cast<float>(k);
    const float uz = __fmaf_rn(kf, c.mz, c.u0z);
    const float uy = __fmaf_rn(kf, c.my, c.u0y);
    const float ux = __fmaf_rn(kf, c.mx, c.u0x);
    if (CHECK) {
      const bool interior = fabsf(uz - mid[0]) <= half[0] - kMaEdge &&
                            fabsf(uy - mid[1]) <= half[1] - kMaEdge &&
                            fabsf(ux - mid[2]) <= half[2] - kMaEdge;
      if (!interior) {
        rest |= 1u << k;
        a_up = 0xffffffffu;
        continue;
      }
    }
    // floor via the magic constant in round-down mode (FADD.RM): the low mantissa bits of
    // t = RD(u + kMaMagic) hold floor(u); weight w = u - (t - kMaMagic)
    const float tz = __fadd_rd(uz, kMaMagic), ty = __fadd_rd(uy, kMaMagic), tx = __fadd_rd(ux, kMaMagic);
    const float wz = uz - (tz - kMaMagic);
    const float wy = uy - (ty - kMaMagic);
    const float wx = ux - (tx - kMaMagic);
    const uint32_t zb = static_cast<uint32_t>(__float_as_int(tz));
    const uint32_t a = zb * c.slot_b + (static_cast<uint32_t>(__float_as_int(ty)) * c.row_b +
                                        (static_cast<uint32_t>(__float_as_int(tx)) * es + c.abase));
    // ring wrap: plane (zb - magic) lives in slot (zb - magic - wrap base), minus R when >= R
    const uint32_t a00 = a - (zb >= c.thr ? c.ring_b : 0u);
    const uint32_t a10 = a + c.slot_b - (zb + 1u >= c.thr ? c.ring_b : 0u);
    ring_quad_if<T>(a00 != a_up, a00, a00 + c.row_b, r0, r1, r2, r3);
    const float v000 = r0, v001 = r1, v010 = r2, v011 = r3;
    r0 = ring_elem<T>(a10);
    r1 = ring_elem<T>(a10 + es);
    r2 = ring_elem<T>(a10 + c.row_b);
    r3 = ring_elem<T>(a10 + c.row_b + es);
    a_up = a10;
    float v;
    if (SCRUB || sizeof(T) == 2) {
      // finite taps (uint16, or verified below): v0 + w * (v1 - v0), 2 instructions per lerp
      const float q00 = __fmaf_rn(wx, v001 - v000, v000), q01 = __fmaf_rn(wx, v011 - v010, v010);
      const float q10 = __fmaf_rn(wx, r1 - r0, r0), q11 = __fmaf_rn(wx, r3 - r2, r2);
      const float q0 = __fmaf_rn(wy, q01 - q00, q00), q1 = __fmaf_rn(wy, q11 - q10, q10);
      v = __fmaf_rn(wz, q1 - q0, q0);
      // a NaN/inf tap makes v non-finite: the exact path applies the scrub per tap
      if (sizeof(T) == 4 && !(fabsf(v) <= FLT_MAX)) {
        rest |= 1u << k;
        continue;
      }
    } else {
      v = lerp_w(lerp_w(lerp_w(v000, v001, wx), lerp_w(v010, v011, wx), wy),
                 lerp_w(lerp_w(r0, r1, wx), lerp_w(r2, r3, wx), wy), wz);
    }
    st_global_cs(o, v);
  }
  return rest;
}

template <typename T, int BOUNDARY, bool SCRUB>
__global__ void __launch_bounds__(kMaThreads, 4)
    affine_march_kernel(const __grid_constant__ CUtensorMap src_map,
                        const __grid_constant__ AffineParams p,
                        const __grid_constant__ MarchGeom g, const int tiles_x) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar[2];    // "full": the planes of the chunk of this parity have landed
  __shared__ uint64_t done[2];   // "empty": every warp has finished the chunk of this parity
  // b0z, b0y, b0x, bits(c0l[3]), flags (1 = in-plane brick fits, 2 = in-plane hull strictly inside
  // the source, 4 = in-plane hull outside the source)
  __shared__ int s_geo[8];
  __shared__ int s_chunk[2][2];  // per chunk parity: {ring wrap base, class 0 mixed / 1 inside / 2 outside}
  constexpr int kVec = 16 / static_cast<int>(sizeof(T));
  constexpr uint32_t es = static_cast<uint32_t>(sizeof(T));
  const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;

  const int tx_i = blockIdx.x % tiles_x, ty_i = blockIdx.x / tiles_x;
  const int y0 = ty_i * kMaTY, x0 = tx_i * kMaTX;
  const int z0 = blockIdx.y * g.zchunks * kMaCZ;
  const int nz_cta = min(g.zchunks * kMaCZ, p.oz - z0);
  const int nchunks = (nz_cta + kMaCZ - 1) / kMaCZ;

  // ---- producer state (meaningful in thread 0 only)
  double cz0 = 0.0;      // float64 source z coordinate of the CTA's origin voxel
  int t_b0z = 0, t_b0y = 0, t_b0x = 0;
  int loaded_upto = -1;  // highest ring plane (relative to b0z) already requested
  int inplane = 0;       // 1 = in-plane hull strictly inside, 2 = outside

  // planes of chunk c: [lo, hi] relative to b0z; issues the TMA loads of the not-yet-requested
  // ones on the barrier of the chunk's parity and publishes the chunk's wrap base + class
  auto issue_chunk = [&](int c) {
    const double zc = __dadd_rn(cz0, __dmul_rn(p.m[0], static_cast<double>(kMaCZ * c)));
    const int lo = __double2int_rd(zc + (g.negc[0] - 1e-6)) - t_b0z;
    // +2: the upper tap, and the neighbour an ITK edge clamp reads with weight 0
    const int hi = __double2int_rd(zc + (g.posc[0] + 1e-6)) + 2 - t_b0z;
    const int first = max(loaded_upto + 1, lo);
    const int n = max(hi - first + 1, 0);
    uint64_t* b = &bar[c & 1];
    mbar_expect_tx(b, static_cast<uint32_t>(n) * static_cast<uint32_t>(g.BY * g.BX) * es);
    for (int q = first; q <= hi; ++q) {
      const uint32_t slot = static_cast<uint32_t>(q) % static_cast<uint32_t>(g.R);
      tma_load_3d(ring + slot * static_cast<uint32_t>(g.slot_bytes), &src_map, b, t_b0x, t_b0y,
                  t_b0z + q);
    }
    loaded_upto = max(loaded_upto, hi);
    int cls = 0;
    const int alo = lo + t_b0z, ahi = hi + t_b0z;  // absolute plane range (ahi includes the +2)
    if (inplane == 2 || ahi <= -1 || alo >= p.sz + 1) {
      cls = 2;  // the chunk's hull misses the source (and its half-voxel ITK band) entirely
    } else if (inplane == 1 && alo >= 1 && ahi <= p.sz - 1) {
      cls = 1;  // every tap of every voxel is a valid source index, >= 1 voxel from the border
    }
    s_chunk[c & 1][0] = (lo / g.R) * g.R;
    s_chunk[c & 1][1] = cls;
  };

  if (threadIdx.x == 0) {
    const double zf = static_cast<double>(z0 + p.cz), yf = static_cast<double>(y0 + p.cy),
                 xf = static_cast<double>(x0 + p.cx);
    cz0 = march_coord(p.m, zf, yf, xf);
    const double cy0 = march_coord(p.m + 4, zf, yf, xf);
    const double cx0 = march_coord(p.m + 8, zf, yf, xf);
    t_b0z = __double2int_rd(cz0 + (g.negc[0] - 1e-6));
    const int rb0y = __double2int_rd(cy0 + (g.negt[1] - 1e-6));
    const int rb0x = __double2int_rd(cx0 + (g.negt[2] - 1e-6));
    const int bhiy = __double2int_rd(cy0 + (g.post[1] + 1e-6)) + 2;
    const int bhix = __double2int_rd(cx0 + (g.post[2] + 1e-6)) + 2;
    t_b0y = rb0y;
    t_b0x = rb0x & ~(kVec - 1);  // innermost TMA coordinate must be 16-byte aligned
    const bool ok = (bhiy - t_b0y) < g.BY && (bhix - t_b0x) < g.BX;
    if (rb0y >= 1 && bhiy <= p.sy - 1 && rb0x >= 1 && bhix <= p.sx - 1) inplane = 1;
    if (bhiy <= -1 || rb0y >= p.sy + 1 || bhix <= -1 || rb0x >= p.sx + 1) inplane = 2;
    s_geo[0] = t_b0z;
    s_geo[1] = t_b0y;
    s_geo[2] = t_b0x;
    s_geo[3] = __float_as_int(static_cast<float>(cz0 - static_cast<double>(t_b0z)));
    s_geo[4] = __float_as_int(static_cast<float>(cy0 - static_cast<double>(t_b0y)));
    s_geo[5] = __float_as_int(static_cast<float>(cx0 - static_cast<double>(t_b0x)));
    s_geo[6] = ok ? 1 : 0;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&done[0], kMaThreads / 32);
    mbar_init(&done[1], kMaThreads / 32);
    fence_mbar_init();
    if (ok) {
      issue_chunk(0);
      if (nchunks > 1) issue_chunk(1);
    }
  }
  __syncthreads();

  const int lx = threadIdx.x % kMaTX, ly = threadIdx.x / kMaTX;
  const int x = x0 + lx;
  if (!(s_geo[6] & 1)) {  // host bound too tight for this tile (never expected): from global
#pragma unroll
    for (int c = 0; c < kMaCols; ++c) {
      const int y = y0 + ly + c * kMaRowStep;
      if (x < p.ox && y < p.oy)
        for (int k = 0; k < nz_cta; ++k)
          p.dst[(static_cast<int64_t>(z0 + k) * p.oy + y) * p.dpitch + x] =
              affine_sample_generic<T, 1, BOUNDARY, SCRUB>(p, z0 + k, y, x);
    }
    return;
  }

  MarchRing rg;
  rg.ring = ring;
  rg.slot_b = static_cast<uint32_t>(g.slot_bytes);
  rg.row_b = static_cast<uint32_t>(g.BX) * es;
  rg.R = g.R;
  rg.b0z = s_geo[0];
  rg.b0y = s_geo[1];
  rg.b0x = s_geo[2];
  rg.BX = g.BX;
  const float c0l[3] = {__int_as_float(s_geo[3]), __int_as_float(s_geo[4]), __int_as_float(s_geo[5])};

  // interior  <=>  |u - mid| <= half - kMaEdge  (brick-local u; both taps valid on the axis)
  float mid[3], half[3];
  {
    const int n[3] = {p.sz, p.sy, p.sx};
    const int b0[3] = {rg.b0z, rg.b0y, rg.b0x};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      half[d] = 0.5f * static_cast<float>(n[d] - 1);
      mid[d] = half[d] - static_cast<float>(b0[d]);
    }
  }
  const int64_t out_plane = static_cast<int64_t>(p.oy) * p.dpitch;
  const bool x_ok = x < p.ox;

  // column starts (fp32, brick-local): tile origin + yy*col1 + lx*col2
  float u0[kMaCols][3];
#pragma unroll
  for (int c = 0; c < kMaCols; ++c) {
    const float yy = static_cast<float>(ly + c * kMaRowStep);
#pragma unroll
    for (int d = 0; d < 3; ++d)
      u0[c][d] = __fmaf_rn(static_cast<float>(lx), g.mcol[3 * d + 2],
                           __fmaf_rn(yy, g.mcol[3 * d + 1], c0l[d]));
  }

  MarchCol cc;
  cc.ring_b = static_cast<uint32_t>(g.R) * rg.slot_b;
  cc.slot_b = rg.slot_b;
  cc.row_b = rg.row_b;
  cc.out_plane = out_plane;
  cc.mz = g.mcol[0];
  cc.my = g.mcol[3];
  cc.mx = g.mcol[6];

  // output address of this thread's first column in the chunk; advanced by kMaCZ planes per chunk
  float* out_col0 = p.dst + (static_cast<int64_t>(z0) * p.oy + (y0 + ly)) * p.dpitch + x;
  const int64_t col_step = static_cast<int64_t>(kMaRowStep) * p.dpitch;
  for (int c = 0; c < nchunks; ++c, out_col0 += kMaCZ * out_plane) {
    mbar_wait(&bar[c & 1], static_cast<uint32_t>(c >> 1) & 1u);
    const int wb = s_chunk[c & 1][0];
    const int cls = s_chunk[c & 1][1];
    const int zc0 = z0 + c * kMaCZ;
    const int nzc = min(kMaCZ, p.oz - zc0);
    const float cf = static_cast<float>(c * kMaCZ);
    cc.abase = ring - (kMaMagicBits + static_cast<uint32_t>(wb)) * rg.slot_b -
               kMaMagicBits * (rg.row_b + es);
    cc.thr = kMaMagicBits + static_cast<uint32_t>(wb) + static_cast<uint32_t>(g.R);
    if (x_ok) {
#pragma unroll
      for (int col = 0; col < kMaCols; ++col) {
        const int y = y0 + ly + col * kMaRowStep;
        if (y >= p.oy) continue;
        float* __restrict__ out = out_col0 + static_cast<int64_t>(col) * col_step;
        if (cls == 2) {
          for (int k = 0; k < nzc; ++k) st_global_cs(out + k * out_plane, 0.0f);
          continue;
        }
        cc.u0z = __fmaf_rn(cf, cc.mz, u0[col][0]);
        cc.u0y = __fmaf_rn(cf, cc.my, u0[col][1]);
        cc.u0x = __fmaf_rn(cf, cc.mx, u0[col][2]);
        uint32_t rest;
        if (nzc != kMaCZ) {
          rest = (1u << nzc) - 1u;  // ragged last chunk: every voxel on the exact path
        } else if (cls == 1) {
          rest = march_column<T, SCRUB, false>(cc, mid, half, out);
        } else {
          rest = march_column<T, SCRUB, true>(cc, mid, half, out);
        }
        while (rest) {
          const int k = __ffs(rest) - 1;
          rest &= rest - 1;
          const float kf = static_cast<float>(k);
          const float dz = fabsf(__fmaf_rn(kf, cc.mz, cc.u0z) - mid[0]);
          const float dy = fabsf(__fmaf_rn(kf, cc.my, cc.u0y) - mid[1]);
          const float dx = fabsf(__fmaf_rn(kf, cc.mx, cc.u0x) - mid[2]);
          const bool outside = dz > half[0] + 0.5f + kMaEdge || dy > half[1] + 0.5f + kMaEdge ||
                               dx > half[2] + 0.5f + kMaEdge;
          const float v =
              outside ? 0.0f : march_sample_exact<T, BOUNDARY, SCRUB>(p, rg, zc0 + k, y, x);
          st_global_cs(out + k * out_plane, v);
        }
      }
    }
    // Once EVERY warp is done with chunk c the planes below chunk c+1's range are dead and their
    // slots may be overwritten by the planes of chunk c+2.  Warps only signal (no CTA barrier):
    // the other warps run on into chunk c+1, whose planes are already there; thread 0 alone waits
    // for the stragglers and issues the loads.
    if (c + 2 < nchunks) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&done[c & 1]);
      if (threadIdx.x == 0) {
        mbar_wait(&done[c & 1], static_cast<uint32_t>(c >> 1) & 1u);
        issue_chunk(c + 2);
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool march_geometry(const AffineParams& p, MarchGeom* g, size_t* smem_bytes, int* grid_z) {
  if (p.order != 1) return false;
  if (!(p.m[0] > 0.0)) return false;  // the march needs a monotone source-plane demand
  if (reinterpret_cast<uintptr_t>(p.src) % 16 != 0) return false;
  if ((static_cast<int64_t>(p.spitch) * sizeof(T)) % 16 != 0) return false;
  const int vec = 16 / sizeof(T);
  int sms = 148;
  sm_count(&sms);
  const int64_t tiles = static_cast<int64_t>((p.oy + kMaTY - 1) / kMaTY) * ((p.ox + kMaTX - 1) / kMaTX);
  const int chunks_total = (p.oz + kMaCZ - 1) / kMaCZ;
  // enough CTAs for several waves of 4 CTAs per SM; otherwise one CTA marches the whole depth
  int zsplit = 1;
  const int64_t target = static_cast<int64_t>(sms) * 16;
  if (tiles < target) zsplit = static_cast<int>((target + tiles - 1) / tiles);
  int zchunks = (chunks_total + zsplit - 1) / zsplit;
  if (zchunks < 2) zchunks = chunks_total < 2 ? chunks_total : 2;
  // the in-plane brick grows with the depth a CTA marches (m10, m20): cap that growth
  const double drift = fabs(p.m[4]) > fabs(p.m[8]) ? fabs(p.m[4]) : fabs(p.m[8]);
  // ... and brick-local z coordinates grow over the march: keep them below 250 (fp32 ulp 1.5e-5)
  const double zside = fabs(p.m[1]) * kMaTY + fabs(p.m[2]) * kMaTX + 8.0;
  while (zchunks > 2 &&
         (drift * (zchunks * kMaCZ) > 12.0 || p.m[0] * (zchunks * kMaCZ) + zside >= 250.0))
    zchunks = (zchunks + 1) / 2;
  g->zchunks = zchunks;
  *grid_z = (chunks_total + zchunks - 1) / zchunks;
  if (*grid_z > 65535) return false;

  const int tc[3] = {kMaCZ - 1, kMaTY - 1, kMaTX - 1};
  const int tt[3] = {zchunks * kMaCZ - 1, kMaTY - 1, kMaTX - 1};
  for (int d = 0; d < 3; ++d) {
    double nc = 0.0, pc = 0.0, nt = 0.0, pt = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double vc = p.m[4 * d + j] * tc[j], vt = p.m[4 * d + j] * tt[j];
      (vc < 0.0 ? nc : pc) += vc;
      (vt < 0.0 ? nt : pt) += vt;
    }
    g->negc[d] = nc; g->posc[d] = pc;
    g->negt[d] = nt; g->post[d] = pt;
    if (!((pt - nt) < 240.0)) return false;
  }
  for (int i = 0; i < 9; ++i) g->mcol[i] = static_cast<float>(p.m[4 * (i / 3) + (i % 3)]);
  for (int d = 0; d < 3; ++d) {  // |coordinate| bound: keeps the device-side int conversions defined
    const double reach = fabs(p.m[4 * d]) * (p.oz + fabs((double)p.cz)) +
                         fabs(p.m[4 * d + 1]) * (p.oy + fabs((double)p.cy)) +
                         fabs(p.m[4 * d + 2]) * (p.ox + fabs((double)p.cx)) + fabs(p.m[4 * d + 3]);
    if (!(reach < 1.0e9)) return false;
  }
  if (!(p.m[0] * (zchunks * kMaCZ) + (g->posc[0] - g->negc[0]) < 256.0)) return false;
  const int BY = static_cast<int>(g->post[1] - g->negt[1]) + 5;
  int BX = static_cast<int>(g->post[2] - g->negt[2]) + 5 + (vec - 1);
  BX = (BX + vec - 1) / vec * vec;
  if (BY > 256 || BX > 256) return false;
  // planes alive at once: from the first plane of chunk c+1 to the last (+2) of chunk c+2
  const double span2 = (g->posc[0] - g->negc[0]) + p.m[0] * kMaCZ;
  const int R = static_cast<int>(span2) + 6;
  if (R > 64) return false;
  const int slot = (BY * BX * static_cast<int>(sizeof(T)) + 127) / 128 * 128;
  const int64_t bytes = static_cast<int64_t>(R) * slot;
  if (bytes > 72 * 1024) return false;  // >= 3 CTAs per SM, else the brick kernel decides
  g->R = R;
  g->BY = BY;
  g->BX = BX;
  g->slot_bytes = slot;
  *smem_bytes = static_cast<size_t>(bytes) + 128;
  return true;
}

template <typename T, int BOUNDARY, bool SCRUB>
static int launch_march(const AffineParams& p, const MarchGeom& g, size_t smem_bytes, int grid_z,
                        cudaStream_t stream) {
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return B2_ERR_NO_DEVICE;
  }
  CUtensorMap map;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p.sx), static_cast<cuuint64_t>(p.sy),
                              static_cast<cuuint64_t>(p.sz)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(p.spitch) * sizeof(T),
                                 static_cast<cuuint64_t>(p.spitch) * p.sy * sizeof(T)};
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
  const int tiles_y = (p.oy + kMaTY - 1) / kMaTY;
  const int tiles_x = (p.ox + kMaTX - 1) / kMaTX;
  const int64_t tiles = static_cast<int64_t>(tiles_y) * tiles_x;
  if (tiles > 2147483647LL) return B2_ERR_UNSUPPORTED;
  auto kern = affine_march_kernel<T, BOUNDARY, SCRUB>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_bytes)));
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(grid_z), 1);
  kern<<<grid, kMaThreads, smem_bytes, stream>>>(map, p, g, tiles_x);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

template <typename T>
static int march_typed(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  MarchGeom g{};
  size_t smem = 0;
  int grid_z = 1;
  *eligible = march_geometry<T>(p, &g, &smem, &grid_z);
  if (!*eligible) return B2_ERR_UNSUPPORTED;
  const bool scrub = p.scrub && sizeof(T) == 4;
  int rc;
  if (p.boundary == B2_BOUNDARY_CONSTANT)
    rc = scrub ? launch_march<T, B2_BOUNDARY_CONSTANT, true>(p, g, smem, grid_z, stream)
               : launch_march<T, B2_BOUNDARY_CONSTANT, false>(p, g, smem, grid_z, stream);
  else
    rc = scrub ? launch_march<T, B2_BOUNDARY_ITK, true>(p, g, smem, grid_z, stream)
               : launch_march<T, B2_BOUNDARY_ITK, false>(p, g, smem, grid_z, stream);
  if (rc == B2_ERR_UNSUPPORTED) *eligible = false;
  return rc;
}

int affine_march_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible) {
  if (src_dtype == B2_DTYPE_U16) return march_typed<uint16_t>(p, stream, eligible);
  return march_typed<float>(p, stream, eligible);
}

}  // namespace b2
