// TMA-staged affine pull warp for GENERIC 3x4 matrices (out-of-plane rotations, shears, flips):
// the case `affine_zsep_kernel` cannot take.  One CTA = one output tile (TZ x TY x TX); the
// tile's back-projected bounding box is a 3-D brick of the source that is staged in shared
// memory with ONE 3-D TMA box load (out-of-bounds zero fill, origin rounded down to 16 bytes).
// Tiles are rasterised z-fastest so that tiles sharing z-halo planes run concurrently and the
// halos are served from L2.
//
// Coordinates: each thread owns one (y, x) column of the tile and walks TZ output planes.  Its
// first coordinate is evaluated in float64 in the oracle's op order and expressed relative to the
// brick origin; along z it advances by fp32 increments (|error| ~1e-5 voxel, well inside the
// 1e-4-of-range tolerance).  Any voxel whose coordinate comes within kEdge of a decision edge
// (volume border, ITK half-voxel band, the k+0.5 rounding point of nearest-neighbour) is
// re-evaluated exactly in float64 through `resolve_axis`, so inside/outside and nearest-neighbour
// index decisions are identical to the float64 oracle (order 0 stays bit-exact).
#include <algorithm>
#include <cstdlib>

#include "b2_affine.cuh"

namespace b2 {

// Tile depth along z.  (16-deep tiles — less z halo per output voxel, 3 CTAs/SM — and a persistent
// double-buffered variant with a producer warp were built in round 2, are bit-compatible, and
// measure the same 1.2 ms on the C3-sized volume as 8-deep tiles: the kernel is bound by the
// per-warp dependency chain, not by brick traffic, the TMA wait or the issue rate; an L2 prefetch of
// the tile one wave ahead (cp.async.bulk.prefetch.tensor) measured flat as well; git history.)
constexpr int kBrTZ = 8;
// In-plane tile: 16 (y) x 32 (x) with lanes along x; LY variant 32 (y) x 16 (x) with lanes along y
// for matrices that map output y onto source x (90-degree in-plane rotations: with lanes along x
// a warp's taps walk a brick column, 8-way bank conflicts) — its output goes through a padded
// shared-memory tile so that the global stores stay row-contiguous.
constexpr int kBrLanes = 32;  // tile extent along the lane axis
constexpr int kBrOther = 16;  // tile extent along the other in-plane axis
constexpr int kBrThreads = 256;
constexpr int kBrCols = (kBrLanes * kBrOther) / kBrThreads;  // (y, x) columns per thread
constexpr int kBrRowStep = kBrThreads / kBrLanes;            // distance between a thread's columns
constexpr int kBrOutPitch = kBrOther + 1;                    // LY: padded row pitch of the staged output
constexpr int kBrStageBytes = kBrTZ * kBrLanes * kBrOutPitch * 4;  // LY tiles stay 8 deep
// slack (voxels) added to both ends of a tile's back-projected hull before it is floored: covers
// the rounding of the host's hull sums and the fp32 coordinate error of the fast path (~1e-5)
constexpr double kBrGuard = 1.0e-4;
constexpr float kEdge = 2.0e-3f;
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: (v + kMagic) - kMagic rounds v to nearest

struct BrickGeom {
  int BZ, BY, BX;  // brick extent (elements)
  int bytes;       // BZ*BY*BX*sizeof(T)
  // hull of a full tile relative to its origin voxel: sum of the negative / positive parts of
  // m_dj * (T_j - 1) per source axis d (host, float64)
  double neg[3], pos[3];
  float mcol[9];   // float32 copy of the 3x3 linear part (row-major) for the fp32 increments
  float half[3];   // 0.5 * (n_d - 1) of the source axes: interior <=> |u - mid| <= half - edge
  long long out_plane;  // elements between output planes (oy * dpitch), formed once on the host
  unsigned out_plane_bytes;  // the same in bytes (the launch admits only kBrTZ * this < 2^32)
};

template <typename T>
__device__ __forceinline__ float brick_elem(uint32_t addr);
template <>
__device__ __forceinline__ float brick_elem<float>(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <>
__device__ __forceinline__ float brick_elem<uint16_t>(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return static_cast<float>(v);
}

// four taps of one source plane (rows a0 / a1, columns +0 / +1), loaded only when `need`;
// otherwise the registers keep their value.  Predicated LDS: no branch, no wavefronts when off.
template <typename T>
__device__ __forceinline__ void brick_quad_if(bool need, uint32_t a0, uint32_t a1, float& v00,
                                              float& v01, float& v10, float& v11);
template <>
__device__ __forceinline__ void brick_quad_if<float>(bool need, uint32_t a0, uint32_t a1, float& v00,
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
__device__ __forceinline__ void brick_quad_if<uint16_t>(bool need, uint32_t a0, uint32_t a1,
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

__device__ __forceinline__ double coord_full(const double* m, double zf, double yf, double xf) {
  return __dadd_rn(__dadd_rn(__dadd_rn(m[3], __dmul_rn(zf, m[0])), __dmul_rn(yf, m[1])),
                   __dmul_rn(xf, m[2]));
}

// exact (float64) evaluation of one voxel with taps read from the brick
template <typename T, int ORDER, int BOUNDARY, bool SCRUB>
__device__ __noinline__ float brick_sample_exact(const AffineParams& p, uint32_t brick, int bz0,
                                                 int by0, int bx0, int BZ, int BY, int BX, int z,
                                                 int y, int x) {
  const double zf = static_cast<double>(z + p.cz), yf = static_cast<double>(y + p.cy),
               xf = static_cast<double>(x + p.cx);
  const AxisTap tz = resolve_axis<ORDER, BOUNDARY>(coord_full(p.m, zf, yf, xf), p.sz);
  const AxisTap ty = resolve_axis<ORDER, BOUNDARY>(coord_full(p.m + 4, zf, yf, xf), p.sy);
  const AxisTap tx = resolve_axis<ORDER, BOUNDARY>(coord_full(p.m + 8, zf, yf, xf), p.sx);
  if (!(tz.inside && ty.inside && tx.inside)) return 0.0f;
  // a tap with zero weight (the dropped neighbour of the ITK half-voxel band) may sit one element
  // past the tight brick: clamp into it (the value is multiplied by 0, it only has to be loadable)
  auto tap = [&](int iz, int iy, int ix) {
    iz = min(max(iz - bz0, 0), BZ - 1) + bz0;
    iy = min(max(iy - by0, 0), BY - 1) + by0;
    ix = min(max(ix - bx0, 0), BX - 1) + bx0;
    const uint32_t off = static_cast<uint32_t>(((iz - bz0) * BY + (iy - by0)) * BX + (ix - bx0));
    B2_SMEM_CHECK(off, 0u, static_cast<uint32_t>(BZ * BY * BX));
    float v = brick_elem<T>(brick + off * static_cast<uint32_t>(sizeof(T)));
    if (SCRUB && sizeof(T) == 4) v = scrub_value(v);
    return v;
  };
  if (ORDER == 0) return tap(tz.i0, ty.i0, tx.i0);
  const float p0 = lerp_w(lerp_w(tap(tz.i0, ty.i0, tx.i0), tap(tz.i0, ty.i0, tx.i1), tx.w),
                          lerp_w(tap(tz.i0, ty.i1, tx.i0), tap(tz.i0, ty.i1, tx.i1), tx.w), ty.w);
  const float p1 = lerp_w(lerp_w(tap(tz.i1, ty.i0, tx.i0), tap(tz.i1, ty.i0, tx.i1), tx.w),
                          lerp_w(tap(tz.i1, ty.i1, tx.i0), tap(tz.i1, ty.i1, tx.i1), tx.w), ty.w);
  return lerp_w(p0, p1, tz.w);
}

// one output voxel: straight to global memory (streaming), or — LY — into the staged output tile
// in shared memory (generic store)
template <bool LY>
__device__ __forceinline__ void brick_put(float* o, float v) {
  if (LY) {
    *o = v;
  } else {
    st_global_cs(o, v);
  }
}

struct BrickCol {
  uint32_t brick, plane_b, row_b;  // `brick` = brick base minus the magic-floor index biases (packed routine)
  uint32_t lo, hi;                 // shared-memory extent of the brick (B2_BOUNDS_CHECK builds)
  int64_t out_plane;
  float u0z, u0y, u0x;  // brick-local coordinate of the column's first voxel
  float mz, my, mx;     // coordinate step per output plane (first column of the matrix)
};

// One (y, x) column of a full-depth tile, order 1.  The voxels are walked along output z; every
// strictly interior voxel (both taps valid on every axis, kEdge away from every decision edge) is
// interpolated from the brick in fp32 and stored.  Returns the bit mask of the voxels that were NOT
// written (not strictly interior, or a non-finite tap turned up): the caller finishes those on the
// exact path.  CHECK = false: the whole tile is known to be strictly interior (decided once per
// CTA from the brick hull) and the per-voxel test is compiled out.  ~40 instructions per voxel.
template <typename T, bool SCRUB, bool CHECK, bool LY>
__device__ __forceinline__ uint32_t brick_column_linear(const BrickCol& c, const float (&mid)[3],
                                                        const float (&half)[3],
                                                        float* __restrict__ out) {
  constexpr uint32_t es = static_cast<uint32_t>(sizeof(T));
  // floor via the magic constant in round-down mode: t = RD(u + kMagic) holds floor(u) in its low
  // mantissa bits (FADD.RM is a full-rate FADD), weight w = u - (t - kMagic).  The sum of the
  // three index biases (0x4B400000 each) is folded into the base address.
  const uint32_t abase = c.brick - 0x4B400000u * (c.plane_b + c.row_b + es);
  uint32_t rest = 0;
  // taps of the upper source plane (iz+1) of the previous voxel: when the next voxel of the column
  // sits in the same (y, x) cell one plane further - the usual case for m00 ~ 1 and small
  // out-of-plane terms - they are its lower-plane taps and are not read again (halves the
  // shared-memory wavefronts)
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f;
  uint32_t a_up = 0xffffffffu;
  float* __restrict__ o = out;
#pragma unroll
  for (int k = 0; k < kBrTZ; ++k, o += c.out_plane) {
    const float kf = static_cast<float>(k);
    const float uz = __fmaf_rn(kf, c.mz, c.u0z);
    const float uy = __fmaf_rn(kf, c.my, c.u0y);
    const float ux = __fmaf_rn(kf, c.mx, c.u0x);
    if (CHECK) {
      const bool interior = fabsf(uz - mid[0]) <= half[0] - kEdge &&
                            fabsf(uy - mid[1]) <= half[1] - kEdge &&
                            fabsf(ux - mid[2]) <= half[2] - kEdge;
      if (!interior) {
        rest |= 1u << k;
        a_up = 0xffffffffu;
        continue;
      }
    }
    const float tz = __fadd_rd(uz, kMagic), ty = __fadd_rd(uy, kMagic), tx = __fadd_rd(ux, kMagic);
    const float wz = uz - (tz - kMagic);
    const float wy = uy - (ty - kMagic);
    const float wx = ux - (tx - kMagic);
    const uint32_t a00 = static_cast<uint32_t>(__float_as_int(tz)) * c.plane_b +
                         (static_cast<uint32_t>(__float_as_int(ty)) * c.row_b +
                          (static_cast<uint32_t>(__float_as_int(tx)) * es + abase));
    const uint32_t a10 = a00 + c.plane_b;
    B2_SMEM_CHECK(a00, c.lo, c.hi);
    B2_SMEM_CHECK(a10 + c.row_b + es, c.lo, c.hi);
    // lower plane: r0..r3 keep their value unless the cell changed
    brick_quad_if<T>(a00 != a_up, a00, a00 + c.row_b, r0, r1, r2, r3);
    const float v000 = r0, v001 = r1, v010 = r2, v011 = r3;
    r0 = brick_elem<T>(a10);
    r1 = brick_elem<T>(a10 + es);
    r2 = brick_elem<T>(a10 + c.row_b);
    r3 = brick_elem<T>(a10 + c.row_b + es);
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
    brick_put<LY>(o, v);
  }
  return rest;
}

// the four in-plane taps at `a` (rows a / a + row_b, columns +0 / +1), loaded only when the cell
// address differs from `a_prev`; otherwise the registers keep their value.  The compare lives inside
// the asm block so that the predicate never round-trips through a register.
template <typename T>
__device__ __forceinline__ void brick_quad_moved(uint32_t a, uint32_t a_prev, uint32_t row_b, float& v00,
                                                 float& v01, float& v10, float& v11);
template <>
__device__ __forceinline__ void brick_quad_moved<float>(uint32_t a, uint32_t a_prev, uint32_t row_b,
                                                        float& v00, float& v01, float& v10, float& v11) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 a1;\n"
      "setp.ne.u32 p, %4, %5;\n"
      "add.u32 a1, %4, %6;\n"
      "@p ld.shared.f32 %0, [%4];\n"
      "@p ld.shared.f32 %1, [%4+4];\n"
      "@p ld.shared.f32 %2, [a1];\n"
      "@p ld.shared.f32 %3, [a1+4];\n"
      "}\n"
      : "+f"(v00), "+f"(v01), "+f"(v10), "+f"(v11)
      : "r"(a), "r"(a_prev), "r"(row_b));
}
template <>
__device__ __forceinline__ void brick_quad_moved<uint16_t>(uint32_t a, uint32_t a_prev, uint32_t row_b,
                                                           float& v00, float& v01, float& v10,
                                                           float& v11) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 a1;\n"
      ".reg .b16 h0, h1, h2, h3;\n"
      "setp.ne.u32 p, %4, %5;\n"
      "add.u32 a1, %4, %6;\n"
      "@p ld.shared.u16 h0, [%4];\n"
      "@p ld.shared.u16 h1, [%4+2];\n"
      "@p ld.shared.u16 h2, [a1];\n"
      "@p ld.shared.u16 h3, [a1+2];\n"
      "@p cvt.rn.f32.u16 %0, h0;\n"
      "@p cvt.rn.f32.u16 %1, h1;\n"
      "@p cvt.rn.f32.u16 %2, h2;\n"
      "@p cvt.rn.f32.u16 %3, h3;\n"
      "}\n"
      : "+f"(v00), "+f"(v01), "+f"(v10), "+f"(v11)
      : "r"(a), "r"(a_prev), "r"(row_b));
}

// NC (y, x) columns of a tile walked in LOCKSTEP, order 1, FINITE taps (uint16, or float32 with
// the scrub: a non-finite result sends the column to the exact path) — the packed-arithmetic form
// of brick_column_linear.
//  * Per column the four in-plane taps of the two source planes sit in four register PAIRS (lo
//    half / hi half = the two planes); the upper plane of voxel k is the lower plane of voxel k+1,
//    so the halves swap roles from step to step (the loop is fully unrolled: the parity is static)
//    and only the upper plane is loaded — the lower one again only when the (y, x) cell moved.
//  * Per voxel: 2 coordinate adds, 3 floors, 5 weight ops (y and x packed), 3 address, 1 compare,
//    4 + 4 predicated LDS, 8 lerp instructions (3 FADD2 + 3 FFMA2 + 2 scalar), check + store.
//  * Why lockstep: one column is ONE dependency chain (coordinates -> address -> LDS -> lerps, and
//    the tap hand-down serialises consecutive voxels); ncu showed every variant of this kernel
//    (4, 3 or 2 CTAs/SM; 61, 53 or 46 instructions per voxel; TMA wait hidden or not) at the same
//    1.2 ms with ~10 cycles between a warp's instructions.  Two independent chains in the same
//    basic block give the scheduler something to issue in between.
// Partial-depth tiles (nz < kBrTZ) compute every voxel (the brick always covers a full tile) and
// only predicate the store.  `c[i].brick` has the magic-floor index biases folded in.
template <typename T, bool CHECK, bool LY, int BOUNDARY, int NC>
__device__ __forceinline__ void brick_columns_packed(const BrickCol (&c)[NC], const float (&mid)[3],
                                                     const float (&half)[3],
                                                     float* const (&out)[NC], const int nz,
                                                     uint32_t (&rest)[NC], const uint32_t plane_bytes) {
  constexpr uint32_t es = static_cast<uint32_t>(sizeof(T));
  float bad[NC];      // turns NaN as soon as one voxel of the column is non-finite (v * 0 accumulates)
  float qa[NC][4];    // lo halves: taps (y0,x0) (y0,x1) (y1,x0) (y1,x1) of one plane
  float qb[NC][4];    // hi halves: the same taps of the other plane
  uint32_t a_up[NC];
  float uz[NC];
  f32x2 uyx[NC], myx[NC];
  char* o[NC];        // byte pointers: one 64-bit add per plane, no index scaling
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    bad[i] = 0.0f;
    rest[i] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) qa[i][j] = qb[i][j] = 0.0f;
    a_up[i] = 0xffffffffu;
    uz[i] = c[i].u0z;
    uyx[i] = pk2(c[i].u0y, c[i].u0x);
    myx[i] = pk2(c[i].my, c[i].mx);
    o[i] = reinterpret_cast<char*>(out[i]);
  }
  const f32x2 magic2 = pk2(kMagic, kMagic);
  // plane_bytes: byte offset between output planes, formed on the host (the launch admits only
  // outputs with kBrTZ * plane_bytes < 2^32): k * plane_bytes is one uniform multiply per store
#pragma unroll
  for (int k = 0; k < kBrTZ; ++k) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      bool interior = true;
      float cz = uz[i];       // the coordinate this voxel samples at (clamped in the ITK band)
      f32x2 cyx = uyx[i];
      if (CHECK) {
        float uy, ux;
        upk2(uyx[i], uy, ux);
        if (BOUNDARY == B2_BOUNDARY_ITK) {
          // ITK: inside iff -0.5 <= c < n - 0.5, and inside the half-voxel band the interpolator
          // clamps to the edge voxel (base clamped, neighbour dropped) = trilinear at the CLAMPED
          // coordinate, continuous across c = 0 and c = n-1: only the +-0.5 edges are decisions
          interior = fabsf(uz[i] - mid[0]) <= half[0] + 0.5f - kEdge &&
                     fabsf(uy - mid[1]) <= half[1] + 0.5f - kEdge &&
                     fabsf(ux - mid[2]) <= half[2] + 0.5f - kEdge;
          cz = fminf(fmaxf(uz[i], mid[0] - half[0]), mid[0] + half[0]);
          cyx = pk2(fminf(fmaxf(uy, mid[1] - half[1]), mid[1] + half[1]),
                    fminf(fmaxf(ux, mid[2] - half[2]), mid[2] + half[2]));
        } else {
          interior = fabsf(uz[i] - mid[0]) <= half[0] - kEdge && fabsf(uy - mid[1]) <= half[1] - kEdge &&
                     fabsf(ux - mid[2]) <= half[2] - kEdge;
        }
      }
      if (interior) {
        const float tz = __fadd_rd(cz, kMagic);
        const f32x2 tyx = add2_rd(cyx, magic2);
        const float wz = cz - (tz - kMagic);
        const f32x2 wyx = sub2(cyx, sub2(tyx, magic2));
        float ty, tx, wy, wx;
        upk2(tyx, ty, tx);
        upk2(wyx, wy, wx);
        const uint32_t a00 = static_cast<uint32_t>(__float_as_int(tz)) * c[i].plane_b +
                             (static_cast<uint32_t>(__float_as_int(ty)) * c[i].row_b +
                              (static_cast<uint32_t>(__float_as_int(tx)) * es + c[i].brick));
        const uint32_t a10 = a00 + c[i].plane_b;
        B2_SMEM_CHECK(a00, c[i].lo, c[i].hi);
        B2_SMEM_CHECK(a10 + c[i].row_b + es, c[i].lo, c[i].hi);
        float (&lo)[4] = (k & 1) ? qb[i] : qa[i];  // lower source plane of this voxel
        float (&up)[4] = (k & 1) ? qa[i] : qb[i];  // upper plane: loaded now, lower plane next step
        brick_quad_moved<T>(a00, a_up[i], c[i].row_b, lo[0], lo[1], lo[2], lo[3]);
        up[0] = brick_elem<T>(a10);
        up[1] = brick_elem<T>(a10 + es);
        up[2] = brick_elem<T>(a10 + c[i].row_b);
        up[3] = brick_elem<T>(a10 + c[i].row_b + es);
        a_up[i] = a10;
        const f32x2 q00 = pk2(qa[i][0], qb[i][0]), q01 = pk2(qa[i][1], qb[i][1]);
        const f32x2 q10 = pk2(qa[i][2], qb[i][2]), q11 = pk2(qa[i][3], qb[i][3]);
        const f32x2 wx2 = pk2(wx, wx), wy2 = pk2(wy, wy);
        const f32x2 x0 = fma2(wx2, sub2(q01, q00), q00);  // x lerp on row y0 of both planes
        const f32x2 x1 = fma2(wx2, sub2(q11, q10), q10);  // ... on row y1
        const f32x2 yy = fma2(wy2, sub2(x1, x0), x0);     // in-plane bilinear value of both planes
        float pa, pb;
        upk2(yy, pa, pb);
        const float vlo = (k & 1) ? pb : pa, vup = (k & 1) ? pa : pb;
        const float v = __fmaf_rn(wz, vup - vlo, vlo);
        // a NaN/inf tap makes v non-finite; the voxel is stored anyway and the whole column is
        // redone on the exact path (which applies the scrub per tap) after the loop
        if (sizeof(T) == 4) bad[i] = __fmaf_rn(v, 0.0f, bad[i]);
        if (k < nz)
          brick_put<LY>(reinterpret_cast<float*>(o[i] + static_cast<uint32_t>(k) * plane_bytes), v);
      } else {
        if (k < nz) rest[i] |= 1u << k;
        a_up[i] = 0xffffffffu;
      }
      uz[i] += c[i].mz;
      uyx[i] = add2(uyx[i], myx[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i)
    if (bad[i] != bad[i]) rest[i] = (1u << nz) - 1u;
}

#ifndef B2_BRICK_MINB
#define B2_BRICK_MINB 4  // CTAs per SM the register allocation aims at (64 registers)
#endif
template <typename T, int ORDER, int BOUNDARY, bool SCRUB, bool LY>
__global__ void __launch_bounds__(kBrThreads, B2_BRICK_MINB)
    affine_brick_kernel(const __grid_constant__ CUtensorMap src_map,
                        const __grid_constant__ AffineParams p,
                        const __grid_constant__ BrickGeom g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  constexpr int kVec = 16 / static_cast<int>(sizeof(T));
  constexpr int kBrTY = LY ? kBrLanes : kBrOther, kBrTX = LY ? kBrOther : kBrLanes;
  // 128-byte aligned brick, in shared-window addresses (no generic pointer arithmetic per thread)
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t brick = (smem_base + 127u) & ~127u;
  // LY: staged output tile [kBrTZ][kBrTY][kBrOutPitch] behind the brick
  float* stage_out =
      LY ? reinterpret_cast<float*>(smem_raw + (brick - smem_base) + ((g.bytes + 127) / 128) * 128)
         : nullptr;

  // z-fastest rasterisation: consecutive CTAs share z-halo planes
  // 3-D grid (z tiles, x tiles, y tiles): x is the fastest launch dimension, so the rasterisation
  // is z-fastest without the two runtime integer divisions a flat tile index costs every warp
  const int tz_i = blockIdx.x, tx_i = blockIdx.y, ty_i = blockIdx.z;
  const int z0 = tz_i * kBrTZ, y0 = ty_i * kBrTY, x0 = tx_i * kBrTX;
  const int nz = min(kBrTZ, p.oz - z0);

  // ---- brick origin: exact float64 coordinate of the tile origin + the host-computed hull of a
  //      full tile (the kBrGuard slack absorbs the rounding of the hull sums and the fp32 error of
  //      the fast path's coordinates; the host guarantees
  //      |coordinate| < 1e9 over the whole output so the int conversions are defined).
  //      CTA-uniform, so ONE thread evaluates it (float64, ~170 instructions), issues the TMA
  //      load and publishes the result through shared memory.
  // b0[3], bits(c0l[3]), flags (1 = brick ok, 2 = tile strictly interior, 4 = tile outside),
  // bits(mid[3]) = source centre in brick-local coordinates
  __shared__ __align__(16) int s_geo[12];
  if (threadIdx.x == 0) {
    int tb0[3], tbhi[3];
    const double zf = static_cast<double>(z0 + p.cz), yf = static_cast<double>(y0 + p.cy),
                 xf = static_cast<double>(x0 + p.cx);
    const int n[3] = {p.sz, p.sy, p.sx};
    bool inside = true, outside = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double c = coord_full(p.m + 4 * d, zf, yf, xf);
      tb0[d] = __double2int_rd(c + (g.neg[d] - kBrGuard));
      // the hull of the tile misses the source (and its half-voxel ITK band) on this axis
      outside = outside || tb0[d] >= n[d] + 1 || __double2int_rd(c + (g.pos[d] + kBrGuard)) <= -2;
      // every tap of every voxel of a full tile is a valid source index, with >= 1 voxel margin
      inside = inside && tb0[d] >= 1;
      if (d == 2) tb0[d] &= ~(kVec - 1);  // innermost TMA coordinate must be 16-byte aligned
      tbhi[d] = __double2int_rd(c + (g.pos[d] + kBrGuard)) + 1;  // index of the last tap
      inside = inside && tbhi[d] <= n[d] - 2;
      s_geo[d] = tb0[d];
      s_geo[3 + d] = __float_as_int(static_cast<float>(c - static_cast<double>(tb0[d])));
    }
    const bool ok = (tbhi[0] - tb0[0]) < g.BZ && (tbhi[1] - tb0[1]) < g.BY && (tbhi[2] - tb0[2]) < g.BX;
    s_geo[6] = (ok ? 1 : 0) | (inside ? 2 : 0) | (outside ? 4 : 0);
    mbar_init(&bar, 1);
    fence_mbar_init();
    if (ok && !outside) {
      mbar_expect_tx(&bar, static_cast<uint32_t>(g.bytes));
      tma_load_3d(brick, &src_map, &bar, tb0[2], tb0[1], tb0[0]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) s_geo[8 + d] = __float_as_int(g.half[d] - static_cast<float>(tb0[d]));
  }
  __syncthreads();
  // three 16-byte loads instead of ten 4-byte ones
  const int4 ga = *reinterpret_cast<const int4*>(s_geo);
  const int4 gb = *reinterpret_cast<const int4*>(s_geo + 4);
  const int4 gc = *reinterpret_cast<const int4*>(s_geo + 8);
  const int b0[3] = {ga.x, ga.y, ga.z};
  // tile-origin coordinate relative to the brick origin
  const float c0l[3] = {__int_as_float(ga.w), __int_as_float(gb.x), __int_as_float(gb.y)};
  const int geo_flags = gb.z;
  const bool brick_ok = (geo_flags & 1) != 0;
  const bool tile_in = (geo_flags & 2) != 0;
  if (geo_flags & 4) {  // the whole tile maps outside the source: zeros, nothing to load
    if (x0 + kBrTX <= p.ox && y0 + kBrTY <= p.oy && (p.dpitch & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0) {
      // full tile, 16-byte aligned rows: one 16-byte store per 4 voxels, the row pointer advanced
      // by additions (this path is 12 % of the tiles of a rotated volume)
      constexpr int kQ = kBrTX / 4;  // 16-byte pieces per tile row
      static_assert(kBrThreads % (kQ * kBrTY) == 0, "zero fill: whole planes per pass");
      constexpr int kPlanesPerPass = kBrThreads / (kQ * kBrTY);
      const int q = threadIdx.x % kQ, yy = (threadIdx.x / kQ) % kBrTY, k0 = threadIdx.x / (kQ * kBrTY);
      const int64_t plane = static_cast<int64_t>(p.oy) * p.dpitch;
      float* o = p.dst + (static_cast<int64_t>(z0 + k0) * p.oy + y0 + yy) * p.dpitch + x0 + 4 * q;
      for (int k = k0; k < nz; k += kPlanesPerPass, o += kPlanesPerPass * plane)
        st_global_cs4(o, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
      return;
    }
    for (int i = threadIdx.x; i < nz * kBrTY * kBrTX; i += kBrThreads) {
      const int xx = i % kBrTX, yy = (i / kBrTX) % kBrTY, k = i / (kBrTX * kBrTY);
      if (x0 + xx < p.ox && y0 + yy < p.oy)
        st_global_cs(p.dst + (static_cast<int64_t>(z0 + k) * p.oy + y0 + yy) * p.dpitch + x0 + xx, 0.0f);
    }
    return;
  }

  const int lane = threadIdx.x % kBrLanes, oth = threadIdx.x / kBrLanes;

  if (!brick_ok) {  // host bound too tight for this tile (never expected): straight from global
#pragma unroll
    for (int c = 0; c < kBrCols; ++c) {
      const int o2 = oth + c * kBrRowStep;
      const int y = y0 + (LY ? lane : o2), x = x0 + (LY ? o2 : lane);
      if (x < p.ox && y < p.oy)
        for (int k = 0; k < nz; ++k)
          p.dst[(static_cast<int64_t>(z0 + k) * p.oy + y) * p.dpitch + x] =
              affine_sample_generic<T, ORDER, BOUNDARY, SCRUB>(p, z0 + k, y, x);
    }
    return;
  }

  // ---- per-axis constants in brick-local coordinates
  // interior  <=>  |u - mid| <= half - kEdge   (both taps valid, away from every volume edge)
  // outside   <=>  |u - mid| >  half + 0.5 + kEdge on some axis
  float mcol[3][3], mid[3], half[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int j = 0; j < 3; ++j) mcol[d][j] = g.mcol[3 * d + j];
    half[d] = g.half[d];
  }
  mid[0] = __int_as_float(gc.x);
  mid[1] = __int_as_float(gc.y);
  mid[2] = __int_as_float(gc.z);
  const uint32_t es = static_cast<uint32_t>(sizeof(T));
  const uint32_t row_b = static_cast<uint32_t>(g.BX) * es;
  const uint32_t plane_b = static_cast<uint32_t>(g.BY) * row_b;
  mbar_wait(&bar, 0);

  // output voxel (k, yy, xx) of the tile goes to out + k * out_plane: global memory, or the
  // staged tile in shared memory (LY)
  const int64_t out_plane = LY ? static_cast<int64_t>(kBrTY * kBrOutPitch)
                               : static_cast<int64_t>(g.out_plane);
  const uint32_t out_plane_bytes = LY ? static_cast<uint32_t>(kBrTY * kBrOutPitch * 4) : g.out_plane_bytes;
  // the thread's kBrCols columns: (yy, xx), validity, brick-local start coordinate, output pointer
  bool col_ok[kBrCols];
  int col_y[kBrCols], col_x[kBrCols];
  float u0s[kBrCols][3];
  float* outs[kBrCols];
  {
    // column 0 from the tile origin; the following ones step kBrRowStep positions along the
    // non-lane axis (3 FADD + one pointer add each instead of 6 FFMA + a 64-bit index product)
    const int yy = LY ? lane : oth, xx = LY ? oth : lane;
    col_y[0] = y0 + yy;
    col_x[0] = x0 + xx;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      u0s[0][d] = __fmaf_rn(static_cast<float>(xx), mcol[d][2],
                            __fmaf_rn(static_cast<float>(yy), mcol[d][1], c0l[d]));
    // global: CTA-uniform tile origin (uniform datapath) + a 32-bit offset inside the tile layer
    outs[0] = LY ? stage_out + yy * kBrOutPitch + xx
                 : (p.dst + static_cast<int64_t>(z0) * g.out_plane +
                    static_cast<int64_t>(y0) * p.dpitch + x0) +
                       (yy * p.dpitch + xx);
    const int64_t ostep = LY ? static_cast<int64_t>(kBrRowStep)
                             : static_cast<int64_t>(kBrRowStep) * p.dpitch;
#pragma unroll
    for (int c = 1; c < kBrCols; ++c) {
      col_y[c] = col_y[c - 1] + (LY ? 0 : kBrRowStep);
      col_x[c] = col_x[c - 1] + (LY ? kBrRowStep : 0);
#pragma unroll
      for (int d = 0; d < 3; ++d)
        u0s[c][d] = __fmaf_rn(static_cast<float>(kBrRowStep), mcol[d][LY ? 2 : 1], u0s[c - 1][d]);
      outs[c] = outs[c - 1] + ostep;
    }
#pragma unroll
    for (int c = 0; c < kBrCols; ++c) col_ok[c] = col_x[c] < p.ox && col_y[c] < p.oy;
  }
  // voxels the fast path left: near a decision edge, outside the source, or non-finite taps
  auto finish_exact = [&](int c, uint32_t rest) {
    while (rest) {
      const int k = __ffs(rest) - 1;
      rest &= rest - 1;
      const float kf = static_cast<float>(k);
      const float dz = fabsf(__fmaf_rn(kf, mcol[0][0], u0s[c][0]) - mid[0]);
      const float dy = fabsf(__fmaf_rn(kf, mcol[1][0], u0s[c][1]) - mid[1]);
      const float dx = fabsf(__fmaf_rn(kf, mcol[2][0], u0s[c][2]) - mid[2]);
      const bool outside = dz > half[0] + 0.5f + kEdge || dy > half[1] + 0.5f + kEdge ||
                           dx > half[2] + 0.5f + kEdge;
      const float v = outside ? 0.0f
                              : brick_sample_exact<T, ORDER, BOUNDARY, SCRUB>(
                                    p, brick, b0[0], b0[1], b0[2], g.BZ, g.BY, g.BX, z0 + k, col_y[c],
                                    col_x[c]);
      brick_put<LY>(outs[c] + k * out_plane, v);
    }
  };
  constexpr bool kPacked = ORDER == 1 && (SCRUB || sizeof(T) == 2);  // finite taps
  // A tile that is not strictly inside by its hull (`tile_in`) still has warps whose columns are:
  // coordinates are linear in k, so a column is strictly interior iff its two end voxels are
  // (all kBrTZ of them — the brick always covers a full tile).  Decided per WARP (vote), so the
  // unchecked routine is taken without divergence, e.g. by the whole first z layer of a volume
  // whose z shift keeps its taps inside but closer than one voxel to plane 0.
  bool warp_in = tile_in;
  if (kPacked && !tile_in) {
    bool mine = true;
#pragma unroll
    for (int c = 0; c < kBrCols; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float ue = __fmaf_rn(static_cast<float>(kBrTZ - 1), mcol[d][0], u0s[c][d]);
        mine = mine && fabsf(u0s[c][d] - mid[d]) <= half[d] - kEdge &&
               fabsf(ue - mid[d]) <= half[d] - kEdge;
      }
    warp_in = __all_sync(0xffffffffu, mine);
  }
  if (kPacked) {
    const uint32_t biased = brick - 0x4B400000u * (plane_b + row_b + es);  // magic-floor index biases
    bool all_ok = true;
#pragma unroll
    for (int c = 0; c < kBrCols; ++c) all_ok = all_ok && col_ok[c];
    if (all_ok) {
      // ---- both columns in lockstep (every tile but the ragged ones at the volume's y / x end)
      BrickCol cc[kBrCols];
#pragma unroll
      for (int c = 0; c < kBrCols; ++c)
        cc[c] = BrickCol{biased, plane_b, row_b, brick, brick + static_cast<uint32_t>(g.bytes), out_plane,
                         u0s[c][0], u0s[c][1], u0s[c][2], mcol[0][0], mcol[1][0], mcol[2][0]};
      uint32_t rest[kBrCols];
      if (warp_in) {
        brick_columns_packed<T, false, LY, BOUNDARY, kBrCols>(cc, mid, half, outs, nz, rest, out_plane_bytes);
      } else {
        brick_columns_packed<T, true, LY, BOUNDARY, kBrCols>(cc, mid, half, outs, nz, rest, out_plane_bytes);
      }
#pragma unroll
      for (int c = 0; c < kBrCols; ++c) finish_exact(c, rest[c]);
    } else {
#pragma unroll
      for (int c = 0; c < kBrCols; ++c) {
        if (!col_ok[c]) continue;
        const BrickCol cc[1] = {BrickCol{biased, plane_b, row_b, brick,
                                         brick + static_cast<uint32_t>(g.bytes), out_plane, u0s[c][0],
                                         u0s[c][1], u0s[c][2], mcol[0][0], mcol[1][0], mcol[2][0]}};
        float* const o1[1] = {outs[c]};
        uint32_t rest[1];
        brick_columns_packed<T, true, LY, BOUNDARY, 1>(cc, mid, half, o1, nz, rest, out_plane_bytes);
        finish_exact(c, rest[0]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kBrCols; ++c) {
    if (kPacked || !col_ok[c]) continue;
    const int y = col_y[c], x = col_x[c];
    const float (&u0)[3] = u0s[c];
    float* __restrict__ out = outs[c];
    uint32_t todo = 0;  // bit k: voxel k is not strictly interior -> handled after the hot loop
    float* __restrict__ o = out;
    if (ORDER == 1 && nz == kBrTZ) {
      // ---- float32 taps that may be non-finite (scrub off): brick_column_linear
      const BrickCol cc{brick, plane_b, row_b, brick, brick + static_cast<uint32_t>(g.bytes), out_plane,
                        u0[0], u0[1], u0[2], mcol[0][0], mcol[1][0], mcol[2][0]};
      const uint32_t rest = tile_in ? brick_column_linear<T, SCRUB, false, LY>(cc, mid, half, out)
                                    : brick_column_linear<T, SCRUB, true, LY>(cc, mid, half, out);
      finish_exact(c, rest);
      continue;
    }
#pragma unroll 2
    for (int k = 0; k < nz; ++k, o += out_plane) {
      const float kf = static_cast<float>(k);
      const float uz = __fmaf_rn(kf, mcol[0][0], u0[0]);
      const float uy = __fmaf_rn(kf, mcol[1][0], u0[1]);
      const float ux = __fmaf_rn(kf, mcol[2][0], u0[2]);
      const bool interior = fabsf(uz - mid[0]) <= half[0] - kEdge &&
                            fabsf(uy - mid[1]) <= half[1] - kEdge &&
                            fabsf(ux - mid[2]) <= half[2] - kEdge;
      if (!interior) {
        todo |= 1u << k;
        continue;
      }
      float v;
      if (ORDER == 0) {
        // round to nearest via the magic constant; the tie k+0.5 is re-decided exactly
        const float rz = (uz + kMagic) - kMagic, ry = (uy + kMagic) - kMagic,
                    rx = (ux + kMagic) - kMagic;
        if (fabsf(fabsf(uz - rz) - 0.5f) < kEdge || fabsf(fabsf(uy - ry) - 0.5f) < kEdge ||
            fabsf(fabsf(ux - rx) - 0.5f) < kEdge) {
          todo |= 1u << k;
          continue;
        }
        const uint32_t a = brick + static_cast<uint32_t>(static_cast<int>(rz)) * plane_b +
                           static_cast<uint32_t>(static_cast<int>(ry)) * row_b +
                           static_cast<uint32_t>(static_cast<int>(rx)) * es;
        B2_SMEM_CHECK(a, brick, brick + static_cast<uint32_t>(g.bytes));
        v = brick_elem<T>(a);
        if (SCRUB && sizeof(T) == 4) v = scrub_value(v);
      } else {
        // floor via round-to-nearest of (u - 0.5): an exact integer u may land on u-1 with
        // weight 1, which blends to the same value (both taps are valid in the interior)
        const float tz = (uz - 0.5f) + kMagic, ty = (uy - 0.5f) + kMagic, tx = (ux - 0.5f) + kMagic;
        const float wz = uz - (tz - kMagic), wy = uy - (ty - kMagic), wx = ux - (tx - kMagic);
        const uint32_t a00 =
            brick + static_cast<uint32_t>(__float_as_int(tz) - 0x4B400000) * plane_b +
            static_cast<uint32_t>(__float_as_int(ty) - 0x4B400000) * row_b +
            static_cast<uint32_t>(__float_as_int(tx) - 0x4B400000) * es;
        const uint32_t a01 = a00 + row_b, a10 = a00 + plane_b, a11 = a10 + row_b;
        B2_SMEM_CHECK(a00, brick, brick + static_cast<uint32_t>(g.bytes));
        B2_SMEM_CHECK(a11 + es, brick, brick + static_cast<uint32_t>(g.bytes));
        const float v000 = brick_elem<T>(a00), v001 = brick_elem<T>(a00 + es);
        const float v010 = brick_elem<T>(a01), v011 = brick_elem<T>(a01 + es);
        const float v100 = brick_elem<T>(a10), v101 = brick_elem<T>(a10 + es);
        const float v110 = brick_elem<T>(a11), v111 = brick_elem<T>(a11 + es);
        const float bxw = 1.0f - wx, byw = 1.0f - wy;
        const float r00 = __fmaf_rn(wx, v001, bxw * v000), r01 = __fmaf_rn(wx, v011, bxw * v010);
        const float r10 = __fmaf_rn(wx, v101, bxw * v100), r11 = __fmaf_rn(wx, v111, bxw * v110);
        const float q0 = __fmaf_rn(wy, r01, byw * r00), q1 = __fmaf_rn(wy, r11, byw * r10);
        v = __fmaf_rn(wz, q1, (1.0f - wz) * q0);
        // NaN/inf taps (possibly with zero weight): the exact path applies the scrub per tap
        if (sizeof(T) == 4 && SCRUB && !(fabsf(v) <= FLT_MAX)) {
          todo |= 1u << k;
          continue;
        }
      }
      brick_put<LY>(o, v);
    }
    // ---- voxels near a decision edge / outside the source: exact float64 path
    while (todo) {
      const int k = __ffs(todo) - 1;
      todo &= todo - 1;
      const float kf = static_cast<float>(k);
      const float dz = fabsf(__fmaf_rn(kf, mcol[0][0], u0[0]) - mid[0]);
      const float dy = fabsf(__fmaf_rn(kf, mcol[1][0], u0[1]) - mid[1]);
      const float dx = fabsf(__fmaf_rn(kf, mcol[2][0], u0[2]) - mid[2]);
      const bool outside = dz > half[0] + 0.5f + kEdge || dy > half[1] + 0.5f + kEdge ||
                           dx > half[2] + 0.5f + kEdge;
      const float v = outside ? 0.0f
                              : brick_sample_exact<T, ORDER, BOUNDARY, SCRUB>(
                                    p, brick, b0[0], b0[1], b0[2], g.BZ, g.BY, g.BX, z0 + k, y, x);
      brick_put<LY>(out + k * out_plane, v);
    }
  }
  if (LY) {
    // staged tile -> global memory: lanes along x again (16 columns = 64-byte row segments).
    // Each thread owns 4 consecutive x of one row and every other plane; its global pointer is
    // formed once and advanced by additions (the flat-index version of this loop — two integer
    // divisions, a 64-bit index product and two bounds tests per element — was 16 % of the
    // kernel's instructions).
    __syncthreads();
    static_assert(kBrTX % 4 == 0 && kBrThreads % ((kBrTX / 4) * kBrTY) == 0, "LY copy-out mapping");
    constexpr int kQ = kBrTX / 4;                             // 16-byte pieces per tile row
    constexpr int kPlanesPerPass = kBrThreads / (kQ * kBrTY);  // planes written per pass
    const int q = threadIdx.x % kQ, yy = (threadIdx.x / kQ) % kBrTY, k0 = threadIdx.x / (kQ * kBrTY);
    const int xx = 4 * q;
    const bool row_ok = y0 + yy < p.oy;
    const bool vec_ok = row_ok && x0 + xx + 4 <= p.ox && (p.dpitch & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0;
    float* o = p.dst + static_cast<int64_t>(z0 + k0) * g.out_plane +
               static_cast<int64_t>(y0 + yy) * p.dpitch + x0 + xx;
    const float* st = stage_out + (k0 * kBrTY + yy) * kBrOutPitch + xx;
    for (int k = k0; k < nz; k += kPlanesPerPass, o += kPlanesPerPass * g.out_plane,
             st += kPlanesPerPass * kBrTY * kBrOutPitch) {
      if (vec_ok) {
        st_global_cs4(o, make_float4(st[0], st[1], st[2], st[3]));
      } else if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (x0 + xx + j < p.ox) st_global_cs(o + j, st[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool brick_geometry_tz(const AffineParams& p, bool ly, int tz, int64_t max_bytes, BrickGeom* g,
                              size_t* smem_bytes) {
  const int kBrTY = ly ? kBrLanes : kBrOther, kBrTX = ly ? kBrOther : kBrLanes;
  const int vec = 16 / sizeof(T);
  const int t[3] = {tz - 1, kBrTY - 1, kBrTX - 1};
  int ext[3];
  for (int d = 0; d < 3; ++d) {
    double neg = 0.0, pos = 0.0;
    for (int j = 0; j < 3; ++j) {
      const double v = p.m[4 * d + j] * t[j];
      (v < 0.0 ? neg : pos) += v;
    }
    const double e = pos - neg;
    if (!(e < 240.0)) return false;
    // taps floor(c_min - guard) .. floor(c_max + guard) + 1:  at most int(e + 2 guard) + 3 of them
    ext[d] = static_cast<int>(e + 2.0 * kBrGuard) + 3;
    g->neg[d] = neg;
    g->pos[d] = pos;
  }
  int BX = ext[2] + (vec - 1);  // the brick origin is rounded down to 16 bytes
  BX = (BX + vec - 1) / vec * vec;
  if (ext[0] > 256 || ext[1] > 256 || BX > 256) return false;
  const int64_t bytes = static_cast<int64_t>(ext[0]) * ext[1] * BX * sizeof(T);
  if (bytes > max_bytes) return false;
  g->BZ = ext[0];
  g->BY = ext[1];
  g->BX = BX;
  g->bytes = static_cast<int>(bytes);
  *smem_bytes = static_cast<size_t>(bytes) + 256 + (ly ? kBrStageBytes : 0);
  return true;
}

// float32 matrix columns for the fp32 increments + the |coordinate| bound over the whole output
// (keeps the device-side int conversions defined)
static bool brick_geometry_common(const AffineParams& p, BrickGeom* g) {
  for (int i = 0; i < 9; ++i) g->mcol[i] = static_cast<float>(p.m[4 * (i / 3) + (i % 3)]);
  g->half[0] = 0.5f * static_cast<float>(p.sz - 1);
  g->half[1] = 0.5f * static_cast<float>(p.sy - 1);
  g->half[2] = 0.5f * static_cast<float>(p.sx - 1);
  g->out_plane = static_cast<long long>(p.oy) * p.dpitch;
  g->out_plane_bytes = static_cast<unsigned>(g->out_plane * 4);
  for (int d = 0; d < 3; ++d) {
    const double reach = fabs(p.m[4 * d]) * (p.oz + fabs((double)p.cz)) +
                         fabs(p.m[4 * d + 1]) * (p.oy + fabs((double)p.cy)) +
                         fabs(p.m[4 * d + 2]) * (p.ox + fabs((double)p.cx)) + fabs(p.m[4 * d + 3]);
    if (!(reach < 1.0e9)) return false;
  }
  return true;
}

template <typename T>
static bool brick_geometry(const AffineParams& p, bool ly, BrickGeom* g, size_t* smem_bytes) {
  if (reinterpret_cast<uintptr_t>(p.src) % 16 != 0) return false;
  if ((static_cast<int64_t>(p.spitch) * sizeof(T)) % 16 != 0) return false;
  if (!brick_geometry_common(p, g)) return false;
  // tight 8-deep bricks: typically ~41 KB (4 CTAs/SM, register-limited); up to 72 KB still leaves
  // 3 CTAs/SM; larger footprints (strong out-of-plane rotations or scalings) use the gather path
  return brick_geometry_tz<T>(p, ly, kBrTZ, 72 * 1024, g, smem_bytes);
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB, bool LY>
static int launch_brick(const AffineParams& p, const BrickGeom& g, size_t smem_bytes,
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
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(g.BX), static_cast<cuuint32_t>(g.BY),
                             static_cast<cuuint32_t>(g.BZ)};
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
  constexpr int kBrTY = LY ? kBrLanes : kBrOther, kBrTX = LY ? kBrOther : kBrLanes;
  const int tiles_z = (p.oz + kBrTZ - 1) / kBrTZ;
  const int tiles_y = (p.oy + kBrTY - 1) / kBrTY;
  const int tiles_x = (p.ox + kBrTX - 1) / kBrTX;
  // the kernel forms k * (output plane bytes), k < kBrTZ, in 32 bits
  const bool plane_fits = static_cast<int64_t>(p.oy) * p.dpitch * 4 * kBrTZ < (1LL << 32);
  if (tiles_x > 65535 || tiles_y > 65535 || !plane_fits)
    return affine_gather_launch(p, sizeof(T) == 2 ? B2_DTYPE_U16 : B2_DTYPE_F32, stream);
  auto kern = affine_brick_kernel<T, ORDER, BOUNDARY, SCRUB, LY>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_bytes)));
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  const dim3 grid(static_cast<unsigned>(tiles_z), static_cast<unsigned>(tiles_x),
                  static_cast<unsigned>(tiles_y));
  kern<<<grid, kBrThreads, smem_bytes, stream>>>(map, p, g);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

template <typename T, bool LY>
static int brick_typed_ly(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  BrickGeom g{};
  size_t smem = 0;
  const bool scrub = p.scrub && sizeof(T) == 4;
  *eligible = brick_geometry<T>(p, LY, &g, &smem);
  if (!*eligible) return B2_ERR_UNSUPPORTED;
#define B2_BR(ORD, BND)                                            \
  (scrub ? launch_brick<T, ORD, BND, true, LY>(p, g, smem, stream) \
         : launch_brick<T, ORD, BND, false, LY>(p, g, smem, stream))
  if (p.order == 0)
    return p.boundary == B2_BOUNDARY_CONSTANT ? B2_BR(0, B2_BOUNDARY_CONSTANT)
                                              : B2_BR(0, B2_BOUNDARY_ITK);
  return p.boundary == B2_BOUNDARY_CONSTANT ? B2_BR(1, B2_BOUNDARY_CONSTANT)
                                            : B2_BR(1, B2_BOUNDARY_ITK);
#undef B2_BR
}

template <typename T>
static int brick_typed(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  // lanes follow the output axis along which the SOURCE x coordinate moves fastest
  if (fabs(p.m[9]) > fabs(p.m[10])) {
    const int rc = brick_typed_ly<T, true>(p, stream, eligible);
    if (*eligible) return rc;
  }
  return brick_typed_ly<T, false>(p, stream, eligible);
}

B2_OOB_GETTER(brick_oob_count)

int affine_brick_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible) {
  if (src_dtype == B2_DTYPE_U16) return brick_typed<uint16_t>(p, stream, eligible);
  return brick_typed<float>(p, stream, eligible);
}

}  // namespace b2
