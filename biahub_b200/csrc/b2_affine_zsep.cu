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
// consumer threads hold, per output point, ONE brick offset and FOUR bilinear weights in
// registers (computed once, in float64, in the oracle's op order), reduce each arriving plane
// to one in-plane-interpolated value per point (4 LDS + 4 FMA), and blend consecutive planes
// along z (2 FMA).  The per-z tap table (plane indices + weights) is computed once per CTA into
// shared memory.  Every source plane brick is read from global memory exactly once per CTA and
// every output voxel is written once, coalesced along x.
//
// m00 == 1 (all of the reference's builder matrices and stabilisation shifts) takes the "regular
// run": each arriving plane finishes one output plane (o = fma(w1, v, pend)) and starts the next
// (pend = w0 * v) — same operations and order as the general two-plane blend, bit-identical, 25 %
// fewer instructions.  The ring is 6-8 planes deep (chosen per launch from the plane brick size).
#include "b2_affine.cuh"

namespace b2 {

// Output tile: 16 x 64 points (y x x), lanes along x.  LY variant (matrices that map output y
// onto source x, e.g. the 90-degree rotations of the manual registration): 64 x 16, lanes along
// y, so that a warp's taps run along a brick ROW (conflict-free shared-memory reads; with lanes
// along x they walk a brick COLUMN: 8-way bank conflicts, 0.33 of the roofline); the output
// plane is transposed through shared memory to keep the global stores coalesced.
constexpr int kZsTileLanes = 64;  // tile extent along the lane axis
constexpr int kZsTileOther = 16;  // tile extent along the other in-plane axis
constexpr int kZsPPT = (kZsTileLanes * kZsTileOther) / 256;  // output points per consumer thread
constexpr int kZsOutPitch = kZsTileOther + 1;  // LY: padded pitch of the transposed output tile
constexpr int kZsConsumers = 256;
#ifndef B2_ZSEP_LY_GROUP
#define B2_ZSEP_LY_GROUP 8
#endif
constexpr int kZsLyGroup = B2_ZSEP_LY_GROUP;  // LY rasterisation: tile columns per group (1 = x-fastest)
constexpr int kZsThreads = kZsConsumers + 32;  // + one producer warp
constexpr int kZsMaxStages = 8;  // ring depth is chosen per launch (ZsepGeom::stages, 3..8)
constexpr int kZsRowsPerPass = kZsConsumers / kZsTileLanes;
constexpr int kZsMaxChunk = 128;  // output planes per CTA (size of the z tap table)

struct ZsepGeom {
  int BY, BX;       // brick extent per plane (elements)
  int stage_bytes;  // BY*BX*sizeof(T) rounded up to 128
  int zchunk;       // output planes per CTA (<= kZsMaxChunk)
  int unit_z;       // m00 == 1 exactly: source planes advance one per output plane
  int stages;       // depth of the plane ring (3..kZsMaxStages)
};

template <typename T>
__device__ __forceinline__ float lds_elem(uint32_t addr);
template <>
__device__ __forceinline__ float lds_elem<float>(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <>
__device__ __forceinline__ float lds_elem<uint16_t>(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return static_cast<float>(v);
}

__device__ __forceinline__ double coord_yx(double yf, double xf, const double* m) {
  // ((t_d + z*m_d0) + y*m_d1) + x*m_d2 with m_d0 == 0 (t_d + z*0 == t_d exactly)
  return __dadd_rn(__dadd_rn(m[3], __dmul_rn(yf, m[1])), __dmul_rn(xf, m[2]));
}

// Rare path: the host-side bound on the brick size was too tight for this tile (never expected):
// every thread of the CTA resamples the tile straight from global memory.
template <typename T, int ORDER, int BOUNDARY, bool SCRUB, bool LY>
__device__ __noinline__ void zsep_tile_from_global(const AffineParams& p, int y0, int x0, int zb,
                                                   int ze) {
  constexpr int kZsTY = LY ? kZsTileLanes : kZsTileOther, kZsTX = LY ? kZsTileOther : kZsTileLanes;
  const int n = (ze - zb) * kZsTY * kZsTX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = x0 + i % kZsTX;
    const int y = y0 + (i / kZsTX) % kZsTY;
    const int z = zb + i / (kZsTX * kZsTY);
    if (y < p.oy && x < p.ox)
      p.dst[(static_cast<int64_t>(z) * p.oy + y) * p.dpitch + x] =
          affine_sample_generic<T, ORDER, BOUNDARY, SCRUB>(p, z, y, x);
  }
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB, bool LY>
__global__ void __launch_bounds__(kZsThreads, 3)
    affine_zsep_kernel(const __grid_constant__ CUtensorMap src_map,
                       const __grid_constant__ AffineParams p, const ZsepGeom g) {
  constexpr int kZsTY = LY ? kZsTileLanes : kZsTileOther, kZsTX = LY ? kZsTileOther : kZsTileLanes;
  extern __shared__ uint8_t smem_raw[];

  constexpr int kVec = 16 / static_cast<int>(sizeof(T));
  // dynamic shared memory: [full barriers | empty barriers | pad to 128 B] [z tap table]
  // [stage ring]; the z tap table holds per output plane of this CTA
  // {i0 (or -1 when outside), i1, bits(w0), bits(w1)}; all
  // addresses derive from ONE register (a static __shared__ barrier array makes the compiler
  // re-derive the shared-window address with S2R/LEA inside the plane loop)
  const uint32_t bars = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full0 = bars;
  const uint32_t empty0 = bars + 8u * kZsMaxStages;
  const uint32_t ztab0 = bars + 128u;
  // LY: two transposed output tiles (double-buffered: one consumer barrier per plane)
  constexpr uint32_t kOutTileBytes = LY ? kZsTileLanes * kZsOutPitch * 4u : 0u;
  const uint32_t otile0 = ztab0 + 16u * kZsMaxChunk;
  const uint32_t stage0 = (otile0 + 2u * kOutTileBytes + 127u) & ~127u;
  auto ztab_at = [&](int zl) {
    int4 e;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                 : "r"(ztab0 + 16u * static_cast<uint32_t>(zl)));
    return e;
  };
  const int tid = threadIdx.x;
  // LY: the CTAs of a wave are rasterised in groups of kZsLyGroup tile columns (x) by all tile
  // rows (y).  Their source bricks then line up along source x (= output y) — contiguous segments
  // of the same DRAM rows — while their output rows still form 16 * kZsLyGroup * 4-byte runs;
  // with plain x-fastest order every concurrent brick row opens another DRAM page.
  int tile_x = blockIdx.x, tile_y = blockIdx.y;
  if (LY && kZsLyGroup > 1) {
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    const int per_group = kZsLyGroup * gridDim.y;
    const int grp = lin / per_group, within = lin - grp * per_group;
    const int gsize = min(kZsLyGroup, static_cast<int>(gridDim.x) - grp * kZsLyGroup);
    tile_y = within / gsize;
    tile_x = grp * kZsLyGroup + (within - tile_y * gsize);
  }
  const int y0 = tile_y * kZsTY;
  const int x0 = tile_x * kZsTX;
  const int zb = blockIdx.z * g.zchunk;
  const int ze = min(zb + g.zchunk, p.oz);
  const int nz = ze - zb;

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
  // Clamp far-away tiles so the int conversion is defined (such tiles are entirely outside).
  // The innermost TMA coordinate must be 16-byte aligned (measured on B200 / driver 580: an
  // unaligned inner coordinate raises "illegal instruction"), so the brick starts at the 16-byte
  // boundary at or below the back-projected minimum.  Taps are always (i0, i0+1) with the clamped
  // neighbour's weight forced to 0, so the brick must reach floor(max)+2 when i0 was clamped up.
  const double big = 1.0e9;
  const int by0 = __double2int_rd(fmax(-big, fmin(big, cy_min)));
  const int bx0 = __double2int_rd(fmax(-big, fmin(big, cx_min))) & ~(kVec - 1);
  const int by_hi = __double2int_rd(fmax(-big, fmin(big, cy_max))) + 2;
  const int bx_hi = __double2int_rd(fmax(-big, fmin(big, cx_max))) + 2;
  const bool brick_ok = (by_hi - by0) < g.BY && (bx_hi - bx0) < g.BX;  // CTA-uniform
  if (!brick_ok) {
    zsep_tile_from_global<T, ORDER, BOUNDARY, SCRUB, LY>(p, y0, x0, zb, ze);
    return;
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < g.stages; ++s) {
      mbar_init_u32(full0 + 8u * s, 1);
      mbar_init_u32(empty0 + 8u * s, kZsConsumers / 32);
    }
    fence_mbar_init();
  }
  if (tid < nz) {
    const double cz = __dadd_rn(p.m[3], __dmul_rn(static_cast<double>(zb + tid + p.cz), p.m[0]));
    const AxisTap tz = resolve_axis<ORDER, BOUNDARY>(cz, p.sz);
    int4 e;
    e.x = tz.inside ? tz.i0 : -1;
    e.y = tz.i1;
    e.z = __float_as_int(__fsub_rn(1.0f, tz.w));
    e.w = __float_as_int(tz.w);
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(ztab0 + 16u * static_cast<uint32_t>(tid)),
                 "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w)
                 : "memory");
  }
  __syncthreads();

  if (tid >= kZsConsumers) {
    // ============ producer warp: walks the plane sequence, lane 0 issues the TMA loads ============
    const bool issuer = (tid == kZsConsumers);
    int s_last = INT_MIN;
    uint32_t stage = 0, phase = 0;  // ring position of the next plane, parity of its round
    bool first_round = true;
    for (int zl = 0; zl < nz; ++zl) {
      const int4 e = ztab_at(zl);
      if (e.x < 0) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = h ? e.y : e.x;
        if (s > s_last) {
          if (!first_round) mbar_wait_u32(empty0 + 8u * stage, phase ^ 1u);
          if (issuer) {
            mbar_expect_tx_u32(full0 + 8u * stage, static_cast<uint32_t>(g.BY) * g.BX * sizeof(T));
            tma_load_3d_u32(stage0 + stage * g.stage_bytes, &src_map, full0 + 8u * stage, bx0, by0, s);
          }
          __syncwarp();
          s_last = s;
          if (++stage == static_cast<uint32_t>(g.stages)) {
            stage = 0;
            phase ^= 1u;
            first_round = false;
          }
        }
      }
    }
    return;
  }

  // ================================= consumer threads =================================
  const int ll = tid % kZsTileLanes;  // position along the lane axis (x, or y when LY)
  const int lo = tid / kZsTileLanes;  // first position along the other axis
  const uint32_t pitch = static_cast<uint32_t>(g.BX) * sizeof(T);

  uint32_t off[kZsPPT];  // byte offset of tap (i0y, i0x) inside a stage
  float w00[kZsPPT], w01[kZsPPT], w10[kZsPPT], w11[kZsPPT];
  int ooff[kZsPPT];  // output offset relative to the tile origin, -1 = no voxel
  uint32_t inmask = 0;  // bit i: point i samples inside the source
#pragma unroll
  for (int i = 0; i < kZsPPT; ++i) {
    const int oo = lo + kZsRowsPerPass * i;
    const int yy = LY ? ll : oo, xx = LY ? oo : ll;
    const int y = y0 + yy, x = x0 + xx;
    const bool live = (y < p.oy) && (x < p.ox);
    const double yf = static_cast<double>(y + p.cy);
    const double xf = static_cast<double>(x + p.cx);
    const AxisTap ty = resolve_axis<ORDER, BOUNDARY>(coord_yx(yf, xf, p.m + 4), p.sy);
    const AxisTap tx = resolve_axis<ORDER, BOUNDARY>(coord_yx(yf, xf, p.m + 8), p.sx);
    const bool in = live && ty.inside && tx.inside;
    // !LY: global offset of the point relative to the tile origin (-1 = no voxel);
    //  LY: byte offset of the point in the transposed output tile in shared memory
    ooff[i] = LY ? (yy * kZsOutPitch + xx) * 4 : (live ? yy * p.dpitch + xx : -1);
    inmask |= in ? (1u << i) : 0u;
    off[i] = in ? static_cast<uint32_t>(ty.i0 - by0) * pitch +
                      static_cast<uint32_t>(tx.i0 - bx0) * static_cast<uint32_t>(sizeof(T))
                : 0u;
    // bilinear weights; a dropped / clamped neighbour has weight 0 (resolve_axis), so the
    // neighbour taps can always be read at +1 element / +1 row
    const float ay = ty.w, ax = tx.w;
    const float by = __fsub_rn(1.0f, ay), bx = __fsub_rn(1.0f, ax);
    w00[i] = in ? __fmul_rn(by, bx) : 0.0f;
    w01[i] = in ? __fmul_rn(by, ax) : 0.0f;
    w10[i] = in ? __fmul_rn(ay, bx) : 0.0f;
    w11[i] = in ? __fmul_rn(ay, ax) : 0.0f;
  }

  float p_prev[kZsPPT], p_last[kZsPPT];
#pragma unroll
  for (int i = 0; i < kZsPPT; ++i) p_prev[i] = p_last[i] = 0.0f;
  int s_last = INT_MIN;
  uint32_t stage = 0, phase = 0;  // ring position of the next plane to consume, parity of its round
  const int64_t plane_out = static_cast<int64_t>(p.oy) * p.dpitch;
  float* __restrict__ out_tile =
      p.dst + static_cast<int64_t>(zb) * plane_out + static_cast<int64_t>(y0) * p.dpitch + x0;
  const bool full_tile = (y0 + kZsTY <= p.oy) && (x0 + kZsTX <= p.ox);  // CTA-uniform

  const bool lane0 = (tid & 31) == 0;
  // reduce the next plane of the producer's sequence to one in-plane-interpolated value per point
  auto load_plane = [&](float (&v)[kZsPPT]) {
    mbar_wait_u32(full0 + stage * 8u, phase);
    const uint32_t base = stage0 + stage * g.stage_bytes;
    const uint32_t base1 = base + pitch;
    // every lane's taps have been consumed when the warp-wide vote below completes: it doubles
    // as the convergence point before lane 0 releases the stage (a __syncwarp() here makes ptxas
    // treat the loop as divergent and re-materialise its uniform registers per plane)
    bool bad = false;
    if (ORDER == 0) {
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i) {
        float t = lds_elem<T>(base + off[i]);
        if (SCRUB && sizeof(T) == 4) t = scrub_value(t);
        v[i] = ((inmask >> i) & 1u) ? t : 0.0f;
        bad |= (__float_as_uint(v[i]) == 0xffffffffu);  // never true: ties the vote to the loads
      }
    } else {
      // groups of 4 points: 16 taps in flight, then 4 x (FMUL + 3 FFMA)
#pragma unroll
      for (int i0 = 0; i0 < kZsPPT; i0 += 4) {
        float t00[4], t01[4], t10[4], t11[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t00[j] = lds_elem<T>(base + off[i0 + j]);
          t01[j] = lds_elem<T>(base + off[i0 + j] + sizeof(T));
          t10[j] = lds_elem<T>(base1 + off[i0 + j]);
          t11[j] = lds_elem<T>(base1 + off[i0 + j] + sizeof(T));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          v[i] = __fmaf_rn(w11[i], t11[j],
                           __fmaf_rn(w10[i], t10[j], __fmaf_rn(w01[i], t01[j], __fmul_rn(w00[i], t00[j]))));
          bad |= !(fabsf(v[i]) <= FLT_MAX);
        }
      }
    }
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (ORDER == 1 && sizeof(T) == 4 && any_bad) {
      // a NaN/inf tap was involved (possibly with zero weight): redo with the scrub
      // (np.nan_to_num semantics), re-reading the taps; without SCRUB only the dummy taps of
      // outside points are fixed.  Warp-uniform branch; the stage is released after it.
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i) {
        if (SCRUB) {
          const uint32_t a0 = base + off[i];
          const uint32_t a1 = a0 + pitch;
          v[i] = __fmaf_rn(w11[i], scrub_value(lds_elem<T>(a1 + sizeof(T))),
                           __fmaf_rn(w10[i], scrub_value(lds_elem<T>(a1)),
                                     __fmaf_rn(w01[i], scrub_value(lds_elem<T>(a0 + sizeof(T))),
                                               __fmul_rn(w00[i], scrub_value(lds_elem<T>(a0))))));
        } else if (!((inmask >> i) & 1u)) {
          v[i] = 0.0f;
        }
      }
      __syncwarp();
    }
    if (lane0) mbar_arrive_u32(empty0 + stage * 8u);
    if (++stage == static_cast<uint32_t>(g.stages)) {
      stage = 0;
      phase ^= 1u;
    }
  };
  auto fetch_plane = [&]() {  // general path: keep the last two planes
    float v[kZsPPT];
    load_plane(v);
#pragma unroll
    for (int i = 0; i < kZsPPT; ++i) {
      p_prev[i] = p_last[i];
      p_last[i] = v[i];
    }
  };
  // LY store phase: thread t writes column t % 16 of the rows t / 16 + 16 j (64-byte row segments);
  // the element offsets relative to the tile origin are formed once (-1 = no voxel), not per plane
  uint32_t oparity = 0;
  constexpr int kStoreRows = kZsConsumers / kZsTileOther;   // rows written per pass
  constexpr int kStoreIters = kZsTileLanes / kStoreRows;
  const int sc = tid % kZsTileOther, sr = tid / kZsTileOther;
  int soff[kStoreIters];
#pragma unroll
  for (int j = 0; j < kStoreIters; ++j) {
    const int r = sr + j * kStoreRows;
    soff[j] = (LY && x0 + sc < p.ox && y0 + r < p.oy) ? r * p.dpitch + sc : -1;
  }
  const uint32_t sld = static_cast<uint32_t>((sr * kZsOutPitch + sc) * 4);
  auto store_plane = [&](const float (&o)[kZsPPT]) {
    if (LY) {
      // transpose through shared memory: lanes run along y here, the global rows run along x
      const uint32_t buf = otile0 + oparity * kOutTileBytes;
      oparity ^= 1u;
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(buf + static_cast<uint32_t>(ooff[i])), "f"(o[i]));
      asm volatile("bar.sync 1, %0;" ::"n"(kZsConsumers) : "memory");  // consumers only
#pragma unroll
      for (int j = 0; j < kStoreIters; ++j) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];"
                     : "=f"(v)
                     : "r"(buf + sld + static_cast<uint32_t>(j * kStoreRows * kZsOutPitch * 4)));
        if (full_tile || soff[j] >= 0) st_global_cs(out_tile + soff[j], v);
      }
      return;
    }
    if (full_tile) {
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i) st_global_cs(out_tile + ooff[i], o[i]);
    } else {
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i)
        if (ooff[i] >= 0) st_global_cs(out_tile + ooff[i], o[i]);
    }
  };

  for (int zl = 0; zl < nz; ++zl, out_tile += plane_out) {
    const int4 e = ztab_at(zl);
    float o[kZsPPT];
    if (ORDER == 1 && g.unit_z && e.x == s_last && e.y == e.x + 1) {
      // ---- regular run (m00 == 1: every output plane takes the NEXT source plane as its upper
      // tap and hands it on as the lower tap of the following one).  Per arriving plane and
      // point: 4 LDS + 4 FMA in-plane, o = fma(w1, v, pend), pend' = w0' * v — the same
      // operations in the same order as the general path, without its two-plane register
      // rotation, tap-table branches and per-plane bookkeeping.
      float pend[kZsPPT];
      {
        const float wz0 = __int_as_float(e.z);
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i) pend[i] = __fmul_rn(wz0, p_last[i]);
      }
      float wz1 = __int_as_float(e.w);
      int s = s_last;
      for (;;) {
        float v[kZsPPT];
        load_plane(v);
        ++s;
        const bool more = zl + 1 < nz;
        const int4 en = ztab_at(more ? zl + 1 : zl);
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i) o[i] = __fmaf_rn(wz1, v[i], pend[i]);
        store_plane(o);
        if (!(more && en.x == s && en.y == s + 1)) {  // CTA-uniform
#pragma unroll
          for (int i = 0; i < kZsPPT; ++i) p_prev[i] = p_last[i] = v[i];
          break;
        }
        const float wn0 = __int_as_float(en.z);
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i) pend[i] = __fmul_rn(wn0, v[i]);
        wz1 = __int_as_float(en.w);
        ++zl;
        out_tile += plane_out;
      }
      s_last = s;
      continue;
    }
    if (e.x < 0) {  // this output plane maps outside the source: zeros
#pragma unroll
      for (int i = 0; i < kZsPPT; ++i) o[i] = 0.0f;
    } else {
      if (e.x > s_last) {  // CTA-uniform
        fetch_plane();
        s_last = e.x;
      }
      if (e.y > s_last) {
        fetch_plane();
        s_last = e.y;
      }
      if (ORDER == 0 || e.x == s_last) {  // single plane (nearest, or the +1 plane was clamped away)
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i) o[i] = p_last[i];
      } else {
        const float wz0 = __int_as_float(e.z), wz1 = __int_as_float(e.w);
#pragma unroll
        for (int i = 0; i < kZsPPT; ++i)
          o[i] = __fmaf_rn(wz1, p_last[i], __fmul_rn(wz0, p_prev[i]));
      }
    }
    store_plane(o);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool zsep_geometry(const AffineParams& p, bool ly, ZsepGeom* g, size_t* smem_bytes) {
  const int kZsTY = ly ? kZsTileLanes : kZsTileOther, kZsTX = ly ? kZsTileOther : kZsTileLanes;
  const double* m = p.m;
  if (m[1] != 0.0 || m[2] != 0.0 || m[4] != 0.0 || m[8] != 0.0) return false;
  if (!(m[0] > 0.0)) return false;
  if (reinterpret_cast<uintptr_t>(p.src) % 16 != 0) return false;
  if ((static_cast<int64_t>(p.spitch) * sizeof(T)) % 16 != 0) return false;
  const double ey = fabs(m[5]) * (kZsTY - 1) + fabs(m[6]) * (kZsTX - 1);
  const double ex = fabs(m[9]) * (kZsTY - 1) + fabs(m[10]) * (kZsTX - 1);
  if (!(ey < 240.0) || !(ex < 240.0)) return false;
  const int vec = 16 / sizeof(T);
  // rows by0 .. floor(max) + 2 with by0 = floor(min): at most floor(e) + 3 of them, and the kernel
  // asks for (by_hi - by0) < BY
  const int BY = static_cast<int>(ey) + 4;
  int BX = static_cast<int>(ex) + 4 + (vec - 1);  // + alignment slack of the brick origin
  BX = (BX + vec - 1) / vec * vec;
  // (a bank-aligned row pitch was measured and makes no difference: the lanes of a warp span ~34
  // columns under a 1.07x scale, so every LDS costs 2 wavefronts either way and the kernel is not
  // bound by shared-memory bandwidth)
  auto stage_of = [&](int bx) { return (BY * bx * static_cast<int>(sizeof(T)) + 127) / 128 * 128; };
  if (BY > 256 || BX > 256) return false;
  const int stage = stage_of(BX);
  // ring depth: as deep as 70 KB per CTA allow (3 CTAs per SM), 8 at most; the consumers run at
  // ~0.7 us per plane and a TMA round trip is ~1.5 us, 4 stages left them waiting (ncu: long
  // scoreboard 3.7 per issue)
  int stages = (70 * 1024) / stage;
  if (stages > kZsMaxStages) stages = kZsMaxStages;
  // large bricks (rotation / scale halos): a deeper ring lets neighbouring CTAs drift apart in z
  // and their shared halo rows fall out of L2 (C3: DRAM reads 2.24 -> 2.61 GB at 8 stages);
  // measured best: 6 for C3's 8.8 KB planes, 8 for C4's 4.9 KB planes
  if (stage > 6 * 1024 && stages > 6) stages = 6;
  {
    const char* e = getenv("B2_ZSEP_STAGES");  // sweep override
    const int v = e ? atoi(e) : 0;
    if (v >= 3 && v < stages) stages = v;
  }
  if (stages < 3) return false;
  g->stages = stages;
  if (static_cast<int64_t>(kZsTY) * p.dpitch >= (1LL << 31)) return false;
  g->BY = BY;
  g->BX = BX;
  g->stage_bytes = stage;
  g->unit_z = (m[0] == 1.0) ? 1 : 0;
  *smem_bytes = static_cast<size_t>(stage) * stages + 384 + 16 * kZsMaxChunk +
                (ly ? 2 * kZsTileLanes * kZsOutPitch * 4 : 0);  // + barriers, z table, alignment
  return true;
}

template <typename T, int ORDER, int BOUNDARY, bool SCRUB, bool LY>
static int launch_zsep(const AffineParams& p, ZsepGeom g, size_t smem_bytes, cudaStream_t stream) {
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
  int sms = 148;
  sm_count(&sms);
  constexpr int kZsTY = LY ? kZsTileLanes : kZsTileOther, kZsTX = LY ? kZsTileOther : kZsTileLanes;
  const int tiles_x = (p.ox + kZsTX - 1) / kZsTX;
  const int tiles_y = (p.oy + kZsTY - 1) / kZsTY;
  const int64_t tiles = static_cast<int64_t>(tiles_x) * tiles_y;
  const int64_t target = static_cast<int64_t>(sms) * 16;  // enough CTAs for a few waves
  int zsplit = 1;
  if (tiles < target) zsplit = static_cast<int>((target + tiles - 1) / tiles);
  int zchunk = (p.oz + zsplit - 1) / zsplit;
  if (zchunk < 8) zchunk = 8;
  if (zchunk > kZsMaxChunk) zchunk = kZsMaxChunk;
  if (zchunk > p.oz) zchunk = p.oz;
  g.zchunk = zchunk;
  const int grid_z = (p.oz + zchunk - 1) / zchunk;
  if (tiles_y > 65535 || grid_z > 65535)
    return affine_gather_launch(p, sizeof(T) == 2 ? B2_DTYPE_U16 : B2_DTYPE_F32, stream);

  auto kern = affine_zsep_kernel<T, ORDER, BOUNDARY, SCRUB, LY>;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem_bytes)));
  const dim3 grid(tiles_x, tiles_y, grid_z);
  kern<<<grid, kZsThreads, smem_bytes, stream>>>(map, p, g);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

// identity linear part with an integer translation: linear interpolation degenerates to a
// shifted copy, which the nearest-neighbour instantiation performs bit-exactly with 1 tap
static bool is_integer_translation(const AffineParams& p) {
  const double* m = p.m;
  const bool ident = m[0] == 1.0 && m[1] == 0.0 && m[2] == 0.0 && m[4] == 0.0 && m[5] == 1.0 &&
                     m[6] == 0.0 && m[8] == 0.0 && m[9] == 0.0 && m[10] == 1.0;
  if (!ident) return false;
  for (int d = 0; d < 3; ++d) {
    const double t = m[4 * d + 3];
    if (fabs(t) > 1e9 || t != static_cast<double>(static_cast<long long>(t))) return false;
  }
  return true;
}

template <typename T, bool LY>
static int zsep_typed_ly(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  ZsepGeom g{};
  size_t smem = 0;
  *eligible = zsep_geometry<T>(p, LY, &g, &smem);
  if (!*eligible) return B2_ERR_UNSUPPORTED;
  const bool scrub = p.scrub && sizeof(T) == 4;
  const int order = (p.order == 1 && is_integer_translation(p)) ? 0 : p.order;
#define B2_ZS(ORD, BND)                                           \
  (scrub ? launch_zsep<T, ORD, BND, true, LY>(p, g, smem, stream) \
         : launch_zsep<T, ORD, BND, false, LY>(p, g, smem, stream))
  if (order == 0)
    return p.boundary == B2_BOUNDARY_CONSTANT ? B2_ZS(0, B2_BOUNDARY_CONSTANT)
                                              : B2_ZS(0, B2_BOUNDARY_ITK);
  return p.boundary == B2_BOUNDARY_CONSTANT ? B2_ZS(1, B2_BOUNDARY_CONSTANT)
                                            : B2_ZS(1, B2_BOUNDARY_ITK);
#undef B2_ZS
}

template <typename T>
static int zsep_typed(const AffineParams& p, cudaStream_t stream, bool* eligible) {
  // lanes follow the output axis along which the SOURCE x coordinate moves fastest
  // (m[9] = d src_x / d out_y, m[10] = d src_x / d out_x)
  static const int force = [] {
    const char* e = getenv("B2_ZSEP_LANES_Y");  // A/B switch: 0 / 1, unset = automatic
    return e ? atoi(e) : -1;
  }();
  const bool ly = force >= 0 ? force != 0 : fabs(p.m[9]) > fabs(p.m[10]);
  if (ly) {
    const int rc = zsep_typed_ly<T, true>(p, stream, eligible);
    if (*eligible) return rc;
  }
  return zsep_typed_ly<T, false>(p, stream, eligible);
}

int affine_zsep_launch(const AffineParams& p, int src_dtype, cudaStream_t stream, bool* eligible) {
  if (src_dtype == B2_DTYPE_U16) return zsep_typed<uint16_t>(p, stream, eligible);
  return zsep_typed<float>(p, stream, eligible);
}

}  // namespace b2
