// Flat-field correction for sm_100a: the pipeline stage BEFORE deskew in the mantis workflow
// (reference biahub/flat_field.py:105-122, 152-166).
//
//   pattern[y, x] = median_z src[z, y, x]                 (np.median: mean of the two middle samples)
//   out[z, y, x]  = (double(src[z, y, x]) / pattern[y, x]) * mean(pattern)     -> float32 | float64
//
// Data layout in HBM: src (Z, Y, X) uint16 C-contiguous, dst (Z, Y, X) float32|float64, pattern
// (Y*X) float32 (medians of uint16 samples are multiples of 0.5 below 65536: exact in fp32) and
// one uint64 accumulator of sum(2 * pattern) in the caller's workspace.
//
// Two HBM-bound kernels:
//  * flatfield_median_kernel: one thread owns 2 adjacent pixels (4-byte coalesced loads along x)
//    and finds the middle samples of each by a 2 x 8-bit radix select: a 256-bin histogram of the
//    high bytes (uint16 counters in shared memory, 1 KB per thread, bank-conflict free), then one
//    of the low bytes of the samples in the selected bin; the same sweep keeps the smallest
//    sample above that bin, so the upper middle sample of an even Z needs no further pass.
//    Two sweeps over the column: 2 * Z*Y*X*2 bytes, the second mostly from L2.  Integer work, exact.
//  * flatfield_apply_kernel: exact float64 quotient without the division subroutine: r = RN(1/p)
//    once per pixel, then q0 = v*r, rem = fma(-p, q0, v), q = fma(rem, r, q0) — the correctly
//    rounded quotient for every uint16 v and every half-integer p (checked exhaustively against
//    __ddiv_rn by scripts/flatfield_div_check.cu) — times the mean, rounded once more to float32.
//    p == 0 (a column of zeros) takes the IEEE division (inf / nan as numpy).
// mean(pattern): every partial sum of half-integers below 2^52 is exact, so the integer sum equals
// numpy's pairwise float64 sum bit for bit (tests/test_flatfield_oracle.py).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "b2_common.cuh"

namespace b2 {

constexpr int kFfThreads = 256;
constexpr int kFfPix = 4;  // pixels per thread (one 8-byte load per plane)
constexpr int kFfBatch = 4;      // apply kernel: planes loaded before the first quotient
constexpr int kFfApplyPlanes = 16;  // apply kernel: planes per CTA (the reciprocals are per CTA)

struct FlatfieldParams {
  const uint16_t* src;
  void* dst;
  float* pattern;              // [P] medians
  unsigned long long* sum2;    // sum over pixels of 2 * median
  int64_t P;                   // pixels per plane (Y * X), plane stride of src and dst
  int64_t p0, pn;              // pixel window [p0, p0 + pn) handled by this launch
  int Z;
  int z0, zn;                  // plane window of the apply kernel
};

// ---------------------------------------------------------------------------------------------
// median: two sweeps over the column (high byte, low byte), 256-bin histograms in shared memory
// ---------------------------------------------------------------------------------------------
#ifndef B2_FM_THREADS
#define B2_FM_THREADS 64
#endif
#ifndef B2_FM_RING
#define B2_FM_RING 32
#endif
#ifndef B2_FM_GROUP
#define B2_FM_GROUP 8
#endif
constexpr int kFmThreads = B2_FM_THREADS;  // threads per CTA
constexpr int kFmPix = 2;       // pixels per thread: one 4-byte load per plane
// counter of (bin, pixel j) of thread t: the 16-bit half j of the 32-bit word [bin][t] — a warp's
// 32 threads always hit 32 different banks, whatever bins their samples fall into.  Bin 256 takes
// the samples the second sweep discards (no predicated shared-memory accesses in the loop).
constexpr uint32_t kFmBinStride = kFmThreads * 4u;
constexpr int kFmRing = B2_FM_RING;     // planes of the load ring (4 bytes per thread and plane)
constexpr int kFmGroup = B2_FM_GROUP;     // planes per cp.async group
constexpr int kFmHistBytes = 257 * kFmThreads * 4;
constexpr int kFmSmemBytes = kFmHistBytes + kFmRing * kFmThreads * 4;  // 72.25 KB: three CTAs per SM

__device__ __forceinline__ uint32_t fm_lds16(uint32_t a) {
  unsigned short c;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c) : "r"(a));
  return c;
}
__device__ __forceinline__ uint32_t fm_lds32(uint32_t a) {
  uint32_t c;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(a));
  return c;
}
__device__ __forceinline__ void fm_sts16(uint32_t a, uint32_t c) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(static_cast<unsigned short>(c)));
}

// the thread's two pixels of one plane packed in 32 bits (pixel 0 in the low half).  asm volatile:
// the compiler must keep these loads where the source puts them, a whole batch ahead of their use
// (as plain __ldg it sinks them below the shared-memory code of the previous batch to save
// registers, and every batch then waits for DRAM: measured 35 % of the kernel)
__device__ __forceinline__ uint32_t fm_ldg32(const uint16_t* q) {
  uint32_t w;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(w) : "l"(q));
  return w;
}
__device__ __forceinline__ uint32_t fm_ldg16(const uint16_t* q) {
  unsigned short w;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(w) : "l"(q));
  return w;
}
template <bool VEC>
__device__ __forceinline__ uint32_t fm_load(const uint16_t* q, bool two) {
  if (VEC) return fm_ldg32(q);
  const uint32_t lo = fm_ldg16(q);
  return lo | ((two ? fm_ldg16(q + 1) : lo) << 16);
}

// One sweep over the Z samples of the thread's two pixels.
//  SECOND == false: histogram of the high bytes of all samples.
//  SECOND == true : histogram of the low bytes of the samples whose high byte is the one the
//                   first sweep selected (pre[j] = that byte << 8), the others go to bin 256;
//                   TRACK also keeps min(sample - (pre + 256)) in unsigned arithmetic, i.e. the
//                   smallest sample above the selected bin (the upper middle sample of an even Z
//                   when it lies outside the bin).
// Two consecutive planes are counted together: four independent shared-memory round trips per
// thread instead of two, and when both samples of a pixel fall into one bin each store writes
// count + 2 (the counter updates of one pixel are otherwise an ordered chain, 29 cycles per LDS).
template <bool VEC, bool SECOND, bool TRACK>
__device__ __forceinline__ void fm_sweep(const uint16_t* __restrict__ q, int64_t plane, int Z, bool two,
                                         uint32_t hb, const uint32_t (&pre)[kFmPix],
                                         uint32_t (&above)[kFmPix], uint32_t ring) {
  auto slot = [&](uint32_t w, int j) -> uint32_t {
    if (!SECOND) {
      const uint32_t bin = j == 0 ? ((w >> 8) & 0xffu) : (w >> 24);
      return hb + 2u * j + bin * kFmBinStride;
    }
    const uint32_t v = j == 0 ? (w & 0xffffu) : (w >> 16);
    if (TRACK) above[j] = min(above[j], v - (pre[j] + 256u));
    // v ^ pre < 256 exactly when the high byte matches, and is the low byte then
    return hb + 2u * j + min(v ^ pre[j], 256u) * kFmBinStride;
  };
  auto bump2 = [&](uint32_t wa, uint32_t wb) {
    uint32_t a[kFmPix], b[kFmPix], ca[kFmPix], cb[kFmPix];
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) {
      a[j] = slot(wa, j);
      b[j] = slot(wb, j);
    }
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) {
      ca[j] = fm_lds16(a[j]);
      cb[j] = fm_lds16(b[j]);
    }
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) {
      const uint32_t inc = a[j] == b[j] ? 2u : 1u;
      fm_sts16(a[j], ca[j] + inc);
      fm_sts16(b[j], cb[j] + inc);
    }
  };
  auto bump1 = [&](uint32_t w) {
    uint32_t a[kFmPix], c[kFmPix];
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) a[j] = slot(w, j);
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) c[j] = fm_lds16(a[j]);
#pragma unroll
    for (int j = 0; j < kFmPix; ++j) fm_sts16(a[j], c[j] + 1u);
  };
  if (VEC) {
    // Ring of kFmRing planes in shared memory, filled by cp.async in groups of kFmGroup planes,
    // kFmRing - kFmGroup planes ahead of the counting.  Every thread copies and reads only its own
    // 4 bytes of each plane, so cp.async.wait_group is all the synchronisation there is.  (Plain
    // loads into a register double buffer do not work here: ptxas sinks them below the
    // shared-memory code of the previous batch and every batch then waits for DRAM — with 64 KB
    // of counters per CTA only six warps live on an SM to hide it.)
    constexpr int kSlots = kFmRing / kFmGroup;
    const int ng = Z / kFmGroup;
    const uint16_t* src = q;  // next plane to issue
    auto issue = [&](int slot, bool real) {
      if (real) {
#pragma unroll
        for (int u = 0; u < kFmGroup; ++u) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                           ring + static_cast<uint32_t>(slot * kFmGroup + u) * kFmBinStride),
                       "l"(src)
                       : "memory");
          src += plane;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int g = 0; g < kSlots - 1; ++g) issue(g, g < ng);
#pragma unroll 1
    for (int g0 = 0; g0 < ng; g0 += kSlots) {
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int g = g0 + sl;
        if (g < ng) {
          issue((sl + kSlots - 1) % kSlots, g + kSlots - 1 < ng);
          asm volatile("cp.async.wait_group %0;" ::"n"(kSlots - 1) : "memory");
          uint32_t w[kFmGroup];
#pragma unroll
          for (int u = 0; u < kFmGroup; ++u)
            w[u] = fm_lds32(ring + static_cast<uint32_t>(sl * kFmGroup + u) * kFmBinStride);
#pragma unroll
          for (int u = 0; u < kFmGroup; u += 2) bump2(w[u], w[u + 1]);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
    for (int z = ng * kFmGroup; z < Z; ++z) bump1(fm_load<true>(q + z * plane, two));
  } else {
    // unaligned pixel pairs (odd plane size or odd window start): 2-byte loads, no ring
    int z = 0;
#pragma unroll 1
    for (; z + 1 < Z; z += 2)
      bump2(fm_load<false>(q + z * plane, two), fm_load<false>(q + (z + 1) * plane, two));
    if (z < Z) bump1(fm_load<false>(q + z * plane, two));
  }
}

// Two-level scan of the thread's 256 words: for each pixel the bin where the cumulative count
// passes rank k[j] (bin, rank within the bin, samples in the bin).  Sums of 16 bins are formed for
// both pixels at once (16-bit halves cannot carry: every sum is at most Z <= 65535); "how many
// cumulative sums are <= k" and "k - the largest of them" (an unsigned minimum) need no branch.
__device__ __forceinline__ void fm_scan(uint32_t hb, const int (&k)[kFmPix], uint32_t (&bin)[kFmPix],
                                        int (&rank)[kFmPix], int (&here)[kFmPix]) {
  uint32_t gs[16];
#pragma unroll
  for (uint32_t g = 0; g < 16u; ++g) {
    uint32_t s = 0;
#pragma unroll
    for (uint32_t u = 0; u < 16u; ++u) s += fm_lds32(hb + (g * 16u + u) * kFmBinStride);
    gs[g] = s;
  }
#pragma unroll
  for (int j = 0; j < kFmPix; ++j) {
    const uint32_t kj = static_cast<uint32_t>(k[j]);
    uint32_t acc = 0, rem = kj, grp = 0;
#pragma unroll
    for (uint32_t g = 0; g < 16u; ++g) {
      acc += j == 0 ? (gs[g] & 0xffffu) : (gs[g] >> 16);
      rem = min(rem, kj - acc);
      grp += acc <= kj ? 1u : 0u;
    }
    grp = min(grp, 15u);
    const uint32_t base = hb + 2u * j + grp * 16u * kFmBinStride;
    uint32_t acc2 = 0, rem2 = rem, b = 0;
#pragma unroll
    for (uint32_t u = 0; u < 16u; ++u) {
      acc2 += fm_lds16(base + u * kFmBinStride);
      rem2 = min(rem2, rem - acc2);
      b += acc2 <= rem ? 1u : 0u;
    }
    b = min(b, 15u);
    bin[j] = grp * 16u + b;
    rank[j] = static_cast<int>(rem2);
    here[j] = static_cast<int>(fm_lds16(base + b * kFmBinStride));
  }
}

__global__ void __launch_bounds__(kFmThreads)
    flatfield_median_kernel(const __grid_constant__ FlatfieldParams p) {
  extern __shared__ __align__(16) uint32_t fm_hist[];  // [257][kFmThreads] counters, then the load ring
  __shared__ unsigned long long block_sum;
  const uint32_t t = threadIdx.x;
  const uint32_t hb = smem_u32(fm_hist) + t * 4u;
  const int64_t pend = p.p0 + p.pn;
  const int64_t pix_own = p.p0 + (static_cast<int64_t>(blockIdx.x) * kFmThreads + t) * kFmPix;
  const int npx = pix_own >= pend ? 0 : static_cast<int>(pend - pix_own < kFmPix ? pend - pix_own : kFmPix);
  // every thread runs the whole kernel (the warp votes below need all lanes): threads past the
  // window redo the last pixel and write nothing
  const int64_t pix = npx > 0 ? pix_own : pend - 1;
  const bool two = npx == kFmPix;
  // 4-byte loads need both pixels and an even element offset in every plane; decided per warp
  const bool vec =
      __all_sync(0xffffffffu, two && ((p.P & 1) == 0) && ((pix & 1) == 0)) != 0;
  if (t == 0) block_sum = 0ull;
  const uint16_t* q = p.src + pix;
  const uint32_t ring = smem_u32(fm_hist) + kFmHistBytes + t * 4u;  // [kFmRing][kFmThreads]

  auto clear = [&]() {
#pragma unroll 16
    for (uint32_t b = 0; b < 256u; ++b)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(hb + b * kFmBinStride), "r"(0u));
  };

  const int k1 = (p.Z - 1) >> 1;  // 0-based rank of the lower middle sample
  const bool even = (p.Z & 1) == 0;
  const int kk[kFmPix] = {k1, k1};
  uint32_t pre[kFmPix] = {0u, 0u}, above[kFmPix] = {0xffffffffu, 0xffffffffu};
  uint32_t hi[kFmPix], lo[kFmPix];
  int here[kFmPix], krem[kFmPix], rank[kFmPix];

  // sweep 1: high bytes
  clear();
  if (vec)
    fm_sweep<true, false, false>(q, p.P, p.Z, two, hb, pre, above, ring);
  else
    fm_sweep<false, false, false>(q, p.P, p.Z, two, hb, pre, above, ring);
  fm_scan(hb, kk, hi, krem, here);
  bool outside = false;  // the upper middle sample lies above the selected high-byte bin
#pragma unroll
  for (int j = 0; j < kFmPix; ++j) {
    pre[j] = hi[j] << 8;
    outside = outside || (even && krem[j] + 1 >= here[j]);
  }
  const bool track = __any_sync(0xffffffffu, outside) != 0;

  // sweep 2: low bytes of the samples in the selected bin
  clear();
  if (vec) {
    if (track)
      fm_sweep<true, true, true>(q, p.P, p.Z, two, hb, pre, above, ring);
    else
      fm_sweep<true, true, false>(q, p.P, p.Z, two, hb, pre, above, ring);
  } else {
    if (track)
      fm_sweep<false, true, true>(q, p.P, p.Z, two, hb, pre, above, ring);
    else
      fm_sweep<false, true, false>(q, p.P, p.Z, two, hb, pre, above, ring);
  }
  fm_scan(hb, krem, lo, rank, here);

  unsigned long long local = 0ull;
#pragma unroll
  for (int j = 0; j < kFmPix; ++j) {
    const uint32_t m1 = pre[j] | lo[j];
    uint32_t m2 = m1;
    // upper middle sample (rank k1 + 1, even Z): the same value while the rank stays inside the
    // low-byte bin, else the next occupied low-byte bin, else the smallest sample above the
    // high-byte bin (tracked by the second sweep)
    if (even && rank[j] + 1 >= here[j]) {
      uint32_t b = lo[j] + 1u;
      while (b < 256u && fm_lds16(hb + 2u * j + b * kFmBinStride) == 0u) ++b;
      m2 = b < 256u ? (pre[j] | b) : above[j] + pre[j] + 256u;
    }
    if (j < npx) {
      const uint32_t s2 = m1 + m2;  // 2 * median
      p.pattern[pix_own + j] = 0.5f * static_cast<float>(s2);
      local += s2;
    }
  }
  // block reduction of the integer sum, one global atomic per CTA
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
  __syncthreads();
  if ((t & 31) == 0 && local) atomicAdd(&block_sum, local);
  __syncthreads();
  if (t == 0 && block_sum) atomicAdd(p.sum2, block_sum);
}

// exact float64 (v / pat) * mean, one rounding per operation as numpy does
__device__ __forceinline__ double ff_value(uint32_t v, double pat, double rcp, double mean, bool plain) {
  // 2^52 + v assembled from bits, minus 2^52: exact uint16 -> float64 without a conversion op
  const double dv = __hiloint2double(0x43300000, static_cast<int>(v)) - 4503599627370496.0;
  double q;
  if (plain) {
    q = __ddiv_rn(dv, pat);
  } else {
    const double q0 = __dmul_rn(dv, rcp);
    const double rem = __fma_rn(-pat, q0, dv);
    q = __fma_rn(rem, rcp, q0);
  }
  return __dmul_rn(q, mean);
}

template <typename OUT>
__global__ void __launch_bounds__(kFfThreads)
    flatfield_apply_kernel(const __grid_constant__ FlatfieldParams p, const int planes_per_cta) {
  const int64_t pix = p.p0 + (static_cast<int64_t>(blockIdx.x) * kFfThreads + threadIdx.x) * kFfPix;
  const int64_t pend = p.p0 + p.pn;
  if (pix >= pend) return;
  const int npx = static_cast<int>(pend - pix < kFfPix ? pend - pix : kFfPix);
  const bool vec = npx == kFfPix && ((p.P & 3) == 0) && ((pix & 3) == 0);
  // numpy: pattern.mean() = add.reduce(pattern) / count; the sum is exact (see file header)
  const double mean = __ddiv_rn(__dmul_rn(static_cast<double>(*p.sum2), 0.5), static_cast<double>(p.P));
  double pat[kFfPix], rcp[kFfPix];
  bool plain[kFfPix];
#pragma unroll
  for (int j = 0; j < kFfPix; ++j) {
    pat[j] = j < npx ? static_cast<double>(p.pattern[pix + j]) : 1.0;
    plain[j] = pat[j] == 0.0;
    rcp[j] = plain[j] ? 0.0 : __drcp_rn(pat[j]);
  }
  const int zb = p.z0 + blockIdx.y * planes_per_cta;
  const int ze = min(zb + planes_per_cta, p.z0 + p.zn);
  OUT* __restrict__ dst = static_cast<OUT*>(p.dst);
  auto unpack = [](const uint2 w, uint32_t (&v)[kFfPix]) {
    v[0] = w.x & 0xffffu; v[1] = w.x >> 16; v[2] = w.y & 0xffffu; v[3] = w.y >> 16;
  };
  auto emit = [&](int z, const uint32_t (&v)[kFfPix]) {
    const int64_t off = static_cast<int64_t>(z) * p.P + pix;
    double r[kFfPix];
#pragma unroll
    for (int j = 0; j < kFfPix; ++j) r[j] = ff_value(v[j], pat[j], rcp[j], mean, plain[j]);
    if (sizeof(OUT) == 4) {
      float* o = reinterpret_cast<float*>(dst) + off;
      if (vec) {
        st_global_cs4(o, make_float4(__double2float_rn(r[0]), __double2float_rn(r[1]),
                                     __double2float_rn(r[2]), __double2float_rn(r[3])));
      } else {
#pragma unroll
        for (int j = 0; j < kFfPix; ++j)
          if (j < npx) o[j] = __double2float_rn(r[j]);
      }
    } else {
      double* o = reinterpret_cast<double*>(dst) + off;
#pragma unroll
      for (int j = 0; j < kFfPix; ++j)
        if (j < npx) __stcs(o + j, r[j]);
    }
  };
  int z = zb;
  if (vec) {
    // kFfBatch planes of loads in flight per thread before the first quotient
#pragma unroll 1
    for (; z + kFfBatch <= ze; z += kFfBatch) {
      uint2 w[kFfBatch];
#pragma unroll
      for (int u = 0; u < kFfBatch; ++u)
        w[u] = __ldcs(reinterpret_cast<const uint2*>(p.src + static_cast<int64_t>(z + u) * p.P + pix));
#pragma unroll
      for (int u = 0; u < kFfBatch; ++u) {
        uint32_t v[kFfPix];
        unpack(w[u], v);
        emit(z + u, v);
      }
    }
  }
#pragma unroll 1
  for (; z < ze; ++z) {
    const int64_t off = static_cast<int64_t>(z) * p.P + pix;
    uint32_t v[kFfPix];
    if (vec) {
      unpack(__ldcs(reinterpret_cast<const uint2*>(p.src + off)), v);
    } else {
#pragma unroll
      for (int j = 0; j < kFfPix; ++j) v[j] = j < npx ? __ldg(p.src + off + j) : 0u;
    }
    emit(z, v);
  }
}

// ---------------------------------------------------------------------------------------------
// host side (device pointers, caller's stream, never synchronises)
// ---------------------------------------------------------------------------------------------
size_t flatfield_workspace_bytes(int64_t Y, int64_t X) {
  const size_t pat = (static_cast<size_t>(Y) * X * sizeof(float) + 255) / 256 * 256;
  return pat + 256;
}

static int ff_check(const void* src, int64_t Z, int64_t Y, int64_t X, void* ws, size_t ws_bytes) {
  if (!src || !ws) {
    set_error("flatfield: null pointer");
    return B2_ERR_INVALID;
  }
  if (Z < 1 || Y < 1 || X < 1 || Z > 65535) {
    set_error("flatfield: invalid shape (1 <= Z <= 65535: uint16 histogram counters)");
    return B2_ERR_INVALID;
  }
  if (ws_bytes < flatfield_workspace_bytes(Y, X)) {
    set_error("flatfield: workspace too small (%zu < %zu bytes)", ws_bytes,
              flatfield_workspace_bytes(Y, X));
    return B2_ERR_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(src) % 8 != 0 || reinterpret_cast<uintptr_t>(ws) % 256 != 0) {
    set_error("flatfield: src must be 8-byte aligned and the workspace 256-byte aligned");
    return B2_ERR_INVALID;
  }
  return B2_OK;
}

static FlatfieldParams ff_params(const void* src, int64_t Z, int64_t Y, int64_t X, void* dst, void* ws) {
  FlatfieldParams p{};
  p.src = static_cast<const uint16_t*>(src);
  p.dst = dst;
  p.pattern = static_cast<float*>(ws);
  const size_t pat = (static_cast<size_t>(Y) * X * sizeof(float) + 255) / 256 * 256;
  p.sum2 = reinterpret_cast<unsigned long long*>(static_cast<char*>(ws) + pat);
  p.P = Y * X;
  p.Z = static_cast<int>(Z);
  return p;
}

// zero the pattern-sum accumulator (first step of a volume)
int flatfield_begin(int64_t Y, int64_t X, void* ws, cudaStream_t stream) {
  const size_t pat = (static_cast<size_t>(Y) * X * sizeof(float) + 255) / 256 * 256;
  B2_CUDA(cudaMemsetAsync(static_cast<char*>(ws) + pat, 0, 256, stream));
  return B2_OK;
}

// medians of the pixels [p0, p0 + pn) (all Z planes of those pixels must be resident)
int flatfield_median(const void* src, int64_t Z, int64_t Y, int64_t X, void* ws, size_t ws_bytes,
                     int64_t p0, int64_t pn, cudaStream_t stream) {
  int rc = ff_check(src, Z, Y, X, ws, ws_bytes);
  if (rc) return rc;
  if (p0 < 0 || pn < 0 || p0 + pn > Y * X) {
    set_error("flatfield: invalid pixel window");
    return B2_ERR_INVALID;
  }
  if (pn == 0) return B2_OK;
  FlatfieldParams p = ff_params(src, Z, Y, X, nullptr, ws);
  p.p0 = p0;
  p.pn = pn;
  const int64_t per_cta = static_cast<int64_t>(kFmThreads) * kFmPix;
  const int64_t grid = (pn + per_cta - 1) / per_cta;
  if (grid > 2147483647LL) {
    set_error("flatfield: plane too large");
    return B2_ERR_INVALID;
  }
  B2_CUDA(cudaFuncSetAttribute(flatfield_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kFmSmemBytes));
  B2_CUDA(cudaFuncSetAttribute(flatfield_median_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared));
  flatfield_median_kernel<<<static_cast<unsigned>(grid), kFmThreads, kFmSmemBytes, stream>>>(p);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

// out planes [z0, z0 + zn) from the finished pattern; dst_dtype B2_DTYPE_F32 | B2_DTYPE_F64
int flatfield_apply(const void* src, int64_t Z, int64_t Y, int64_t X, void* dst, int dst_dtype,
                    void* ws, size_t ws_bytes, int64_t z0, int64_t zn, cudaStream_t stream) {
  int rc = ff_check(src, Z, Y, X, ws, ws_bytes);
  if (rc) return rc;
  if (!dst || (dst_dtype != B2_DTYPE_F32 && dst_dtype != B2_DTYPE_F64)) {
    set_error("flatfield: dst must be float32 or float64");
    return B2_ERR_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(dst) % 16 != 0) {
    set_error("flatfield: dst must be 16-byte aligned");
    return B2_ERR_INVALID;
  }
  if (z0 < 0 || zn < 0 || z0 + zn > Z) {
    set_error("flatfield: invalid plane window");
    return B2_ERR_INVALID;
  }
  if (zn == 0) return B2_OK;
  FlatfieldParams p = ff_params(src, Z, Y, X, dst, ws);
  p.p0 = 0;
  p.pn = p.P;
  p.z0 = static_cast<int>(z0);
  p.zn = static_cast<int>(zn);
  const int64_t per_cta = static_cast<int64_t>(kFfThreads) * kFfPix;
  const int64_t gx = (p.P + per_cta - 1) / per_cta;
  // kFfApplyPlanes planes per CTA: the four reciprocals of a thread are amortised over 64 voxels
  // and the grid is tens of waves deep (with a few hundred planes per CTA the grid was 4.05 waves
  // of 4 CTAs per SM: a fifth, almost empty wave cost 20 % of the kernel)
  const int planes = static_cast<int>(std::min<int64_t>(kFfApplyPlanes, zn));
  const int64_t gy = (zn + planes - 1) / planes;
  if (gx > 2147483647LL || gy > 65535) {
    set_error("flatfield: grid too large");
    return B2_ERR_INVALID;
  }
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
  if (dst_dtype == B2_DTYPE_F32)
    flatfield_apply_kernel<float><<<grid, kFfThreads, 0, stream>>>(p, planes);
  else
    flatfield_apply_kernel<double><<<grid, kFfThreads, 0, stream>>>(p, planes);
  B2_CUDA(cudaGetLastError());
  count_launch();
  return B2_OK;
}

}  // namespace b2
