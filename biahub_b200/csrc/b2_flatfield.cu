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
//  * flatfield_median_kernel: one thread owns 4 adjacent pixels (8-byte coalesced loads along x)
//    and finds the k-th smallest of the Z samples of each by a 4 x 4-bit radix select: per pass
//    a 16-bin histogram (uint16 counters in shared memory, 128 B per thread) of the samples that
//    match the prefix found so far; a 5th pass (even Z only) finds the next order statistic.
//    Every pass re-reads the column from L2/HBM: 5 * Z*Y*X*2 bytes.  Integer work, exact.
//  * flatfield_apply_kernel: exact float64 quotient without the division subroutine: r = RN(1/p)
//    once per pixel, then q0 = v*r, rem = fma(-p, q0, v), q = fma(rem, r, q0) — the correctly
//    rounded quotient for every uint16 v and every half-integer p (checked exhaustively against
//    __ddiv_rn by scripts/flatfield_div_check.cu) — times the mean, rounded once more to float32.
//    p == 0 (a column of zeros) takes the IEEE division (inf / nan as numpy).
// mean(pattern): every partial sum of half-integers below 2^52 is exact, so the integer sum equals
// numpy's pairwise float64 sum bit for bit (tests/test_flatfield_oracle.py).
#include <algorithm>
#include <type_traits>

#include "b2_common.cuh"

namespace b2 {

constexpr int kFfThreads = 256;
constexpr int kFfPix = 4;  // pixels per thread (one 8-byte load per plane)

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

// histogram counter of (bin, pixel j) of thread t: [(bin * kFfPix + j) * kFfThreads + t]
__device__ __forceinline__ uint32_t ff_slot(uint32_t bin, uint32_t j, uint32_t t) {
  return ((bin * kFfPix + j) * kFfThreads + t) * 2u;
}

__global__ void __launch_bounds__(kFfThreads)
    flatfield_median_kernel(const __grid_constant__ FlatfieldParams p) {
  __shared__ uint16_t hist[16 * kFfPix * kFfThreads];  // 32 KB
  __shared__ unsigned long long block_sum;
  const uint32_t t = threadIdx.x;
  const uint32_t hbase = smem_u32(hist);
  const int64_t pix = p.p0 + (static_cast<int64_t>(blockIdx.x) * kFfThreads + t) * kFfPix;
  const int64_t pend = p.p0 + p.pn;
  // vector path needs all 4 pixels in range and an 8-byte aligned column start in every plane
  const bool vec = (pix + kFfPix <= pend) && ((p.P & 3) == 0) && ((pix & 3) == 0);
  const int npx = pix >= pend ? 0 : static_cast<int>(pend - pix < kFfPix ? pend - pix : kFfPix);
  if (t == 0) block_sum = 0ull;

  const int k1 = (p.Z - 1) >> 1;  // 0-based rank of the lower middle sample
  const bool even = (p.Z & 1) == 0;
  uint32_t prefix[kFfPix] = {0, 0, 0, 0};  // bits found so far (high to low)
  int krem[kFfPix] = {k1, k1, k1, k1};     // rank within the samples matching the prefix
  int ceq[kFfPix] = {0, 0, 0, 0};          // after the last pass: samples equal to the median

  // VEC: the thread's 4 pixels are one aligned 8-byte load per plane (CTA-interior, P % 4 == 0)
  auto load4 = [&](auto vec_tag, int z, uint32_t (&v)[kFfPix]) {
    const uint16_t* q = p.src + static_cast<int64_t>(z) * p.P + pix;
    if (decltype(vec_tag)::value) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(q));
      v[0] = w.x & 0xffffu; v[1] = w.x >> 16; v[2] = w.y & 0xffffu; v[3] = w.y >> 16;
    } else {
#pragma unroll
      for (int j = 0; j < kFfPix; ++j) v[j] = j < npx ? __ldg(q + j) : 0u;
    }
  };

#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 12 - 4 * pass;
    // clear this thread's counters
#pragma unroll
    for (uint32_t b = 0; b < 16; ++b)
#pragma unroll
      for (uint32_t j = 0; j < kFfPix; ++j)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(hbase + ff_slot(b, j, t)), "h"((unsigned short)0));
    if (npx > 0) {
      auto histogram_pass = [&](auto vec_tag) {
      // kFfAhead planes are loaded before their counters are touched: the counter updates are an
      // ordered chain of shared-memory round trips (two samples of a column may hit the same
      // bin), so without this the kernel waits for one global load at a time per thread
      // (measured: 25 % of the HBM bandwidth, 35 % of the issue slots)
      constexpr int kFfAhead = 8;
      for (int zb = 0; zb < p.Z; zb += kFfAhead) {
        uint32_t vv[kFfAhead][kFfPix];
#pragma unroll
        for (int u = 0; u < kFfAhead; ++u) {
          if (zb + u < p.Z) {
            load4(vec_tag, zb + u, vv[u]);
          } else {
#pragma unroll
            for (int j = 0; j < kFfPix; ++j) vv[u][j] = 0u;
          }
        }
#pragma unroll
        for (int u = 0; u < kFfAhead; ++u) {
          if (zb + u >= p.Z) break;
          uint32_t (&v)[kFfPix] = vv[u];
          // the four pixels own separate counters: their four loads are issued before the four
          // stores (pixel order ld, add, st, ld, ... would chain the shared-memory round trips)
          uint32_t addr[kFfPix];
          bool on[kFfPix];
          unsigned short cnt[kFfPix];
#pragma unroll
          for (uint32_t j = 0; j < kFfPix; ++j) {
            const uint32_t hi = pass == 0 ? 0u : (v[j] >> (shift + 4));
            on[j] = hi == prefix[j];
            addr[j] = hbase + ff_slot((v[j] >> shift) & 15u, j, t);
          }
#pragma unroll
          for (uint32_t j = 0; j < kFfPix; ++j) {
            cnt[j] = 0;
            if (on[j]) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(cnt[j]) : "r"(addr[j]));
          }
#pragma unroll
          for (uint32_t j = 0; j < kFfPix; ++j) {
            const unsigned short c = static_cast<unsigned short>(cnt[j] + 1);
            if (on[j]) asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr[j]), "h"(c));
          }
        }
      }
      };
      if (vec) {
        histogram_pass(std::true_type{});
      } else {
        histogram_pass(std::false_type{});
      }
      // scan: the bin where the cumulative count passes the rank
#pragma unroll
      for (uint32_t j = 0; j < kFfPix; ++j) {
        int acc = 0;
        uint32_t bin = 15;
        int below = 0, here = 0;
        bool found = false;
#pragma unroll
        for (uint32_t b = 0; b < 16; ++b) {
          unsigned short c;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c) : "r"(hbase + ff_slot(b, j, t)));
          if (!found && acc + static_cast<int>(c) > krem[j]) {
            found = true;
            bin = b;
            below = acc;
            here = c;
          }
          acc += c;
        }
        krem[j] -= below;
        prefix[j] = (prefix[j] << 4) | bin;
        ceq[j] = here;
      }
    }
  }

  // m1 = prefix; samples <= m1 among ALL samples: (k1 - krem) below + ceq equal.  The upper middle
  // sample (rank k1 + 1, even Z) equals m1 when krem + 1 < ceq, else it is the smallest sample > m1.
  uint32_t m2[kFfPix];
  bool need_next = false;
#pragma unroll
  for (int j = 0; j < kFfPix; ++j) {
    m2[j] = prefix[j];
    if (even && j < npx && krem[j] + 1 >= ceq[j]) {
      m2[j] = 0xffffffffu;
      need_next = true;
    }
  }
  if (need_next) {
    auto next_pass = [&](auto vec_tag) {
#pragma unroll 8
      for (int z = 0; z < p.Z; ++z) {
        uint32_t v[kFfPix];
        load4(vec_tag, z, v);
#pragma unroll
        for (int j = 0; j < kFfPix; ++j)
          if (v[j] > prefix[j] && m2[j] != prefix[j]) m2[j] = min(m2[j], v[j]);
      }
    };
    if (vec) {
      next_pass(std::true_type{});
    } else {
      next_pass(std::false_type{});
    }
  }
  unsigned long long local = 0ull;
#pragma unroll
  for (int j = 0; j < kFfPix; ++j) {
    if (j < npx) {
      const uint32_t s2 = prefix[j] + m2[j];  // 2 * median
      p.pattern[pix + j] = 0.5f * static_cast<float>(s2);
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
#pragma unroll 2
  for (int z = zb; z < ze; ++z) {
    const int64_t off = static_cast<int64_t>(z) * p.P + pix;
    uint32_t v[kFfPix];
    if (vec) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(p.src + off));
      v[0] = w.x & 0xffffu; v[1] = w.x >> 16; v[2] = w.y & 0xffffu; v[3] = w.y >> 16;
    } else {
#pragma unroll
      for (int j = 0; j < kFfPix; ++j) v[j] = j < npx ? __ldg(p.src + off + j) : 0u;
    }
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
  const int64_t per_cta = static_cast<int64_t>(kFfThreads) * kFfPix;
  const int64_t grid = (pn + per_cta - 1) / per_cta;
  if (grid > 2147483647LL) {
    set_error("flatfield: plane too large");
    return B2_ERR_INVALID;
  }
  flatfield_median_kernel<<<static_cast<unsigned>(grid), kFfThreads, 0, stream>>>(p);
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
  int sms = 148;
  sm_count(&sms);
  // enough CTAs for a few waves; each CTA amortises its reciprocals over planes_per_cta planes
  int64_t gy = (static_cast<int64_t>(sms) * 16 + gx - 1) / gx;
  gy = std::max<int64_t>(1, std::min<int64_t>(gy, zn));
  const int planes = static_cast<int>((zn + gy - 1) / gy);
  gy = (zn + planes - 1) / planes;
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
