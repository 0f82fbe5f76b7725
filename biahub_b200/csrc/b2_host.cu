// Host-buffer pipeline behind b2h_deskew / b2h_affine3d: numpy-in / numpy-out callers
// (reference biahub/deskew.py:551-579, biahub/register.py:202-281, biahub/stabilize.py:32-90)
// hand over pageable or pinned HOST arrays and get a complete HOST result back.
//
// The output volume is cut into slabs along its slowest axis.  Slab i needs only a band of the
// source (deskew: a band of tilt rows of every scan plane; affine: a range of source planes), so
//
//   upload   stream: H2D of the source band of slab i+1        (cudaMemcpy2DAsync / cudaMemcpyAsync,
//                                                                one call per band - never batched)
//   compute  stream: resampling kernel of slab i                (ordered after its band by an event)
//   download stream: D2H of slab i-1                            (ordered after its kernel by an event)
//
// run concurrently: PCIe carries traffic in both directions at once and the kernel time hides
// behind the copies.  Pinned user buffers are used directly; pageable ones are staged through
// small pinned rings by a pool of host threads.  Per-process, per-device state (streams, device
// volumes, pinned rings, events) is cached and grown on demand; b2h_release() frees it.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "b2_common.cuh"

namespace b2 {

int deskew_device(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* dst,
                  int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                  float pxct32, float off32, int path, cudaStream_t stream, const int* slab,
                  int64_t dst_row_pitch);
int affine_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* dst,
                  int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                  int order, int boundary, int scrub, int path, cudaStream_t stream,
                  int64_t src_row_pitch, int64_t dst_row_pitch);

size_t fill_workspace_bytes(int64_t z, int64_t y, int64_t x);
int fill_device(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                int iterations, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t spline_workspace_bytes(int64_t sz, int64_t sy, int64_t sx);
int spline_prefilter_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                            int scrub, void* ws, size_t ws_bytes, const double** coef,
                            cudaStream_t st);
int spline_eval_device(const double* coef, int64_t sz, int64_t sy, int64_t sx, void* dst,
                       int dst_dtype, int64_t oz, int64_t oy, int64_t ox, const double* M12,
                       const int64_t* crop_start, cudaStream_t st);

size_t flatfield_workspace_bytes(int64_t Y, int64_t X);
int flatfield_begin(int64_t Y, int64_t X, void* ws, cudaStream_t stream);
int flatfield_median(const void* src, int64_t Z, int64_t Y, int64_t X, void* ws, size_t ws_bytes,
                     int64_t p0, int64_t pn, cudaStream_t stream);
int flatfield_apply(const void* src, int64_t Z, int64_t Y, int64_t X, void* dst, int dst_dtype,
                    void* ws, size_t ws_bytes, int64_t z0, int64_t zn, cudaStream_t stream);

namespace {

constexpr size_t kSlabBytes = 48u << 20;  // target output bytes per slab
constexpr int kRing = 3;                  // pinned staging ring depth (pageable callers)
constexpr int kOutRing = 3;               // device output ring depth (slab i computes while i-1 travels)
constexpr int kMaxDevices = 16;

// ------------------------------------------------------------------ host thread pool
class HostPool {
 public:
  static HostPool& get() {
    static HostPool pool;
    return pool;
  }
  int size() const { return static_cast<int>(workers_.size()) + 1; }

  // run fn(i) for i in [0, n) on the pool + the calling thread; returns when all are done
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (n == 1 || workers_.empty()) {
      for (int i = 0; i < n; ++i) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      next_ = 0;
      total_ = n;
      pending_ = n;
      ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() {
    unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? static_cast<int>(hc) : 4;
    const char* env = getenv("B2_HOST_THREADS");
    if (env && atoi(env) > 0) t = atoi(env);
    t = std::max(1, std::min(t, 12));
    for (int i = 1; i < t; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  void work() {
    for (;;) {
      int i;
      const std::function<void(int)>* fn;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!fn_ || next_ >= total_) return;
        i = next_++;
        fn = fn_;
      }
      (*fn)(i);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

// copy `rows` rows of `width` bytes between pitched host buffers, split over the pool
void host_copy_2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width,
                  size_t rows) {
  HostPool& pool = HostPool::get();
  const size_t total = width * rows;
  int parts = static_cast<int>(std::min<size_t>(pool.size(), std::max<size_t>(1, total >> 21)));
  if (parts <= 1) {
    for (size_t r = 0; r < rows; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
    return;
  }
  if (rows >= static_cast<size_t>(parts)) {
    pool.parallel_for(parts, [&](int t) {
      const size_t r0 = rows * t / parts, r1 = rows * (t + 1) / parts;
      for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
    });
  } else {  // few long rows: split each row
    for (size_t r = 0; r < rows; ++r) {
      pool.parallel_for(parts, [&](int t) {
        const size_t b0 = (width * t / parts) & ~static_cast<size_t>(63);
        const size_t b1 = (t + 1 == parts) ? width : ((width * (t + 1) / parts) & ~static_cast<size_t>(63));
        memcpy(dst + r * dpitch + b0, src + r * spitch + b0, b1 - b0);
      });
    }
  }
}

// ------------------------------------------------------------------ per-device context
struct DeviceCtx {
  bool init = false;
  cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
  void* d_src = nullptr;
  size_t d_src_bytes = 0;
  void* d_dst = nullptr;
  size_t d_dst_bytes = 0;
  void* d_ws = nullptr;  // flat-field pattern + sum
  size_t d_ws_bytes = 0;
  void* d_mid = nullptr;  // deskewed volume of the chained deskew -> register unit
  size_t d_mid_bytes = 0;
  void* h_in[kRing] = {nullptr, nullptr, nullptr};
  size_t h_in_bytes[kRing] = {0, 0, 0};
  void* h_out[kRing] = {nullptr, nullptr, nullptr};
  size_t h_out_bytes[kRing] = {0, 0, 0};
  std::vector<cudaEvent_t> events;
};

std::mutex g_mu;  // b2h_* calls are serialised per process (one worker process per GPU)
DeviceCtx g_ctx[kMaxDevices];

// Every b2h_* call selects its GPU with cudaSetDevice; the statically linked runtime shares the
// primary context with the caller (torch), so the previous device is restored on every exit path.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      (void)cudaGetLastError();
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) (void)cudaSetDevice(prev);
  }
};

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int grow_device(void** p, size_t* have, size_t need) {
  if (*have >= need) return B2_OK;
  if (*p) B2_CUDA(cudaFree(*p));
  *p = nullptr;
  *have = 0;
  B2_CUDA(cudaMalloc(p, need));
  *have = need;
  return B2_OK;
}

int grow_pinned(void** p, size_t* have, size_t need) {
  if (*have >= need) return B2_OK;
  if (*p) B2_CUDA(cudaFreeHost(*p));
  *p = nullptr;
  *have = 0;
  B2_CUDA(cudaHostAlloc(p, need, cudaHostAllocDefault));
  *have = need;
  return B2_OK;
}

int get_ctx(int device, DeviceCtx** out) {
  if (device < 0 || device >= kMaxDevices) {
    set_error("device index %d out of range", device);
    return B2_ERR_INVALID;
  }
  B2_CUDA(cudaSetDevice(device));
  DeviceCtx& c = g_ctx[device];
  if (!c.init) {
    B2_CUDA(cudaStreamCreateWithFlags(&c.s_up, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c.s_run, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c.s_down, cudaStreamNonBlocking));
    c.init = true;
  }
  *out = &c;
  return B2_OK;
}

int get_event(DeviceCtx& c, size_t i, cudaEvent_t* ev) {
  while (c.events.size() <= i) {
    cudaEvent_t e;
    B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c.events.push_back(e);
  }
  *ev = c.events[i];
  return B2_OK;
}

size_t elem_size(int dtype) { return dtype == B2_DTYPE_U16 ? 2 : 4; }

// A band of the source: `rows` pieces of `width` bytes, piece r at byte offset off + r*pitch of
// both the host volume and its device mirror.
struct Band {
  size_t off = 0, pitch = 0, width = 0, rows = 0;
  size_t bytes() const { return width * rows; }
};

struct Slab {
  std::vector<Band> bands;  // source data to upload before this slab's kernel
  size_t out_off = 0, out_bytes = 0;  // where the slab's result goes in the host output
  // kernel(s) of this slab; `d_out` is the slab's slot of the device output ring (out_bytes long).
  // Empty = nothing to launch (upload-only or download-only slab).
  std::function<int(cudaStream_t, char* d_out)> launch;
  const char* d_from = nullptr;  // download from here instead of the ring slot (resident results)
};

// Run the three-stream pipeline over `slabs`.  Device results live in a ring of kOutRing slots
// (the largest slab each), not in a full-volume mirror: slab i's kernel waits for the download
// of slab i - kOutRing.
int run_pipeline_impl(DeviceCtx& c, const char* h_src, char* h_dst, std::vector<Slab>& slabs) {
  const bool src_pinned = is_pinned(h_src);
  const bool dst_pinned = is_pinned(h_dst);
  const size_t S = slabs.size();
  size_t slot_bytes = 0;
  for (const Slab& sl : slabs)
    if (!sl.d_from) slot_bytes = std::max(slot_bytes, sl.out_bytes);
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  if (slot_bytes) {
    int grc = grow_device(&c.d_dst, &c.d_dst_bytes, slot_bytes * kOutRing);
    if (grc) return grc;
  }
  auto ev = [&](size_t kind, size_t i, cudaEvent_t* e) { return get_event(c, 3 * i + kind, e); };
  int rc;

  // pageable destination: results land in a pinned staging ring first; stage_owner[r] is the slab
  // whose result currently sits in ring slot r (-1 = free)
  long stage_owner[kRing];
  for (int r = 0; r < kRing; ++r) stage_owner[r] = -1;
  size_t n_staged = 0;
  auto unstage = [&](int r) -> int {
    if (stage_owner[r] < 0) return B2_OK;
    const size_t j = static_cast<size_t>(stage_owner[r]);
    cudaEvent_t e;
    if ((rc = ev(2, j, &e))) return rc;
    B2_CUDA(cudaEventSynchronize(e));
    host_copy_2d(h_dst + slabs[j].out_off, slabs[j].out_bytes,
                 static_cast<const char*>(c.h_out[r]), slabs[j].out_bytes, slabs[j].out_bytes, 1);
    stage_owner[r] = -1;
    return B2_OK;
  };

  for (size_t i = 0; i < S; ++i) {
    Slab& sl = slabs[i];
    cudaEvent_t e_up, e_run, e_down;
    if ((rc = ev(0, i, &e_up)) || (rc = ev(1, i, &e_run)) || (rc = ev(2, i, &e_down))) return rc;

    // ---- upload the source band(s) of this slab
    size_t band_total = 0;
    for (const Band& b : sl.bands) band_total += b.bytes();
    if (!src_pinned && band_total) {
      if (i >= static_cast<size_t>(kRing)) {  // ring slot reuse: its previous upload must be done
        cudaEvent_t prev;
        if ((rc = ev(0, i - kRing, &prev))) return rc;
        B2_CUDA(cudaEventSynchronize(prev));
      }
      if ((rc = grow_pinned(&c.h_in[i % kRing], &c.h_in_bytes[i % kRing], band_total))) return rc;
    }
    size_t stage_off = 0;
    for (const Band& b : sl.bands) {
      if (!b.bytes()) continue;
      char* d = static_cast<char*>(c.d_src) + b.off;
      if (src_pinned) {
        if (b.rows == 1)
          B2_CUDA(cudaMemcpyAsync(d, h_src + b.off, b.width, cudaMemcpyHostToDevice, c.s_up));
        else
          B2_CUDA(cudaMemcpy2DAsync(d, b.pitch, h_src + b.off, b.pitch, b.width, b.rows,
                                    cudaMemcpyHostToDevice, c.s_up));
      } else {
        char* st = static_cast<char*>(c.h_in[i % kRing]) + stage_off;
        host_copy_2d(st, b.width, h_src + b.off, b.pitch, b.width, b.rows);
        if (b.rows == 1)
          B2_CUDA(cudaMemcpyAsync(d, st, b.width, cudaMemcpyHostToDevice, c.s_up));
        else
          B2_CUDA(cudaMemcpy2DAsync(d, b.pitch, st, b.width, b.width, b.rows,
                                    cudaMemcpyHostToDevice, c.s_up));
        stage_off += b.bytes();
      }
    }
    B2_CUDA(cudaEventRecord(e_up, c.s_up));

    // ---- kernel of this slab (its ring slot must have been downloaded)
    B2_CUDA(cudaStreamWaitEvent(c.s_run, e_up, 0));
    char* d_slot = static_cast<char*>(c.d_dst) + (i % kOutRing) * slot_bytes;
    if (i >= static_cast<size_t>(kOutRing)) {
      cudaEvent_t prev_down;
      if ((rc = ev(2, i - kOutRing, &prev_down))) return rc;
      B2_CUDA(cudaStreamWaitEvent(c.s_run, prev_down, 0));
    }
    if (sl.launch && (rc = sl.launch(c.s_run, d_slot))) return rc;
    B2_CUDA(cudaEventRecord(e_run, c.s_run));

    // ---- download
    B2_CUDA(cudaStreamWaitEvent(c.s_down, e_run, 0));
    const char* d_out = sl.d_from ? sl.d_from : d_slot;
    if (!sl.out_bytes) {
      // nothing to bring back (upload / compute only)
    } else if (dst_pinned) {
      B2_CUDA(cudaMemcpyAsync(h_dst + sl.out_off, d_out, sl.out_bytes, cudaMemcpyDeviceToHost,
                              c.s_down));
    } else {
      const int r = static_cast<int>(n_staged++ % kRing);
      if ((rc = unstage(r))) return rc;
      if ((rc = grow_pinned(&c.h_out[r], &c.h_out_bytes[r], sl.out_bytes))) return rc;
      B2_CUDA(cudaMemcpyAsync(c.h_out[r], d_out, sl.out_bytes, cudaMemcpyDeviceToHost, c.s_down));
      stage_owner[r] = static_cast<long>(i);
    }
    B2_CUDA(cudaEventRecord(e_down, c.s_down));
  }
  for (size_t k = 0; k < static_cast<size_t>(kRing); ++k)  // oldest first
    if ((rc = unstage(static_cast<int>((n_staged + k) % kRing)))) return rc;
  B2_CUDA(cudaStreamSynchronize(c.s_down));
  B2_CUDA(cudaStreamSynchronize(c.s_up));
  return B2_OK;
}

int run_pipeline(DeviceCtx& c, const char* h_src, char* h_dst, std::vector<Slab>& slabs) {
  const int rc = run_pipeline_impl(c, h_src, h_dst, slabs);
  if (rc) {
    // a failed step must not leave copies in flight that target the caller's arrays (or pooled
    // pinned blocks): drain all three streams, keeping the first error for the caller
    (void)cudaStreamSynchronize(c.s_up);
    (void)cudaStreamSynchronize(c.s_run);
    (void)cudaStreamSynchronize(c.s_down);
    (void)cudaGetLastError();
  }
  return rc;
}

}  // namespace

// fill_mode: 0 = none, 1 = overhang fill with the mean of the un-masked voxels, 2 = with
// fill_value (reference biahub/deskew.py:538-540 -> _fill_overhang_torch).  The fill needs the
// whole deskewed volume (global mask dilation + mean), so in that mode the deskew slabs stay on
// the device, the fill runs once after the last slab, and the download slabs follow; uploads,
// deskew kernels and (afterwards) the downloads still overlap.
int host_deskew_fill(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                     float* h_dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N,
                     float px32, float pxct32, float off32, int fill_mode, float fill_value,
                     int device) {
  if (!h_src || !h_dst) {
    set_error("b2h_deskew: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("b2h_deskew: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (Zi < 2 || Yi < 1 || Xi < 1 || Zavg < 1 || Yo < 1 || Xo < 1 || N < 1 || Zo_full != Yi ||
      Zavg != (Zo_full + N - 1) / N) {
    set_error("b2h_deskew: invalid shape");
    return B2_ERR_INVALID;
  }
  if (fill_mode < 0 || fill_mode > 2) {
    set_error("b2h_deskew_fill: fill_mode must be 0, 1 (mean) or 2 (constant)");
    return B2_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t es = elem_size(src_dtype);
  const size_t in_bytes = static_cast<size_t>(Zi) * Yi * Xi * es;
  const size_t out_bytes = static_cast<size_t>(Zavg) * Yo * Xo * sizeof(float);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, in_bytes))) return rc;
  float* d_full = nullptr;  // fill mode: the whole deskewed volume stays resident
  size_t fill_ws = 0;
  if (fill_mode) {
    fill_ws = fill_workspace_bytes(Zavg, Yo, Xo);
    if ((rc = grow_device(&c->d_mid, &c->d_mid_bytes, out_bytes))) return rc;
    if ((rc = grow_device(&c->d_ws, &c->d_ws_bytes, fill_ws))) return rc;
    d_full = static_cast<float*>(c->d_mid);
  }
  void* d_ws = c->d_ws;

  // slabs of averaged slices [a0, a0+cnt): contiguous in dst; they read tilt rows
  // iy in [Yi - min((a0+cnt)*N, Yi), Yi - 1 - a0*N] of every scan plane
  const size_t slice_bytes = static_cast<size_t>(Yo) * Xo * sizeof(float);
  const int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / slice_bytes));
  const size_t row_bytes = static_cast<size_t>(Xi) * es;
  std::vector<Slab> slabs;
  void* d_src = c->d_src;
  for (int64_t a0 = 0; a0 < Zavg; a0 += per_slab) {
    const int64_t cnt = std::min(per_slab, Zavg - a0);
    const int64_t iy_hi = Yi - 1 - a0 * N;
    const int64_t iy_lo = Yi - std::min<int64_t>((a0 + cnt) * N, Yi);
    Slab s;
    Band b;
    b.off = static_cast<size_t>(iy_lo) * row_bytes;
    b.pitch = static_cast<size_t>(Yi) * row_bytes;
    b.width = static_cast<size_t>(iy_hi - iy_lo + 1) * row_bytes;
    b.rows = static_cast<size_t>(Zi);
    s.bands.push_back(b);
    s.out_off = static_cast<size_t>(a0) * slice_bytes;
    s.out_bytes = fill_mode ? 0 : static_cast<size_t>(cnt) * slice_bytes;
    const bool last = a0 + cnt >= Zavg;
    s.launch = [=](cudaStream_t st, char* d_out) {
      const int slab[4] = {0, static_cast<int>(Yi), static_cast<int>(a0), static_cast<int>(cnt)};
      float* dst = fill_mode ? d_full + a0 * Yo * Xo : reinterpret_cast<float*>(d_out);
      int r = deskew_device(d_src, src_dtype, Zi, Yi, Xi, dst, Zavg, Yo, Xo, Zo_full, N, px32,
                            pxct32, off32, B2_PATH_AUTO, st, slab, 0);
      if (r || !fill_mode || !last) return r;
      return fill_device(d_full, Zavg, Yo, Xo, fill_mode == 1, fill_value, 3, d_ws, fill_ws, st);
    };
    slabs.push_back(std::move(s));
  }
  if (fill_mode) {  // download slabs of the filled, resident volume
    for (int64_t a0 = 0; a0 < Zavg; a0 += per_slab) {
      const int64_t cnt = std::min(per_slab, Zavg - a0);
      Slab s;
      s.out_off = static_cast<size_t>(a0) * slice_bytes;
      s.out_bytes = static_cast<size_t>(cnt) * slice_bytes;
      s.d_from = reinterpret_cast<const char*>(d_full) + s.out_off;
      slabs.push_back(std::move(s));
    }
  }
  return run_pipeline(*c, static_cast<const char*>(h_src), reinterpret_cast<char*>(h_dst), slabs);
}

int host_deskew(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* h_dst,
                int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                float pxct32, float off32, int device) {
  return host_deskew_fill(h_src, src_dtype, Zi, Yi, Xi, h_dst, Zavg, Yo, Xo, Zo_full, N, px32,
                          pxct32, off32, 0, 0.0f, device);
}

int host_affine(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* h_dst,
                int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                int order, int boundary, int scrub, int device) {
  if (!h_src || !h_dst || !M12) {
    set_error("b2h_affine3d: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("b2h_affine3d: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (sz < 1 || sy < 1 || sx < 1 || oz < 0 || oy < 0 || ox < 0) {
    set_error("b2h_affine3d: invalid shape");
    return B2_ERR_INVALID;
  }
  if (oz == 0 || oy == 0 || ox == 0) return B2_OK;
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t es = elem_size(src_dtype);
  const size_t plane_in = static_cast<size_t>(sy) * sx * es;
  const size_t in_bytes = static_cast<size_t>(sz) * plane_in;
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, in_bytes))) return rc;

  const int64_t c0[3] = {crop_start ? crop_start[0] : 0, crop_start ? crop_start[1] : 0,
                         crop_start ? crop_start[2] : 0};
  const size_t plane_out = static_cast<size_t>(oy) * ox * sizeof(float);
  // whole 16-plane tile layers of the generic warp kernel per slab (its bricks are staged per
  // 16-deep tile: a 2-plane slab would stage them for 2 planes each)
  int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / plane_out));
  per_slab = (per_slab + 15) / 16 * 16;
  std::vector<Slab> slabs;
  void* d_src = c->d_src;
  std::vector<double> M(M12, M12 + 12);
  int64_t up_lo = 0, up_hi = 0;  // source planes [up_lo, up_hi) already scheduled for upload
  bool any = false;
  for (int64_t z0 = 0; z0 < oz; z0 += per_slab) {
    const int64_t cnt = std::min(per_slab, oz - z0);
    // source plane range touched by this output slab: back-project its 8 corners
    double lo = 1e300, hi = -1e300;
    for (int k = 0; k < 8; ++k) {
      const double z = static_cast<double>((k & 1 ? z0 + cnt - 1 : z0) + c0[0]);
      const double y = static_cast<double>((k & 2 ? oy - 1 : 0) + c0[1]);
      const double x = static_cast<double>((k & 4 ? ox - 1 : 0) + c0[2]);
      const double cz = M[3] + z * M[0] + y * M[1] + x * M[2];
      lo = std::min(lo, cz);
      hi = std::max(hi, cz);
    }
    int64_t p_lo = static_cast<int64_t>(std::floor(std::max(lo, -1.0e15))) - 1;
    int64_t p_hi = static_cast<int64_t>(std::floor(std::min(hi, 1.0e15))) + 2;  // inclusive
    p_lo = std::max<int64_t>(p_lo, 0);
    p_hi = std::min<int64_t>(p_hi, sz - 1);
    Slab s;
    if (p_lo <= p_hi) {
      auto add = [&](int64_t a, int64_t b) {  // planes [a, b)
        if (a >= b) return;
        Band band;
        band.off = static_cast<size_t>(a) * plane_in;
        band.pitch = band.width = static_cast<size_t>(b - a) * plane_in;
        band.rows = 1;
        s.bands.push_back(band);
      };
      if (!any) {
        add(p_lo, p_hi + 1);
        up_lo = p_lo;
        up_hi = p_hi + 1;
        any = true;
      } else {
        if (p_lo < up_lo) {
          add(p_lo, up_lo);
          up_lo = p_lo;
        }
        if (p_hi + 1 > up_hi) {
          add(up_hi, p_hi + 1);
          up_hi = p_hi + 1;
        }
      }
    }
    s.out_off = static_cast<size_t>(z0) * plane_out;
    s.out_bytes = static_cast<size_t>(cnt) * plane_out;
    s.launch = [=](cudaStream_t st, char* d_out) {
      const int64_t crop[3] = {c0[0] + z0, c0[1], c0[2]};
      return affine_device(d_src, src_dtype, sz, sy, sx, reinterpret_cast<float*>(d_out), cnt, oy,
                           ox, M.data(), crop, order, boundary, scrub, B2_PATH_AUTO, st, 0, 0);
    };
    slabs.push_back(std::move(s));
  }
  return run_pipeline(*c, static_cast<const char*>(h_src), reinterpret_cast<char*>(h_dst), slabs);
}

// Chained deskew -> register with host buffers (BASELINE.json configs[4]):
// the deskewed float32 volume lives only on the device (16-byte aligned row pitch, TMA-eligible
// source of the warp).  One pipeline: the tilt-row band of deskew slab i is uploaded while slab
// i-1 is deskewed; as soon as the deskewed planes an output plane range needs exist, that range
// is warped and its download starts — upload, both kernels and download overlap.
int host_deskew_affine(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                       int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                       float pxct32, float off32, float* h_dst, int64_t oz, int64_t oy, int64_t ox,
                       const double* M12, const int64_t* crop_start, int order, int boundary,
                       int scrub, int device) {
  if (!h_src || !h_dst || !M12) {
    set_error("b2h_deskew_affine3d: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("b2h_deskew_affine3d: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (Zi < 2 || Yi < 1 || Xi < 1 || Zavg < 1 || Yo < 1 || Xo < 1 || N < 1 || Zo_full != Yi ||
      Zavg != (Zo_full + N - 1) / N || oz < 0 || oy < 0 || ox < 0) {
    set_error("b2h_deskew_affine3d: invalid shape");
    return B2_ERR_INVALID;
  }
  if (oz == 0 || oy == 0 || ox == 0) return B2_OK;
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t es = elem_size(src_dtype);
  const int64_t pitch = (Xo + 3) / 4 * 4;  // elements; 16-byte aligned rows
  const size_t in_bytes = static_cast<size_t>(Zi) * Yi * Xi * es;
  const size_t mid_bytes = static_cast<size_t>(Zavg) * Yo * pitch * sizeof(float);
  const size_t plane_out = static_cast<size_t>(oy) * ox * sizeof(float);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, in_bytes))) return rc;
  if ((rc = grow_device(&c->d_mid, &c->d_mid_bytes, mid_bytes))) return rc;
  void* d_src = c->d_src;
  float* d_mid = static_cast<float*>(c->d_mid);

  // need[z]: the last deskewed plane output plane z reads (+2: upper tap and a clamped neighbour)
  const int64_t c0[3] = {crop_start ? crop_start[0] : 0, crop_start ? crop_start[1] : 0,
                         crop_start ? crop_start[2] : 0};
  std::vector<double> M(M12, M12 + 12);
  std::vector<int64_t> need(static_cast<size_t>(oz));
  int64_t run_max = -1;
  for (int64_t z = 0; z < oz; ++z) {
    double hi = -1e300;
    for (int k = 0; k < 4; ++k) {
      const double y = static_cast<double>((k & 1 ? oy - 1 : 0) + c0[1]);
      const double x = static_cast<double>((k & 2 ? ox - 1 : 0) + c0[2]);
      hi = std::max(hi, M[3] + static_cast<double>(z + c0[0]) * M[0] + y * M[1] + x * M[2]);
    }
    int64_t p = static_cast<int64_t>(std::floor(std::min(std::max(hi, -1.0e15), 1.0e15))) + 2;
    p = std::min<int64_t>(std::max<int64_t>(p, -1), Zavg - 1);
    run_max = std::max(run_max, p);  // planes are released in order: prefix maximum
    need[static_cast<size_t>(z)] = run_max;
  }

  const size_t slice_bytes = static_cast<size_t>(Yo) * pitch * sizeof(float);
  const int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / slice_bytes));
  const size_t row_bytes = static_cast<size_t>(Xi) * es;
  std::vector<Slab> slabs;
  int64_t zdone = 0;
  for (int64_t a0 = 0; a0 < Zavg; a0 += per_slab) {
    const int64_t cnt = std::min(per_slab, Zavg - a0);
    const int64_t a1 = a0 + cnt;
    const int64_t iy_hi = Yi - 1 - a0 * N;
    const int64_t iy_lo = Yi - std::min<int64_t>(a1 * N, Yi);
    Slab s;
    Band b;
    b.off = static_cast<size_t>(iy_lo) * row_bytes;
    b.pitch = static_cast<size_t>(Yi) * row_bytes;
    b.width = static_cast<size_t>(iy_hi - iy_lo + 1) * row_bytes;
    b.rows = static_cast<size_t>(Zi);
    s.bands.push_back(b);
    // output planes whose deskewed source planes are complete after this slab
    int64_t znext = zdone;
    while (znext < oz && (need[static_cast<size_t>(znext)] < a1 || a1 == Zavg)) ++znext;
    const int64_t z0 = zdone, zc = znext - zdone;
    zdone = znext;
    s.out_off = static_cast<size_t>(z0) * plane_out;
    s.out_bytes = static_cast<size_t>(zc) * plane_out;
    s.launch = [=](cudaStream_t st, char* d_out) {
      const int slab[4] = {0, static_cast<int>(Yi), static_cast<int>(a0), static_cast<int>(cnt)};
      int r = deskew_device(d_src, src_dtype, Zi, Yi, Xi, d_mid + a0 * Yo * pitch, Zavg, Yo, Xo,
                            Zo_full, N, px32, pxct32, off32, B2_PATH_AUTO, st, slab, pitch);
      if (r || zc == 0) return r;
      const int64_t crop[3] = {c0[0] + z0, c0[1], c0[2]};
      return affine_device(d_mid, B2_DTYPE_F32, Zavg, Yo, Xo, reinterpret_cast<float*>(d_out), zc,
                           oy, ox, M.data(), crop, order, boundary, scrub, B2_PATH_AUTO, st, pitch, 0);
    };
    slabs.push_back(std::move(s));
  }
  return run_pipeline(*c, static_cast<const char*>(h_src), reinterpret_cast<char*>(h_dst), slabs);
}

// Flat-field correction with host buffers.  Phase 1: Y-bands of the source (all Z planes of a
// range of rows: one strided cudaMemcpy2DAsync each) are uploaded while the medians of the
// previous band are computed.  The pattern mean needs every median, so phase 2 starts after
// the last band: Z-slabs of the result are computed and downloaded, slab i+1 computing while slab
// i travels.  Same three-stream pipeline as the resamplers.
int host_flatfield(const void* h_src, int64_t Z, int64_t Y, int64_t X, void* h_dst, int dst_dtype,
                   int device) {
  if (!h_src || !h_dst) {
    set_error("b2h_flatfield_u16: null pointer");
    return B2_ERR_INVALID;
  }
  if (Z < 1 || Y < 1 || X < 1 || Z > 65535) {
    set_error("b2h_flatfield_u16: invalid shape (1 <= Z <= 65535)");
    return B2_ERR_INVALID;
  }
  if (dst_dtype != B2_DTYPE_F32 && dst_dtype != B2_DTYPE_F64) {
    set_error("b2h_flatfield_u16: dst_dtype must be float32 or float64");
    return B2_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t osz = dst_dtype == B2_DTYPE_F32 ? 4 : 8;
  const size_t plane_in = static_cast<size_t>(Y) * X * 2;
  const size_t plane_out = static_cast<size_t>(Y) * X * osz;
  const size_t ws_bytes = flatfield_workspace_bytes(Y, X);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, static_cast<size_t>(Z) * plane_in))) return rc;
  if ((rc = grow_device(&c->d_ws, &c->d_ws_bytes, ws_bytes))) return rc;
  // the apply kernel indexes the whole result volume: it stays resident (d_mid) and the slabs
  // are downloaded from there
  if ((rc = grow_device(&c->d_mid, &c->d_mid_bytes, static_cast<size_t>(Z) * plane_out))) return rc;
  void* d_src = c->d_src;
  void* d_dst = c->d_mid;
  void* d_ws = c->d_ws;

  std::vector<Slab> slabs;
  const size_t row_bytes = static_cast<size_t>(X) * 2;
  const int64_t rows_per_band =
      std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / (row_bytes * static_cast<size_t>(Z))));
  for (int64_t y0 = 0; y0 < Y; y0 += rows_per_band) {
    const int64_t cnt = std::min(rows_per_band, Y - y0);
    Slab s;
    Band b;
    b.off = static_cast<size_t>(y0) * row_bytes;
    b.pitch = plane_in;
    b.width = static_cast<size_t>(cnt) * row_bytes;
    b.rows = static_cast<size_t>(Z);
    s.bands.push_back(b);
    const bool first = y0 == 0;
    s.launch = [=](cudaStream_t st, char*) {
      int r = first ? flatfield_begin(Y, X, d_ws, st) : B2_OK;
      if (r) return r;
      return flatfield_median(d_src, Z, Y, X, d_ws, ws_bytes, y0 * X, cnt * X, st);
    };
    slabs.push_back(std::move(s));
  }
  const int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / plane_out));
  for (int64_t z0 = 0; z0 < Z; z0 += per_slab) {
    const int64_t cnt = std::min(per_slab, Z - z0);
    Slab s;
    s.out_off = static_cast<size_t>(z0) * plane_out;
    s.out_bytes = static_cast<size_t>(cnt) * plane_out;
    s.d_from = static_cast<const char*>(d_dst) + s.out_off;
    s.launch = [=](cudaStream_t st, char*) {
      return flatfield_apply(d_src, Z, Y, X, d_dst, dst_dtype, d_ws, ws_bytes, z0, cnt, st);
    };
    slabs.push_back(std::move(s));
  }
  return run_pipeline(*c, static_cast<const char*>(h_src), static_cast<char*>(h_dst), slabs);
}

// Cubic-spline warp (reference method="scipy") with host buffers: source planes are uploaded in
// ~48 MB pieces, the last piece triggers the three prefilter passes (they need the whole volume),
// then slabs of output planes are evaluated and downloaded, slab i+1 computing while slab i travels.
int host_affine_spline(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                       void* h_dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                       const double* M12, const int64_t* crop_start, int scrub, int device) {
  if (!h_src || !h_dst || !M12) {
    set_error("b2h_affine3d_spline3: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("b2h_affine3d_spline3: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (dst_dtype != B2_DTYPE_U16 && dst_dtype != B2_DTYPE_F32) {
    set_error("b2h_affine3d_spline3: dst_dtype must be uint16 or float32");
    return B2_ERR_INVALID;
  }
  if (sz < 1 || sy < 1 || sx < 1 || oz < 0 || oy < 0 || ox < 0) {
    set_error("b2h_affine3d_spline3: invalid shape");
    return B2_ERR_INVALID;
  }
  if (oz == 0 || oy == 0 || ox == 0) return B2_OK;
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t es = elem_size(src_dtype);
  const size_t plane_in = static_cast<size_t>(sy) * sx * es;
  const size_t ws_bytes = spline_workspace_bytes(sz, sy, sx);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, static_cast<size_t>(sz) * plane_in))) return rc;
  if ((rc = grow_device(&c->d_mid, &c->d_mid_bytes, ws_bytes))) return rc;
  void* d_src = c->d_src;
  void* d_ws = c->d_mid;
  const int64_t c0[3] = {crop_start ? crop_start[0] : 0, crop_start ? crop_start[1] : 0,
                         crop_start ? crop_start[2] : 0};
  std::vector<double> M(M12, M12 + 12);
  // the prefilter publishes the coefficient pointer for the evaluation slabs that follow it in
  // stream order (host-side hand-over inside this call)
  auto coef = std::make_shared<const double*>(nullptr);

  std::vector<Slab> slabs;
  const int64_t up = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / plane_in));
  for (int64_t p0 = 0; p0 < sz; p0 += up) {
    const int64_t cnt = std::min(up, sz - p0);
    Slab s;
    Band b;
    b.off = static_cast<size_t>(p0) * plane_in;
    b.pitch = b.width = static_cast<size_t>(cnt) * plane_in;
    b.rows = 1;
    s.bands.push_back(b);
    if (p0 + cnt >= sz)
      s.launch = [=](cudaStream_t st, char*) {
        return spline_prefilter_device(d_src, src_dtype, sz, sy, sx, scrub, d_ws, ws_bytes,
                                       coef.get(), st);
      };
    slabs.push_back(std::move(s));
  }
  const size_t osz = dst_dtype == B2_DTYPE_U16 ? 2 : 4;
  const size_t plane_out = static_cast<size_t>(oy) * ox * osz;
  const int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>(kSlabBytes / plane_out));
  for (int64_t z0 = 0; z0 < oz; z0 += per_slab) {
    const int64_t cnt = std::min(per_slab, oz - z0);
    Slab s;
    s.out_off = static_cast<size_t>(z0) * plane_out;
    s.out_bytes = static_cast<size_t>(cnt) * plane_out;
    s.launch = [=](cudaStream_t st, char* d_out) {
      const int64_t crop[3] = {c0[0] + z0, c0[1], c0[2]};
      return spline_eval_device(*coef, sz, sy, sx, d_out, dst_dtype, cnt, oy, ox, M.data(), crop, st);
    };
    slabs.push_back(std::move(s));
  }
  return run_pipeline(*c, static_cast<const char*>(h_src), static_cast<char*>(h_dst), slabs);
}

int host_release() {
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceGuard guard;
  for (int d = 0; d < kMaxDevices; ++d) {
    DeviceCtx& c = g_ctx[d];
    if (!c.init) continue;
    if (cudaSetDevice(d) != cudaSuccess) continue;
    cudaDeviceSynchronize();
    for (auto e : c.events) cudaEventDestroy(e);
    c.events.clear();
    if (c.d_src) cudaFree(c.d_src);
    if (c.d_dst) cudaFree(c.d_dst);
    if (c.d_ws) cudaFree(c.d_ws);
    if (c.d_mid) cudaFree(c.d_mid);
    for (int r = 0; r < kRing; ++r) {
      if (c.h_in[r]) cudaFreeHost(c.h_in[r]);
      if (c.h_out[r]) cudaFreeHost(c.h_out[r]);
    }
    cudaStreamDestroy(c.s_up);
    cudaStreamDestroy(c.s_run);
    cudaStreamDestroy(c.s_down);
    c = DeviceCtx();
  }
  return B2_OK;
}

}  // namespace b2
