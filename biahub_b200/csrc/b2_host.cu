// Host-buffer pipeline behind b2h_deskew / b2h_affine3d: numpy-in / numpy-out callers
// (reference biahub/deskew.py:551-579, biahub/register.py:202-281, biahub/stabilize.py:32-90)
// hand over pageable or pinned HOST arrays; the source volume is made resident in HBM, resampled
// slab by slab and copied back while later slabs are still being computed.
//
//   upload  stream : chunked cudaMemcpyAsync H2D (one call per chunk; pageable sources are first
//                    copied into a pinned staging buffer by a small pool of host threads)
//   compute stream : one kernel launch per output slab, ordered after the chunks it needs
//   download stream: cudaMemcpyAsync D2H of each finished slab, un-staged by the host pool
//
// Per-process, per-device state (streams, pinned and device buffers) is cached and grown on
// demand; b2h_release() frees it.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "b2_common.cuh"

namespace b2 {

int deskew_device(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* dst,
                  int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                  float pxct32, float off32, int path, cudaStream_t stream, const int* slab);
int affine_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* dst,
                  int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                  int order, int boundary, int scrub, int path, cudaStream_t stream);

namespace {

constexpr size_t kChunkBytes = 32u << 20;  // granularity of the copy pipeline
constexpr int kMaxDevices = 16;

struct DeviceCtx {
  bool init = false;
  cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;
  void* d_src = nullptr;
  size_t d_src_bytes = 0;
  void* d_dst = nullptr;
  size_t d_dst_bytes = 0;
  void* h_in = nullptr;  // pinned staging
  size_t h_in_bytes = 0;
  void* h_out = nullptr;
  size_t h_out_bytes = 0;
  std::vector<cudaEvent_t> events;
};

std::mutex g_mu;  // b2h_* calls are serialised per process (one worker process per GPU)
DeviceCtx g_ctx[kMaxDevices];

int host_threads() {
  static int n = [] {
    unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? static_cast<int>(hc) : 4;
    const char* env = getenv("B2_HOST_THREADS");
    if (env && atoi(env) > 0) t = atoi(env);
    return std::max(1, std::min(t, 16));
  }();
  return n;
}

// memcpy split over a few threads (pageable <-> pinned staging runs at memory speed this way)
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  const int nt = host_threads();
  if (bytes < (8u << 20) || nt == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (bytes / nt + 4095) & ~static_cast<size_t>(4095);
  for (int t = 0; t < nt; ++t) {
    const size_t off = static_cast<size_t>(t) * per;
    if (off >= bytes) break;
    const size_t len = std::min(per, bytes - off);
    th.emplace_back([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
  }
  for (auto& t : th) t.join();
}

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

// H2D of `bytes` in chunks; returns after everything is ENQUEUED on c.s_up (pageable sources
// have been staged by then).  `ready` is recorded on s_up after the last chunk.
int upload(DeviceCtx& c, void* d_dst, const void* h_src, size_t bytes, cudaEvent_t ready) {
  const bool direct = is_pinned(h_src);
  if (!direct) {
    int rc = grow_pinned(&c.h_in, &c.h_in_bytes, bytes);
    if (rc) return rc;
  }
  for (size_t off = 0; off < bytes; off += kChunkBytes) {
    const size_t len = std::min(kChunkBytes, bytes - off);
    const char* from = static_cast<const char*>(h_src) + off;
    if (!direct) {
      parallel_memcpy(static_cast<char*>(c.h_in) + off, from, len);
      from = static_cast<const char*>(c.h_in) + off;
    }
    B2_CUDA(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, from, len, cudaMemcpyHostToDevice,
                            c.s_up));
  }
  B2_CUDA(cudaEventRecord(ready, c.s_up));
  return B2_OK;
}

// D2H of `bytes` (already ordered after the producing kernel on s_down); blocks until complete.
int download(DeviceCtx& c, void* h_dst, const void* d_src, size_t bytes, size_t ev_base) {
  const bool direct = is_pinned(h_dst);
  if (!direct) {
    int rc = grow_pinned(&c.h_out, &c.h_out_bytes, bytes);
    if (rc) return rc;
  }
  const size_t nchunks = (bytes + kChunkBytes - 1) / kChunkBytes;
  for (size_t i = 0; i < nchunks; ++i) {
    const size_t off = i * kChunkBytes;
    const size_t len = std::min(kChunkBytes, bytes - off);
    char* to = direct ? static_cast<char*>(h_dst) + off : static_cast<char*>(c.h_out) + off;
    B2_CUDA(cudaMemcpyAsync(to, static_cast<const char*>(d_src) + off, len, cudaMemcpyDeviceToHost,
                            c.s_down));
    cudaEvent_t ev;
    int rc = get_event(c, ev_base + i, &ev);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(ev, c.s_down));
  }
  for (size_t i = 0; i < nchunks; ++i) {
    cudaEvent_t ev;
    int rc = get_event(c, ev_base + i, &ev);
    if (rc) return rc;
    B2_CUDA(cudaEventSynchronize(ev));
    if (!direct) {
      const size_t off = i * kChunkBytes;
      const size_t len = std::min(kChunkBytes, bytes - off);
      parallel_memcpy(static_cast<char*>(h_dst) + off, static_cast<char*>(c.h_out) + off, len);
    }
  }
  return B2_OK;
}

size_t elem_size(int dtype) { return dtype == B2_DTYPE_U16 ? 2 : 4; }

}  // namespace

int host_deskew(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* h_dst,
                int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                float pxct32, float off32, int device) {
  if (!h_src || !h_dst) {
    set_error("b2h_deskew: null pointer");
    return B2_ERR_INVALID;
  }
  if (src_dtype != B2_DTYPE_U16 && src_dtype != B2_DTYPE_F32) {
    set_error("b2h_deskew: unknown src_dtype %d", src_dtype);
    return B2_ERR_INVALID;
  }
  if (Zi < 1 || Yi < 1 || Xi < 1 || Zavg < 1 || Yo < 1 || Xo < 1) {
    set_error("b2h_deskew: invalid shape");
    return B2_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t in_bytes = static_cast<size_t>(Zi) * Yi * Xi * elem_size(src_dtype);
  const size_t out_bytes = static_cast<size_t>(Zavg) * Yo * Xo * sizeof(float);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, in_bytes))) return rc;
  if ((rc = grow_device(&c->d_dst, &c->d_dst_bytes, out_bytes))) return rc;

  cudaEvent_t up_done, run_done;
  if ((rc = get_event(*c, 0, &up_done))) return rc;
  if ((rc = get_event(*c, 1, &run_done))) return rc;
  if ((rc = upload(*c, c->d_src, h_src, in_bytes, up_done))) return rc;
  B2_CUDA(cudaStreamWaitEvent(c->s_run, up_done, 0));

  // output slabs along the averaged-slice axis: each slab is contiguous in dst and can start
  // its D2H while the next slab is being computed
  const size_t slice_bytes = static_cast<size_t>(Yo) * Xo * sizeof(float);
  int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>((64u << 20) / slice_bytes));
  size_t ev_next = 2;
  std::vector<std::pair<int64_t, int64_t>> slabs;
  for (int64_t a0 = 0; a0 < Zavg; a0 += per_slab) slabs.emplace_back(a0, std::min(per_slab, Zavg - a0));
  std::vector<cudaEvent_t> slab_ev(slabs.size());
  for (size_t i = 0; i < slabs.size(); ++i) {
    const int slab[4] = {0, static_cast<int>(Yi), static_cast<int>(slabs[i].first),
                         static_cast<int>(slabs[i].second)};
    float* d_out = static_cast<float*>(c->d_dst) + slabs[i].first * Yo * Xo;
    rc = deskew_device(c->d_src, src_dtype, Zi, Yi, Xi, d_out, Zavg, Yo, Xo, Zo_full, N, px32,
                       pxct32, off32, B2_PATH_AUTO, c->s_run, slab);
    if (rc) return rc;
    if ((rc = get_event(*c, ev_next++, &slab_ev[i]))) return rc;
    B2_CUDA(cudaEventRecord(slab_ev[i], c->s_run));
  }
  (void)run_done;
  for (size_t i = 0; i < slabs.size(); ++i) {
    B2_CUDA(cudaStreamWaitEvent(c->s_down, slab_ev[i], 0));
    const size_t off = static_cast<size_t>(slabs[i].first) * slice_bytes;
    const size_t len = static_cast<size_t>(slabs[i].second) * slice_bytes;
    rc = download(*c, reinterpret_cast<char*>(h_dst) + off, static_cast<char*>(c->d_dst) + off, len,
                  ev_next);
    if (rc) return rc;
  }
  B2_CUDA(cudaStreamSynchronize(c->s_down));
  return B2_OK;
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
  DeviceCtx* c = nullptr;
  int rc = get_ctx(device, &c);
  if (rc) return rc;
  const size_t in_bytes = static_cast<size_t>(sz) * sy * sx * elem_size(src_dtype);
  const size_t out_bytes = static_cast<size_t>(oz) * oy * ox * sizeof(float);
  if ((rc = grow_device(&c->d_src, &c->d_src_bytes, in_bytes))) return rc;
  if ((rc = grow_device(&c->d_dst, &c->d_dst_bytes, out_bytes))) return rc;

  cudaEvent_t up_done;
  if ((rc = get_event(*c, 0, &up_done))) return rc;
  if ((rc = upload(*c, c->d_src, h_src, in_bytes, up_done))) return rc;
  B2_CUDA(cudaStreamWaitEvent(c->s_run, up_done, 0));

  const size_t plane_bytes = static_cast<size_t>(oy) * ox * sizeof(float);
  int64_t per_slab = std::max<int64_t>(1, static_cast<int64_t>((64u << 20) / plane_bytes));
  size_t ev_next = 1;
  std::vector<std::pair<int64_t, int64_t>> slabs;
  for (int64_t z0 = 0; z0 < oz; z0 += per_slab) slabs.emplace_back(z0, std::min(per_slab, oz - z0));
  std::vector<cudaEvent_t> slab_ev(slabs.size());
  for (size_t i = 0; i < slabs.size(); ++i) {
    int64_t crop[3] = {crop_start ? crop_start[0] : 0, crop_start ? crop_start[1] : 0,
                       crop_start ? crop_start[2] : 0};
    crop[0] += slabs[i].first;
    float* d_out = static_cast<float*>(c->d_dst) + slabs[i].first * oy * ox;
    rc = affine_device(c->d_src, src_dtype, sz, sy, sx, d_out, slabs[i].second, oy, ox, M12, crop,
                       order, boundary, scrub, B2_PATH_AUTO, c->s_run);
    if (rc) return rc;
    if ((rc = get_event(*c, ev_next++, &slab_ev[i]))) return rc;
    B2_CUDA(cudaEventRecord(slab_ev[i], c->s_run));
  }
  for (size_t i = 0; i < slabs.size(); ++i) {
    B2_CUDA(cudaStreamWaitEvent(c->s_down, slab_ev[i], 0));
    const size_t off = static_cast<size_t>(slabs[i].first) * plane_bytes;
    const size_t len = static_cast<size_t>(slabs[i].second) * plane_bytes;
    rc = download(*c, reinterpret_cast<char*>(h_dst) + off, static_cast<char*>(c->d_dst) + off, len,
                  ev_next);
    if (rc) return rc;
  }
  B2_CUDA(cudaStreamSynchronize(c->s_down));
  return B2_OK;
}

int host_release() {
  std::lock_guard<std::mutex> lock(g_mu);
  for (int d = 0; d < kMaxDevices; ++d) {
    DeviceCtx& c = g_ctx[d];
    if (!c.init) continue;
    if (cudaSetDevice(d) != cudaSuccess) continue;
    cudaDeviceSynchronize();
    for (auto e : c.events) cudaEventDestroy(e);
    c.events.clear();
    if (c.d_src) cudaFree(c.d_src);
    if (c.d_dst) cudaFree(c.d_dst);
    if (c.h_in) cudaFreeHost(c.h_in);
    if (c.h_out) cudaFreeHost(c.h_out);
    cudaStreamDestroy(c.s_up);
    cudaStreamDestroy(c.s_run);
    cudaStreamDestroy(c.s_down);
    c = DeviceCtx();
  }
  return B2_OK;
}

}  // namespace b2
