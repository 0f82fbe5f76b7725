// C ABI of libbiahub_b200.so (see include/biahub_b200.h) + process-wide plumbing.
#include <cstring>
#include <mutex>

#include "b2_common.cuh"

namespace b2 {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return B2_ERR_CUDA_BASE + static_cast<int>(e);
}

EncodeTiledFn get_encode_tiled() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  });
  return fn;
}

int sm_count(int* out) {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B2_ERR_NO_DEVICE;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      return B2_ERR_NO_DEVICE;
    cached = n;
    cached_dev = dev;
  }
  *out = cached;
  return B2_OK;
}

// implemented in the kernel translation units
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
int host_deskew(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* h_dst,
                int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                float pxct32, float off32, int device);
int host_affine(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* h_dst,
                int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                int order, int boundary, int scrub, int device);
int host_deskew_fill(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                     float* h_dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N,
                     float px32, float pxct32, float off32, int fill_mode, float fill_value,
                     int device);
size_t spline_workspace_bytes(int64_t sz, int64_t sy, int64_t sx);
int affine_spline_device(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                         void* dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                         const double* M12, const int64_t* crop_start, int scrub, void* ws,
                         size_t ws_bytes, cudaStream_t st);
int host_affine_spline(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                       void* h_dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                       const double* M12, const int64_t* crop_start, int scrub, int device);
int fill_device_ex(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                   int iterations, int connectivity, void* ws, size_t ws_bytes, cudaStream_t stream);
int average_slices_device(const float* src, int64_t z, int64_t plane, int n, float* dst,
                          cudaStream_t stream);
int host_release();
unsigned long long brick_oob_count();
unsigned long long deskew_oob_count();
size_t flatfield_workspace_bytes(int64_t Y, int64_t X);
int flatfield_begin(int64_t Y, int64_t X, void* ws, cudaStream_t stream);
int flatfield_median(const void* src, int64_t Z, int64_t Y, int64_t X, void* ws, size_t ws_bytes,
                     int64_t p0, int64_t pn, cudaStream_t stream);
int flatfield_apply(const void* src, int64_t Z, int64_t Y, int64_t X, void* dst, int dst_dtype,
                    void* ws, size_t ws_bytes, int64_t z0, int64_t zn, cudaStream_t stream);
int host_deskew_affine(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                       int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int N, float px32,
                       float pxct32, float off32, float* h_dst, int64_t oz, int64_t oy, int64_t ox,
                       const double* M12, const int64_t* crop_start, int order, int boundary,
                       int scrub, int device);
int host_flatfield(const void* h_src, int64_t Z, int64_t Y, int64_t X, void* h_dst, int dst_dtype,
                   int device);

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device available (cudaGetDeviceCount: %s); biahub_b200 has no CPU fallback",
              e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    return B2_ERR_NO_DEVICE;
  }
  return B2_OK;
}

}  // namespace b2

extern "C" {

int b2_abi_version(void) { return B2_ABI_VERSION; }

const char* b2_last_error(void) { return b2::t_err; }

int b2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int b2_check_device(int dev) {
  int rc = b2::require_device();
  if (rc) return rc;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return b2::cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    b2::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major,
                  prop.minor);
    return B2_ERR_NO_DEVICE;
  }
  return B2_OK;
}

int b2_deskew(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* dst,
              int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int average_n_slices,
              float px32, float pxct32, float off32, int path, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::deskew_device(src, src_dtype, Zi, Yi, Xi, dst, Zavg, Yo, Xo, Zo_full,
                           average_n_slices, px32, pxct32, off32, path,
                           static_cast<cudaStream_t>(stream), nullptr, 0);
}

int b2_deskew_pitched(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                      float* dst, int64_t dst_row_pitch, int64_t Zavg, int64_t Yo, int64_t Xo,
                      int64_t Zo_full, int average_n_slices, float px32, float pxct32, float off32,
                      int path, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::deskew_device(src, src_dtype, Zi, Yi, Xi, dst, Zavg, Yo, Xo, Zo_full,
                           average_n_slices, px32, pxct32, off32, path,
                           static_cast<cudaStream_t>(stream), nullptr, dst_row_pitch);
}

int b2_affine3d(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx, float* dst,
                int64_t oz, int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                int order, int boundary, int scrub_nonfinite, int path, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::affine_device(src, src_dtype, sz, sy, sx, dst, oz, oy, ox, M12, crop_start, order,
                           boundary, scrub_nonfinite, path, static_cast<cudaStream_t>(stream), 0, 0);
}

int b2_affine3d_pitched(const void* src, int src_dtype, int64_t src_row_pitch, int64_t sz,
                        int64_t sy, int64_t sx, float* dst, int64_t dst_row_pitch, int64_t oz,
                        int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                        int order, int boundary, int scrub_nonfinite, int path, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::affine_device(src, src_dtype, sz, sy, sx, dst, oz, oy, ox, M12, crop_start, order,
                           boundary, scrub_nonfinite, path, static_cast<cudaStream_t>(stream),
                           src_row_pitch, dst_row_pitch);
}

size_t b2_overhang_fill_workspace(int64_t z, int64_t y, int64_t x) {
  return b2::fill_workspace_bytes(z, y, x);
}

int b2_overhang_fill(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                     int iterations, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::fill_device(vol, z, y, x, use_mean, fill_value, iterations, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream));
}

int b2h_deskew(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi, float* h_dst,
               int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int average_n_slices,
               float px32, float pxct32, float off32, int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_deskew(h_src, src_dtype, Zi, Yi, Xi, h_dst, Zavg, Yo, Xo, Zo_full,
                         average_n_slices, px32, pxct32, off32, device);
}

int b2h_affine3d(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                 float* h_dst, int64_t oz, int64_t oy, int64_t ox, const double* M12,
                 const int64_t* crop_start, int order, int boundary, int scrub_nonfinite,
                 int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_affine(h_src, src_dtype, sz, sy, sx, h_dst, oz, oy, ox, M12, crop_start, order,
                         boundary, scrub_nonfinite, device);
}

size_t b2_flatfield_workspace(int64_t y, int64_t x) { return b2::flatfield_workspace_bytes(y, x); }

int b2_flatfield_u16(const void* src, int64_t z, int64_t y, int64_t x, void* dst, int dst_dtype,
                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // argument checks happen in the first step; the accumulator is cleared only once they passed
  if ((rc = b2::flatfield_median(src, z, y, x, workspace, workspace_bytes, 0, 0, st))) return rc;
  if ((rc = b2::flatfield_begin(y, x, workspace, st))) return rc;
  if ((rc = b2::flatfield_median(src, z, y, x, workspace, workspace_bytes, 0, y * x, st))) return rc;
  return b2::flatfield_apply(src, z, y, x, dst, dst_dtype, workspace, workspace_bytes, 0, z, st);
}

int b2h_flatfield_u16(const void* h_src, int64_t z, int64_t y, int64_t x, void* h_dst, int dst_dtype,
                      int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_flatfield(h_src, z, y, x, h_dst, dst_dtype, device);
}

int b2h_deskew_affine3d(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                        int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int average_n_slices,
                        float px32, float pxct32, float off32, float* h_dst, int64_t oz, int64_t oy,
                        int64_t ox, const double* M12, const int64_t* crop_start, int order,
                        int boundary, int scrub_nonfinite, int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_deskew_affine(h_src, src_dtype, Zi, Yi, Xi, Zavg, Yo, Xo, Zo_full,
                                average_n_slices, px32, pxct32, off32, h_dst, oz, oy, ox, M12,
                                crop_start, order, boundary, scrub_nonfinite, device);
}

int b2_overhang_fill_ex(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                        int iterations, int connectivity, void* workspace, size_t workspace_bytes,
                        void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::fill_device_ex(vol, z, y, x, use_mean, fill_value, iterations, connectivity, workspace,
                            workspace_bytes, static_cast<cudaStream_t>(stream));
}

int b2_average_slices(const float* src, int64_t z, int64_t plane_elems, int n, float* dst,
                      void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::average_slices_device(src, z, plane_elems, n, dst, static_cast<cudaStream_t>(stream));
}

int b2h_deskew_fill(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                    float* h_dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full,
                    int average_n_slices, float px32, float pxct32, float off32, int fill_mode,
                    float fill_value, int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_deskew_fill(h_src, src_dtype, Zi, Yi, Xi, h_dst, Zavg, Yo, Xo, Zo_full,
                              average_n_slices, px32, pxct32, off32, fill_mode, fill_value, device);
}

size_t b2_spline3_workspace(int64_t sz, int64_t sy, int64_t sx) {
  return b2::spline_workspace_bytes(sz, sy, sx);
}

int b2_affine3d_spline3(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                        void* dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                        const double* M12, const int64_t* crop_start, int scrub_nonfinite,
                        void* workspace, size_t workspace_bytes, void* stream) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::affine_spline_device(src, src_dtype, sz, sy, sx, dst, dst_dtype, oz, oy, ox, M12,
                                  crop_start, scrub_nonfinite, workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
}

int b2h_affine3d_spline3(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                         void* h_dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                         const double* M12, const int64_t* crop_start, int scrub_nonfinite,
                         int device) {
  int rc = b2::require_device();
  if (rc) return rc;
  return b2::host_affine_spline(h_src, src_dtype, sz, sy, sx, h_dst, dst_dtype, oz, oy, ox, M12,
                                crop_start, scrub_nonfinite, device);
}

int b2h_release(void) { return b2::host_release(); }

uint64_t b2_debug_oob_count(void) {
  cudaDeviceSynchronize();
  return b2::brick_oob_count() + b2::deskew_oob_count();
}

int b2_debug_bounds_check_build(void) {
#ifdef B2_BOUNDS_CHECK
  return 1;
#else
  return 0;
#endif
}

uint64_t b2_launch_count(void) { return b2::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
