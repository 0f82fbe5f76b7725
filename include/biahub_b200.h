/*
 * biahub_b200.h — C ABI of the B200-native 3-D affine resampling path of biahub.
 *
 * One shared library (libbiahub_b200.so, built from biahub_b200/csrc/ for sm_100a) exports
 * exactly the entry points a binding for the reference's array-compute layer needs.  The
 * reference is pure Python, so the "FFI" it would bind is ctypes; each
 * entry point below names the reference function whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C: pointers, integers, floats; no C++/torch types.
 *  - b2_*  (device API): every data pointer is a DEVICE pointer owned by the caller (e.g. a torch
 *    tensor's data_ptr()).  The call enqueues work on `stream` (a cudaStream_t passed as void*,
 *    NULL = legacy default stream), never synchronises, allocates nothing persistent.
 *  - b2h_* (host API): data pointers are HOST pointers; the library stages through its own pinned
 *    ring buffers, overlaps H2D / kernel / D2H on side streams and returns when `dst` is complete.
 *  - volumes are C-contiguous (Z, Y, X); output is float32 unless an entry point says otherwise.
 *  - return value: 0 on success, otherwise a B2_ERR_* code (or a cudaError_t value + 1000);
 *    b2_last_error() returns a thread-local human-readable message for the last failure.
 *  - there is NO CPU fallback: without a usable sm_100 device every compute entry point fails.
 */
#ifndef BIAHUB_B200_H
#define BIAHUB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 2

#if defined(__GNUC__)
#define B2_API __attribute__((visibility("default")))
#else
#define B2_API
#endif

/* source element types (reference casts everything to float32: biahub/deskew.py:578,
 * biahub/register.py:266; uint16 is shipped as uint16 and converted in registers) */
#define B2_DTYPE_U16 0
#define B2_DTYPE_F32 1
#define B2_DTYPE_F64 2 /* flat-field results only */

/* boundary rule of the affine warp */
#define B2_BOUNDARY_CONSTANT 0 /* scipy.ndimage mode="constant", cval=0 (biahub/register.py:272) */
#define B2_BOUNDARY_ITK 1      /* ITK ResampleImageFilter rule = method="ants" (biahub/register.py:259-269) */

/* kernel selection (for tests/benchmarks; AUTO is what production uses) */
#define B2_PATH_AUTO 0
#define B2_PATH_GATHER 1 /* plain LDG gather: any shape/alignment */
#define B2_PATH_TMA 2    /* TMA-staged source bricks: fails with B2_ERR_UNSUPPORTED if ineligible */

#define B2_OK 0
#define B2_ERR_INVALID 1     /* bad argument */
#define B2_ERR_UNSUPPORTED 2 /* requested path not eligible for this shape/alignment */
#define B2_ERR_NO_DEVICE 3   /* no CUDA device / not sm_100 */
#define B2_ERR_CUDA_BASE 1000

B2_API int b2_abi_version(void);
B2_API const char* b2_last_error(void);
/* number of visible CUDA devices (0 when none); never fails */
B2_API int b2_device_count(void);
/* 0 if device `dev` exists and has compute capability 10.x */
B2_API int b2_check_device(int dev);

/*
 * Oblique-plane deskew: fused axis flip/transpose + 1-D fp32 lerp along the scan axis +
 * N-slice average.  Replaces the arithmetic of reference `fast_deskew_zyx`
 * (biahub/deskew.py:456-536: _rearrange_axes :99-110, edge pad :517-519, _build_deskew_grid
 * :113-154, F.grid_sample :531-533, mean :536).
 *
 *   src      (Zi, Yi, Xi)  uint16 or float32  — (scan, tilt, coverslip)
 *   dst      (Zavg, Yo, Xo) float32, Yo == Xi, Zavg == ceil(Zo_full / N), Zo_full == Yi
 *   px32, pxct32, off32: float32(px), float32(px*cos(theta)), float32(offset) computed on the host
 *                        in float64 exactly as biahub/deskew.py:136-138.
 */
B2_API int b2_deskew(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
              float* dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full,
              int average_n_slices, float px32, float pxct32, float off32,
              int path, void* stream);

/* Same as b2_deskew with an explicit output row pitch (elements, >= Xo; planes are Yo*pitch
 * apart).  A pitch that is a multiple of 4 keeps the rows 16-byte aligned so that a following
 * b2_affine3d_pitched can stage the deskewed volume with TMA (chained deskew -> register). */
B2_API int b2_deskew_pitched(const void* src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                      float* dst, int64_t dst_row_pitch, int64_t Zavg, int64_t Yo, int64_t Xo,
                      int64_t Zo_full, int average_n_slices, float px32, float pxct32, float off32,
                      int path, void* stream);

/*
 * Affine pull warp, order 0 (nearest) or 1 (trilinear), with NaN/inf scrub on load.
 * Replaces the arithmetic of reference `apply_affine_transform` (biahub/register.py:202-281:
 * nan_to_num :254, ANTs/ITK resample :259-269 or scipy :271-272, crop :278-279) and
 * `apply_stabilization_transform` (biahub/stabilize.py:32-90).
 *
 *   src        (sz, sy, sx) uint16 or float32
 *   dst        (oz, oy, ox) float32 = the CROPPED output box
 *   M12        row-major 3x4: source_index = M[:, :3] @ out_index + M[:, 3], out_index counted in
 *              the UNCROPPED output frame (biahub/register.py:161-162 convention)
 *   crop_start first uncropped output index of dst on each axis (NULL = {0,0,0})
 */
B2_API int b2_affine3d(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                float* dst, int64_t oz, int64_t oy, int64_t ox,
                const double* M12, const int64_t* crop_start,
                int order, int boundary, int scrub_nonfinite, int path, void* stream);

/* Same as b2_affine3d with explicit source / output row pitches in elements (0 = dense). */
B2_API int b2_affine3d_pitched(const void* src, int src_dtype, int64_t src_row_pitch, int64_t sz,
                        int64_t sy, int64_t sx, float* dst, int64_t dst_row_pitch, int64_t oz,
                        int64_t oy, int64_t ox, const double* M12, const int64_t* crop_start,
                        int order, int boundary, int scrub_nonfinite, int path, void* stream);

/*
 * Overhang fill (reference `_fill_overhang_torch`, biahub/deskew.py:339-368): mask = (vol == 0)
 * dilated `iterations` times with a 3x3x3 cube, then vol = mask ? fill : vol, where fill is the
 * fp32 mean of the un-masked voxels (use_mean != 0) or `fill_value`.
 * `workspace` must hold at least b2_overhang_fill_workspace(z,y,x) bytes of device memory.
 */
B2_API size_t b2_overhang_fill_workspace(int64_t z, int64_t y, int64_t x);
B2_API int b2_overhang_fill(float* vol, int64_t z, int64_t y, int64_t x, int use_mean, float fill_value,
                     int iterations, void* workspace, size_t workspace_bytes, void* stream);

/* Same with the structuring element chosen: connectivity 26 = 3x3x3 cube per iteration (the torch
 * variant above), 6 = 3-D cross per iteration = scipy.ndimage.binary_dilation's default, which the
 * numpy variant `_fill_overhang_with_mean` of the legacy `deskew_zyx` uses
 * (biahub/deskew.py:277-336, 448-451). */
B2_API int b2_overhang_fill_ex(float* vol, int64_t z, int64_t y, int64_t x, int use_mean,
                        float fill_value, int iterations, int connectivity, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Legacy averaging of the DESKEWED stack (reference `_average_n_slices_torch`,
 * biahub/deskew.py:71-96, called by `deskew_zyx` :438-440): dst[a] = mean of the n slices
 * src[a*n .. a*n+n-1] of plane_elems floats each, the last slice repeated to fill the last
 * group; dst holds ceil(z / n) slices. */
B2_API int b2_average_slices(const float* src, int64_t z, int64_t plane_elems, int n, float* dst,
                      void* stream);

/*
 * Host-buffer pipeline (what the reference's per-(t,c) callables see: numpy in, numpy out —
 * biahub/deskew.py:551-579, biahub/register.py:202-281).  The volume is split into slabs;
 * slab i+1 is copied H2D (one cudaMemcpyAsync per slab from a pinned ring buffer) while slab i
 * is resampled and slab i-1 is copied back.  `device` selects the GPU.
 */
B2_API int b2h_deskew(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
               float* h_dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full,
               int average_n_slices, float px32, float pxct32, float off32, int device);

B2_API int b2h_affine3d(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                 float* h_dst, int64_t oz, int64_t oy, int64_t ox,
                 const double* M12, const int64_t* crop_start,
                 int order, int boundary, int scrub_nonfinite, int device);

/* b2h_deskew followed by the overhang fill of the whole deskewed volume (reference
 * biahub/deskew.py:538-540: `keep_overhang and overhang_fill != 0` -> _fill_overhang_torch
 * :339-368).  fill_mode 0 = none (same as b2h_deskew), 1 = mean of the un-masked voxels,
 * 2 = fill_value.  Uploads and deskew slabs overlap; the fill runs once on the resident volume;
 * its slabs are then downloaded. */
B2_API int b2h_deskew_fill(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                    float* h_dst, int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full,
                    int average_n_slices, float px32, float pxct32, float off32, int fill_mode,
                    float fill_value, int device);

/*
 * Cubic B-spline affine pull warp = reference `apply_affine_transform(method="scipy")`
 * (biahub/register.py:271-272: scipy.ndimage.affine_transform with scipy's defaults order=3,
 * mode="constant", cval=0, prefilter on).  float64 prefilter (mirror initialisation, as scipy)
 * and float64 evaluation of the 4x4x4 mirror-extended taps; coordinates outside [0, n-1] -> 0.
 *
 *   src   (sz, sy, sx) uint16 or float32 (NaN/inf scrubbed on load when scrub_nonfinite)
 *   dst   (oz, oy, ox) uint16 (scipy's integer conversion: round half up, clamped) or float32
 *   workspace: b2_spline3_workspace(sz, sy, sx) bytes of device memory, 256-byte aligned
 *   M12 / crop_start as in b2_affine3d.
 */
B2_API size_t b2_spline3_workspace(int64_t sz, int64_t sy, int64_t sx);
B2_API int b2_affine3d_spline3(const void* src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                        void* dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                        const double* M12, const int64_t* crop_start, int scrub_nonfinite,
                        void* workspace, size_t workspace_bytes, void* stream);
B2_API int b2h_affine3d_spline3(const void* h_src, int src_dtype, int64_t sz, int64_t sy, int64_t sx,
                         void* h_dst, int dst_dtype, int64_t oz, int64_t oy, int64_t ox,
                         const double* M12, const int64_t* crop_start, int scrub_nonfinite,
                         int device);

/*
 * Chained unit of BASELINE configs[4]: deskew, then warp the deskewed float32 volume — the
 * reference runs `biahub deskew` then `biahub register` with a zarr round trip in between
 * (biahub/deskew.py:739-748, biahub/register.py:561-572).  The deskewed volume stays on the device;
 * upload, both kernels and the download of finished output planes overlap.  The matrix / output
 * shape refer to the DESKEWED volume (Zavg, Yo, Xo) exactly as in b2h_affine3d.
 */
B2_API int b2h_deskew_affine3d(const void* h_src, int src_dtype, int64_t Zi, int64_t Yi, int64_t Xi,
                        int64_t Zavg, int64_t Yo, int64_t Xo, int64_t Zo_full, int average_n_slices,
                        float px32, float pxct32, float off32, float* h_dst, int64_t oz, int64_t oy,
                        int64_t ox, const double* M12, const int64_t* crop_start, int order,
                        int boundary, int scrub_nonfinite, int device);

/*
 * Flat-field correction (reference biahub/flat_field.py:105-122 `flat_field_zyx`, :152-166
 * `_flat_field_czyx`; the pipeline stage before deskew): pattern = median over Z of every (y, x)
 * pixel, out = (double(src) / pattern) * mean(pattern), rounded to float32 (what the czyx adapter
 * stores) or kept as float64 (what `flat_field_zyx` returns).  uint16 sources, 1 <= z <= 65535.
 * Bit-identical to numpy.  `workspace`: b2_flatfield_workspace(y, x) bytes of device memory,
 * 256-byte aligned (holds the float32 pattern and the pattern sum).
 */
B2_API size_t b2_flatfield_workspace(int64_t y, int64_t x);
B2_API int b2_flatfield_u16(const void* src, int64_t z, int64_t y, int64_t x, void* dst, int dst_dtype,
                     void* workspace, size_t workspace_bytes, void* stream);
/* host-buffer variant: Y-bands are uploaded while the medians of the previous band are computed,
 * Z-slabs of the result are downloaded while the next slab is computed */
B2_API int b2h_flatfield_u16(const void* h_src, int64_t z, int64_t y, int64_t x, void* h_dst,
                      int dst_dtype, int device);

/* release the per-process pinned/device staging pools of the b2h_* calls */
B2_API int b2h_release(void);

/* Debug aid.  A library built with -DB2_BOUNDS_CHECK compares every computed shared-memory tap
 * address of the TMA deskew / generic-affine kernels with the extent of its brick and counts the
 * violations (current device; synchronises).  Release builds return 0 from both calls. */
B2_API uint64_t b2_debug_oob_count(void);
B2_API int b2_debug_bounds_check_build(void);

/* number of kernels this library has launched in the calling process (all threads) */
B2_API uint64_t b2_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BIAHUB_B200_H */
