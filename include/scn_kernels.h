/*
 * scn_kernels.h -- C-ABI kernel layer of scanner-b200 (libscn_kernels.so).
 *
 * The thin layer the C++ host pipeline (and any foreign binding: ctypes, cgo, JNI ...) calls
 * to run the per-frame pixel ops of Scanner's decode->evaluate->save hot path on a B200.
 * Plain pointers and sizes only -- no torch, no C++ types.  Each entry point names the
 * reference interface it replaces (paths relative to scanner-research/scanner @ 04a0c4b).
 *
 * Conventions
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every call
 *     only ENQUEUES work on that stream and returns; nothing synchronises, nothing allocates.
 *   - Pointers are DEVICE pointers on the current CUDA device unless the parameter name
 *     starts with `host_`.  `host_*_ptrs` arrays are read synchronously during the call
 *     (they are copied into the kernel's launch parameters) and may be freed on return.
 *   - Frames are dense HWC, no row padding, exactly Scanner's `Frame` layout
 *     (scanner/api/frame.h:34-82: shape = {H, W, C}, u8* data).
 *   - Return value: 0 on success, a positive cudaError_t from the launch, or a negative
 *     SCN_E_* for bad arguments.  Never throws, never aborts.
 *   - There is NO CPU fallback: on a box without a usable CUDA device every launch returns
 *     the CUDA error.
 */
#ifndef SCN_KERNELS_H_
#define SCN_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SCN_API __attribute__((visibility("default")))
#else
#define SCN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SCN_E_BADARG (-1)
#define SCN_E_UNSUPPORTED (-2)
#define SCN_E_PLAN (-3)

/* Library/ABI version (bumped on any signature change). */
SCN_API int scn_abi_version(void);

/* Number of kernels this library has launched since load (all streams, this process).
 * bench.py reports it as `gpu_launches`. */
SCN_API uint64_t scn_launch_count(void);

/* Per-kernel device timing (CUDA events recorded on the launch stream around every kernel this
 * library launches).  Off by default; bench.py turns it on for the roofline leg only.
 *   scn_prof_enable(1|0)  start / stop recording (resets the records when turned on)
 *   scn_prof_report(buf, cap)  synchronises the recorded events and writes a JSON object
 *       {"<kernel>": {"launches": n, "ms": total_ms}, ...}; returns the length written
 *       (0 if nothing was recorded, negative on error). */
SCN_API void scn_prof_enable(int on);
SCN_API int scn_prof_report(char* host_buf, size_t cap);

/* ---------------------------------------------------------------------------------------
 * Histogram  (replaces HistogramKernelCPU::execute, tests/test_ops.cpp:19-49:
 *             cv::calcHist 16 bins x 3 channels, CV_32S, channel-major)
 * For each of `n` u8 HWC 3-channel frames writes int32[3][16] (192 bytes) to out + i*48.
 * `out` is fully overwritten (no pre-zeroing needed).  n may be 0.  Bit-exact.
 */
SCN_API int scn_hist16_u8c3(const uint8_t* const* host_frame_ptrs, int n, int width, int height,
                    int32_t* out, void* stream);
/* Same, frames at base + i*stride_bytes (a block from new_block_buffer_size,
 * scanner/util/memory.h:59-62). */
SCN_API int scn_hist16_u8c3_strided(const uint8_t* base, size_t stride_bytes, int n, int width,
                            int height, int32_t* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Resize  (replaces ResizeKernel::execute, tests/test_ops.cpp:124-162: cv::resize(img, out,
 *          Size(w,h)) == INTER_LINEAR on CV_8UC3, OpenCV's 11-bit fixed-point path,
 *          including its exact-2x INTER_AREA re-route).  Bit-exact vs cv2.resize.
 *
 * scn_resize_target: the output size rule of test_ops.cpp:126-147 (ResizeArgs width, height,
 * min, preserve_aspect; `interpolation` is ignored by the reference).
 *
 * A "plan" is the per-(src size, dst size) coefficient table (x/y tap offsets and 11-bit
 * weights).  The caller owns its storage: ask for the size, fill it on the host, copy it to the
 * device once, pass the device copy to every launch with those sizes.
 */
SCN_API void scn_resize_target(int src_w, int src_h, int arg_w, int arg_h, int arg_min,
                       int arg_preserve_aspect, int* out_w, int* out_h);
SCN_API size_t scn_resize_plan_bytes(int dst_w, int dst_h);
SCN_API int scn_resize_plan_fill(int src_w, int src_h, int dst_w, int dst_h, void* host_plan);
SCN_API int scn_resize_bilinear_u8c3(const uint8_t* const* host_src_ptrs, int n, int src_w, int src_h,
                             uint8_t* const* host_dst_ptrs, int dst_w, int dst_h,
                             const void* plan, void* stream);
SCN_API int scn_resize_bilinear_u8c3_strided(const uint8_t* src, size_t src_stride, int n, int src_w,
                                     int src_h, uint8_t* dst, size_t dst_stride, int dst_w,
                                     int dst_h, const void* plan, void* stream);

/* ---------------------------------------------------------------------------------------
 * Blur  (replaces BlurKernel::execute, tests/test_ops.cpp:265-294: integer box filter of
 *        BlurArgs.kernel_size, interior pixels only; sigma is unused by the reference).
 * Border pixels (which the reference leaves uninitialised) are written as 0.
 * kernel_size in [1, 31].  Bit-exact on the interior.
 */
SCN_API int scn_box_blur_u8c3(const uint8_t* const* host_src_ptrs, int n, int width, int height,
                      int kernel_size, uint8_t* const* host_dst_ptrs, void* stream);
SCN_API int scn_box_blur_u8c3_strided(const uint8_t* src, size_t stride_bytes, int n, int width,
                              int height, int kernel_size, uint8_t* dst, void* stream);

/* ---------------------------------------------------------------------------------------
 * NV12 -> RGB24  (replaces convertNV12toRGBA, scanner/util/image.cu:229-239 and its kernel
 *                 :109-200; called from NVIDIAVideoDecoder::get_frame,
 *                 scanner/video/nvidia/nvidia_video_decoder.cpp:288-296)
 * One decoder surface per frame: luma plane `pitch` x height, interleaved CbCr plane
 * `pitch` x height/2.  width and height must be even.  Output dense RGB24 (rgb_pitch >= 3*w).
 * Bit-exact vs the reference arithmetic (float matrix, clamp, truncate, >>2).
 */
SCN_API int scn_nv12_to_rgb24(const uint8_t* const* host_luma_ptrs,
                      const uint8_t* const* host_chroma_ptrs, size_t pitch, int n, int width,
                      int height, uint8_t* const* host_rgb_ptrs, size_t rgb_pitch,
                      void* stream);

/* Pitched decoder surface -> packed NV12 frame element: `width` x `height` luma rows followed by
 * `width` x height/2 interleaved CbCr rows, no padding (the FrameLayout::NV12 element of
 * include/scanner/api/frame.h; what the decode stage hands to kernels that registered for it
 * instead of the RGB24 frame of nvidia_video_decoder.cpp:288-296).  Byte copy. */
SCN_API int scn_nv12_pack(const uint8_t* const* host_luma_ptrs, const uint8_t* const* host_chroma_ptrs,
                          size_t pitch, int n, int width, int height, uint8_t* const* host_dst_ptrs,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused decode-side DAG of BASELINE.json configs[1]:
 *      NV12 surface -> RGB24 -> { Histogram , Resize(dst_w, dst_h) }
 * without materialising the RGB24 frame (what the reference does in three passes:
 * image.cu NV12_to_RGB, then test_ops.cpp Histogram and Resize kernels over the RGB frame).
 * hist_out: int32[n][3][16] (fully overwritten; NULL skips the histogram); resized: n dense RGB24
 * dst_h x dst_w frames at host_dst_ptrs[i] (NULL skips the resize) -- so the same entry point is
 * also "Histogram on an NV12 frame" and "Resize on an NV12 frame".  Results are bit-identical to
 * running scn_nv12_to_rgb24 + scn_hist16_u8c3 + scn_resize_bilinear_u8c3.
 */
SCN_API int scn_nv12_hist_resize(const uint8_t* const* host_luma_ptrs,
                         const uint8_t* const* host_chroma_ptrs, size_t pitch, int n,
                         int width, int height, int32_t* hist_out,
                         uint8_t* const* host_dst_ptrs, int dst_w, int dst_h, const void* plan,
                         void* stream);

/* ---------------------------------------------------------------------------------------
 * OpticalFlow  (replaces OpticalFlowKernelCPU::execute, tests/test_ops.cpp:79-98:
 *   cvtColor(COLOR_BGR2GRAY) on each RGB frame of the {0,1} stencil window, then
 *   cv::FarnebackOpticalFlow(numLevels, pyrScale, fastPyramids=false, winSize, numIters, polyN,
 *   polySigma, flags=0)->calc; the reference constructs it with (3, 0.5, 15, 3, 5, 1.2), :68-69)
 * For each of n frame pairs writes a dense flow field float32[H][W][2] (dx, dy) to
 * host_flow_ptrs[i].  Floating point: matches cv2.calcOpticalFlowFarneback within a tolerance
 * (max |d| <= 2e-3 px on well-conditioned input; see tests/test_flow_gpu.py), not bit-exact.
 * `workspace` is caller-owned device scratch of scn_farneback_workspace_bytes(w, h) bytes,
 * reusable across calls on the same stream.
 */
SCN_API size_t scn_farneback_workspace_bytes(int width, int height);
SCN_API int scn_farneback_u8c3(const uint8_t* const* host_prev_ptrs, const uint8_t* const* host_next_ptrs,
                               int n, int width, int height, float* const* host_flow_ptrs, int num_levels,
                               double pyr_scale, int win_size, int num_iters, int poly_n, double poly_sigma,
                               void* workspace, size_t workspace_bytes, void* stream);
/* The same for a caller that walks a clip pair by pair and keeps the workspace between calls (the OpticalFlow op,
 * stencil [0, 1]: one pair per call).  Within a call, a pair whose `prev` pointer equals the previous pair's `next`
 * reuses that frame's pyramid of polynomial expansions (gray conversion, the full-resolution Gaussians, resizes and
 * expansions are per frame, about 40 % of a pair's work).  `chain` (host int owned by the caller with the workspace,
 * 0 initially and after anything else used the workspace) carries this over calls: with reuse_prev != 0 and
 * *chain != 0 the caller asserts that pair 0's `prev` frame has the content of the last call's last `next` frame (the
 * same table row), and it is not expanded again.  Results are identical to scn_farneback_u8c3's. */
SCN_API int scn_farneback_u8c3_chain(const uint8_t* const* host_prev_ptrs, const uint8_t* const* host_next_ptrs,
                                     int n, int width, int height, float* const* host_flow_ptrs, int num_levels,
                                     double pyr_scale, int win_size, int num_iters, int poly_n, double poly_sigma,
                                     void* workspace, size_t workspace_bytes, int reuse_prev, int* chain,
                                     void* stream);

/* ---------------------------------------------------------------------------------------
 * FrameDigest: a 16-byte fingerprint of each of n equally sized buffers (`bytes` a multiple of 4):
 *     out[2i]   = sum of the buffer's little-endian 32-bit words            (mod 2^64)
 *     out[2i+1] = sum of word[k] * (k mod 65521 + 1)                         (mod 2^64)
 * Exact integer arithmetic on the raw bytes: two runs agree iff (up to collisions) their buffers are
 * byte-identical.  No counterpart in the reference; it exists so that frame-valued results too large to
 * bring back (a 1080p flow field is 16.6 MB per row) can be compared between a run on one GPU and a run
 * sharded over several (bench.py --config 3, SURVEY 8d configs[3] "must be identical").  `out` is fully
 * overwritten. */
SCN_API int scn_frame_digest(const uint8_t* const* host_ptrs, int n, size_t bytes, uint64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCN_KERNELS_H_ */
