/*
 * comfystereo_b200.h -- C ABI of libcomfystereo_b200.so
 *
 * B200 (sm_100a) implementation of ComfyStereo's depth-image-based stereo hot path.
 * Every entry point is extern "C", takes plain pointers and sizes (no torch types), is
 * asynchronous on the caller's cudaStream_t (passed as void*), returns 0 on success or a
 * negative cs_status, and never throws.  cs_last_error() gives the thread-local message.
 *
 * "Replaces" cites the reference interface each call stands in for:
 *   SIG = /root/reference/stereoimage_generation.py,  GS = /root/reference/GenerateStereo.py
 *
 * Layouts (ComfyUI conventions, GS:126-171, GS:298-307), all float32, dense:
 *   image      [n][h][w][3]      0..1
 *   depth      [n][h][w][c]      c = 1 or 3 (any c >= 1 for GPU Warp: channel 0 is used)
 *   stereo     [n][ho][wo][3]    (ho,wo) = (h,2w) SBS, (2h,w) top/bottom, (h,w) anaglyph / single eye
 *   depth_l/_r [n][h][w][3]
 *   mask       [n][hm][wm]       CPU techniques: (hm,wm) = (ho,wo); GPU Warp: (h,w)
 */
#ifndef COMFYSTEREO_B200_H
#define COMFYSTEREO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_ABI_VERSION 4   /* 4: CS_FILL_GPU_WARP_MESH, cs_forward_warp_mesh, *_scratch_bytes; 3: progress callback, blur_flavor */

#if defined(__GNUC__)
#define CS_API __attribute__((visibility("default")))
#else
#define CS_API
#endif

typedef enum cs_status {
    CS_OK = 0,
    CS_ERR_ARG = -1,     /* bad argument (null pointer, size, enum out of range) */
    CS_ERR_CUDA = -2,    /* a CUDA runtime call or kernel launch failed */
    CS_ERR_DEVICE = -3,  /* no sm_100 device / wrong architecture */
    CS_ERR_WORKSPACE = -4, /* workspace too small */
    CS_ERR_UNSUPPORTED = -5, /* parameter combination the reference also rejects (e.g. blur box = 0) */
    CS_ERR_MODE = -6     /* unknown composition mode: reference raises, SIG:1120 / SIG:1562 */
} cs_status;

/* fill_technique keys, GS:88-100 / dispatch SIG:1605-1618 */
typedef enum cs_fill {
    CS_FILL_NONE = 0,            /* 'none'                 SIG:1909-1910 */
    CS_FILL_NAIVE = 1,           /* 'naive'                SIG:1893-1908 */
    CS_FILL_NAIVE_INTERP = 2,    /* 'naive_interpolating'  SIG:1871-1892 */
    CS_FILL_POLYLINES_SOFT = 3,  /* 'polylines_soft'       SIG:1912-1992 */
    CS_FILL_POLYLINES_SHARP = 4, /* 'polylines_sharp'      SIG:1912-1992 */
    CS_FILL_INVERSE = 5,         /* 'inverse'              SIG:1715-1737 */
    CS_FILL_HYBRID_EDGE = 6,     /* 'hybrid_edge'          SIG:1837-1848 */
    CS_FILL_GPU_WARP = 7,        /* 'gpu_warp' = forward_warp_gpu, SIG:277-450 */
    /* reachable through create_stereoimages(fill_technique=...) and three dropdown names the node still maps
     * (GS:97-99) but no longer lists: */
    CS_FILL_NONE_POST = 8,       /* 'none_post'         SIG:1804-1818: naive mapping + np.interp row fill */
    CS_FILL_INVERSE_POST = 9,    /* 'inverse_post'      SIG:1820-1833: reverse projection + np.interp row fill */
    CS_FILL_HYBRID_EDGE_PLUS = 10, /* 'hybrid_edge_plus' SIG:1778-1802: hybrid edge, black pixels from polylines_soft */
    /* 'GPU Warp (Fast)' on a host with ModernGL: forward_warp_mesh, SIG:453-689 (selected at SIG:1068-1071).  The
     * reference rasterises the mesh with OpenGL, whose fragment coverage and interpolation rounding are
     * implementation-defined; this is a software rasteriser with a fixed rule set (oracle/stereo_oracle.c:orc_mesh_raster,
     * bit-exact against it).  Everything else -- blur, sub-batch coupling, composition, mask -- is CS_FILL_GPU_WARP's. */
    CS_FILL_GPU_WARP_MESH = 11
} cs_fill;

/* composition modes, SIG:1543-1562 / SIG:1093-1120 */
typedef enum cs_mode {
    CS_MODE_LEFT_RIGHT = 0,
    CS_MODE_RIGHT_LEFT = 1,
    CS_MODE_TOP_BOTTOM = 2,
    CS_MODE_BOTTOM_TOP = 3,
    CS_MODE_RED_CYAN = 4,
    CS_MODE_LEFT_ONLY = 5,
    CS_MODE_ONLY_RIGHT = 6,
    CS_MODE_CYAN_RED = 7
} cs_mode;

/* The node's widget values (GS:61-71) as the python host received them, plus the two integers
 * python derives from depth_blur_strength (SIG:1208-1209; banker's rounding stays host-side). */
typedef struct cs_params {
    int32_t fill;               /* cs_fill */
    int32_t mode;               /* cs_mode */
    double divergence;          /* percent of width */
    double separation;          /* percent of width */
    double stereo_balance;
    double convergence_point;
    double stereo_offset_exponent;
    int32_t blur_enabled;       /* depth_map_blur && strength > 0 */
    int32_t blur_box;           /* bs = int(round(depth_blur_strength)), must be >= 1 */
    int32_t blur_radius;        /* R  = int(depth_blur_strength) */
    int32_t blur_vert_smooth;   /* v, 0..15 */
    double blur_edge_threshold;
    double blur_falloff;
    int32_t group_size;         /* GPU Warp only: frames per reference sub-batch (batch_size, GS:119);
                                   the "is depth 0..1 or 0..255" tests are sub-batch wide (SIG:1045, 315, 1125) */
    int32_t depth_h;            /* size of the depth frames when it differs from the image's (GS:141-148, GS:214-220): */
    int32_t depth_w;            /* the gray depth is resized bilinearly (align_corners=False) first; 0 = same as the image */
    int32_t blur_flavor;        /* 0: directional_motion_blur_gpu (torch, zero padding; what the node runs), SIG:1171-1251
                                   1: directional_motion_blur (scipy: reflected Sobel, nearest-border float64 box sums), the
                                      blur create_stereoimages applies to non-tensor inputs, SIG:1346-1419 */
} cs_params;

CS_API int cs_abi_version(void);
CS_API const char *cs_last_error(void);

/* Verifies that the current CUDA device is compute capability 10.x; the library has no
 * other code path.  Returns CS_ERR_DEVICE otherwise. */
CS_API int cs_device_check(void);

/* Output geometry for (fill, mode, h, w). */
CS_API int cs_output_dims(const cs_params *p, int h, int w, int *ho, int *wo, int *hm, int *wm);

/* Bytes of device scratch cs_stereo_batch needs for `chunk` frames in flight. */
CS_API size_t cs_workspace_bytes(const cs_params *p, int chunk, int h, int w);

/* ---- stage entry points (used by the stage-wise parity tests and by cs_stereo_batch) ---- */

/* N1 + L1.  Replaces GS:201-212 / GS:134-139 (RGB -> gray) and the per-frame min/max that
 * SIG:1475 / SIG:1045 / SIG:1587-1588 need.  gray [n][h][w]; minmax [n][2] = {min, max}. */
CS_API int cs_depth_prepare(const float *depth, int n, int h, int w, int c, float *gray, float *minmax,
                     void *stream);

/* N1 resize.  Replaces the gray conversion + torch.nn.functional.interpolate(gray, size=(h, w),
 * mode='bilinear', align_corners=False) of GS:132-148 / GS:206-220 for a depth batch whose frames are
 * dh x dw: depth [n][dh][dw][c] -> gray [n][h][w].  Arithmetic: torch's CPU kernel in strict float32
 * (see cs_prep.cu); c == 3 is converted to gray per tap, other channel counts use channel 0. */
CS_API int cs_depth_resize(const float *depth, int n, int dh, int dw, int c, int h, int w, float *gray,
                    void *stream);

/* B1.  Replaces directional_motion_blur_gpu(depth, s, thr, s, falloff, vert), SIG:1171-1251.
 * depth255 [n][h][w] is on the 0..255 scale; writes blur_l / blur_r [n][h][w] and, if minmax
 * is not NULL, per frame {min_l, max_l, min_r, max_r}.  dist_scratch: 2*n*h*w bytes. */
CS_API int cs_blur(const float *depth255, int n, int h, int w, const cs_params *p, float *blur_l,
            float *blur_r, float *minmax, uint8_t *dist_scratch, void *stream);

/* Debug export of the integer shift indices the bit-exact contract is stated on.
 * nd [n][h][w] is the normalised, convergence-shifted depth (SIG:1594-1600).
 * kind 0: naive col_d = col + int(off + sep)  (SIG:1865);
 * kind 1: j = floor(col + 0.5 + off + sep)    (SIG:1638-1639, SIG:1725-1727). */
CS_API int cs_shift_indices(const float *nd, int n, int h, int w, double div_px, double sep_px,
                     double exponent, int kind, int32_t *out, void *stream);

/* D1 + W1/F0/F1/F2/P/I/H1/H2 for ONE eye.  Replaces apply_stereo_divergence(original_image,
 * depth, divergence, separation, exponent, fill, convergence), SIG:1576-1620, for the CPU
 * techniques.  image_u8 [n][h][w][4] (RGBX, as produced by cs_quantize_image); depth [n][h][w]
 * float32; divergence/separation are the SIGNED per-eye percentages the reference passes
 * (left: +div*(1+bal), -sep; right: -div*(1-bal), +sep).  out_u8 [n][h][w][4].
 * scratch: cs_workspace_bytes(p, n, h, w) bytes. */
CS_API int cs_warp_fill(const uint8_t *image_u8, const float *depth, int n, int h, int w, int fill,
                 double divergence, double separation, double exponent, double convergence,
                 uint8_t *out_u8, void *scratch, size_t scratch_bytes, void *stream);

/* Bytes of scratch cs_warp_fill needs. */
CS_API size_t cs_warp_fill_scratch_bytes(int n, int h, int w);

/* G1 for ONE eye at function level.  Replaces forward_warp_gpu(image, depth, divergence_px,
 * separation_px, stereo_offset_exponent, convergence_point), SIG:277-450 (the "GPU Warp (Fast)"
 * warp when moderngl is absent).  image [n][h][w][3] float32 (NHWC; the reference's NCHW permute is
 * a view), depth [n][h][w] as given (divided by 255 when any frame's max > 1, SIG:314-316).
 * warped [n][h][w][3], mask [n][h][w] = the PRE-fill "unfilled" map (1.0 / 0.0).
 * scratch: cs_forward_warp_scratch_bytes(n, h, w) -- the per-frame statistics, plus global row state for rows too wide
 * for a CTA's shared memory (beyond ~9200 columns; up to 24000). */
CS_API int cs_forward_warp(const float *image, const float *depth, int n, int h, int w, double div_px,
                    double sep_px, double exponent, double convergence, float *warped, float *mask,
                    void *scratch, size_t scratch_bytes, void *stream);
CS_API size_t cs_forward_warp_scratch_bytes(int n, int h, int w);

/* The same for the ModernGL variant: replaces forward_warp_mesh(image, depth, divergence_px, separation_px,
 * stereo_offset_exponent, convergence_point), SIG:453-689 -- the mesh of per-pixel vertices, culled with the
 * reference's gradient rule over the whole batch (SIG:522-537), drawn by a software rasteriser with a fixed rule set
 * (see CS_FILL_GPU_WARP_MESH), gaps smeared from the eye's fill side (SIG:655-683).  Same argument meaning and layouts
 * as cs_forward_warp; scratch: cs_forward_warp_mesh_scratch_bytes(n, h, w). */
CS_API int cs_forward_warp_mesh(const float *image, const float *depth, int n, int h, int w, double div_px,
                    double sep_px, double exponent, double convergence, float *warped, float *mask,
                    void *scratch, size_t scratch_bytes, void *stream);
CS_API size_t cs_forward_warp_mesh_scratch_bytes(int n, int h, int w);

/* O1 input side.  Replaces SIG:1506-1508: clip(x*255, 0, 255).astype(uint8) (truncation). */
CS_API int cs_quantize_image(const float *image, int n, int h, int w, uint8_t *image_u8, void *stream);

/* C1 + M1 + O1.  Replaces SIG:1543-1562, GS:355-361, GS:365-378 for the CPU techniques:
 * composes the two uint8 eyes, converts to float32 /255 and emits the "pure black" mask. */
CS_API int cs_compose(const uint8_t *left_u8, const uint8_t *right_u8, int n, int h, int w, int mode,
               float *stereo, float *mask, void *stream);

/* ---- the hot path in one call ---- */

/* Replaces the body of StereoImageNode.generate (GS:117-269) for n frames that are already
 * resident on the device: prep, blur, per-eye warp + fill, composition, depth outputs, mask.
 * All pointers are device pointers owned by the caller.  workspace >= cs_workspace_bytes(p,
 * chunk, h, w) for some chunk >= 1; the call picks the largest chunk that fits.
 * Asynchronous on `stream`.  Small jobs (n*h*w <= two 1080p frames) that repeat with the SAME pointers, sizes and
 * parameters -- a streaming caller reusing its buffers -- are captured into a CUDA graph on the second call and replayed
 * with one launch afterwards (the graph holds the pointers, not the contents; COMFYSTEREO_GRAPHS=0 disables it,
 * cs_host_release drops the cached graphs). */
CS_API int cs_stereo_batch(const cs_params *p, const float *image, const float *depth, int n, int h,
                    int w, int c, float *stereo, float *depth_l, float *depth_r, float *mask,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Same, with HOST buffers (the tensors ComfyUI hands the node live on the CPU, GS:126,
 * GS:161-171): stages chunks through pinned memory, overlaps H2D, kernels and D2H on internal
 * streams, returns when the outputs are complete in host memory.  `device` selects the GPU. */
CS_API int cs_stereo_batch_host(const cs_params *p, const float *image, const float *depth, int n,
                         int h, int w, int c, float *stereo, float *depth_l, float *depth_r,
                         float *mask, int device);

/* Same, reporting progress: `progress(frames, user)` is called on the calling thread, in frame order, each time
 * another chunk of `frames` frames is complete in the caller's output buffers.  Replaces the reference's
 * pbar.update(actual_batch_size) per sub-batch (GS:173) and pbar.update(1) per frame (GS:262); progress may be NULL. */
typedef void (*cs_progress_fn)(int frames, void *user);
CS_API int cs_stereo_batch_host_progress(const cs_params *p, const float *image, const float *depth, int n,
                                  int h, int w, int c, float *stereo, float *depth_l, float *depth_r,
                                  float *mask, int device, cs_progress_fn progress, void *user);

/* Frees the device buffers / streams cs_stereo_batch_host caches between calls. */
CS_API void cs_host_release(void);

/* 1 when cs_stereo_batch_host would currently move the depth outputs (three identical channels) and the
 * mask (0/1) over PCIe compacted -- one channel, one byte per pixel -- and expand them on the host: it does
 * when this process may use >= 8 host threads per pipeline (hardware threads / LOCAL_WORLD_SIZE), or as
 * COMFYSTEREO_COMPACT_D2H=0/1 says.  The tensors the caller receives are the same either way. */
CS_API int cs_host_compact_enabled(void);

/* Measurement aid for bench.py (e2e.host_frac): GB/s (read + write) that `threads` host threads (0 = what one
 * pipeline of cs_stereo_batch_host may use) reach copying `bytes` with the non-temporal stores of the host path's
 * expansion loops.  No GPU involved. */
CS_API double cs_host_stream_bandwidth(size_t bytes, int threads);

/* Number of kernel launches issued by this library (process-wide) since the last reset
 * (bench.py reports it as gpu_launches). */
CS_API long long cs_launch_count(int reset);

/* Telemetry for the Polylines technique (synchronous): after cs_stereo_batch on `workspace` with
 * `chunk` frames in the last chunk, returns the status word (bit 0: the active list outgrew the
 * reference's own capacity 5*int(|div_px|)+25, SIG:1947) and how many rows needed the exact
 * sequential replay because their result depends on the reference's list order (quirk Q7). */
CS_API int cs_polylines_status(const cs_params *p, int chunk, int h, int w, const void *workspace,
                        int *status_out, int *flagged_rows);

/* Optional per-kernel timing for bench.py: when enabled every kernel launch is bracketed by CUDA events on
 * its own stream; cs_profile_collect (after a synchronize) sums the elapsed milliseconds and launch counts per
 * kernel id into ms[cs_profile_kernel_count()] / launches[...] and clears the record. */
CS_API int cs_profile_kernel_count(void);
CS_API const char *cs_profile_kernel_name(int id);
CS_API void cs_profile_enable(int on);
CS_API int cs_profile_collect(double *ms, long long *launches);

/* Test hook.  bit 0: Polylines replays EVERY row with the exact sequential sweep.
 * bit 2: Polylines uses 64-column tiles even when a whole row fits one CTA.  bit 3: every Polylines column through the
 * FP64 exact path.  bit 4: Polylines writes RGBX8 eyes and k_compose composes them (instead of composing in the sweep).
 * bit 5: the blur's edge-distance pass one pixel per thread with the IEEE division, even where the 4-pixel form applies.
 * bits 8-15: warps per Polylines CTA (0 = choose). */
CS_API void cs_set_test_flags(int flags);

#ifdef __cplusplus
}
#endif
#endif /* COMFYSTEREO_B200_H */
