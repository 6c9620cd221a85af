/*
 * wowsr.h — C ABI of libwowsr.so, the B200-native WOW super-resolution hot path.
 *
 * The reference (fieldin/sentinel2-super-resolution-poc) is pure Python and has no FFI; this is
 * the boundary a maintainer binds with ctypes.  Each entry point names the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions: plain C, opaque handle, caller-allocated buffers, explicit stream (a
 * cudaStream_t passed as void*; NULL = legacy default stream).  Every function returns 0 on
 * success or a negative wowsr_status; wowsr_last_error(ctx) gives the message.  A handle is
 * re-entrant with respect to other handles (no global mutable state); one handle must not be
 * used from two threads at once.  There is NO CPU fallback: without a CUDA device every compute
 * entry point returns WOWSR_ERR_CUDA.
 */
#ifndef WOWSR_H_
#define WOWSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define WOWSR_ABI_VERSION 1

typedef enum wowsr_status {
  WOWSR_OK = 0,
  WOWSR_ERR_ARG = -1,      /* bad argument                                   */
  WOWSR_ERR_CUDA = -2,     /* CUDA runtime / driver error (message has it)   */
  WOWSR_ERR_STATE = -3,    /* e.g. forward before load                       */
  WOWSR_ERR_NOMEM = -4,    /* device workspace could not be allocated        */
  WOWSR_ERR_UNSUPPORTED = -5
} wowsr_status;

typedef struct wowsr_ctx wowsr_ctx;

/* ------------------------------------------------------------------------------------------ */
/* handle                                                                                      */
/* ------------------------------------------------------------------------------------------ */

int wowsr_abi_version(void);
/* Creates a handle bound to CUDA device `device`.  Replaces the per-request construction of
 * `RealESRGAN(...)` (server/app/cnn_super_resolution.py:164-215, built and deleted per job at
 * server/app/wow_sr.py:93-97): keep one handle per worker instead. */
int wowsr_create(int device, wowsr_ctx** out);
void wowsr_destroy(wowsr_ctx* ctx);
const char* wowsr_last_error(const wowsr_ctx* ctx);   /* ctx may be NULL: last create() error */
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
uint64_t wowsr_launch_count(const wowsr_ctx* ctx);
/* Opaque tuning knobs (debug / bring-up); unknown keys return WOWSR_ERR_ARG. */
int wowsr_set_option(wowsr_ctx* ctx, const char* key, int64_t value);
int wowsr_get_option(const wowsr_ctx* ctx, const char* key, int64_t* value);

/* Copies `rows` rows of `row_bytes` bytes from device memory (ordered after everything queued on `stream`) into a PAGEABLE host
 * buffer through the handle's pinned staging ring (chunked cudaMemcpyAsync on a copy stream + a few CPU threads); synchronous.
 * The file entry points use it to bring the finished image down before encoding (server/app/wow_sr.py:126-164). */
int wowsr_download(wowsr_ctx* ctx, const void* dev, int64_t dev_pitch, int64_t row_bytes, int32_t rows, void* host,
                   int64_t host_pitch, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* post-process: wow_sr._enhance_for_crops (server/app/wow_sr.py:187-209) and the farm trio     */
/* enhance_local_contrast / apply_unsharp_mask / enhance_vegetation as called at                */
/* server/app/farm_sr.py:170-178.  RGB uint8, HWC.                                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct wowsr_post_params {
  double clip_limit;  /* cv2.createCLAHE(clipLimit=...)        wow_sr.py:191   (2.5)            */
  double sigma;       /* cv2.GaussianBlur(.., (0,0), sigma)    wow_sr.py:196   (1.2 | farm 1.5) */
  float alpha;        /* cv2.addWeighted(enh, alpha, ..)       wow_sr.py:197   (1.4 | 2.2)      */
  float beta;         /* cv2.addWeighted(.., blur, beta, 0)    wow_sr.py:197   (-0.4 | -1.2)    */
  float sat_boost;    /* S *= sat_boost inside the mask        wow_sr.py:205   (1.2 | 1.3)      */
  int32_t grid;       /* tileGridSize=(grid,grid)              wow_sr.py:191   (8)              */
  int32_t hue_lo;     /* green mask hue > hue_lo               wow_sr.py:202   (35)             */
  int32_t hue_hi;     /* green mask hue < hue_hi               wow_sr.py:202   (85)             */
  int32_t stages;     /* bit0 CLAHE, bit1 unsharp, bit2 vegetation; 7 = whole pipeline          */
  int32_t reserved;
} wowsr_post_params;

#define WOWSR_STAGE_CLAHE 1
#define WOWSR_STAGE_UNSHARP 2
#define WOWSR_STAGE_VEG 4
#define WOWSR_STAGE_ALL 7

void wowsr_post_params_wow(wowsr_post_params* p);   /* constants of wow_sr.py:190-207   */
void wowsr_post_params_farm(wowsr_post_params* p);  /* constants of farm_sr.py:170-178  */

/* A horizontal band of a (possibly larger) image held in device memory.  `data` points at the
 * band's first stored row, which is global row `y0`; `rows` rows are stored.  For a whole image
 * y0 = 0 and rows = H.  Multi-GPU sharding gives each rank one band (+ halo rows). */
typedef struct wowsr_image {
  void* data;         /* device pointer, uint8 HWC, 3 channels */
  int64_t pitch;      /* bytes between stored rows              */
  int32_t W, H;       /* full image size in pixels              */
  int32_t y0, rows;   /* stored band                            */
} wowsr_image;

/* Pass A: per-tile 256-bin histograms of L (RGB->Lab L plane) — the histogram part of
 * cv2 CLAHE::apply (wow_sr.py:191-192).  Accumulates (+=) into hist[grid*grid*256] (uint32,
 * device) the pixels of padded rows [prow0, prow1); zero `hist` first.  Padded rows/cols beyond
 * the image are REFLECT_101 copies and belong to the caller that holds the
 * reflected source rows. */
int wowsr_clahe_hist(wowsr_ctx* ctx, const wowsr_image* rgb, int32_t grid, int32_t prow0,
                     int32_t prow1, uint32_t* hist_dev, void* stream);
/* Clip / redistribute / cumulative sum -> uint8 LUTs [grid*grid*256] (device). */
int wowsr_clahe_luts(wowsr_ctx* ctx, const uint32_t* hist_dev, int32_t grid, int32_t tile_w,
                     int32_t tile_h, double clip_limit, uint8_t* luts_dev, void* stream);
/* Pass B: LUT interpolation -> Lab->RGB -> Gaussian unsharp -> HSV green boost -> RGB for output
 * rows [row0,row1).  `rgb` must hold rows [row0-r, row1+r) clipped to the image, r = blur radius.
 * `out` is a band descriptor for the destination (same W,H). */
int wowsr_post_apply(wowsr_ctx* ctx, const wowsr_image* rgb, const uint8_t* luts_dev,
                     const wowsr_post_params* p, int32_t row0, int32_t row1,
                     const wowsr_image* out, void* stream);
/* Whole pipeline on a device-resident image (hist -> luts -> apply), in == out allowed: no. */
int wowsr_post_process_dev(wowsr_ctx* ctx, const wowsr_image* rgb, const wowsr_post_params* p,
                           const wowsr_image* out, void* stream);
/* Drop-in for _enhance_for_crops(img) with HOST buffers (H2D, kernels, D2H, synchronous). */
int wowsr_post_process_host(wowsr_ctx* ctx, const uint8_t* rgb_host, int32_t H, int32_t W,
                            const wowsr_post_params* p, uint8_t* out_host);
/* Geometry helpers (cv2 CLAHE pads right/bottom to a multiple of grid). */
void wowsr_clahe_geometry(int32_t H, int32_t W, int32_t grid, int32_t* tile_w, int32_t* tile_h,
                          int32_t* padded_w, int32_t* padded_h);
/* Host copies of the integer tables the kernels use, for CPU-side verification against the
 * oracle without a GPU.  id: 0 gam(u16x256) 1 cbrt(u16x3072) 2 lab_y(u16x256) 3 lab_ify(u16x256)
 * 4 invgam(u8x4096) 5 sdiv(u32x256) 6 hdiv(u32x256).  Returns bytes written or <0. */
int64_t wowsr_get_table(int32_t id, void* out, int64_t cap);
/* 8-bit fixed-point Gaussian taps cv2 derives for sigma (ksize=(0,0)); returns ksize. */
int32_t wowsr_gaussian_taps(double sigma, int32_t* taps, int32_t cap);

/* HSV vegetation mask — compute_green_mask_hsv (server/app/vector_extraction.py:222-270):
 * cv2.cvtColor(rgb, COLOR_RGB2HSV) (H in [0,180)), cv2.inRange per colour range (bounds inclusive),
 * bitwise_or of the range masks, (> 0).astype(float32).  The reference uses two ranges: green
 * [hue_min..hue_max, sat_min..255, val_min..255] (:255-259) and brown [10..35, 20..200, 40..200]
 * (:262-264).  `rgb` is a device band (rows are band-relative in `mask_dev`, a float32 [rows][W]
 * array with `mask_pitch` BYTES between rows); 1 <= n_ranges <= 4. */
typedef struct wowsr_hsv_range {
  uint8_t lo[3];   /* lower (H, S, V), inclusive */
  uint8_t hi[3];   /* upper (H, S, V), inclusive */
} wowsr_hsv_range;
int wowsr_green_mask(wowsr_ctx* ctx, const wowsr_image* rgb, const wowsr_hsv_range* ranges,
                     int32_t n_ranges, float* mask_dev, int64_t mask_pitch, void* stream);
/* Same with HOST buffers (H2D, kernel, D2H, synchronous): rgb [H,W,3] uint8 -> mask [H,W] float32. */
int wowsr_green_mask_host(wowsr_ctx* ctx, const uint8_t* rgb_host, int32_t H, int32_t W,
                          const wowsr_hsv_range* ranges, int32_t n_ranges, float* mask_host);

/* ------------------------------------------------------------------------------------------ */
/* XYZ tile pyramid: the resampling of gdal2tiles.py --xyz --resampling average                  */
/* (server/app/tiling.py:147-186) from the device-resident image.  PARITY UNPINNED (no GDAL).     */
/* ------------------------------------------------------------------------------------------ */

/* Resamples one strip of a zoom level's tile mosaic: mosaic pixel (X, Y) covers the source rectangle
 * [sx0 + X sxp, sx0 + (X+1) sxp) x [sy0 + Y syp, sy0 + (Y+1) syp) (source pixel units, may lie outside the image).
 * out: RGBA uint8 [OH][OW][4], alpha 255 where the mosaic pixel's centre is inside the raster (area-weighted mean of the
 * covered source pixels, rounded half up; nearest pixel when sxp, syp <= 1), 0 elsewhere. */
int wowsr_tiles_resample(wowsr_ctx* ctx, const wowsr_image* rgb, double sx0, double sy0, double sxp, double syp,
                         uint8_t* out_rgba_dev, int64_t out_pitch, int32_t OW, int32_t OH, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* window planner: RealESRGAN._tile_process geometry (cnn_super_resolution.py:244-278)         */
/* ------------------------------------------------------------------------------------------ */

typedef struct wowsr_window {
  int32_t x0, y0, x1, y1;        /* LR window [x0,x1) x [y0,y1) fed to the network            */
  int32_t ox0, oy0, ox1, oy1;    /* LR-pixel rectangle of the output this window OWNS after     */
                                 /* last-writer-wins resolution (may be empty: ox1<=ox0)        */
} wowsr_window;

/* Pure host function.  Returns the window count (tiles_y*tiles_x, row-major like the reference
 * loop) and fills up to `cap` entries; when H*W <= 4*tile*tile the reference does not tile
 * (cnn_super_resolution.py:226) and a single whole-image window is returned. */
int32_t wowsr_plan_windows(int32_t H, int32_t W, int32_t tile, int32_t pad, wowsr_window* out,
                           int32_t cap);

/* ------------------------------------------------------------------------------------------ */
/* RRDBNet: cnn_super_resolution.py:73-158 (network) and :217-280 (enhance / _tile_process)     */
/* ------------------------------------------------------------------------------------------ */

#define WOWSR_PREC_BF16 0      /* bf16 operands, fp32 accumulate, fp32 residual trunk */
#define WOWSR_PREC_FP16 1      /* fp16 operands, fp32 accumulate, fp32 residual trunk */
#define WOWSR_PREC_MIXED 2     /* bf16 RRDB trunk (92 % of the FLOPs) + fp16 for the 5 tail convs; default */

/* Loads weights.  `tensors` are host fp32 pointers in the order of the reference state_dict
 * (conv_first.weight, conv_first.bias, body.0.rdb1.conv1.weight, ... conv_last.bias; weights
 * OIHW) — what `RealESRGAN.__init__` loads at cnn_super_resolution.py:205-211.  Weights are
 * repacked once into the tensor-core layout. */
int wowsr_load_rrdbnet(wowsr_ctx* ctx, int32_t num_block, int32_t num_feat, int32_t num_grow,
                       const float* const* tensors, int32_t n_tensors, int32_t precision);
/* Runs the network on `n` windows of a device-resident BGR uint8 image and writes each window's
 * owned rectangle (x4) into `out` (uint8 BGR, truncating quantisation of :232). All windows must
 * have the same size.  Optional `out_f32` (may be NULL) receives the pre-quantisation float
 * output, HWC fp32 with pitch out_f32_pitch bytes, for parity checks. */
int wowsr_rrdbnet_forward_windows(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W,
                                  int64_t pitch, const wowsr_window* windows, int32_t n,
                                  uint8_t* out_dev, int64_t out_pitch, float* out_f32,
                                  int64_t out_f32_pitch, void* stream);
/* Drop-in for RealESRGAN.enhance(img) with HOST buffers: plan, H2D, forward, D2H. tile_size as
 * in the reference constructor (cnn_super_resolution.py:168). out_host is [4H,4W,3]. */
int wowsr_enhance_host(wowsr_ctx* ctx, const uint8_t* img_host, int32_t H, int32_t W,
                       int32_t tile_size, uint8_t* out_host, float* out_f32_host);
/* Same with device buffers (img [H,W,3] pitch W*3, out [4H,4W,3] pitch 4W*3). */
int wowsr_enhance_dev(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W,
                      int32_t tile_size, uint8_t* out_dev, float* out_f32_dev, void* stream);
/* One 3x3 convolution layer on NHWC fp32 host data through the tensor-core kernel (bring-up and
 * parity tests): in [n,h,w,cin], weight OIHW [cout,cin,3,3], bias [cout], out [n,h,w,cout].
 * act: 0 none, 1 LeakyReLU(0.2).  Operands are rounded to the handle's precision. */
int wowsr_conv3x3_host(wowsr_ctx* ctx, const float* in, int32_t n, int32_t h, int32_t w,
                       int32_t cin, const float* weight, const float* bias, int32_t cout,
                       int32_t act, int32_t precision, float* out);
/* Timing of the last forward (CUDA events on the handle's stream): milliseconds per phase.
 * phases: 0 total, 1 head, 2 trunk (RRDBs), 3 tail (HR convs); 4 = HOST milliseconds the calling thread spent enqueuing the
 * forward's launches (wowsr_rrdbnet_forward_windows / wowsr_enhance_*; 0 for EDSR).  Returns count written (<= cap). */
int32_t wowsr_get_timing(const wowsr_ctx* ctx, float* ms, int32_t cap);
/* Debug: per-tile clock64 stamps [tile][mma_start, mma_issued, epilogue_start, epilogue_end] of CTA 0 of the conv
 * launch selected with wowsr_set_option("tc_trace_layer", k) (k = 1-based launch index). Returns count. */
int32_t wowsr_debug_trace(wowsr_ctx* ctx, int64_t* out, int32_t cap);

/* Debug / test introspection of the rolling conv kernel's work list (csrc/roll_kernel.cuh; pure host function): the column
 * segments that one launch over `n_win` windows of h x w assigns to `max_units` CTAs (pair = 0) or CTA pairs (pair = 1);
 * columns x >= strip_x0 (strip_x0 = w: none) are covered by vertical tasks.  Each task is 8 int32:
 * {window of rank 0, window of rank 1 (-1: dummy), run origin of rank 0, of rank 1, first output row, rows, 0, 0}; `off`
 * receives units + 1 task offsets; info[4] = {units, horizontal units, offsets written, 0}.  Returns the task count. */
int32_t wowsr_debug_roll_plan(int32_t n_win, int32_t h, int32_t w, int32_t strip_x0, int32_t pair, int32_t max_units,
                              int32_t* tasks, int32_t cap_tasks, int32_t* off, int32_t cap_off, int32_t* info);

/* ------------------------------------------------------------------------------------------ */
/* EDSR-baseline x4 "farm SR" variant (super_resolution.py:92-124,196: cv2.dnn_superres          */
/* DnnSuperResImpl.upsample with EDSR_x4.pb — third-party, parity unpinned)       */
/* ------------------------------------------------------------------------------------------ */

int wowsr_load_edsr(wowsr_ctx* ctx, int32_t num_block, int32_t num_feat, float res_scale,
                    const float* const* tensors, int32_t n_tensors, int32_t precision);
int wowsr_edsr_upsample_host(wowsr_ctx* ctx, const uint8_t* img_host, int32_t H, int32_t W,
                             uint8_t* out_host, float* out_f32_host);
/* Same with device buffers (img [H,W,3] BGR u8 pitch W*3, out [4H,4W,3] pitch 4W*3, out_f32_dev may be NULL); returns after
 * the stream has drained (the kernels' watchdog flag is read back).  wowsr_get_timing: 1 head, 2 resblocks, 3 upsampler + tail. */
int wowsr_edsr_upsample_dev(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, uint8_t* out_dev,
                            float* out_f32_dev, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* WOWSR_H_ */
