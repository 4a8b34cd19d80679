/* C ABI of the B200-native Tuatara OCR hot path (libtuatara_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/OpenCV types, no
 * exceptions.  Every function returns 0 on success and non-zero on failure; the message is
 * available from tt_last_error() (thread-local).  The reference has no FFI of its own -- its
 * boundary is the C++ function image_to_data (tuatara.h:13) and the pybind11 module
 * (bindings/python.cpp:54-58); include/tuatara.h and tuatara_b200/bindings/python.cpp re-create
 * those on top of this ABI.  Each entry point cites the reference code it replaces.
 */
#ifndef TUATARA_C_H
#define TUATARA_C_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TT_API __attribute__((visibility("default")))
#else
#define TT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TT_MAX_LABEL_LEN 26  /* PARSeq positions per crop (max_label_length + 1) */
#define TT_NUM_CLASSES 95    /* logits per position */

typedef struct tt_engine tt_engine;

/* A caller-owned 8-bit, 3-channel, row-major image (the cv::Mat of tuatara.h:13 without OpenCV).
 * The buffer is only read (the reference swaps channels in place at tuatara.cpp:349; here the
 * swap happens inside the kernels). */
typedef struct tt_image {
  const uint8_t* data;
  int rows, cols, channels; /* channels must be 3 */
  size_t step;              /* bytes per row */
} tt_image;

/* Every literal of the reference gathered in one struct (the TODO at tuatara.cpp:396). */
typedef struct tt_config {
  float canvas_size;    /* 1024  tuatara.cpp:352 */
  float mag_ratio;      /* 1.0   tuatara.cpp:353 */
  float text_threshold; /* 0.7   tuatara.cpp:397 */
  float link_threshold; /* 0.4   tuatara.cpp:398 */
  float low_text;       /* 0.4   tuatara.cpp:399 */
  int min_area;         /* 10    tuatara.cpp:148 */
  int max_batch_pages;  /* pages whose crops form one PARSeq batch in a slot (0 = default 32); detection runs in units of <= 8
                           equally sized pages, crops of pages of any size share the recognition batch */
  int slots_per_gpu;    /* concurrent execution slots (streams + host threads) per GPU: 0 = default 3, max 4.  More slots
                           overlap one batch's host phases and latency-bound kernels with another batch's kernels
                           (2 slots +16 %, 3 slots +19 % over 1 on the 512-page bench); 1 = strictly serial kernels */
  int rectify;          /* 0 (default) = the reference's axis-aligned boundingRect crop (tuatara.cpp:416).  1 = opt-in for the
                           TODO at tuatara.cpp:411-415: each crop is the perspective warp of its rotated box to 128 x 32
                           (cv::getPerspectiveTransform + cv::warpPerspective semantics); boxes / bboxes are unchanged */
} tt_config;

typedef struct tt_item {
  char* text;    /* NUL-terminated, owned by the result */
  float bbox[4]; /* min_x, min_y, max_x, max_y (tuatara.cpp:256-274) */
} tt_item;
typedef struct tt_page_result {
  int n_items;
  tt_item* items; /* in CCL label order (tuatara.cpp:146) */
} tt_page_result;
typedef struct tt_result {
  int n_pages;
  tt_page_result* pages;
} tt_result;

/* ---------------------------------------------------------------- whole path */
TT_API void tt_config_default(tt_config* cfg);
TT_API const char* tt_last_error(void);
/* Engine = weights resident on each listed device + streams + workspaces.  Replaces the two
 * torch::jit::load calls per image (tuatara.cpp:333-336, :423-428): weights_dir must hold
 * craft.ttw and parseq.ttw (tuatara_b200.weights exporter).  devices == NULL -> the devices named by the TT_DEVICES
 * environment variable ("0,1,2" or "all"), else {0}. */
TT_API int tt_engine_create(const char* weights_dir, const int* devices, int n_devices, const tt_config* cfg,
                     tt_engine** out);
TT_API void tt_engine_destroy(tt_engine* e);
/* CUDA devices visible to this process (0 when there is none: the library has no CPU fallback). */
TT_API int tt_device_count(void);
/* image_to_data (tuatara.cpp:314-512) for a batch of pages.  The pages are cut into detection units that the execution
 * slots of all the engine's GPUs pull from a shared work queue (no collective: pages are independent); results are
 * gathered on the host in page order. */
TT_API int tt_ocr_pages(tt_engine* e, const tt_image* pages, int n_pages, tt_result** out);
/* Same, with options the benchmark and the parity tests need.
 *  pages_on_device: tt_image.data are device pointers (page i on the GPU that gets page i, i.e.
 *    engine device i % n_devices) -- throughput with inputs already resident in HBM.
 *  score_override: NULL, or n_pages pointers (NULL entries allowed) to fp32 [h32/2][w32/2][2] maps
 *    that replace CRAFT's output for that page AFTER CRAFT has run in full (random-init weights
 *    give near-constant maps).  override_on_device: those pointers are device memory.
 *  Zero-initialise the struct: fields added later default to 0. */
typedef struct tt_ocr_options {
  int pages_on_device;
  int override_on_device;
  const float* const* score_override;
  int detect_only;   /* 1: stop after the word boxes (tuatara.cpp:349-418): items carry bbox and an empty text
                        (BASELINE config "CRAFT detection only") */
} tt_ocr_options;
TT_API int tt_ocr_pages_ex(tt_engine* e, const tt_image* pages, int n_pages, const tt_ocr_options* opt, tt_result** out);
TT_API void tt_result_free(tt_result* r);
/* Kernel launches issued by this library so far (bench.py's gpu_launches). */
TT_API unsigned long long tt_launch_count(void);

/* Change slots_per_gpu of a live engine (1 = strictly serial kernels: per-launch event timing is only
 * meaningful then). */
TT_API void tt_engine_set_slots(tt_engine* e, int slots);
/* cudaStream_t the engine's idx-th device works on (bench.py records its CUDA events there). */
TT_API void* tt_engine_stream(tt_engine* e, int idx);
/* Host<->device bytes moved by tt_ocr_pages* so far (bench.py's h2d/d2h_bytes_per_step). */
TT_API void tt_io_bytes(unsigned long long* h2d, unsigned long long* d2h);
/* Per-launch CUDA-event timing of the tensor-core GEMM/conv kernel. tt_profile_collect waits for the
 * recorded launches, returns their summed duration, algorithmic FLOPs and count, and clears the log. */
TT_API void tt_profile_enable(int on);
TT_API void tt_profile_collect(double* total_ms, double* total_flops, unsigned long long* launches);
/* Same, and appends one CSV line per launch ("tag,flops,ms") to `path` (per-layer tables in profiles/). */
TT_API void tt_profile_dump(const char* path, double* total_ms, double* total_flops, unsigned long long* launches);

/* Stage-level timings recorded while the profile is enabled: "name,count,ms,flops,bytes\n" per stage of
 * image_to_data (preprocess tuatara.cpp:206-234, craft :376, postprocess :119-204, crop_resize :416-448,
 * parseq_encoder / parseq_decoder :307), flops / bytes being the stage's algorithmic work.  Writes at most
 * cap-1 characters + NUL, returns the full length, clears the log. */
TT_API int tt_profile_stages(char* buf, int cap);

/* Development aid (TT_TRACE=1 in the environment, off otherwise): where every CTA of the TMEM-allocating kernels stands,
 * read from a host-mapped progress buffer -- callable from a watchdog thread while a kernel is stuck on the GPU.
 * Writes at most cap-1 characters + NUL, returns the full length. */
TT_API int tt_debug_trace_report(char* buf, int cap);

/* ---------------------------------------------------- stage level, host memory */
/* Size arithmetic of resize_aspect_ratio (tuatara.cpp:211-226), fp32 like the reference. */
TT_API int tt_resize_plan(int rows, int cols, float canvas_size, float mag_ratio, int* target_h, int* target_w, int* h32,
                   int* w32, float* ratio);
/* Channel swap + resize + zero pad (tuatara.cpp:349, :206-234). out: h32*w32*3 bytes. */
TT_API int tt_preprocess(const tt_image* image, float canvas_size, float mag_ratio, uint8_t* out);
/* CRAFT forward (tuatara.cpp:363-394). craft_input = tt_preprocess output; maps_out: [h32/2][w32/2][2] fp32. */
TT_API int tt_craft_forward(tt_engine* e, const uint8_t* craft_input, int h32, int w32, float* maps_out);
/* get_detected_boxes (tuatara.cpp:119-204) on one score map [H][W][2].
 *  labels_out  (nullable) [H*W] int32          == cv::connectedComponentsWithStats labels
 *  stats_out   (nullable) [stats_cap][5] int32 == its stats rows (left, top, width, height, area), row 0 = background
 *  rects_out   [rect_cap][5] fp32 (cx, cy, w, h, angle) of the kept components, rect_labels_out their labels */
TT_API int tt_postprocess(const float* maps, int H, int W, const tt_config* cfg, int32_t* labels_out, int32_t* stats_out,
                   int stats_cap, int* n_labels, float* rects_out, int32_t* rect_labels_out, int rect_cap,
                   int* n_rects);
/* Crop + resize to 128x32 (tuatara.cpp:416, :440-441). rects_xywh already clamped to the image.
 * out: [n][32][128][3] u8 in the caller's channel order. */
TT_API int tt_crop_resize(const tt_image* image, const int32_t* rects_xywh, int n, uint8_t* out);
/* Per-slice parity of the CRAFT forward: a named activation of the LAST tt_craft_forward call on this
 * engine as fp32 [H][W][C] (names: relu2_2, relu3_2, relu4_3, relu5_3, fc7, up1, up2, up3, up4 -- upstream CRAFT's
 * skip / U-net tensors).  out == NULL only reports dims_out = {H, W, C}.  Test entry point: not thread safe against
 * other calls on the same engine. */
TT_API int tt_craft_tap(tt_engine* e, const char* name, float* out, long long capacity, int dims_out[3]);
/* Rectified crops (tt_config.rectify): quads [n][4][2] fp32 = top-left, top-right, bottom-right, bottom-left corners in
 * image coordinates (may leave the image: border pixels are replicated).  out: [n][32][128][3] u8. */
TT_API int tt_crop_warp(const tt_image* image, const float* quads, int n, uint8_t* out);
/* The quad tt_config.rectify uses for a RotatedRect {cx, cy, w, h, angle}: 8 floats, corner order as above. */
TT_API int tt_rect_to_quad(const float rect[5], float quad_out[8]);
/* PARSeq forward (tuatara.cpp:307, 26 AR steps + 1 refinement). crops: [n][32][128][3] u8.
 * forced_tokens (nullable) [n][25]: teacher-forced AR context (parity tests).
 * logits_out (nullable) [n][26][95] fp32; ids_out (nullable) [n][26] argmax (== softmax + max, :486,:103). */
TT_API int tt_parseq_forward(tt_engine* e, const uint8_t* crops, int n, const int32_t* forced_tokens, float* logits_out,
                      int32_t* ids_out);
/* Tokenizer::decode + truncation (tuatara.cpp:61-116, :497-502). ids [n][len]; out [n][out_stride] chars. */
TT_API int tt_decode(const int32_t* ids, int n, int len, char* out, int out_stride);
/* itos_out: >= 99 bytes. */
TT_API int tt_tokenizer_table(char* itos_out, int* eos_id, int* bos_id, int* pad_id);

/* ------------------------------------------------- host geometry (no GPU needed) */
TT_API int tt_convex_hull_i32(const int32_t* xy, int n, int32_t* idx_out, int* n_out);
TT_API int tt_convex_hull_f32(const float* xy, int n, int32_t* idx_out, int* n_out);
TT_API int tt_min_area_rect_i32(const int32_t* xy, int n, float rect_out[5]);
TT_API int tt_min_area_rect_f32(const float* xy, int n, float rect_out[5]);
TT_API int tt_rect_points(const float rect[5], float pts_out[8]);
TT_API int tt_rect_bounding(const float rect[5], int32_t xywh_out[4]);
TT_API int tt_adjust_rect(const float rect[5], float ratio_w, float ratio_h, float ratio_net, float rect_out[5]);
TT_API int tt_rect_to_bbox(const float rect[5], float bbox_out[4]);

/* --------------------------------------- stage level, device memory (bench / kernel tests) */
/* out[M][N] = act(A[M][K] * W[N][K]^T + bias) (+ residual). All pointers are device pointers;
 * A, W bf16; bias fp32; out bf16 (out_f32 == 0) or fp32. act: 0 none, 1 relu, 2 gelu.
 * BN: N tile (0 = planner decides); resident: schedule flags -- with BN given: bit 0 = weight-resident,
 * bit 1 = CTA-pair (cta_group::2, 256-row tiles); with BN == 0: 4 = pair forced, 8 = pair forbidden, else planner. */
TT_API int tt_linear_dev(const void* A, int lda, int M, int K, const void* W, int N, const float* bias, int act,
                  const void* residual, int res_f32, int ldr, void* out, int out_f32, int ldc, int BN, int resident,
                  void* stream);
/* Kernel test of the fused LayerNorm pair the PARSeq encoder runs (x += A W1^T + b1; y = act(Linear(LN(x)))).  The fp32
 * residual stream x is kept split: x_hi = bf16(x), x_lo = bf16(x - x_hi), both bf16 [M][D], updated in place (x_hi is
 * the second GEMM's A operand).  stats fp32 [M][8] scratch, W2f = W2 * gamma (bf16 [N2][D]), c0[n] = b2[n] + beta . W2[n],
 * c1[n] = sum_k W2f[n][k]; out bf16 [M][N2].  All device pointers. */
TT_API int tt_linear_ln_pair_dev(const void* A, int M, int K1, const void* W1, const float* b1, int D, void* x_hi, void* x_lo,
                          float* stats, const void* W2f, const float* c0, const float* c1, int N2, int act, float eps,
                          void* out, void* stream);
/* Kernel test of the fused second half of a PARSeq-base encoder block (timm Block, reached by the reference through
 * TorchScript at tuatara.cpp:307):  x += att Wp^T + bp  (the attention output projection; skipped when att is NULL), then
 * x += fc2(GELU(fc1(LN2(x)))) -- one kernel, neither x1 nor the hidden activations leave the SM.  The residual stream is the
 * split pair of tt_linear_ln_pair_dev (bf16 [M][384] each, updated in place); stats fp32 [M][4]: with att == NULL two
 * partial (sum x, sum x^2) pairs per row on entry; on return the two partials of the new x.  att bf16 [M][384], Wp bf16
 * [384][384], bp fp32 [384]; W1f = fc1 * gamma bf16 [1536][384], c0 / c1 as above, W2 bf16 [384][1536], b2 fp32 [384].
 * All device pointers. */
TT_API int tt_enc_mlp_dev(void* x_hi, void* x_lo, float* stats, long long M, const void* att, const void* Wp, const float* bp,
                   const void* W1f, const float* c0, const float* c1, const void* W2, const float* b2, float eps, void* stream);
/* NHWC bf16 stride-1 "same" convolution as implicit GEMM; src1 may be NULL (else channel concat).
 * weight bf16 [Cout][taps][C0+C1]; out bf16 [batch][H][W][Cout]. */
TT_API int tt_conv_dev(const void* src0, int C0, const void* src1, int C1, int batch, int H, int W, int taps, int dil,
                const void* weight, const float* bias, int Cout, int relu, void* out, int BN, int resident,
                void* stream);
/* Batched device-side post-processing (bench): maps_dev [batch][H][W][2] fp32. Returns total rects. */
TT_API int tt_postprocess_dev(tt_engine* e, const float* maps_dev, int batch, int H, int W, int* n_rects_total,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TUATARA_C_H */
