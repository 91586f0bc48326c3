/*
 * rndense.h -- C ABI of librndense.so: the RetinaNet dense-head data path on B200 (sm_100a).
 *
 * The reference (vshmyhlo/retinanet-tensorflow) has no FFI of its own: its boundary for this
 * path is the Python module surface of levels.py / dataset.py / utils.py / losses.py.  Each
 * entry point below names the reference function(s) it replaces (file:line, relative to the
 * reference checkout).  The Python host in retinanet_tensorflow_b200/ keeps those function
 * signatures and forwards to these symbols through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / DLPack types.
 *   - Pointers named d_* are DEVICE pointers, caller-allocated, never freed or retained.
 *     rn_pyramid_t / rn_rows_t / rn_maps_t are small HOST structs passed by pointer.
 *   - Every launch goes to the caller's `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising.  The caller selects the device (cudaSetDevice / rn_set_device).
 *   - Return value: RN_OK (0) or a negative rn_status; rn_last_error() gives a thread-local
 *     message.  Data-dependent violations (the reference's tf.assert_* inside utils.iou,
 *     utils.py:65-68,94) set bits in a caller-supplied device status word (d_status, may be
 *     NULL) -- RN_FLAG_*.
 *   - Boxes are [ymin, xmin, ymax, xmax] normalised to [0,1] by the image size
 *     (dataset.py:161); centre form is [cy, cx, h, w] (dataset.py:28-39).  All arithmetic
 *     is float32 with the reference's op order and no FMA contraction where indices
 *     depend on it.
 *
 * Level-major "maps".  The reference keeps one tensor per pyramid level, shaped
 * [B, H_l, W_l, A, K] (retinanet.py:68-69).  An rn_maps_t holds one base pointer per level
 * to exactly that layout.  A single merged buffer (concatenation over levels P3..P7, the
 * order of utils.merge_outputs, utils.py:151-153) is the special case built by
 * rn_maps_from_merged().  Row r of level l is (b * cells_l + y * W_l + x) * A + a.
 */
#ifndef RNDENSE_H_
#define RNDENSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RN_API __declspec(dllexport)
#else
#define RN_API __attribute__((visibility("default")))
#endif

#define RN_VERSION 100 /* 0.1.0 */
#define RN_MAX_LEVELS 8
#define RN_MAX_ANCHORS 16

typedef enum rn_status {
  RN_OK = 0,
  RN_ERR_BAD_ARG = -1,     /* NULL pointer, negative size, unknown enum */
  RN_ERR_BAD_SHAPE = -2,   /* shapes inconsistent with the pyramid / limits exceeded */
  RN_ERR_CAPACITY = -3,    /* workspace or output capacity too small (host-detectable) */
  RN_ERR_CUDA = -4,        /* a CUDA runtime call failed; see rn_last_error() */
  RN_ERR_UNSUPPORTED = -5
} rn_status;

/* bits of the device status word */
#define RN_FLAG_BOX_ORDER 1u      /* a box has min > max  (utils.py:66-67 assert_less_equal) */
#define RN_FLAG_IOU_RANGE 2u      /* IoU NaN or outside [0,1]  (utils.py:94)                  */
#define RN_FLAG_CAND_OVERFLOW 4u  /* more candidates than `capacity`; extra ones dropped      */
#define RN_FLAG_DET_OVERFLOW 8u   /* more detections than `out_capacity`; extra ones dropped  */
#define RN_FLAG_EXCHANGE_TIMEOUT 32u /* a peer's message did not arrive in time (rn_stats_exchange): fatal, NaN */
#define RN_FLAG_EMPTY_GT 16u      /* an image had 0 GT boxes (reference asserts >= 1,
                                     data_loaders/inferred.py:30); handled as all-background  */

/* how `level.anchor_sizes / image_size` is evaluated (dataset.py:53, utils.py:264) */
#define RN_SIZE_TF1_INT_TRUNC 0 /* TF1: ndarray -> int32 (truncation), f64 divide, f32 cast   */
#define RN_SIZE_FLOAT 1         /* no truncation                                              */

/* compact per-anchor label values written by rn_assign_encode */
#define RN_LABEL_BACKGROUND (-1) /* max IoU < neg_thr (or positive with class id out of range) */
#define RN_LABEL_IGNORE (-2)     /* neg_thr <= max IoU < pos_thr: excluded from the loss       */

typedef struct rn_level {
  int32_t grid_h, grid_w;    /* ceil(image / stride)                       dataset.py:55   */
  int32_t num_anchors;       /* anchors per cell                                            */
  int32_t tiles_x;           /* assign-kernel tiling (filled by the builders)               */
  int32_t tiles_y;
  int32_t tile_start;        /* first tile of this level within one image                   */
  int64_t anchor_offset;     /* anchors of earlier levels within ONE image                  */
  float y0, y_step;          /* centre_y(i) = y0 + y_step * i  (tf.linspace, dataset.py:19) */
  float x0, x_step;
  float max_half_h, max_half_w; /* max over anchors of size / 2 (patch bounds of the assign kernel) */
  float anchor_h[RN_MAX_ANCHORS]; /* normalised sizes, float32                              */
  float anchor_w[RN_MAX_ANCHORS];
} rn_level_t;

typedef struct rn_pyramid {
  int32_t num_levels;
  int32_t image_h, image_w;
  int32_t tiles_per_image;
  int64_t anchors_per_image;
  rn_level_t level[RN_MAX_LEVELS];
} rn_pyramid_t;

/* Row segments for kernels that need no geometry (losses, select on decoded boxes).
 * Segment s has rows[s] rows; for a batched pyramid rows[l] = B * cells_l * A. */
typedef struct rn_rows {
  int32_t num_segments;
  int32_t images;                       /* B; rows[s] must be divisible by it             */
  int64_t rows[RN_MAX_LEVELS];
} rn_rows_t;

typedef struct rn_maps {
  void* p[RN_MAX_LEVELS];
} rn_maps_t;

/* classification-loss modes (losses.py:115-141; the commented-out alternatives are modes) */
#define RN_CLS_NONE 0
#define RN_CLS_FOCAL 1        /* sum(focal) / max(num_fg, 1)            losses.py:6-15,119-122 */
#define RN_CLS_BCE 2          /* mean(BCE)                              losses.py:124-125      */
#define RN_CLS_BALANCED_BCE 3 /* mean(balanced BCE, axis=0)             losses.py:96-110,127   */
#define RN_CLS_DICE_FLAG 16   /* + mean_C(dice(smooth=0, axis=0))       losses.py:50-60,130    */
#define RN_CLS_GENERIC_FLAG 32 /* per call: take the generic kernel (libm, IEEE division, any C / focus) even where the
                                 streaming kernel applies -- the parity tests run both; no process-global switch */

/* element type of the logits / dlogits maps of the classification-loss entry points */
#define RN_DTYPE_F32 0
#define RN_DTYPE_BF16 1  /* bfloat16 in and out (128-bit bf16x8 accesses), float32 math; fast kernel only:
                            compact labels, C % 8 == 0, focal (focus 2) / BCE / balanced BCE without dice */

/* layout of the double-precision statistics vector (device), length RN_STATS_HEADER + 3*C.
 * It is the ONLY data exchanged between GPUs (one all-reduce(sum) per phase). */
#define RN_STATS_HEADER 8
#define RN_STAT_TRAINABLE 0   /* M: anchors that are not ignored                            */
#define RN_STAT_NUM_FG 1      /* anchors with a one-hot class (fg_mask, losses.py:157)      */
#define RN_STAT_CLS_SUM 2     /* sum of the elementwise classification loss (focal / BCE)   */
#define RN_STAT_HUBER_SUM 3   /* sum over fg anchors and 4 coords of Huber                  */
#define RN_STAT_NPOS(c) (RN_STATS_HEADER + (c))               /* sum_m z[m,c]              */
#define RN_STAT_INTER(C, c) (RN_STATS_HEADER + (C) + (c))     /* sum_m z*p   (dice I_c)    */
#define RN_STAT_PSUM(C, c) (RN_STATS_HEADER + 2 * (C) + (c))  /* sum_m p                   */

/* ---- library ------------------------------------------------------------------------- */
RN_API int rn_version(void);
RN_API const char* rn_last_error(void);
RN_API const char* rn_status_string(int status);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
RN_API int64_t rn_launch_count(void);
RN_API int rn_set_device(int device);
/* Process-wide switches.  "fast_focal" (default 1): use the specialised focal kernel when it
 * applies (focal, focus == 2, compact labels, C % 4 == 0); 0 forces the generic kernel. */
RN_API int rn_set_option(const char* name, int value);

/* ---- pyramid description (host only, no launch) -------------------------------------- */
/* levels.build_levels + the per-level constants of dataset.level_labels (dataset.py:53-57):
 * anchor sizes -> normalised float32 (size_mode), grid = ceil(image/stride), linspace
 * start/step.  anchor_sizes_px is [num_levels][num_anchors][2] float64 (h, w) in pixels,
 * as produced by levels.Level.anchor_sizes (levels.py:38-44). */
RN_API int rn_pyramid_build(int image_h, int image_w, int num_levels, const int32_t* strides,
                            const double* anchor_sizes_px, int num_anchors, int size_mode,
                            rn_pyramid_t* out);
/* Same, from explicit grid sizes and already-normalised float32 anchor sizes
 * [num_levels][num_anchors][2] -- the form utils.anchor_boxmap / regression_postprocess
 * receive (utils.py:47,108). */
RN_API int rn_pyramid_from_grids(int num_levels, const int32_t* grid_hw,
                                 const float* anchor_sizes_norm, int num_anchors,
                                 rn_pyramid_t* out);
RN_API int rn_rows_from_pyramid(const rn_pyramid_t* pyr, int batch, rn_rows_t* out);
RN_API int rn_maps_from_merged(const rn_pyramid_t* pyr, int batch, void* d_base,
                               size_t bytes_per_anchor, rn_maps_t* out);

/* ---- anchors ------------------------------------------------------------------------- */
/* utils.anchor_boxmap (utils.py:47-58) / dataset.position_grid (dataset.py:16-25) for every
 * level: d_boxes [anchors_per_image, 4] corner boxes and (optional) d_centres
 * [anchors_per_image, 4] (cy, cx, h, w), level-major. */
RN_API int rn_anchor_generate(const rn_pyramid_t* pyr, float* d_boxes, float* d_centres, void* stream);

/* ---- assignment + encoding ----------------------------------------------------------- */
/* dataset.level_labels / build_labels (dataset.py:43-142) for a batch, compact outputs.
 *   d_gt_boxes [B, max_gt, 4] f32 corner boxes, d_gt_classes [B, max_gt] i32, d_gt_counts [B].
 *   Outputs are level-major maps of int32 / int32 / float[4] / float per anchor:
 *   matched = argmax_n IoU (first maximum; 0 when every IoU is 0), label = class id |
 *   RN_LABEL_BACKGROUND | RN_LABEL_IGNORE, deltas = (dy, dx, log dh, log dw) against the
 *   matched box for EVERY anchor (dataset.py:121), iou_max optional (p[0] == NULL to skip).
 *   d_stats (optional): adds M, num_fg and n_pos[c] of this batch into the stats vector. */
RN_API int rn_assign_encode(const rn_pyramid_t* pyr, int batch, const float* d_gt_boxes,
                            const int32_t* d_gt_classes, const int32_t* d_gt_counts, int max_gt,
                            int num_classes, float neg_iou_thr, float pos_iou_thr,
                            const rn_maps_t* matched, const rn_maps_t* label,
                            const rn_maps_t* deltas, const rn_maps_t* iou_max,
                            double* d_stats, uint32_t* d_status, void* stream);
/* Same with an explicit work decomposition: one CTA handles patch_sy x patch_sx sub-patches of 8 x 8 cells and
 * shares the per-image work (GT conversion + culling, encoding table) between them.  0, 0 = automatic (what
 * rn_assign_encode uses: 2 x 2 for launches of many waves, 2 x 1 for medium ones, 1 x 1 for small ones).  The
 * results do not depend on the decomposition; the parity tests run several. */
RN_API int rn_assign_encode_patch(const rn_pyramid_t* pyr, int batch, const float* d_gt_boxes,
                                  const int32_t* d_gt_classes, const int32_t* d_gt_counts, int max_gt,
                                  int num_classes, float neg_iou_thr, float pos_iou_thr,
                                  const rn_maps_t* matched, const rn_maps_t* label,
                                  const rn_maps_t* deltas, const rn_maps_t* iou_max,
                                  double* d_stats, uint32_t* d_status, int patch_sy, int patch_sx,
                                  void* stream);

/* Dense reference layout from compact labels: classification one-hot f32 [.., C] (zero rows
 * for background and ignored anchors, dataset.py:93-100) and trainable mask u8 [..]
 * (dataset.py:87). */
RN_API int rn_labels_dense(const rn_rows_t* rows, int num_classes, const rn_maps_t* label,
                           const rn_maps_t* cls_onehot, const rn_maps_t* trainable_mask, void* stream);

/* augmentation.flip (augmentation.py:5-22) on compact labels: reverse the W axis of every
 * level and negate dx. in/out must not alias. */
RN_API int rn_flip_labels(const rn_pyramid_t* pyr, int batch, const rn_maps_t* label_in,
                          const rn_maps_t* deltas_in, const rn_maps_t* label_out,
                          const rn_maps_t* deltas_out, void* stream);

/* utils.iou (utils.py:62-97) elementwise on n pre-broadcast pairs. */
RN_API int rn_iou_pairwise(int64_t n, const float* d_a, const float* d_b, float* d_out,
                           uint32_t* d_status, void* stream);

/* ---- losses -------------------------------------------------------------------------- */
/* Label statistics (M, num_fg, n_pos[c]) from compact labels; ADDS into d_stats. */
RN_API int rn_label_stats(const rn_rows_t* rows, int num_classes, const rn_maps_t* label,
                          double* d_stats, void* stream);

/* Same from dense float labels [rows, C] (the reference's own label format, every row
 * trainable): M = rows, num_fg = #rows with max_c z > 0.5 (utils.classmap_decode,
 * utils.py:171-179), n_pos[c] = sum z.  ADDS into d_stats. */
RN_API int rn_label_stats_dense(int64_t rows, int num_classes, const float* d_labels,
                                double* d_stats, void* stream);

RN_API size_t rn_loss_workspace_bytes(int num_classes);

/* Per-class sums needed before a dice gradient (losses.py:53-56): WRITES I_c and sum p into
 * d_stats.  dense_labels == 0: label maps are int32 compact; 1: float [rows, C] (every row
 * trainable).  logits_dtype: RN_DTYPE_F32, or RN_DTYPE_BF16 (compact labels, C % 8 == 0;
 * float32 math on bfloat16 logits). */
RN_API int rn_cls_sums(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits,
                       const rn_maps_t* label, int dense_labels, int logits_dtype, double* d_stats,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* Fused classification loss forward + backward in one pass over the logits
 * (losses.focal_sigmoid_cross_entropy_with_logits, balanced_sigmoid_cross_entropy_with_logits,
 * dice_loss, classification_loss; utils.process_labels_and_logits' sigmoid and trainable
 * masking, utils.py:249,270).  Reads the normalisers (M, num_fg, n_pos, I, sum p) from
 * d_stats, writes d(loss)/d(logits) for every element (0 on ignored rows) and WRITES
 * d_stats[RN_STAT_CLS_SUM].  dlogits may be NULL maps (p[0]==NULL) for forward only. */
RN_API int rn_cls_loss_fwd_bwd(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits,
                               const rn_maps_t* label, int dense_labels, int mode, float focus,
                               float alpha, float eps, const rn_maps_t* dlogits, int logits_dtype,
                               double* d_stats, void* d_workspace, size_t workspace_bytes, void* stream);

/* losses.regression_loss (losses.py:144-152): Huber(delta=1) on fg anchors, gradient
 * clip(e,-1,1)/(4 num_fg) (num_fg read from d_stats), WRITES d_stats[RN_STAT_HUBER_SUM].
 * fg = label >= 0 (compact) or d_fg_mask u8 (when label maps are NULL). */
RN_API int rn_huber_fwd_bwd(const rn_rows_t* rows, const rn_maps_t* pred, const rn_maps_t* target,
                            const rn_maps_t* label, const rn_maps_t* fg_mask, const rn_maps_t* dpred,
                            double* d_stats, void* d_workspace, size_t workspace_bytes, void* stream);

/* The loss stage of a training step in ONE launch (train.py:218-220): classification loss forward + backward
 * (as rn_cls_loss_fwd_bwd, compact labels), Huber regression loss forward + backward (as rn_huber_fwd_bwd,
 * fg = label >= 0), and -- capacity > 0 -- the candidate pre-selection of rn_cls_loss_select
 * (utils.boxes_decode, utils.py:183-190; finish with rn_select_finalize).  d_loss_out (optional, float[4]): the
 * final scalars of rn_loss_finalize, written by the same launch -- single GPU only, where no statistics exchange
 * has to happen between the sums and the normalisation.  d_stats_next (optional): another statistics vector
 * (RN_STATS_HEADER + 3 * num_classes doubles) that the same launch leaves zeroed for the next step's
 * rn_assign_encode, so a training loop that rotates its statistics vectors needs no memset.  The first 256 bytes of d_loss_workspace must be zero
 * before the first use; every kernel leaves them zero.  Shapes the streaming kernel does not take (C % 4 != 0,
 * focus != 2) run as separate launches; the fused selection is then RN_ERR_UNSUPPORTED
 * (rn_fused_select_supported). */
RN_API int rn_head_loss(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits, const rn_maps_t* label,
                        int mode, float focus, float alpha, float eps, const rn_maps_t* dlogits, int logits_dtype,
                        const rn_maps_t* reg_pred, const rn_maps_t* reg_target, const rn_maps_t* dreg,
                        double* d_stats, double* d_stats_next, void* d_loss_workspace, size_t loss_workspace_bytes,
                        float* d_loss_out, float score_thr, int capacity, int32_t* d_cand_count, uint32_t* d_status,
                        void* d_select_workspace, size_t select_workspace_bytes, void* stream);

/* The same loss for the reference's own call shape: [rows, width] arrays of any width and a
 * u8 fg mask [rows] (losses_test.py:17-27 uses width 1).  Reads num_fg from d_stats, writes
 * the scalar loss to d_loss and (optional) the gradient to d_dpred. */
RN_API int rn_huber_rows(int64_t rows, int width, const float* d_pred, const float* d_target,
                         const uint8_t* d_fg_mask, float* d_dpred, const double* d_stats,
                         float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream);

/* Final scalars from the (all-reduced) stats: d_out[0] = classification loss, d_out[1] =
 * regression loss, d_out[2] = num_fg, d_out[3] = M  (losses.loss, losses.py:155-175). */
RN_API int rn_loss_finalize(int num_classes, int mode, const double* d_stats, float* d_out, void* stream);

/* Elementwise, unreduced losses with the reference's return shape [rows, C]
 * (losses.py:6-15 focal; tf.nn.sigmoid_cross_entropy_with_logits; balanced BCE with
 * per-class weights d_w_pos/d_w_neg [C] or NULL).  d_grad (optional) = d loss / d logits. */
RN_API int rn_cls_loss_elementwise(int64_t rows, int num_classes, const float* d_labels,
                                   const float* d_logits, int mode, float focus, float alpha,
                                   float eps, const float* d_w_pos, const float* d_w_neg,
                                   float* d_loss, float* d_grad, void* stream);

/* losses.focal_softmax_cross_entropy_with_logits (losses.py:19-34): per-row loss [rows] and
 * (optional) d loss[row] / d logits [rows, C]. */
RN_API int rn_focal_softmax(int64_t rows, int num_classes, const float* d_labels,
                            const float* d_logits, float focus, float alpha, float eps,
                            float* d_loss, float* d_grad, void* stream);

/* Backward of a loss of the per-class sums I_c = sum_m z p and P_c = sum_m p (dice_loss,
 * jaccard_loss, fixed_iou_loss; losses.py:37-73): grad[m,c] = (z*a[c] + b[c]) * p (1 - p),
 * with a = dL/dI and b = dL/dP supplied per class. */
RN_API int rn_colsum_backward(int64_t rows, int num_classes, const float* d_labels,
                              const float* d_logits, const float* d_a, const float* d_b,
                              float* d_grad, void* stream);

/* ---- trainable masking ------------------------------------------------------------------- */
/* tf.boolean_mask(x, trainable_mask) per level + concat over levels (utils.postprocess_and_mask,
 * utils.py:258-284; utils.merge_outputs, utils.py:151-153): order-preserving compaction of the rows whose
 * u8 mask is non-zero, (level, b, y, x, a) order.  rn_mask_scan counts and scans once per mask (d_count:
 * device int64, M); rn_mask_gather copies one tensor ([rows, row_bytes] per level -> d_out [M, row_bytes]) and
 * can be repeated for every tensor that shares the mask (labels, logits, probabilities, regressions). */
RN_API size_t rn_mask_workspace_bytes(const rn_rows_t* rows);
RN_API int rn_mask_scan(const rn_rows_t* rows, const rn_maps_t* mask, void* d_workspace, size_t workspace_bytes,
                        int64_t* d_count, void* stream);
RN_API int rn_mask_gather(const rn_rows_t* rows, int row_bytes, const rn_maps_t* src, const rn_maps_t* mask,
                          const void* d_workspace, void* d_out, void* stream);

/* ---- the two ends of the path on the device --------------------------------------- */
/* data_loaders/shapes.py:26-53,113-131: rasterise the synthetic Shapes images into an image batch
 * d_image [B, H, W, 3] u8 (RGB as the reference's colour tuples): background colour d_background [B, 3], then
 * d_counts[b] shapes in list order (later ones paint over earlier ones).  d_specs [B, max_shapes, 8] int32 =
 * (kind, x, y, s, r, g, b, 0), kind 0 square / 1 triangle / 2 circle (the reference's class order).  Squares and
 * circles match cv2.rectangle / cv2.circle pixel for pixel; triangles match cv2.fillPoly up to isolated edge pixels. */
RN_API int rn_shapes_render(int batch, int height, int width, int max_shapes, const int32_t* d_specs,
                            const int32_t* d_counts, const uint8_t* d_background, uint8_t* d_image, void* stream);

/* utils.draw_bounding_boxes (utils.py:120-148) on the device detection lists: 1-pixel outlines of
 * d_boxes [B, capacity, 4] (normalised [ymin, xmin, ymax, xmax]; pixel = round(box * [H, W, H, W])) in the colour
 * d_palette [num_classes, 3] of d_classes, for the first d_counts[b] entries of every image.  No text labels. */
RN_API int rn_draw_boxes(int batch, int height, int width, int capacity, const float* d_boxes,
                         const int32_t* d_classes, const int32_t* d_counts, const uint8_t* d_palette,
                         int num_classes, uint8_t* d_image, void* stream);

/* ---- multi-GPU exchange over NVLink peer memory -------------------------------------------- */
/* All-reduce(sum) of d_stats[offset : offset + len) across `world` ranks WITHOUT a collective
 * library: every rank owns a mailbox of rn_stats_mailbox_bytes() in peer-mapped memory (zeroed
 * once); d_peer_mailboxes is a DEVICE array of the `world` mailbox base pointers as mapped into
 * this process (e.g. torch symmetric memory's buffer_ptrs, cuMem/IPC mappings).  One small
 * kernel stores the slice into every peer, publishes `epoch` (must increase by 1 per use of a
 * phase; use slot = 2 * phase + (epoch & 1) so that a slot is never rewritten before it is read),
 * waits for all ranks' messages and sums them in rank order: results are bit-identical on all
 * ranks.  A peer that does not arrive within the limit (default 10 min, rn_set_option("exchange_timeout_ms", ms))
 * is FATAL: the slice becomes NaN and RN_FLAG_EXCHANGE_TIMEOUT is raised -- never local-only statistics.
 * rn_stats_exchange_auto keeps the epochs in device memory (d_epochs: `phases` uint64 counters, zeroed once,
 * incremented by the kernel; mailbox sized for slots = 2 * phases), which makes the call graph-capturable. */
RN_API size_t rn_stats_mailbox_bytes(int world, int len_max, int slots);
RN_API int rn_stats_exchange(double* d_stats, int offset, int len, int len_max, int world, int rank,
                             const void* const* d_peer_mailboxes, int slot, int slots, uint64_t epoch,
                             uint32_t* d_status, void* stream);
RN_API int rn_stats_exchange_auto(double* d_stats, int offset, int len, int len_max, int world, int rank,
                                  const void* const* d_peer_mailboxes, int phase, int phases, uint64_t* d_epochs,
                                  uint32_t* d_status, void* stream);

/* ---- decode + select + NMS ------------------------------------------------------------ */
/* tf.nn.sigmoid (utils.py:249) with the same device function the fused kernels use. */
RN_API int rn_sigmoid(int64_t n, const float* d_in, float* d_out, void* stream);

/* utils.regression_postprocess (utils.py:108-117) over whole maps: deltas -> corner boxes. */
RN_API int rn_decode_boxes(const rn_pyramid_t* pyr, int batch, const rn_maps_t* deltas,
                           const rn_maps_t* boxes, void* stream);

/* The other elementwise box-map functions of the reference, one float4 per anchor, in != out
 * not required. */
#define RN_BOXOP_DECODE 0            /* utils.regression_postprocess            utils.py:108-117 */
#define RN_BOXOP_ADD_CENTRE 1        /* utils.boxmap_anchor_relative_to_image_relative  :22-36   */
#define RN_BOXOP_SCALE 2             /* utils.scale_regression                  utils.py:100-105 */
#define RN_BOXOP_CENTRE_TO_CORNER 3  /* utils.boxmap_center_relative_to_corner_relative :39-44,
                                        dataset.from_center_box                 dataset.py:35-39 */
#define RN_BOXOP_CORNER_TO_CENTRE 4  /* dataset.to_center_box                   dataset.py:28-32 */
RN_API int rn_boxmap_op(const rn_pyramid_t* pyr, int batch, int op, const rn_maps_t* in,
                        const rn_maps_t* out, void* stream);

/* utils.classmap_decode (utils.py:171-179): fg_mask[r] = max_c classmap[r, c] > thr. */
RN_API int rn_classmap_fg(int64_t rows, int num_classes, const float* d_classmap, float thr,
                          uint8_t* d_fg_mask, void* stream);

#define RN_SCORES_LOGITS 1
#define RN_SCORES_BF16 2

RN_API size_t rn_select_workspace_bytes(int batch, int capacity);

/* utils.boxes_decode + merge_boxes_decoded (utils.py:183-195,223-227) for a batch:
 * per anchor score = max_c prob, class = argmax_c prob (first maximum), keep score >
 * score_thr (strict), candidates of image b in ascending level-major anchor order.
 *   scores maps: [.., C]; score_format = RN_SCORES_LOGITS (apply sigmoid in-kernel) |
 *                RN_SCORES_BF16 (bfloat16 instead of float32; C % 8 == 0); 0 = f32 probabilities.
 *   regs maps:   f32 [.., 4]; regs_are_deltas != 0 decodes them on the fly (needs pyr),
 *                otherwise they are already corner boxes (pyr may describe geometry only).
 *   topk_per_level > 0: NOT in the reference -- keep the k best of each (image, level)
 *                (ties -> lower anchor), identity when k >= #candidates.
 *   topk_per_level == -1: select pass only -- the candidates stay unordered in d_workspace and d_cand_count holds the
 *                appended count; finish with rn_detect_tail (one launch) or rn_select_finalize.
 * Outputs [B, capacity, .]; d_cand_anchor = anchor index within the image; d_cand_count [B]. */
RN_API int rn_decode_select(const rn_pyramid_t* pyr, int batch, int num_classes,
                            const rn_maps_t* scores, int score_format, const rn_maps_t* regs,
                            int regs_are_deltas, float score_thr, int topk_per_level, int capacity,
                            float* d_cand_boxes, float* d_cand_scores, int32_t* d_cand_classes,
                            int32_t* d_cand_anchor, int32_t* d_cand_count, uint32_t* d_status,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* Second half of rn_decode_select on its own: order the (unordered) candidates a selection pass
 * left in the workspace, optionally drop duplicate anchors (dedup != 0, needed after
 * rn_cls_loss_select), apply the per-level top-k, decode / gather the boxes. */
RN_API int rn_select_finalize(const rn_pyramid_t* pyr, int batch, const rn_maps_t* regs,
                              int regs_are_deltas, int topk_per_level, int capacity, int dedup,
                              float* d_cand_boxes, float* d_cand_scores, int32_t* d_cand_classes,
                              int32_t* d_cand_anchor, int32_t* d_cand_count, void* d_workspace,
                              size_t workspace_bytes, void* stream);

/* Training step with detections from the same logits (train.py:218-220 + 68-75): ONE pass over
 * the logits computes the focal loss forward + backward (as rn_cls_loss_fwd_bwd) AND pre-selects
 * the detection candidates (as the first half of rn_decode_select with logits as scores), so
 * the [rows, C] tensor is read once instead of twice.  Follow with rn_select_finalize(dedup=1).
 * Available when rn_fused_select_supported() != 0 (focal, focus == 2, compact labels,
 * C % 4 == 0, < 2^32 float4 per level); 0 < score_thr < 1. */
RN_API int rn_fused_select_supported(int num_classes, int mode, float focus, const rn_rows_t* rows);
RN_API int rn_cls_loss_select(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits,
                              const rn_maps_t* label, int mode, float focus, float alpha, float eps,
                              const rn_maps_t* dlogits, int logits_dtype, double* d_stats,
                              void* d_loss_workspace,
                              size_t loss_workspace_bytes, float score_thr, int capacity,
                              int32_t* d_cand_count, uint32_t* d_status, void* d_select_workspace,
                              size_t select_workspace_bytes, void* stream);

RN_API size_t rn_nms_workspace_bytes(int batch, int num_classes, int capacity);

/* utils.nms_classwise + utils.nms (utils.py:198-220): per class greedy NMS with
 * tf.image.non_max_suppression semantics (descending score, ties -> lower index, suppress
 * iff IoU > iou_thr, at most max_output_per_class per class), output class-major, selection
 * order within a class.  Inputs are the [B, capacity, .] candidate arrays of
 * rn_decode_select (any source with d_cand_count valid entries per image works).
 * d_det_index = index into the image's candidate list. */
RN_API int rn_nms_classwise(int batch, int num_classes, int capacity, const float* d_cand_boxes,
                            const float* d_cand_scores, const int32_t* d_cand_classes,
                            const int32_t* d_cand_count, float iou_thr, int max_output_per_class,
                            int out_capacity, float* d_det_boxes, float* d_det_scores,
                            int32_t* d_det_classes, int32_t* d_det_index, int32_t* d_det_count,
                            uint32_t* d_status, void* d_workspace, size_t workspace_bytes, void* stream);

/* The detection tail of a step in ONE launch (one CTA per image): rn_select_finalize (order by anchor, drop
 * duplicate anchors when dedup != 0, decode / gather the boxes -> d_cand_*, d_cand_count) followed by
 * rn_nms_classwise on those candidates (-> d_det_*), i.e. utils.boxes_decode + merge_boxes_decoded +
 * nms_classwise (utils.py:183-227, train.py:68-75) after a select pass (rn_decode_select's first half,
 * rn_cls_loss_select or rn_head_loss) has left the unordered candidates in d_select_workspace.  Results are
 * identical to the two separate calls (topk_per_level as in rn_decode_select; 0 = off).  Available when
 * rn_detect_tail_supported() != 0 (capacity <= 8192); otherwise call the separate entry points.
 * rn_decode_select(topk_per_level = -1) runs only the select pass and leaves the candidates for this call. */
RN_API int rn_detect_tail_supported(int num_classes, int capacity);
RN_API int rn_detect_tail(const rn_pyramid_t* pyr, int batch, int num_classes, const rn_maps_t* regs,
                          int regs_are_deltas, int capacity, int dedup, int topk_per_level,
                          float* d_cand_boxes, float* d_cand_scores, int32_t* d_cand_classes,
                          int32_t* d_cand_anchor,
                          int32_t* d_cand_count, float iou_thr, int max_output_per_class, int out_capacity,
                          float* d_det_boxes, float* d_det_scores, int32_t* d_det_classes,
                          int32_t* d_det_index, int32_t* d_det_count, uint32_t* d_status,
                          void* d_select_workspace, size_t select_workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RNDENSE_H_ */
