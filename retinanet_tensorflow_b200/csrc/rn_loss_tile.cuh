// Host-side interface of the tile-streaming classification-loss kernel (rn_loss_tile.cu).
#pragma once

#include "rn_common.cuh"

namespace rn {

struct TileParams {
  Ptrs logits, label, dlogits;   // label: compact int32 [rows]; dense float [rows, C] when `dense`
  Ptrs pred, target, dpred;      // fused Huber regression loss (pred.p[0] == nullptr: off); needs compact labels
  float alpha, eps;
  const double* stats;
  double* stats_out;
  void* workspace;               // rn_loss_workspace_bytes(): counters in the first 256 bytes, per-CTA partials after
  int num_classes;
  int bf16;                      // logits / dlogits are bfloat16 (bf16x8 units)
  int dense;                     // dense float labels, every row trainable (the reference's own loss signatures)
  int write_grad;
  float* loss_out;               // optional [4]: final scalars (as rn_loss_finalize) written by the last CTA
  int final_mode;                //   ... for this RN_CLS_* | RN_CLS_DICE_FLAG mode
  double* stats_next;            // optional: another statistics vector (RN_STATS_HEADER + 3C), zeroed by the last CTA
  // fused candidate pre-selection (rn_cls_loss_select, utils.boxes_decode, utils.py:183-190)
  float sel_trigger, sel_thr;
  int sel_images, sel_capacity;
  int32_t* sel_anchor;
  float* sel_score;
  int32_t* sel_class;
  int32_t* sel_count;
  uint32_t* status;
};

// true when the tile kernel can take this shape (C % 4 == 0 resp. % 8 for bf16, a row fits a stage)
bool cls_tile_supported(int num_classes, bool bf16, bool dense);

// mode: RN_CLS_* | RN_CLS_DICE_FLAG.  Returns RN_OK or an rn_status (error text set).
int launch_cls_tile(const Segs& segs, const TileParams& tp, int mode, bool select, cudaStream_t st);

}  // namespace rn
