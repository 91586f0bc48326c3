// K1: anchor-vs-GT IoU matching + box-delta encoding (dataset.level_labels, dataset.py:43-123),
// plus the small label-format kernels around it (dense one-hot, flip, pairwise IoU, label stats).
//
// One CTA owns a kTileH x kTileW patch of cells of one (image, level).  Anchors are regenerated
// in registers (never read from memory); the image's GT boxes are converted corner -> centre ->
// corner exactly as the reference does (dataset.py:48,74), culled against the patch's bounding
// box (a GT that is disjoint from the patch has IoU == 0 with every anchor of it, so it can
// neither win nor tie above 0) and staged in shared memory in their original order, which keeps
// tf.argmax's first-maximum rule.  Each thread then scans the staged boxes for its anchors with
// the reference's float32 op sequence (no FMA contraction, IEEE division).
#include "rn_common.cuh"
#include "rn_tiles.cuh"

namespace rn {

constexpr int kApt = 3;          // anchors per thread
constexpr int kMaxGtSmem = 2048; // GT boxes per image the assign kernel can stage

struct AssignParams {
  const float4* gt_boxes;   // [B, max_gt]
  const int32_t* gt_classes;
  const int32_t* gt_counts;
  int max_gt;
  int num_classes;
  float neg_thr, pos_thr;
  Ptrs matched, label, deltas, iou_max;
  double* stats;
  uint32_t* status;
  int tpc;  // threads per cell
};

__global__ void __launch_bounds__(kTileH * kTileW * 6)

assign_encode_kernel(const __grid_constant__ rn_pyramid_t pyr, const __grid_constant__ AssignParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_box = reinterpret_cast<float4*>(smem_raw);                       // [max_gt] recomputed corners
  float4* s_ctr = s_box + prm.max_gt;                                        // [max_gt] (cy, cx, h, w) of the raw box
  float* s_area = reinterpret_cast<float*>(s_ctr + prm.max_gt);              // [max_gt]
  int* s_idx = reinterpret_cast<int*>(s_area + prm.max_gt);                  // [max_gt] original index
  int* s_cls = s_idx + prm.max_gt;                                           // [max_gt] class id
  __shared__ int s_warp_cnt[32];
  __shared__ float4 s_ctr0;   // GT 0 (the match of anchors that overlap nothing, dataset.py:76)
  __shared__ int s_cls0;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / pyr.tiles_per_image;
  const int tile = blockIdx.x - b * pyr.tiles_per_image;
  int l = 0;
  for (int k = 1; k < pyr.num_levels; ++k) l += (tile >= pyr.level[k].tile_start) ? 1 : 0;
  const rn_level_t& L = pyr.level[l];
  const int t_local = tile - L.tile_start;
  const int ty0 = (t_local / L.tiles_x) * kTileH;
  const int tx0 = (t_local % L.tiles_x) * kTileW;
  const int ty1 = min(ty0 + kTileH, L.grid_h) - 1;  // inclusive
  const int tx1 = min(tx0 + kTileW, L.grid_w) - 1;

  // ---- patch bounding box: fsub/fadd are monotone, so the extremes come from the corner cells
  // and the level's largest half sizes (precomputed on the host)
  const float T0 = fsub(cell_centre(L.y0, L.y_step, ty0), L.max_half_h);
  const float T1 = fsub(cell_centre(L.x0, L.x_step, tx0), L.max_half_w);
  const float T2 = fadd(cell_centre(L.y0, L.y_step, ty1), L.max_half_h);
  const float T3 = fadd(cell_centre(L.x0, L.x_step, tx1), L.max_half_w);

  // ---- stage the GT boxes that touch the patch, in original order.  The count and the boxes
  // are loaded together (padding rows are simply masked afterwards).
  const float4* gt = prm.gt_boxes + (size_t)b * prm.max_gt;
  const int32_t* gcls = prm.gt_classes + (size_t)b * prm.max_gt;
  const int n_gt = min(max(__ldg(prm.gt_counts + b), 0), prm.max_gt);
  unsigned flags = 0;
  int n_kept = 0;
  // (the first chunk is processed unconditionally so that its loads do not wait for the count)
  for (int base = 0; base < prm.max_gt && (base == 0 || base < n_gt); base += blockDim.x) {
    const int i = base + tid;
    bool keep = false;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), ctr = g;
    int cls = -1;
    if (i < prm.max_gt) {
      const float4 raw = __ldg(gt + i);
      cls = __ldg(gcls + i);
      // to_center_box (dataset.py:28-32) then from_center_box (dataset.py:35-39)
      const float sh = fsub(raw.z, raw.x), sw = fsub(raw.w, raw.y);
      const float cy = fadd(raw.x, sh * 0.5f), cx = fadd(raw.y, sw * 0.5f);
      ctr = make_float4(cy, cx, sh, sw);
      g = make_float4(fsub(cy, sh * 0.5f), fsub(cx, sw * 0.5f), fadd(cy, sh * 0.5f), fadd(cx, sw * 0.5f));
      if (i < n_gt) {
        if (!(g.x <= g.z && g.y <= g.w)) flags |= RN_FLAG_BOX_ORDER;  // utils.py:67
        // utils.py:94 (0 <= iou <= 1): anchors have positive finite area, so an IoU can only leave
        // [0, 1] or be NaN when a GT coordinate is not finite -- checked here once per box instead of
        // once per pair.
        if (!(fabsf(g.x) <= 3.0e38f && fabsf(g.y) <= 3.0e38f && fabsf(g.z) <= 3.0e38f && fabsf(g.w) <= 3.0e38f))
          flags |= RN_FLAG_IOU_RANGE;
        keep = !(g.z < T0 || g.x > T2 || g.w < T1 || g.y > T3);
        if (i == 0) {
          s_ctr0 = ctr;
          s_cls0 = cls;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = n_kept, tot = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) {
      const int c = s_warp_cnt[w];
      off += w < warp ? c : 0;
      tot += c;
    }
    if (keep) {
      const int pos = off + __popc(m & ((1u << lane) - 1u));
      s_box[pos] = g;
      s_ctr[pos] = ctr;
      s_area[pos] = fmul(fsub(g.z, g.x), fsub(g.w, g.y));  // box_b_area (utils.py:84-85)
      s_idx[pos] = i;
      s_cls[pos] = cls;
    }
    n_kept += tot;
    __syncthreads();
  }

  // ---- this thread's anchors
  const int cell_local = tid / prm.tpc;
  const int sub = tid - cell_local * prm.tpc;
  const int cy_i = ty0 + cell_local / kTileW;
  const int cx_i = tx0 + cell_local % kTileW;
  const bool cell_ok = cell_local < kTileH * kTileW && cy_i < L.grid_h && cx_i < L.grid_w;
  const float cy = cell_centre(L.y0, L.y_step, cy_i);
  const float cx = cell_centre(L.x0, L.x_step, cx_i);

  float a0[kApt], a1[kApt], a2[kApt], a3[kApt], aa[kApt], best[kApt];
  int bestj[kApt];
  bool ok[kApt];
#pragma unroll
  for (int k = 0; k < kApt; ++k) {
    const int a = sub * kApt + k;
    ok[k] = cell_ok && a < L.num_anchors;
    const float h = L.anchor_h[min(a, RN_MAX_ANCHORS - 1)], w = L.anchor_w[min(a, RN_MAX_ANCHORS - 1)];
    a0[k] = fsub(cy, h * 0.5f);  // from_center_box(anchor), dataset.py:74
    a1[k] = fsub(cx, w * 0.5f);
    a2[k] = fadd(cy, h * 0.5f);
    a3[k] = fadd(cx, w * 0.5f);
    aa[k] = fmul(fsub(a2[k], a0[k]), fsub(a3[k], a1[k]));  // box_a_area (utils.py:82-83)
    best[k] = 0.f;
    bestj[k] = -1;
  }

  // ---- IoU scan (utils.iou, utils.py:70-92), strict '>' keeps the first maximum.  The IEEE
  // division only runs when the pair can beat the running best: rcp.approx bounds the quotient
  // to 4e-7 relative, so `approx * (1 + 2e-6) <= best` proves fl(inter / uni) <= best.
  // Two cheap rejections before the per-anchor work (a rejected pair has IoU == 0, which can neither win nor
  // tie above 0, so skipping it cannot change the arg-max or its first-maximum rule):
  //   * per warp: the union of the warp's anchor boxes (shuffle-reduced once) against the GT -- uniform branch;
  //   * per thread: the thread's three anchors are concentric (one aspect ratio, three scales), so a GT that
  //     misses the hull of the three misses all of them.
  float h0 = a0[0], h1 = a1[0], h2 = a2[0], h3 = a3[0];
#pragma unroll
  for (int k = 1; k < kApt; ++k) {
    h0 = fminf(h0, a0[k]);
    h1 = fminf(h1, a1[k]);
    h2 = fmaxf(h2, a2[k]);
    h3 = fmaxf(h3, a3[k]);
  }
  const float w0 = warp_min(h0), w1 = warp_min(h1), w2 = warp_max(h2), w3 = warp_max(h3);
  for (int j = 0; j < n_kept; ++j) {
    const float4 g = s_box[j];
    if (g.z < w0 || g.x > w2 || g.w < w1 || g.y > w3) continue;   // misses every anchor of this warp
    if (g.z < h0 || g.x > h2 || g.w < h1 || g.y > h3) continue;   // misses this thread's anchors
    const float ga = s_area[j];
#pragma unroll
    for (int k = 0; k < kApt; ++k) {
      const float yt = fmaxf(a0[k], g.x), xl = fmaxf(a1[k], g.y);
      const float yb = fminf(a2[k], g.z), xr = fminf(a3[k], g.w);
      if (!(yb < yt || xr < xl)) {
        const float inter = fmul(fsub(yb, yt), fsub(xr, xl));
        const float uni = fsub(fadd(aa[k], ga), inter);
        const float approx = inter * rcp_approx(uni);
        // (best == 0: always take the exact division -- a denormal intersection flushes `approx` to 0, but the
        //  reference's argmax still prefers that tiny positive IoU over "no overlap")
        if (approx * 1.000002f > best[k] || best[k] == 0.f) {
          const float v = fdiv(inter, uni);
          if (v > best[k]) {
            best[k] = v;
            bestj[k] = j;
          }
        }
      }
    }
  }

  // ---- labels + encoding for the matched box (all anchors, dataset.py:105-121)
  int n_ign = 0, n_fg = 0;   // both rare: the common warp issues no atomic at all
  if (n_gt == 0 && tid == 0) flags |= RN_FLAG_EMPTY_GT;
  const long long cell_row = ((long long)b * L.grid_h * L.grid_w + (long long)cy_i * L.grid_w + cx_i) * L.num_anchors;
#pragma unroll
  for (int k = 0; k < kApt; ++k) {
    if (!ok[k]) continue;
    const int a = sub * kApt + k;
    const long long row = cell_row + a;
    int lab = RN_LABEL_BACKGROUND;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    int matched = -1;
    if (n_gt > 0) {
      const bool hit = bestj[k] >= 0;
      matched = hit ? s_idx[bestj[k]] : 0;
      const float4 t = hit ? s_ctr[bestj[k]] : s_ctr0;   // (cy, cx, h, w) of the matched raw box
      const bool pos = best[k] >= prm.pos_thr;                 // !bg_mask, dataset.py:83
      const bool ign = !pos && !(best[k] < prm.neg_thr);       // !trainable_mask, dataset.py:87
      if (pos) {
        const int c = hit ? s_cls[bestj[k]] : s_cls0;
        lab = (c >= 0 && c < prm.num_classes) ? c : RN_LABEL_BACKGROUND;  // one_hot of an out-of-range id is 0
      } else if (ign) {
        lab = RN_LABEL_IGNORE;
      }
      // deltas are compared within 1e-5 relative, not bit-exactly.  The shifts take a reciprocal (rcp.approx + one
      // Newton step, <= 1 ulp of the quotient); the scale ratios keep the IEEE division, because log() of a ratio
      // near 1 turns a 2-ulp ratio error into an arbitrarily large RELATIVE error of a near-zero delta.
      const float h = L.anchor_h[a], w = L.anchor_w[a];
      d.x = fsub(t.x, cy) * rcp_newton(h);   // shifts, dataset.py:108
      d.y = fsub(t.y, cx) * rcp_newton(w);
      const float rh = fdiv(t.z, h), rw = fdiv(t.w, w);   // scales, dataset.py:110
      d.z = log_fast(rh);                                 // log(scales), dataset.py:112
      d.w = log_fast(rw);
      // log_fast is for positive normal finite arguments (sign 0, exponent field 1..254); anything else -- a
      // zero-area, denormal-sized, negative or non-finite box -- takes libm in one shared rare branch
      const bool nh = (unsigned)(__float_as_int(rh) - 0x00800000) < 0x7f000000u;
      const bool nw = (unsigned)(__float_as_int(rw) - 0x00800000) < 0x7f000000u;
      if (!(nh && nw)) {
        if (!nh) d.z = logf(rh);
        if (!nw) d.w = logf(rw);
      }
    }
    reinterpret_cast<int32_t*>(prm.matched.p[l])[row] = matched;
    reinterpret_cast<int32_t*>(prm.label.p[l])[row] = lab;
    reinterpret_cast<float4*>(prm.deltas.p[l])[row] = d;
    if (prm.iou_max.p[0]) reinterpret_cast<float*>(prm.iou_max.p[l])[row] = best[k];
    n_ign += lab == RN_LABEL_IGNORE;
    n_fg += lab >= 0;
    if (prm.stats && lab >= 0) atomicAdd(prm.stats + RN_STAT_NPOS(lab), 1.0);
  }

  if (prm.stats) {
    // M = all anchors - ignored: the constant part is added once, ignored anchors are rare
    if (blockIdx.x == 0 && tid == 0)
      atomicAdd(prm.stats + RN_STAT_TRAINABLE, (double)pyr.anchors_per_image * (double)(gridDim.x / pyr.tiles_per_image));
    if (__any_sync(0xffffffffu, (n_ign | n_fg) != 0)) {
      n_ign = warp_sum(n_ign);
      n_fg = warp_sum(n_fg);
      if (lane == 0) {
        if (n_ign) atomicAdd(prm.stats + RN_STAT_TRAINABLE, -(double)n_ign);
        if (n_fg) atomicAdd(prm.stats + RN_STAT_NUM_FG, (double)n_fg);
      }
    }
  }
  if (prm.status) {
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(prm.status, flags);
  }
}

// ---- compact labels -> dense one-hot + trainable mask (dataset.py:87-100) ----------------------
__global__ void labels_dense_kernel(const __grid_constant__ Segs segs, int num_classes,
                                    const __grid_constant__ Ptrs label, const __grid_constant__ Ptrs onehot,
                                    const __grid_constant__ Ptrs mask) {
  const long long total = segs.start[RN_MAX_LEVELS];
  const long long elems = total * num_classes;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (onehot.p[0]) {
    for (long long e = t0; e < elems; e += stride) {
      const long long r = e / num_classes;
      const int c = (int)(e - r * num_classes);
      const int s = seg_of(segs, r);
      const long long rl = r - segs.start[s];
      const int lab = reinterpret_cast<const int32_t*>(label.p[s])[rl];
      reinterpret_cast<float*>(onehot.p[s])[rl * num_classes + c] = (lab == c) ? 1.0f : 0.0f;
    }
  }
  if (mask.p[0]) {
    for (long long r = t0; r < total; r += stride) {
      const int s = seg_of(segs, r);
      const long long rl = r - segs.start[s];
      const int lab = reinterpret_cast<const int32_t*>(label.p[s])[rl];
      reinterpret_cast<unsigned char*>(mask.p[s])[rl] = lab != RN_LABEL_IGNORE;
    }
  }
}

// ---- augmentation.flip on compact labels (augmentation.py:5-22) --------------------------------
__global__ void flip_labels_kernel(const __grid_constant__ rn_pyramid_t pyr, int batch,
                                   const __grid_constant__ Ptrs lab_in, const __grid_constant__ Ptrs del_in,
                                   const __grid_constant__ Ptrs lab_out, const __grid_constant__ Ptrs del_out) {
  const long long total = pyr.anchors_per_image * (long long)batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // merged index i = batch * anchor_offset_l + local
    int l = 0;
    for (int k = 1; k < pyr.num_levels; ++k) l += (i >= pyr.level[k].anchor_offset * batch) ? 1 : 0;
    const rn_level_t& L = pyr.level[l];
    const long long local = i - L.anchor_offset * batch;
    const int a = (int)(local % L.num_anchors);
    const long long cell = local / L.num_anchors;
    const int x = (int)(cell % L.grid_w);
    const long long by = cell / L.grid_w;  // b * H + y
    const long long src = (by * L.grid_w + (L.grid_w - 1 - x)) * L.num_anchors + a;
    if (lab_in.p[0]) reinterpret_cast<int32_t*>(lab_out.p[l])[local] = reinterpret_cast<const int32_t*>(lab_in.p[l])[src];
    if (del_in.p[0]) {
      float4 d = reinterpret_cast<const float4*>(del_in.p[l])[src];
      d.y = -d.y;
      reinterpret_cast<float4*>(del_out.p[l])[local] = d;
    }
  }
}

// ---- utils.iou, elementwise (utils.py:62-97) ----------------------------------------------------
__global__ void iou_pairwise_kernel(long long n, const float4* __restrict__ a, const float4* __restrict__ b,
                                    float* __restrict__ out, uint32_t* status) {
  unsigned flags = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = a[i], q = b[i];
    if (!(p.x <= p.z && p.y <= p.w) || !(q.x <= q.z && q.y <= q.w)) flags |= RN_FLAG_BOX_ORDER;
    const float yt = fmaxf(p.x, q.x), xl = fmaxf(p.y, q.y);
    const float yb = fminf(p.z, q.z), xr = fminf(p.w, q.w);
    const bool invalid = yb < yt || xr < xl;
    const float inter = fmul(fsub(yb, yt), fsub(xr, xl));
    const float area_a = fmul(fsub(p.z, p.x), fsub(p.w, p.y));
    const float area_b = fmul(fsub(q.z, q.x), fsub(q.w, q.y));
    float v = fdiv(inter, fsub(fadd(area_a, area_b), inter));
    v = invalid ? 0.f : v;
    if (!(v >= 0.f && v <= 1.f)) flags |= RN_FLAG_IOU_RANGE;
    out[i] = v;
  }
  if (status) {
    flags = __reduce_or_sync(__activemask(), flags);
    if ((threadIdx.x & 31) == 0 && flags) atomicOr(status, flags);
  }
}

// ---- label statistics ----------------------------------------------------------------------------
__global__ void label_stats_kernel(const __grid_constant__ Segs segs, int num_classes,
                                   const __grid_constant__ Ptrs label, double* stats) {
  const long long total = segs.start[RN_MAX_LEVELS];
  int n_ign = 0, n_fg = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total;
       r += (long long)gridDim.x * blockDim.x) {
    const int s = seg_of(segs, r);
    const int lab = reinterpret_cast<const int32_t*>(label.p[s])[r - segs.start[s]];
    n_ign += lab == RN_LABEL_IGNORE;
    if (lab >= 0 && lab < num_classes) {
      n_fg += 1;
      atomicAdd(stats + RN_STAT_NPOS(lab), 1.0);
    }
  }
  n_ign = warp_sum(n_ign);
  n_fg = warp_sum(n_fg);
  if ((threadIdx.x & 31) == 0) {
    if (n_ign) atomicAdd(stats + RN_STAT_TRAINABLE, -(double)n_ign);
    if (n_fg) atomicAdd(stats + RN_STAT_NUM_FG, (double)n_fg);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + RN_STAT_TRAINABLE, (double)total);
}

static unsigned grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace rn

using namespace rn;

extern "C" int rn_assign_encode(const rn_pyramid_t* pyr, int batch, const float* d_gt_boxes,
                                const int32_t* d_gt_classes, const int32_t* d_gt_counts, int max_gt,
                                int num_classes, float neg_iou_thr, float pos_iou_thr, const rn_maps_t* matched,
                                const rn_maps_t* label, const rn_maps_t* deltas, const rn_maps_t* iou_max,
                                double* d_stats, uint32_t* d_status, void* stream) {
  RN_REQUIRE(pyr && d_gt_boxes && d_gt_classes && d_gt_counts, RN_ERR_BAD_ARG, "rn_assign_encode: NULL input");
  RN_REQUIRE(maps_present(matched) && maps_present(label) && maps_present(deltas), RN_ERR_BAD_ARG,
             "rn_assign_encode: matched, label and deltas maps are required");
  RN_REQUIRE(batch >= 0 && num_classes >= 1, RN_ERR_BAD_ARG, "rn_assign_encode: bad batch/num_classes");
  RN_REQUIRE(max_gt >= 1 && max_gt <= kMaxGtSmem, RN_ERR_BAD_SHAPE, "max_gt must be in 1..%d", kMaxGtSmem);
  RN_REQUIRE(neg_iou_thr <= pos_iou_thr, RN_ERR_BAD_ARG, "neg_iou_thr must be <= pos_iou_thr");
  RN_REQUIRE(((uintptr_t)d_gt_boxes & 15) == 0, RN_ERR_BAD_ARG, "d_gt_boxes must be 16-byte aligned");
  if (batch == 0 || pyr->anchors_per_image == 0) return RN_OK;
  int na = 0;
  for (int l = 0; l < pyr->num_levels; ++l) {
    na = pyr->level[l].num_anchors > na ? pyr->level[l].num_anchors : na;
    RN_REQUIRE(matched->p[l] && label->p[l] && deltas->p[l], RN_ERR_BAD_ARG, "missing map pointer for level %d", l);
    RN_REQUIRE(((uintptr_t)deltas->p[l] & 15) == 0, RN_ERR_BAD_ARG, "deltas maps must be 16-byte aligned");
  }
  AssignParams prm;
  prm.gt_boxes = reinterpret_cast<const float4*>(d_gt_boxes);
  prm.gt_classes = d_gt_classes;
  prm.gt_counts = d_gt_counts;
  prm.max_gt = max_gt;
  prm.num_classes = num_classes;
  prm.neg_thr = neg_iou_thr;
  prm.pos_thr = pos_iou_thr;
  prm.matched = make_ptrs(matched);
  prm.label = make_ptrs(label);
  prm.deltas = make_ptrs(deltas);
  prm.iou_max = make_ptrs(maps_present(iou_max) ? iou_max : nullptr);
  prm.stats = d_stats;
  prm.status = d_status;
  prm.tpc = (na + kApt - 1) / kApt;
  const int threads = kTileH * kTileW * prm.tpc;
  const size_t smem = (size_t)max_gt * (2 * sizeof(float4) + sizeof(float) + 2 * sizeof(int));
  const long long blocks = (long long)batch * pyr->tiles_per_image;
  RN_REQUIRE(blocks < (1ll << 31), RN_ERR_BAD_SHAPE, "too many tiles");
  static size_t smem_opted = 48 * 1024;   // more than 48 KB of dynamic shared memory needs an opt-in (max_gt > ~1100)
  if (smem > smem_opted) {
    if (cudaFuncSetAttribute(assign_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      rn::set_error("rn_assign_encode: %zu bytes of shared memory for max_gt = %d refused", smem, max_gt);
      return RN_ERR_UNSUPPORTED;
    }
    smem_opted = smem;
  }
  static bool carveout_set = false;
  if (!carveout_set) {
    // Same shared-memory carve-out as the streaming loss kernel (which needs the maximum): two kernels only share an
    // SM when they agree on the L1 / shared split, and prefetch_assign runs this kernel beside that one.
    cudaFuncSetAttribute(assign_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  assign_encode_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(*pyr, prm);
  RN_CUDA_CHECK_LAUNCH("rn_assign_encode");
  return RN_OK;
}

extern "C" int rn_labels_dense(const rn_rows_t* rows, int num_classes, const rn_maps_t* label,
                               const rn_maps_t* cls_onehot, const rn_maps_t* trainable_mask, void* stream) {
  RN_REQUIRE(rows && maps_present(label), RN_ERR_BAD_ARG, "rn_labels_dense: NULL input");
  RN_REQUIRE(num_classes >= 1, RN_ERR_BAD_ARG, "rn_labels_dense: num_classes must be >= 1");
  RN_REQUIRE(maps_present(cls_onehot) || maps_present(trainable_mask), RN_ERR_BAD_ARG, "rn_labels_dense: no output");
  const Segs segs = make_segs(rows);
  const long long total = segs.start[RN_MAX_LEVELS];
  if (total == 0) return RN_OK;
  const long long work = maps_present(cls_onehot) ? total * num_classes : total;
  labels_dense_kernel<<<grid_for(work, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      segs, num_classes, make_ptrs(label), make_ptrs(maps_present(cls_onehot) ? cls_onehot : nullptr),
      make_ptrs(maps_present(trainable_mask) ? trainable_mask : nullptr));
  RN_CUDA_CHECK_LAUNCH("rn_labels_dense");
  return RN_OK;
}

extern "C" int rn_flip_labels(const rn_pyramid_t* pyr, int batch, const rn_maps_t* label_in,
                              const rn_maps_t* deltas_in, const rn_maps_t* label_out, const rn_maps_t* deltas_out,
                              void* stream) {
  RN_REQUIRE(pyr, RN_ERR_BAD_ARG, "rn_flip_labels: NULL pyramid");
  RN_REQUIRE(maps_present(label_in) == maps_present(label_out) && maps_present(deltas_in) == maps_present(deltas_out),
             RN_ERR_BAD_ARG, "rn_flip_labels: in/out maps must be given in pairs");
  RN_REQUIRE(maps_present(label_in) || maps_present(deltas_in), RN_ERR_BAD_ARG, "rn_flip_labels: nothing to do");
  const long long total = pyr->anchors_per_image * (long long)batch;
  if (total == 0) return RN_OK;
  flip_labels_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      *pyr, batch, make_ptrs(maps_present(label_in) ? label_in : nullptr),
      make_ptrs(maps_present(deltas_in) ? deltas_in : nullptr),
      make_ptrs(maps_present(label_out) ? label_out : nullptr),
      make_ptrs(maps_present(deltas_out) ? deltas_out : nullptr));
  RN_CUDA_CHECK_LAUNCH("rn_flip_labels");
  return RN_OK;
}

extern "C" int rn_iou_pairwise(int64_t n, const float* d_a, const float* d_b, float* d_out, uint32_t* d_status,
                               void* stream) {
  RN_REQUIRE(n >= 0, RN_ERR_BAD_ARG, "rn_iou_pairwise: negative n");
  if (n == 0) return RN_OK;
  RN_REQUIRE(d_a && d_b && d_out, RN_ERR_BAD_ARG, "rn_iou_pairwise: NULL pointer");
  RN_REQUIRE((((uintptr_t)d_a | (uintptr_t)d_b) & 15) == 0, RN_ERR_BAD_ARG, "boxes must be 16-byte aligned");
  iou_pairwise_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      n, reinterpret_cast<const float4*>(d_a), reinterpret_cast<const float4*>(d_b), d_out, d_status);
  RN_CUDA_CHECK_LAUNCH("rn_iou_pairwise");
  return RN_OK;
}

extern "C" int rn_label_stats(const rn_rows_t* rows, int num_classes, const rn_maps_t* label, double* d_stats,
                              void* stream) {
  RN_REQUIRE(rows && maps_present(label) && d_stats, RN_ERR_BAD_ARG, "rn_label_stats: NULL input");
  const Segs segs = make_segs(rows);
  const long long total = segs.start[RN_MAX_LEVELS];
  if (total == 0) return RN_OK;
  label_stats_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(segs, num_classes, make_ptrs(label),
                                                                               d_stats);
  RN_CUDA_CHECK_LAUNCH("rn_label_stats");
  return RN_OK;
}
