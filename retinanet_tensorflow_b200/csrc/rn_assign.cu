// K1: anchor-vs-GT IoU matching + box-delta encoding (dataset.level_labels, dataset.py:43-123),
// plus the small label-format kernels around it (dense one-hot, flip, pairwise IoU, label stats).
//
// One CTA owns a rectangular patch of cells of one (image, level): SY x SX sub-patches of kTileH x kTileW cells,
// visited one after the other by the same threads (3 anchors per thread).  Everything that does not depend on the
// anchor is done once per CTA: the image's GT boxes are converted corner -> centre -> corner exactly as the
// reference does (dataset.py:48,74), culled against the patch's bounding box (a GT that is disjoint from the patch
// has IoU == 0 with every anchor of it, so it can neither win nor tie above 0) and staged in shared memory in their
// original order, which keeps tf.argmax's first-maximum rule; the log size ratios of the encoding
// (dataset.py:110-112) depend only on (GT box, anchor shape), so they are tabulated once per CTA for the staged
// boxes instead of being evaluated per anchor.  Anchors are regenerated in registers (never read from memory); each
// thread scans the staged boxes for its anchors with the reference's float32 op sequence (no FMA contraction, IEEE
// division).  Round 2 profile of the one-sub-patch form (profiles/r02_k1_breakdown.md): 1327 instructions per warp
// of which only 29 % were the IoU scan -- 37 % were per-CTA fixed cost and 31 % the per-anchor encoding.
#include "rn_common.cuh"
#include "rn_tiles.cuh"

namespace rn {

constexpr int kApt = 3;          // anchors per thread
constexpr int kMaxGtSmem = 2048; // GT boxes per image the assign kernel can stage
#ifndef RN_ASSIGN_LUT
#define RN_ASSIGN_LUT 576
#endif
constexpr int kLutEntries = RN_ASSIGN_LUT;   // (staged GT + 1) x anchors-per-cell table of log size ratios (else: per anchor)

#ifndef RN_ASSIGN_SUPER_Y
#define RN_ASSIGN_SUPER_Y 2
#endif
#ifndef RN_ASSIGN_SUPER_X
#define RN_ASSIGN_SUPER_X 2
#endif

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
  // CTA patches: sy x sx sub-patches of kTileH x kTileW cells
  int sy, sx;
  int tiles_x[RN_MAX_LEVELS];     // CTA patches per grid row of the level
  int tile_start[RN_MAX_LEVELS];  // first CTA patch of the level within one image
};

// (log(t_h / a_h), log(t_w / a_w)), dataset.py:110-112.  The ratios keep the IEEE division: log() of a ratio near 1
// turns a 2-ulp ratio error into an arbitrarily large RELATIVE error of a near-zero delta.  log_fast is for positive
// normal finite arguments (sign 0, exponent field 1..254); anything else -- a zero-area, denormal-sized, negative or
// non-finite box -- takes libm in one shared rare branch.
__device__ __noinline__ float log_libm(float x) { return logf(x); }

__device__ __forceinline__ float2 log_size_ratios(float th, float tw, float h, float w) {
  const float rh = fdiv(th, h), rw = fdiv(tw, w);
  float2 d = make_float2(log_fast(rh), log_fast(rw));
  const bool nh = (unsigned)(__float_as_int(rh) - 0x00800000) < 0x7f000000u;
  const bool nw = (unsigned)(__float_as_int(rw) - 0x00800000) < 0x7f000000u;
  if (!(nh && nw)) {
    if (!nh) d.x = log_libm(rh);
    if (!nw) d.y = log_libm(rw);
  }
  return d;
}

// (the per-anchor form, taken only when too many boxes touch the patch for the table: kept out of line so that the
//  common path keeps its registers)
__device__ __noinline__ float2 log_size_ratios_direct(float th, float tw, float h, float w) {
  return log_size_ratios(th, tw, h, w);
}

// 64 registers (one 12-byte spill): five 192-thread CTAs per SM.  The compiler's own choice (80) leaves four and is
// 8 % slower, 56 spills into the scan loop (profiles/r02_k1_breakdown.md).
#ifndef RN_ASSIGN_MAXREG
#define RN_ASSIGN_MAXREG 64
#endif
#define RN_ASSIGN_REGS __maxnreg__(RN_ASSIGN_MAXREG)

__global__ void RN_ASSIGN_REGS
assign_encode_kernel(const __grid_constant__ rn_pyramid_t pyr, const __grid_constant__ AssignParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // slot 0 is GT 0 (the match of anchors that overlap nothing, dataset.py:76), slots 1.. the staged boxes
  const int slots = prm.max_gt + 1;
  float4* s_box = reinterpret_cast<float4*>(smem_raw);                       // [slots] recomputed corners
  float4* s_ctr = s_box + slots;                                             // [slots] (cy, cx, h, w) of the raw box
  float* s_area = reinterpret_cast<float*>(s_ctr + slots);                   // [slots]
  int* s_idx = reinterpret_cast<int*>(s_area + slots);                       // [slots] original index
  int* s_cls = s_idx + slots;                                                // [slots] class id
  __shared__ float2 s_lut[kLutEntries];
  __shared__ float4 s_shape[RN_MAX_ANCHORS];   // per anchor shape: (h / 2, w / 2, 1 / h, 1 / w)
  __shared__ int s_warp_cnt[32];
  __shared__ char* s_out[4];   // this image's first row in the level's matched / label / deltas / iou maps

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // grid: x = image, y = patch, coarse levels first -- their patches see most of the image's boxes and take longest,
  // so they start first and the many light fine-level patches fill the end of the launch
  const int b = blockIdx.x;
  const int tile = gridDim.y - 1 - blockIdx.y;
  int l = 0;
  for (int k = 1; k < pyr.num_levels; ++k) l += (tile >= prm.tile_start[k]) ? 1 : 0;
  const rn_level_t& L = pyr.level[l];
  const int t_local = tile - prm.tile_start[l];
  const int py = t_local / prm.tiles_x[l];
  const int ty0 = py * (kTileH * prm.sy);
  const int tx0 = (t_local - py * prm.tiles_x[l]) * (kTileW * prm.sx);
  const int ty1 = min(ty0 + kTileH * prm.sy, L.grid_h) - 1;  // inclusive
  const int tx1 = min(tx0 + kTileW * prm.sx, L.grid_w) - 1;
  const int A = L.num_anchors;

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
  if (tid == 0) {
    const long long img_row0 = (long long)b * L.grid_h * L.grid_w * A;
    s_out[0] = reinterpret_cast<char*>(prm.matched.p[l]) + img_row0 * 4;
    s_out[1] = reinterpret_cast<char*>(prm.label.p[l]) + img_row0 * 4;
    s_out[2] = reinterpret_cast<char*>(prm.deltas.p[l]) + img_row0 * 16;
    s_out[3] = prm.iou_max.p[0] ? reinterpret_cast<char*>(prm.iou_max.p[l]) + img_row0 * 4 : nullptr;
  }
  if (tid < A) {
    const float h = L.anchor_h[tid], w = L.anchor_w[tid];
    // the shifts take a reciprocal (rcp.approx + one Newton step, <= 1 ulp of the quotient); deltas are compared
    // within 1e-5 relative, not bit-exactly
    s_shape[tid] = make_float4(h * 0.5f, w * 0.5f, rcp_newton(h), rcp_newton(w));
  }
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
          s_ctr[0] = ctr;
          s_idx[0] = 0;
          s_cls[0] = cls;
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
      const int pos = 1 + off + __popc(m & ((1u << lane) - 1u));
      s_box[pos] = g;
      s_ctr[pos] = ctr;
      s_area[pos] = fmul(fsub(g.z, g.x), fsub(g.w, g.y));  // box_b_area (utils.py:84-85)
      s_idx[pos] = i;
      s_cls[pos] = cls;
    }
    n_kept += tot;
    __syncthreads();
  }

  // ---- log size ratios of every (staged box, anchor shape) pair, once per CTA
  const bool use_lut = n_gt > 0 && (n_kept + 1) * A <= kLutEntries;   // CTA-uniform
  if (use_lut) {
    for (int e = tid; e < (n_kept + 1) * A; e += blockDim.x) {
      const int j = e / A, a = e - j * A;
      const float4 t = s_ctr[j];
      s_lut[e] = log_size_ratios(t.z, t.w, L.anchor_h[a], L.anchor_w[a]);
    }
    __syncthreads();
  }

  // ---- this thread's anchor shapes (the same for every sub-patch)
  const int cell_local = tid / prm.tpc;
  const int sub = tid - cell_local * prm.tpc;
  const int cl_y = cell_local / kTileW, cl_x = cell_local % kTileW;
  const int a_first = min(sub * kApt, A - 1);   // (threads past the last anchor shape recompute it and store nothing)
  // cells of this warp within a sub-patch (warp-uniform): rows r_lo..r_hi, columns x_lo..x_hi
  const int c_lo = (warp * 32) / prm.tpc, c_hi = min((warp * 32 + 31) / prm.tpc, kTileH * kTileW - 1);
  const int r_lo = c_lo / kTileW, r_hi = c_hi / kTileW;
  const int x_lo = r_lo == r_hi ? c_lo % kTileW : 0, x_hi = r_lo == r_hi ? c_hi % kTileW : kTileW - 1;

  int n_ign = 0, n_fg = 0;   // both rare: the common warp issues no atomic at all
  if (n_gt == 0 && tid == 0) flags |= RN_FLAG_EMPTY_GT;
  double* const npos = prm.stats ? prm.stats + RN_STAT_NPOS(0) : nullptr;
  const bool keep_iou = prm.iou_max.p[0] != nullptr;

  // (warp-uniform loop bounds: sub-patches and warps that lie outside the grid have nothing to do; there is no CTA
  //  barrier below)
#pragma unroll 1
  for (int sy0 = ty0; sy0 < ty0 + prm.sy * kTileH && sy0 + r_lo < L.grid_h; sy0 += kTileH) {
#pragma unroll 1
    for (int sx0 = tx0; sx0 < tx0 + prm.sx * kTileW && sx0 + x_lo < L.grid_w; sx0 += kTileW) {
      const int cy_i = sy0 + cl_y, cx_i = sx0 + cl_x;
      const bool cell_ok = cy_i < L.grid_h && cx_i < L.grid_w;
      const float cy = cell_centre(L.y0, L.y_step, cy_i);
      const float cx = cell_centre(L.x0, L.x_step, cx_i);

      float a0[kApt], a1[kApt], a2[kApt], a3[kApt], aa[kApt], best[kApt];
      int bestj[kApt];
#pragma unroll
      for (int k = 0; k < kApt; ++k) {
        const float4 sh = s_shape[min(a_first + k, A - 1)];
        a0[k] = fsub(cy, sh.x);  // from_center_box(anchor), dataset.py:74
        a1[k] = fsub(cx, sh.y);
        a2[k] = fadd(cy, sh.x);
        a3[k] = fadd(cx, sh.y);
        aa[k] = fmul(fsub(a2[k], a0[k]), fsub(a3[k], a1[k]));  // box_a_area (utils.py:82-83)
        best[k] = 0.f;
        bestj[k] = 0;   // slot 0 = GT 0: tf.argmax of an all-zero column
      }

      // ---- IoU scan (utils.iou, utils.py:70-92), strict '>' keeps the first maximum.
      // Per warp, 32 staged boxes at a time are tested (one per lane) against the bounding box of the warp's cells
      // grown by the level's largest half sizes; only the boxes that touch it are visited, in their original order.
      // A skipped box has IoU == 0 with every anchor of the warp, which can neither win nor tie above 0, so the
      // arg-max and its first-maximum rule are unchanged.  (Cells outside the grid compute along and store nothing.)
      // The IEEE division only runs when the pair can beat the running best: fl(inter / uni) > best implies
      // inter > best * uni, and the two products are compared with a 2e-6 margin (their roundings are 6e-8 each); a
      // denormal-range intersection always takes the division.
      const float w0 = fsub(cell_centre(L.y0, L.y_step, sy0 + r_lo), L.max_half_h);
      const float w1 = fsub(cell_centre(L.x0, L.x_step, sx0 + x_lo), L.max_half_w);
      const float w2 = fadd(cell_centre(L.y0, L.y_step, min(sy0 + r_hi, L.grid_h - 1)), L.max_half_h);
      const float w3 = fadd(cell_centre(L.x0, L.x_step, min(sx0 + x_hi, L.grid_w - 1)), L.max_half_w);
      for (int j0 = 1; j0 <= n_kept; j0 += 32) {
        bool touches = false;
        if (j0 + lane <= n_kept) {
          const float4 g = s_box[j0 + lane];
          touches = !(g.z < w0 || g.x > w2 || g.w < w1 || g.y > w3);
        }
        unsigned todo = __ballot_sync(0xffffffffu, touches);
        while (todo) {
          const int j = j0 + __ffs(todo) - 1;
          todo &= todo - 1;
          const float4 g = s_box[j];
          const float ga = s_area[j];
#pragma unroll
          for (int k = 0; k < kApt; ++k) {
            // IoU <= min(area) / max(area): the intersection is at most the smaller box and the union at least the
            // larger (also for the rounded products, rounding being monotone), so a box whose bound cannot exceed
            // the running best is skipped before any of the pair's arithmetic (crowded scenes: -10 %)
            if (best[k] * fmaxf(aa[k], ga) > fminf(aa[k], ga) * 1.000002f) continue;
            const float yt = fmaxf(a0[k], g.x), xl = fmaxf(a1[k], g.y);
            const float yb = fminf(a2[k], g.z), xr = fminf(a3[k], g.w);
            if (!(yb < yt || xr < xl)) {
              const float inter = fmul(fsub(yb, yt), fsub(xr, xl));
              const float uni = fsub(fadd(aa[k], ga), inter);
              if (inter * 1.000002f > best[k] * uni || inter < 1e-30f) {
                const float v = fdiv(inter, uni);
                if (v > best[k]) {
                  best[k] = v;
                  bestj[k] = j;
                }
              }
            }
          }
        }
      }

      // ---- labels + encoding for the matched box (all anchors, dataset.py:105-121)
      if (cell_ok) {
        const int row0 = (cy_i * L.grid_w + cx_i) * A + sub * kApt;
        int32_t* const o_matched = reinterpret_cast<int32_t*>(s_out[0]) + row0;
        int32_t* const o_label = reinterpret_cast<int32_t*>(s_out[1]) + row0;
        float4* const o_deltas = reinterpret_cast<float4*>(s_out[2]) + row0;
        float* const o_iou = reinterpret_cast<float*>(s_out[3]) + row0;
#pragma unroll
        for (int k = 0; k < kApt; ++k) {
          const int a = sub * kApt + k;
          if (a < A) {
            int lab = RN_LABEL_BACKGROUND;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            int matched = -1;
            if (n_gt > 0) {
              const int j = bestj[k];
              matched = s_idx[j];
              const float4 t = s_ctr[j];   // (cy, cx, h, w) of the matched raw box
              const bool pos = best[k] >= prm.pos_thr;                 // !bg_mask, dataset.py:83
              const bool ign = !pos && !(best[k] < prm.neg_thr);       // !trainable_mask, dataset.py:87
              if (pos) {
                const int c = s_cls[j];
                lab = (c >= 0 && c < prm.num_classes) ? c : RN_LABEL_BACKGROUND;  // one_hot of an out-of-range id is 0
              } else if (ign) {
                lab = RN_LABEL_IGNORE;
              }
              const float4 sh = s_shape[a];
              d.x = fsub(t.x, cy) * sh.z;   // shifts, dataset.py:108
              d.y = fsub(t.y, cx) * sh.w;
              const float2 zw = use_lut ? s_lut[j * A + a]
                                        : log_size_ratios_direct(t.z, t.w, L.anchor_h[a], L.anchor_w[a]);
              d.z = zw.x;                    // log(scales), dataset.py:110-112
              d.w = zw.y;
            }
            o_matched[k] = matched;
            o_label[k] = lab;
            o_deltas[k] = d;
            if (keep_iou) o_iou[k] = best[k];
            n_ign += lab == RN_LABEL_IGNORE;
            n_fg += lab >= 0;
            if (npos && lab >= 0) atomicAdd(npos + lab, 1.0);
          }
        }
      }
    }
  }

  if (prm.stats) {
    // M = all anchors - ignored: the constant part is added once, ignored anchors are rare
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
      atomicAdd(prm.stats + RN_STAT_TRAINABLE, (double)pyr.anchors_per_image * (double)gridDim.x);
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
  return rn_assign_encode_patch(pyr, batch, d_gt_boxes, d_gt_classes, d_gt_counts, max_gt, num_classes, neg_iou_thr,
                                pos_iou_thr, matched, label, deltas, iou_max, d_stats, d_status, 0, 0, stream);
}

extern "C" int rn_assign_encode_patch(const rn_pyramid_t* pyr, int batch, const float* d_gt_boxes,
                                      const int32_t* d_gt_classes, const int32_t* d_gt_counts, int max_gt,
                                      int num_classes, float neg_iou_thr, float pos_iou_thr, const rn_maps_t* matched,
                                      const rn_maps_t* label, const rn_maps_t* deltas, const rn_maps_t* iou_max,
                                      double* d_stats, uint32_t* d_status, int patch_sy, int patch_sx, void* stream) {
  RN_REQUIRE(patch_sy >= 0 && patch_sy <= 64 && patch_sx >= 0 && patch_sx <= 64 && (patch_sy == 0) == (patch_sx == 0),
             RN_ERR_BAD_ARG, "rn_assign_encode_patch: patch_sy / patch_sx must both be 0 (automatic) or in 1..64");
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
  // CTA patch = sy x sx sub-patches: the per-CTA work (GT staging, the log-ratio table) is shared by all of them.
  // Small launches keep one sub-patch per CTA so that they still fill the device.
  long long tiles8 = 0;
  for (int l = 0; l < pyr->num_levels; ++l) tiles8 += (long long)pyr->level[l].tiles_x * pyr->level[l].tiles_y;
  // (measured: COCO batch 64 -- 24 k sub-patches -- 2 x 2: 0.164 ms, 2 x 1: 0.174, 1 x 1: 0.210; VOC batch 32 -- 3.9 k --
  //  2 x 1: 0.038 ms, 1 x 1: 0.042, 2 x 2: 0.046)
  const long long work = tiles8 * batch, sms = num_sms();
  prm.sy = patch_sy ? patch_sy : (work >= sms * 8 ? RN_ASSIGN_SUPER_Y : 1);
  prm.sx = patch_sx ? patch_sx : (work >= sms * 32 ? RN_ASSIGN_SUPER_X : 1);
  int tiles = 0;
  for (int l = 0; l < RN_MAX_LEVELS; ++l) {
    prm.tiles_x[l] = 1;
    prm.tile_start[l] = tiles;
    if (l < pyr->num_levels) {
      const rn_level_t& L = pyr->level[l];
      prm.tiles_x[l] = (L.grid_w + kTileW * prm.sx - 1) / (kTileW * prm.sx);
      tiles += prm.tiles_x[l] * ((L.grid_h + kTileH * prm.sy - 1) / (kTileH * prm.sy));
    }
  }
  const size_t smem = (size_t)(max_gt + 1) * (2 * sizeof(float4) + sizeof(float) + 2 * sizeof(int));
  RN_REQUIRE(tiles <= 65535, RN_ERR_BAD_SHAPE, "rn_assign_encode: more than 65535 patches per image");
  static size_t smem_opted = 44 * 1024;   // more than 48 KB of shared memory (static + dynamic) needs an opt-in (max_gt > ~1000)
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
  assign_encode_kernel<<<dim3((unsigned)batch, (unsigned)tiles), threads, smem, (cudaStream_t)stream>>>(*pyr, prm);
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
