// K4/K5: box decoding, candidate selection and class-wise NMS
// (utils.regression_postprocess utils.py:108-117, boxes_decode 183-195, merge_boxes_decoded
// 223-227, nms 213-220, nms_classwise 198-210).
//
//   select   : streams the [rows, C] scores once; a group of LPR lanes owns a row, keeps its
//              slice in registers, reduces (max, first argmax) with shuffles and appends the rare
//              rows with score > thr to the image's candidate list (one atomic per candidate).
//   finalize : one CTA per image sorts the (unordered) candidates by anchor index -> the
//              reference's row-major / level-major order, applies the optional per-level top-k,
//              decodes or gathers the boxes.
//   nms      : one CTA per image sorts by (class asc, score desc, index asc); one warp per
//              (image, class) runs the greedy pass -- 32 candidates at a time against the kept
//              boxes staged in shared memory, then resolves the 32 among themselves with
//              __ballot_sync masks; one CTA per image compacts class-major.
#include <cuda_bf16.h>
#include "rn_common.cuh"
#include "rn_tiles.cuh"

namespace rn {

constexpr int kSortSmemElems = 16384;  // elements a CTA sorts in shared memory (12 B each)
constexpr int kSortThreads = 1024;

__device__ __forceinline__ unsigned next_pow2(unsigned v) {
  unsigned p = 1;
  while (p < v) p <<= 1;
  return p;
}

// float -> unsigned whose ascending order is the float's ascending order
__device__ __forceinline__ unsigned sortable_bits(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// CTA-wide bitonic sort of (key, pos) pairs, ascending, lexicographic.  n_pad is a power of two;
// padding entries must hold the maximum key.  Works on shared or global memory.
__device__ void bitonic_sort_pairs(unsigned long long* key, unsigned* pos, unsigned n_pad) {
  for (unsigned k = 2; k <= n_pad; k <<= 1) {
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {
        const unsigned ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long ka = key[i], kb = key[ixj];
          const unsigned pa = pos[i], pb = pos[ixj];
          const bool gt = ka > kb || (ka == kb && pa > pb);
          const bool up = (i & k) == 0;
          if (gt == up) {
            key[i] = kb; key[ixj] = ka;
            pos[i] = pb; pos[ixj] = pa;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- utils.regression_postprocess for one anchor ------------------------------------------------
__device__ __forceinline__ float4 decode_one(const rn_level_t& L, int y, int x, int a, float4 d) {
  const float sy = fmul(d.x, L.anchor_h[a]);          // scale_regression, utils.py:100-105
  const float sx = fmul(d.y, L.anchor_w[a]);
  const float sh = fmul(expf(d.z), L.anchor_h[a]);    // exp on sizes, utils.py:110-111
  const float sw = fmul(expf(d.w), L.anchor_w[a]);
  const float cy = fadd(sy, cell_centre(L.y0, L.y_step, y));  // utils.py:24-36
  const float cx = fadd(sx, cell_centre(L.x0, L.x_step, x));
  const float hh = sh * 0.5f, hw = sw * 0.5f;         // utils.py:39-44
  return make_float4(fsub(cy, hh), fsub(cx, hw), fadd(cy, hh), fadd(cx, hw));
}

// One kernel for every elementwise box-map op of utils.py / dataset.py (selected by `op`).
__global__ void boxmap_op_kernel(const __grid_constant__ rn_pyramid_t pyr, int batch, int op,
                                 const __grid_constant__ Ptrs in, const __grid_constant__ Ptrs out) {
  const long long total = pyr.anchors_per_image * (long long)batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    for (int k = 1; k < pyr.num_levels; ++k) l += (i >= pyr.level[k].anchor_offset * batch) ? 1 : 0;
    const rn_level_t& L = pyr.level[l];
    const long long local = i - L.anchor_offset * batch;
    const int a = (int)(local % L.num_anchors);
    const long long cell = local / L.num_anchors;
    const int x = (int)(cell % L.grid_w);
    const int y = (int)((cell / L.grid_w) % L.grid_h);
    const float4 d = __ldcs(reinterpret_cast<const float4*>(in.p[l]) + local);
    float4 r;
    switch (op) {
      case RN_BOXOP_DECODE:  // utils.regression_postprocess, utils.py:108-117
        r = decode_one(L, y, x, a, d);
        break;
      case RN_BOXOP_ADD_CENTRE:  // utils.boxmap_anchor_relative_to_image_relative, utils.py:22-36
        r = make_float4(fadd(d.x, cell_centre(L.y0, L.y_step, y)), fadd(d.y, cell_centre(L.x0, L.x_step, x)), d.z, d.w);
        break;
      case RN_BOXOP_SCALE:  // utils.scale_regression, utils.py:100-105
        r = make_float4(fmul(d.x, L.anchor_h[a]), fmul(d.y, L.anchor_w[a]), fmul(d.z, L.anchor_h[a]),
                        fmul(d.w, L.anchor_w[a]));
        break;
      case RN_BOXOP_CENTRE_TO_CORNER: {  // utils.py:39-44, dataset.from_center_box dataset.py:35-39
        const float hh = d.z * 0.5f, hw = d.w * 0.5f;
        r = make_float4(fsub(d.x, hh), fsub(d.y, hw), fadd(d.x, hh), fadd(d.y, hw));
        break;
      }
      default: {  // RN_BOXOP_CORNER_TO_CENTRE: dataset.to_center_box, dataset.py:28-32
        const float sh = fsub(d.z, d.x), sw = fsub(d.w, d.y);
        r = make_float4(fadd(d.x, sh * 0.5f), fadd(d.y, sw * 0.5f), sh, sw);
        break;
      }
    }
    __stcs(reinterpret_cast<float4*>(out.p[l]) + local, r);
  }
}

// utils.classmap_decode (utils.py:171-179): fg = max_c classmap > thr, one thread per row.
__global__ void classmap_fg_kernel(long long rows, int C, const float* __restrict__ in, float thr,
                                   unsigned char* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, in[r * C + c]);
    out[r] = mx > thr;
  }
}

// ---- select -----------------------------------------------------------------------------------------
struct SelectParams {
  Ptrs scores;
  int num_classes;
  int units_per_row;     // C / VEC
  int images;
  int capacity;
  float thr;
  int is_logits;
  int32_t* ws_anchor;    // [B, capacity] unordered
  float* ws_score;
  int32_t* ws_class;
  int32_t* count;        // [B]
  uint32_t* status;
};

template <int LPR, int VEC, int UPL>
__global__ void __launch_bounds__(256) select_kernel(const __grid_constant__ Segs segs,
                                                     const __grid_constant__ SelectParams prm) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;                          // lane within the row group
  const long long group = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / LPR;
  const long long num_groups = (long long)gridDim.x * blockDim.x / LPR;
  const long long total_rows = segs.start[RN_MAX_LEVELS];
  const int C = prm.num_classes;
  const int upr = prm.units_per_row;
  // every lane of a warp runs the same number of iterations (shuffles below are full-warp)
  const long long iters = (total_rows + num_groups - 1) / num_groups;
  unsigned flags = 0;

  // bf16x8 rows are half the bytes of float32 rows, so one row group per iteration does not keep enough
  // bytes in flight: the next iteration's units are requested before this one's are used.
  constexpr int kRaw = VEC == 8 ? UPL : 1;
  uint4 cur[kRaw];
  auto fetch = [&](long long it, uint4 (&raw)[kRaw]) {
    const long long r = group + it * num_groups;
    const bool ok = r < total_rows;
    const uint4* base = nullptr;
    if (ok) {
      const int s = seg_of(segs, r);
      base = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(prm.scores.p[s]) +
                                            (r - segs.start[s]) * C);
    }
#pragma unroll
    for (int j = 0; j < kRaw; ++j) {
      const int u = sub + j * LPR;
      raw[j] = (ok && u < upr) ? __ldcs(base + u) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // -inf
    }
  };
  if constexpr (VEC == 8) {
    if (iters > 0) fetch(0, cur);
  }

  for (long long it = 0; it < iters; ++it) {
    const long long r = group + it * num_groups;
    const bool row_ok = r < total_rows;
    uint4 nxt[kRaw];
    if constexpr (VEC == 8) {
      if (it + 1 < iters) fetch(it + 1, nxt);
    }
    int s = 0;
    long long rl = 0;
    const float* base = nullptr;   // (a bfloat16 row when VEC == 8: C * 2 bytes)
    if (row_ok) {
      s = seg_of(segs, r);
      rl = r - segs.start[s];
      base = VEC == 8 ? reinterpret_cast<const float*>(reinterpret_cast<const __nv_bfloat16*>(prm.scores.p[s]) + rl * C)
                      : reinterpret_cast<const float*>(prm.scores.p[s]) + rl * C;
    }
    float v[UPL][VEC];
#pragma unroll
    for (int j = 0; j < UPL; ++j) {
      const int u = sub + j * LPR;
      if constexpr (VEC == 8) {   // bf16x8: one 128-bit access (already in flight), widened to float32
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h[k]);
          v[j][2 * k] = f.x;
          v[j][2 * k + 1] = f.y;
        }
      } else if (row_ok && u < upr) {
        if constexpr (VEC == 4) {
          const float4 t = ld_stream4(base + u * 4);
          v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
        } else {
          v[j][0] = __ldcs(base + u);
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[j][k] = -INFINITY;
      }
    }
    // (max, first argmax) over this lane's slice, then over the group
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < UPL; ++j) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int c = (sub + j * LPR) * VEC + k;
        if (v[j][k] > best) {  // ascending c within a lane: strict '>' keeps the first maximum
          best = v[j][k];
          besti = c;
        }
      }
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
      }
    }
    float score = best;
    int cls = besti;
    bool fire;
    if (prm.is_logits) {
      // The reference takes max/argmax of sigmoid(x) (utils.py:185-186,249).  sigmoid is
      // monotone up to rounding, so only logits within a rounding window of the max can tie
      // or win in probability space: evaluate sigmoid for exactly those.
      const float pmax = sigmoid_f32(best);
      fire = row_ok && pmax > prm.thr - 4e-7f;
      if (__any_sync(0xffffffffu, fire)) {
        const float pq = pmax * (1.0f - pmax);
        const float lo = pq > 0.f ? best - (2.5e-7f / pq + 1e-6f) : -INFINITY;
        float bp = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            if (fire && v[j][k] >= lo) {
              const float p = sigmoid_f32(v[j][k]);
              const int c = (sub + j * LPR) * VEC + k;
              if (p > bp) {
                bp = p;
                bi = c;
              }
            }
          }
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, bp, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > bp || (ob == bp && oi < bi)) {
            bp = ob;
            bi = oi;
          }
        }
        score = bp;
        cls = bi;
      }
      fire = fire && score > prm.thr;
    } else {
      fire = row_ok && score > prm.thr;  // strict, utils.py:187
    }
    if (fire && sub == 0) {
      const long long per_image = segs.rows[s] / prm.images;
      const int b = (int)(rl / per_image);
      const int anchor = (int)(segs.start[s] / prm.images + (rl - (long long)b * per_image));
      const int slot = reserve_slot(prm.count, b, prm.capacity);
      if (slot < prm.capacity) {
        const size_t o = (size_t)b * prm.capacity + slot;
        prm.ws_anchor[o] = anchor;
        prm.ws_score[o] = score;
        prm.ws_class[o] = cls;
      } else {
        flags |= RN_FLAG_CAND_OVERFLOW;
      }
    }
    if constexpr (VEC == 8) {
#pragma unroll
      for (int j = 0; j < kRaw; ++j) cur[j] = nxt[j];
    }
  }
  if (prm.status) {
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(prm.status, flags);
  }
}

// ---- finalize: order candidates, optional per-level top-k, boxes ---------------------------------
struct FinalizeParams {
  Ptrs regs;
  int regs_are_deltas;
  int capacity;
  int topk;
  int dedup;
  const int32_t* ws_anchor;
  const float* ws_score;
  const int32_t* ws_class;
  unsigned long long* g_key;  // [B, pad] global sort buffers (used when the list exceeds smem)
  unsigned* g_pos;
  unsigned pad;               // next_pow2(capacity)
  float4* out_boxes;
  float* out_scores;
  int32_t* out_classes;
  int32_t* out_anchor;
  int32_t* count;             // in: appended, out: final
};

// Count entries of key[0..n) that carry no drop mark (bits 62 / 63 clear; CTA-wide); result in every thread.
// (Both marks matter after the top-k pass: the duplicates dropped earlier sort in front of the top-k drops and must
//  not be counted back in.)
__device__ int count_kept(const unsigned long long* key, int n, int* s_cnt) {
  if (threadIdx.x == 0) *s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (key[i] >> 62) == 0;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(s_cnt, c);
  __syncthreads();
  const int r = *s_cnt;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kSortThreads)
cand_finalize_kernel(const __grid_constant__ rn_pyramid_t pyr, int batch, const __grid_constant__ FinalizeParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  const int n_in = min(max(prm.count[b], 0), prm.capacity);
  const size_t base = (size_t)b * prm.capacity;
  const unsigned n_pad = max(next_pow2((unsigned)n_in), 2u);
  unsigned long long* key;
  unsigned* pos;
  if (n_pad <= (unsigned)kSortSmemElems) {
    key = reinterpret_cast<unsigned long long*>(smem_raw);
    pos = reinterpret_cast<unsigned*>(key + n_pad);
  } else {
    key = prm.g_key + (size_t)b * prm.pad;
    pos = prm.g_pos + (size_t)b * prm.pad;
  }
  // 1) ascending anchor index == the reference's (level, y, x, a) order
  for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {
    key[i] = i < (unsigned)n_in ? (unsigned long long)(unsigned)prm.ws_anchor[base + i] : ~0ull;
    pos[i] = i;
  }
  __syncthreads();
  bitonic_sort_pairs(key, pos, n_pad);
  int n = n_in;

  if (prm.dedup) {
    // 1b) The fused loss+select pass appends one entry per firing float4 UNIT, so an anchor with
    // several classes above the threshold appears more than once: keep its best entry (max
    // probability, then lowest class == the first argmax of utils.py:185-186).
    // Every entry decides for itself whether another entry of its anchor beats it (reads only), and the marks are
    // written after a barrier -- a group is at most C/4 entries, so the scan is short.
    bool mine = false;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {   // same trip count for every thread
      const int i = c0 + (int)threadIdx.x;
      bool drop = false;
      if (i < n) {
        const unsigned a = (unsigned)key[i];
        const float si = prm.ws_score[base + pos[i]];
        const int ci = prm.ws_class[base + pos[i]];
        for (int t = i - 1; t >= 0 && (unsigned)key[t] == a && !drop; --t) {   // an earlier entry also wins full ties
          const float st = prm.ws_score[base + pos[t]];
          drop = st > si || (st == si && prm.ws_class[base + pos[t]] <= ci);
        }
        for (int t = i + 1; t < n && (unsigned)key[t] == a && !drop; ++t) {
          const float st = prm.ws_score[base + pos[t]];
          drop = st > si || (st == si && prm.ws_class[base + pos[t]] < ci);
        }
      }
      __syncthreads();
      if (drop) {
        key[i] |= 1ull << 62;   // dropped duplicate: sorts behind every kept entry
        mine = true;
      }
      __syncthreads();
    }
    if (__syncthreads_or(mine)) {
      bitonic_sort_pairs(key, pos, n_pad);
      n = count_kept(key, n, &s_cnt);
    }
  }

  if (prm.topk > 0) {
    // 2) per-level top-k (extension, not in the reference): within each level's contiguous range
    // rank by (score desc, anchor asc), drop everything past k, then restore anchor order.
    bool any = false;
    for (int l = 0; l < pyr.num_levels; ++l) {
      const unsigned lo_a = (unsigned)pyr.level[l].anchor_offset;
      const unsigned hi_a = l + 1 < pyr.num_levels ? (unsigned)pyr.level[l + 1].anchor_offset
                                                   : (unsigned)pyr.anchors_per_image;
      // range of this level in the sorted list (uniform across the CTA: every thread searches)
      int lo = 0, hi = n;
      while (lo < hi) { const int m = (lo + hi) >> 1; if ((unsigned)key[m] < lo_a) lo = m + 1; else hi = m; }
      const int start = lo;
      hi = n;
      while (lo < hi) { const int m = (lo + hi) >> 1; if ((unsigned)key[m] < hi_a) lo = m + 1; else hi = m; }
      const int end = lo;
      if (end - start <= prm.topk) continue;
      any = true;
      __syncthreads();
      // O(m^2 / threads) ranking; only runs for levels that actually exceed k
      for (int i = start + threadIdx.x; i < end; i += blockDim.x) {
        const float si = prm.ws_score[base + pos[i]];
        int rank = 0;
        for (int j = start; j < end; ++j) {
          const float sj = prm.ws_score[base + pos[j]];
          rank += (sj > si) || (sj == si && j < i);
        }
        if (rank >= prm.topk) key[i] |= 1ull << 63;
      }
      __syncthreads();
    }
    __syncthreads();
    if (any) {
      bitonic_sort_pairs(key, pos, n_pad);
      n = count_kept(key, n, &s_cnt);
    }
  }

  // 3) outputs
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned p = pos[i];
    const int anchor = prm.ws_anchor[base + p];
    prm.out_scores[base + i] = prm.ws_score[base + p];
    prm.out_classes[base + i] = prm.ws_class[base + p];
    prm.out_anchor[base + i] = anchor;
    int l = 0;
    for (int k = 1; k < pyr.num_levels; ++k) l += (anchor >= pyr.level[k].anchor_offset) ? 1 : 0;
    const rn_level_t& L = pyr.level[l];
    const int local = anchor - (int)L.anchor_offset;
    const long long rl = (long long)b * L.grid_h * L.grid_w * L.num_anchors + local;
    float4 d = __ldg(reinterpret_cast<const float4*>(prm.regs.p[l]) + rl);
    if (prm.regs_are_deltas) {
      const int a = local % L.num_anchors;
      const int cell = local / L.num_anchors;
      d = decode_one(L, cell / L.grid_w, cell % L.grid_w, a, d);
    }
    prm.out_boxes[base + i] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) prm.count[b] = n;
}

// ---- NMS ------------------------------------------------------------------------------------------
struct NmsParams {
  int batch, num_classes, capacity, out_capacity, max_out;
  float iou_thr;
  const float4* boxes;
  const float* scores;
  const int32_t* classes;
  const int32_t* count;
  unsigned long long* g_key;
  unsigned* g_pos;
  unsigned pad;
  unsigned* perm;       // [B, capacity] candidate index by sorted rank
  int32_t* seg;         // [B, C + 1] class segment starts in sorted order
  unsigned* kept;       // [B, capacity] sorted ranks kept, stored at seg start + j
  int32_t* kept_count;  // [B, C]
  float4* det_boxes;
  float* det_scores;
  int32_t* det_classes;
  int32_t* det_index;
  int32_t* det_count;
  uint32_t* status;
  int kept_cap;         // kept boxes a warp can stage in shared memory
};

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const __grid_constant__ NmsParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int n = min(max(prm.count[b], 0), prm.capacity);
  const size_t base = (size_t)b * prm.capacity;
  const unsigned n_pad = max(next_pow2((unsigned)n), 2u);
  unsigned long long* key;
  unsigned* pos;
  if (n_pad <= (unsigned)kSortSmemElems) {
    key = reinterpret_cast<unsigned long long*>(smem_raw);
    pos = reinterpret_cast<unsigned*>(key + n_pad);
  } else {
    key = prm.g_key + (size_t)b * prm.pad;
    pos = prm.g_pos + (size_t)b * prm.pad;
  }
  for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {
    if (i < (unsigned)n) {
      const unsigned c = (unsigned)prm.classes[base + i];
      const unsigned s = ~sortable_bits(prm.scores[base + i]);  // descending score
      key[i] = ((unsigned long long)c << 32) | s;
    } else {
      key[i] = ~0ull;
    }
    pos[i] = i;  // ties -> lower candidate index first (stable order of tf.image.non_max_suppression)
  }
  __syncthreads();
  bitonic_sort_pairs(key, pos, n_pad);
  for (int i = threadIdx.x; i < n; i += blockDim.x) prm.perm[base + i] = pos[i];
  // class c starts at the first sorted rank whose class >= c
  for (int c = threadIdx.x; c <= prm.num_classes; c += blockDim.x) {
    int lo = 0, hi = n;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if ((unsigned)(key[m] >> 32) < (unsigned)c) lo = m + 1; else hi = m;
    }
    prm.seg[(size_t)b * (prm.num_classes + 1) + c] = lo;
  }
}

struct NmsBox {
  float y0, x0, y1, x1, area;
};

// Box normalisation and area exactly as TF's CPU NonMaxSuppression kernel does it.
__device__ __forceinline__ NmsBox nms_box(float4 r) {
  NmsBox q;
  q.y0 = fminf(r.x, r.z); q.x0 = fminf(r.y, r.w);
  q.y1 = fmaxf(r.x, r.z); q.x1 = fmaxf(r.y, r.w);
  q.area = fmul(fsub(q.y1, q.y0), fsub(q.x1, q.x0));
  return q;
}

// iou(a, b) > thr with TF's arithmetic: area <= 0 => IoU 0, intersection = max(dy, 0) * max(dx, 0), IEEE division.
// Disjoint boxes (most pairs) have intersection 0, hence IoU 0, and 0 > thr is false for every valid threshold --
// no division.  For the others fl(inter / uni) > thr is decided from inter and thr * uni whenever they differ by more
// than 2e-6 relative (their roundings are 6e-8 each); only pairs inside that band, and denormal-range
// intersections, take the division.
__device__ __forceinline__ bool nms_suppresses(const NmsBox& a, const NmsBox& b, float thr) {
  if (a.area <= 0.f || b.area <= 0.f) return false;
  const float iy0 = fmaxf(a.y0, b.y0), ix0 = fmaxf(a.x0, b.x0);
  const float iy1 = fminf(a.y1, b.y1), ix1 = fminf(a.x1, b.x1);
  const float ih = fsub(iy1, iy0), iw = fsub(ix1, ix0);
  if (!(ih > 0.f && iw > 0.f)) return false;
  const float inter = fmul(ih, iw);
  const float uni = fsub(fadd(a.area, b.area), inter);
  const float rhs = thr * uni;
  if (inter >= 1e-30f) {
    if (inter > rhs * 1.000002f) return true;
    if (inter * 1.000002f < rhs) return false;
  }
  return fdiv(inter, uni) > thr;
}

constexpr int kNmsWarps = 4;

__global__ void __launch_bounds__(kNmsWarps * 32) nms_greedy_kernel(const __grid_constant__ NmsParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * kNmsWarps + warp;   // (image, class) pair
  if (w >= (long long)prm.batch * prm.num_classes) return;  // whole warp leaves together
  const int b = (int)(w / prm.num_classes);
  const int c = (int)(w % prm.num_classes);
  NmsBox* s_kept = reinterpret_cast<NmsBox*>(smem_raw) + (size_t)warp * prm.kept_cap;
  const int* seg = prm.seg + (size_t)b * (prm.num_classes + 1);
  const int start = seg[c], end = seg[c + 1];
  const size_t base = (size_t)b * prm.capacity;
  int n_kept = 0;
  for (int chunk = start; chunk < end && n_kept < prm.max_out; chunk += 32) {
    const int i = chunk + lane;
    const bool valid = i < end;
    NmsBox me = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) me = nms_box(__ldg(prm.boxes + base + prm.perm[base + i]));
    bool alive = valid;
    // against boxes kept by earlier chunks
    for (int k = 0; k < n_kept && __any_sync(0xffffffffu, alive); ++k) {
      if (alive && nms_suppresses(me, s_kept[k], prm.iou_thr)) alive = false;
    }
    // among the 32 of this chunk, in score order
    unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
    for (int t = 0; t < 32; ++t) {
      if (!((alive_mask >> t) & 1u)) continue;  // uniform: alive_mask is warp-wide
      if (n_kept >= prm.max_out) {
        alive_mask &= (1u << t) - 1u;           // output full: drop t and everything after it
        break;
      }
      NmsBox kb;
      kb.y0 = __shfl_sync(0xffffffffu, me.y0, t);
      kb.x0 = __shfl_sync(0xffffffffu, me.x0, t);
      kb.y1 = __shfl_sync(0xffffffffu, me.y1, t);
      kb.x1 = __shfl_sync(0xffffffffu, me.x1, t);
      kb.area = __shfl_sync(0xffffffffu, me.area, t);
      if (lane == 0) {
        s_kept[n_kept] = kb;
        prm.kept[base + start + n_kept] = (unsigned)(chunk + t);
      }
      ++n_kept;
      const bool dies = lane > t && ((alive_mask >> lane) & 1u) && nms_suppresses(me, kb, prm.iou_thr);
      alive_mask &= ~__ballot_sync(0xffffffffu, dies);
    }
    __syncwarp();
  }
  if (lane == 0) prm.kept_count[(size_t)b * prm.num_classes + c] = n_kept;
}

__global__ void __launch_bounds__(256) nms_compact_kernel(const __grid_constant__ NmsParams prm) {
  extern __shared__ int s_off[];  // [C + 1] exclusive prefix of the per-class kept counts
  const int b = blockIdx.x;
  const int C = prm.num_classes;
  const int* kc = prm.kept_count + (size_t)b * C;
  const int* seg = prm.seg + (size_t)b * (C + 1);
  const size_t base = (size_t)b * prm.capacity;
  const size_t obase = (size_t)b * prm.out_capacity;
  // prefix over classes: warp 0 scans C in chunks of 32
  if (threadIdx.x < 32) {
    int carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + (int)threadIdx.x;
      const int v = c < C ? kc[c] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += t;
      }
      if (c < C) s_off[c] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_off[C] = carry;
  }
  __syncthreads();
  const int total = s_off[C];
  const int n_out = min(total, prm.out_capacity);
  // one output slot per thread: class by binary search in the prefix (class-major order)
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
    int lo = 0, hi = C - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= o) lo = mid; else hi = mid - 1;
    }
    const int j = o - s_off[lo];
    const unsigned rank = prm.kept[base + seg[lo] + j];
    const unsigned ci = prm.perm[base + rank];
    prm.det_boxes[obase + o] = prm.boxes[base + ci];
    prm.det_scores[obase + o] = prm.scores[base + ci];
    prm.det_classes[obase + o] = prm.classes[base + ci];
    prm.det_index[obase + o] = (int)ci;
  }
  if (threadIdx.x == 0) {
    if (total > prm.out_capacity && prm.status) atomicOr(prm.status, RN_FLAG_DET_OVERFLOW);
    prm.det_count[b] = n_out;
  }
}

// ---- fused detection tail: finalize + class-wise NMS in ONE launch, one CTA per image -----------------------------
// The tail of a step is latency, not work: ~1-2 k candidates per image, one CTA per image.  The separate kernels
// (cand_finalize, nms_sort, nms_greedy, nms_compact) pay four launches and sort 12-byte (key, position) pairs with a
// barrier per bitonic stage.  This kernel keeps everything of one image in shared memory: 64-bit keys that carry
// their payload in the low bits (no position array), the 2..32 sub-sequences of the bitonic network in registers
// (warp shuffles, no barrier), order-preserving compaction for the duplicate anchors instead of a second sort, and
// the per-class greedy pass by the same CTA's warps.  Same ordering rules and the same arithmetic as the separate
// kernels, which stay for lists longer than kTailMaxCand and for the per-level top-k extension
// (tests/test_gpu_decode.py compares the two paths bit for bit).
constexpr int kTailThreads = 1024;
constexpr int kTailMaxCand = 8192;   // candidates per image (13 index bits in the keys)
constexpr int kTailIdxBits = 13;
constexpr int kTailKeptMin = 40;     // kept boxes per warp staged in shared memory (more: re-read from global);
constexpr int kTailKeptMax = 256;    // the launch takes as many as fit next to the sort buffers
constexpr int kTailRounds = kTailMaxCand / kTailThreads;

struct TailParams {
  Ptrs regs;
  int regs_are_deltas;
  int capacity;
  int dedup;
  const int32_t* ws_anchor;
  const float* ws_score;
  const int32_t* ws_class;
  float4* cand_boxes;
  float* cand_scores;
  int32_t* cand_classes;
  int32_t* cand_anchor;
  int32_t* cand_count;   // in: appended, out: final
  int num_classes, out_capacity, max_out;
  int kept_smem;         // kept boxes per warp staged in shared memory
  int topk;              // per-level top-k (extension, not in the reference); 0 = off
  float iou_thr;
  float4* det_boxes;
  float* det_scores;
  int32_t* det_classes;
  int32_t* det_index;
  int32_t* det_count;
  uint32_t* status;
};

// CTA-wide bitonic sort of distinct 64-bit keys in shared memory, ascending; n_pad is a power of two >= 32 (padding
// = ~0).  Strides below 32 run in registers with shuffles; strides >= 32 in shared memory, one pair per thread.
__device__ void bitonic_sort_u64(unsigned long long* key, unsigned n_pad) {
  const unsigned lane = threadIdx.x & 31;
  for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {   // whole warps: n_pad and blockDim are multiples of 32
    unsigned long long v = key[i];
#pragma unroll
    for (unsigned k = 2; k <= 32; k <<= 1) {
      const bool up = (i & k) == 0;
#pragma unroll
      for (unsigned j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
        const bool take_min = ((lane & j) == 0) == up;
        v = take_min ? (o < v ? o : v) : (o > v ? o : v);
      }
    }
    key[i] = v;
  }
  __syncthreads();
  for (unsigned k = 64; k <= n_pad; k <<= 1) {
    for (unsigned j = k >> 1; j >= 32; j >>= 1) {
      for (unsigned t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
        const unsigned i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // t with a zero inserted at bit log2(j)
        const unsigned long long a = key[i], b = key[i | j];
        if ((a > b) == ((i & k) == 0)) {
          key[i] = b;
          key[i | j] = a;
        }
      }
      __syncthreads();
    }
    for (unsigned i = threadIdx.x; i < n_pad; i += blockDim.x) {
      unsigned long long v = key[i];
      const bool up = (i & k) == 0;
#pragma unroll
      for (unsigned j = 16; j > 0; j >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
        const bool take_min = ((lane & j) == 0) == up;
        v = take_min ? (o < v ? o : v) : (o > v ? o : v);
      }
      key[i] = v;
    }
    __syncthreads();
  }
}

// Does entry i of the anchor-ordered keys survive the duplicate-anchor rule?  (The fused loss+select pass appends one
// entry per firing float4 unit: an entry is dropped when another entry of its anchor beats it -- higher probability,
// then lower class == the first argmax of utils.py:185-186; an earlier entry also wins full ties.)
__device__ __forceinline__ bool tail_keeps(const unsigned long long* key, int i, int n_in, size_t base, float si, int ci,
                                           const TailParams& prm) {
  const unsigned idx_mask = (1u << kTailIdxBits) - 1u;
  const int anchor = (int)(key[i] >> kTailIdxBits);
  bool keep = true;
  for (int t = i - 1; t >= 0 && (int)(key[t] >> kTailIdxBits) == anchor && keep; --t) {
    const float st = prm.ws_score[base + (key[t] & idx_mask)];
    keep = !(st > si || (st == si && prm.ws_class[base + (key[t] & idx_mask)] <= ci));
  }
  for (int t = i + 1; t < n_in && (int)(key[t] >> kTailIdxBits) == anchor && keep; ++t) {
    const float st = prm.ws_score[base + (key[t] & idx_mask)];
    keep = !(st > si || (st == si && prm.ws_class[base + (key[t] & idx_mask)] < ci));
  }
  return keep;
}

// 32 registers (two CTAs' worth per SM; a 40-byte spill outside the loops): the tail runs beside the next step's
// assignment kernel, and with half of an SM's registers per CTA that kernel's CTAs share the 64 SMs the tail sits on
// (assignment in the step 0.214 -> 0.203 ms, tail unchanged; profiles/r02_tail.md).
#ifndef RN_TAIL_MINB
#define RN_TAIL_MINB 2
#endif
__global__ void __launch_bounds__(kTailThreads, RN_TAIL_MINB)
detect_tail_kernel(const __grid_constant__ rn_pyramid_t pyr, const __grid_constant__ TailParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = prm.num_classes;
  const unsigned cap_pad = max(next_pow2((unsigned)min(prm.capacity, kTailMaxCand)), 32u);
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem_raw);                   // [cap_pad]
  unsigned long long* key2 = key + cap_pad;                                                     // [cap_pad] iff topk > 0
  NmsBox* s_kbox = reinterpret_cast<NmsBox*>(key2 + (prm.topk > 0 ? cap_pad : 0));              // [32][kept_smem]
  int* seg = reinterpret_cast<int*>(s_kbox + (kTailThreads / 32) * prm.kept_smem);              // [C + 1]
  int* koff = seg + (C + 1);                                                                    // [C + 1]
  unsigned short* kept = reinterpret_cast<unsigned short*>(koff + (C + 1));                     // [cap_pad]
  __shared__ int s_warp[kTailThreads / 32];
  __shared__ int s_total, s_next_class;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int n_in = min(max(prm.cand_count[b], 0), min(prm.capacity, kTailMaxCand));
  const size_t base = (size_t)b * prm.capacity;
  const unsigned idx_mask = (1u << kTailIdxBits) - 1u;
  if (tid == 0) s_next_class = 0;

  // ---- 1) ascending anchor index == the reference's (level, y, x, a) order
  const unsigned n_pad = max(next_pow2((unsigned)n_in), 32u);
  for (unsigned i = tid; i < n_pad; i += kTailThreads)
    key[i] = i < (unsigned)n_in ? ((unsigned long long)(unsigned)prm.ws_anchor[base + i] << kTailIdxBits) | i : ~0ull;
  __syncthreads();
  bitonic_sort_u64(key, n_pad);

  // ---- 1b) per-level top-k (extension, not in the reference): the survivors are marked first (`kept` is free until
  // the NMS phase); a level that still holds more than k entries ranks them by (score desc, anchor asc) with one more
  // sort and drops everything past k
  unsigned short* flag = kept;
  if (prm.topk > 0) {
    for (int i = tid; i < n_in; i += kTailThreads) {
      bool keep = true;
      if (prm.dedup) {
        const unsigned long long me = key[i];
        keep = tail_keeps(key, i, n_in, base, prm.ws_score[base + (me & idx_mask)], prm.ws_class[base + (me & idx_mask)],
                          prm);
      }
      flag[i] = keep;
    }
    __syncthreads();
    for (int l = 0; l < pyr.num_levels; ++l) {   // CTA-uniform control flow throughout
      const unsigned lo_a = (unsigned)pyr.level[l].anchor_offset;
      const unsigned hi_a = l + 1 < pyr.num_levels ? (unsigned)pyr.level[l + 1].anchor_offset
                                                   : (unsigned)pyr.anchors_per_image;
      int lo = 0, hi = n_in;
      while (lo < hi) { const int m = (lo + hi) >> 1; if ((unsigned)(key[m] >> kTailIdxBits) < lo_a) lo = m + 1; else hi = m; }
      const int start = lo;
      hi = n_in;
      while (lo < hi) { const int m = (lo + hi) >> 1; if ((unsigned)(key[m] >> kTailIdxBits) < hi_a) lo = m + 1; else hi = m; }
      const int end = lo;
      if (end - start <= prm.topk) continue;
      int alive = 0;
      for (int c0 = start; c0 < end; c0 += kTailThreads) {
        const int i = c0 + tid;
        alive += __syncthreads_count(i < end && flag[i]);
      }
      if (alive <= prm.topk) continue;
      const unsigned m_pad = max(next_pow2((unsigned)(end - start)), 32u);
      for (unsigned t = tid; t < m_pad; t += kTailThreads) {
        const int i = start + (int)t;
        key2[t] = (i < end && flag[i])
                      ? ((unsigned long long)(~sortable_bits(prm.ws_score[base + (key[i] & idx_mask)])) << kTailIdxBits) | t
                      : ~0ull;
      }
      __syncthreads();
      bitonic_sort_u64(key2, m_pad);
      for (unsigned t = prm.topk + tid; t < (unsigned)alive; t += kTailThreads) flag[start + (int)(key2[t] & idx_mask)] = 0;
      __syncthreads();
    }
  }

  // ---- 2) duplicate anchors dropped (tail_keeps) by an order-preserving compaction; 3) ordered outputs + the NMS
  // keys (class asc, score desc, candidate index asc)
  unsigned long long k2[kTailRounds];
  int dst[kTailRounds];
  int done = 0;   // kept entries of earlier rounds (CTA-uniform)
#pragma unroll
  for (int r = 0; r < kTailRounds; ++r) {
    dst[r] = -1;
    if ((unsigned)(r * kTailThreads) < n_pad) {   // CTA-uniform
      const int i = r * kTailThreads + tid;
      bool keep = i < n_in;
      float si = 0.f;
      int ci = 0, anchor = 0;
      if (keep) {
        const unsigned long long me = key[i];
        anchor = (int)(me >> kTailIdxBits);
        si = prm.ws_score[base + (me & idx_mask)];
        ci = prm.ws_class[base + (me & idx_mask)];
        if (prm.topk > 0) keep = flag[i] != 0;
        else if (prm.dedup) keep = tail_keeps(key, i, n_in, base, si, ci, prm);
      }
      // order-preserving compaction: position = kept entries before this one
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int before = done, total = 0;
#pragma unroll 8
      for (int w = 0; w < kTailThreads / 32; ++w) {
        const int c = s_warp[w];
        before += w < warp ? c : 0;
        total += c;
      }
      __syncthreads();
      if (keep) {
        const int o = before + __popc(m & ((1u << lane) - 1u));
        dst[r] = o;
        int l = 0;
        for (int q = 1; q < pyr.num_levels; ++q) l += (anchor >= pyr.level[q].anchor_offset) ? 1 : 0;
        const rn_level_t& L = pyr.level[l];
        const int local = anchor - (int)L.anchor_offset;
        const long long rl = (long long)b * L.grid_h * L.grid_w * L.num_anchors + local;
        float4 d = __ldg(reinterpret_cast<const float4*>(prm.regs.p[l]) + rl);
        if (prm.regs_are_deltas) {
          const int a = local % L.num_anchors;
          const int cell = local / L.num_anchors;
          d = decode_one(L, cell / L.grid_w, cell % L.grid_w, a, d);
        }
        prm.cand_boxes[base + o] = d;
        prm.cand_scores[base + o] = si;
        prm.cand_classes[base + o] = ci;
        prm.cand_anchor[base + o] = anchor;
        k2[r] = ((unsigned long long)(unsigned)ci << (32 + kTailIdxBits)) |
                ((unsigned long long)(~sortable_bits(si)) << kTailIdxBits) | (unsigned)o;
      }
      done += total;
    }
  }
  const int n = done;
  const unsigned n_pad2 = max(next_pow2((unsigned)n), 32u);   // <= n_pad
  // (every read of the anchor-ordered keys happened before the last barrier of the loop above)
#pragma unroll
  for (int r = 0; r < kTailRounds; ++r)
    if (dst[r] >= 0) key[dst[r]] = k2[r];
  for (unsigned i = n + tid; i < n_pad2; i += kTailThreads) key[i] = ~0ull;
  if (tid == 0) prm.cand_count[b] = n;
  __syncthreads();

  // ---- 4) class-wise NMS (utils.nms_classwise, utils.py:198-210): sort, class segments, greedy pass per class
  bitonic_sort_u64(key, n_pad2);
  for (int c = tid; c <= C; c += kTailThreads) {   // class c starts at the first sorted rank whose class >= c
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((unsigned)(key[mid] >> (32 + kTailIdxBits)) < (unsigned)c) lo = mid + 1; else hi = mid;
    }
    seg[c] = lo;
  }
  __syncthreads();   // (also orders the candidate boxes written above before the reads below)
  NmsBox* my_kbox = s_kbox + warp * prm.kept_smem;
  for (;;) {   // classes are handed out one at a time: their lists differ in length
    int c = 0;
    if (lane == 0) c = atomicAdd(&s_next_class, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= C) break;
    const int start = seg[c], end = seg[c + 1];
    int n_kept = 0;
    for (int chunk = start; chunk < end && n_kept < prm.max_out; chunk += 32) {
      const int i = chunk + lane;
      const bool valid = i < end;
      NmsBox me = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) me = nms_box(prm.cand_boxes[base + (key[i] & idx_mask)]);
      bool alive = valid;
      // against boxes kept by earlier chunks
      for (int q = 0; q < n_kept && __any_sync(0xffffffffu, alive); ++q) {
        const NmsBox kb = q < prm.kept_smem ? my_kbox[q]
                                            : nms_box(prm.cand_boxes[base + (key[kept[start + q]] & idx_mask)]);
        if (alive && nms_suppresses(me, kb, prm.iou_thr)) alive = false;
      }
      // among the 32 of this chunk, in score order
      unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
      for (int t = 0; t < 32; ++t) {
        if (!((alive_mask >> t) & 1u)) continue;  // uniform: alive_mask is warp-wide
        if (n_kept >= prm.max_out) {
          alive_mask &= (1u << t) - 1u;           // output full: drop t and everything after it
          break;
        }
        NmsBox kb;
        kb.y0 = __shfl_sync(0xffffffffu, me.y0, t);
        kb.x0 = __shfl_sync(0xffffffffu, me.x0, t);
        kb.y1 = __shfl_sync(0xffffffffu, me.y1, t);
        kb.x1 = __shfl_sync(0xffffffffu, me.x1, t);
        kb.area = __shfl_sync(0xffffffffu, me.area, t);
        if (lane == 0) {
          if (n_kept < prm.kept_smem) my_kbox[n_kept] = kb;
          kept[start + n_kept] = (unsigned short)(chunk + t);
        }
        ++n_kept;
        const bool dies = lane > t && ((alive_mask >> lane) & 1u) && nms_suppresses(me, kb, prm.iou_thr);
        alive_mask &= ~__ballot_sync(0xffffffffu, dies);
      }
      __syncwarp();
    }
    if (lane == 0) koff[c] = n_kept;
  }
  __syncthreads();

  // ---- 5) class-major compaction of the kept candidates
  if (warp == 0) {   // exclusive prefix of the per-class counts
    int carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      const int v = c < C ? koff[c] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      __syncwarp();
      if (c < C) koff[c] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      koff[C] = carry;
      s_total = carry;
    }
  }
  __syncthreads();
  const int total = s_total;
  const int n_out = min(total, prm.out_capacity);
  const size_t obase = (size_t)b * prm.out_capacity;
  for (int o = tid; o < n_out; o += kTailThreads) {   // class by binary search in the prefix
    int lo = 0, hi = C - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (koff[mid] <= o) lo = mid; else hi = mid - 1;
    }
    const unsigned ci = (unsigned)(key[kept[seg[lo] + (o - koff[lo])]] & idx_mask);
    prm.det_boxes[obase + o] = prm.cand_boxes[base + ci];
    prm.det_scores[obase + o] = prm.cand_scores[base + ci];
    prm.det_classes[obase + o] = prm.cand_classes[base + ci];
    prm.det_index[obase + o] = (int)ci;
  }
  if (tid == 0) {
    if (total > prm.out_capacity && prm.status) atomicOr(prm.status, RN_FLAG_DET_OVERFLOW);
    prm.det_count[b] = n_out;
  }
}

static unsigned host_next_pow2(unsigned v) {
  unsigned p = 2;
  while (p < v) p <<= 1;
  return p;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_smem_bytes(int capacity) {
  unsigned pad = host_next_pow2((unsigned)capacity);
  if (pad > (unsigned)kSortSmemElems) pad = kSortSmemElems;
  return (size_t)pad * 12;
}

static bool needs_global_sort(int capacity) { return host_next_pow2((unsigned)capacity) > (unsigned)kSortSmemElems; }

template <int LPR, int VEC>
static void launch_select(int upl, unsigned grid, cudaStream_t st, const Segs& segs, const SelectParams& prm) {
  switch (upl) {
    case 1: select_kernel<LPR, VEC, 1><<<grid, 256, 0, st>>>(segs, prm); break;
    case 2: select_kernel<LPR, VEC, 2><<<grid, 256, 0, st>>>(segs, prm); break;
    case 3: select_kernel<LPR, VEC, 3><<<grid, 256, 0, st>>>(segs, prm); break;
    case 4: select_kernel<LPR, VEC, 4><<<grid, 256, 0, st>>>(segs, prm); break;
    case 5: select_kernel<LPR, VEC, 5><<<grid, 256, 0, st>>>(segs, prm); break;
    case 6: select_kernel<LPR, VEC, 6><<<grid, 256, 0, st>>>(segs, prm); break;
    case 7: select_kernel<LPR, VEC, 7><<<grid, 256, 0, st>>>(segs, prm); break;
    default: select_kernel<LPR, VEC, 8><<<grid, 256, 0, st>>>(segs, prm); break;
  }
}

template <int VEC>
static void launch_select_lpr(int lpr, int upl, unsigned grid, cudaStream_t st, const Segs& segs,
                              const SelectParams& prm) {
  switch (lpr) {
    case 1: launch_select<1, VEC>(upl, grid, st, segs, prm); break;
    case 2: launch_select<2, VEC>(upl, grid, st, segs, prm); break;
    case 4: launch_select<4, VEC>(upl, grid, st, segs, prm); break;
    case 8: launch_select<8, VEC>(upl, grid, st, segs, prm); break;
    case 16: launch_select<16, VEC>(upl, grid, st, segs, prm); break;
    default: launch_select<32, VEC>(upl, grid, st, segs, prm); break;
  }
}

}  // namespace rn

using namespace rn;

extern "C" int rn_boxmap_op(const rn_pyramid_t* pyr, int batch, int op, const rn_maps_t* in, const rn_maps_t* out,
                            void* stream) {
  RN_REQUIRE(pyr && maps_present(in) && maps_present(out), RN_ERR_BAD_ARG, "rn_boxmap_op: NULL input");
  RN_REQUIRE(batch >= 0, RN_ERR_BAD_ARG, "rn_boxmap_op: negative batch");
  RN_REQUIRE(op >= RN_BOXOP_DECODE && op <= RN_BOXOP_CORNER_TO_CENTRE, RN_ERR_BAD_ARG, "rn_boxmap_op: unknown op %d", op);
  for (int l = 0; l < pyr->num_levels; ++l)
    RN_REQUIRE(in->p[l] && out->p[l] && ((((uintptr_t)in->p[l]) | ((uintptr_t)out->p[l])) & 15) == 0, RN_ERR_BAD_ARG,
               "rn_boxmap_op: level %d map missing or not 16-byte aligned", l);
  const long long total = pyr->anchors_per_image * (long long)batch;
  if (total == 0) return RN_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  boxmap_op_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*pyr, batch, op, make_ptrs(in), make_ptrs(out));
  RN_CUDA_CHECK_LAUNCH("rn_boxmap_op");
  return RN_OK;
}

extern "C" int rn_decode_boxes(const rn_pyramid_t* pyr, int batch, const rn_maps_t* deltas, const rn_maps_t* boxes,
                               void* stream) {
  return rn_boxmap_op(pyr, batch, RN_BOXOP_DECODE, deltas, boxes, stream);
}

extern "C" int rn_classmap_fg(int64_t rows, int num_classes, const float* d_classmap, float thr, uint8_t* d_fg_mask,
                              void* stream) {
  RN_REQUIRE(rows >= 0 && num_classes >= 1, RN_ERR_BAD_ARG, "rn_classmap_fg: bad shape");
  if (rows == 0) return RN_OK;
  RN_REQUIRE(d_classmap && d_fg_mask, RN_ERR_BAD_ARG, "rn_classmap_fg: NULL pointer");
  long long blocks = (rows + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  classmap_fg_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, num_classes, d_classmap, thr, d_fg_mask);
  RN_CUDA_CHECK_LAUNCH("rn_classmap_fg");
  return RN_OK;
}

extern "C" size_t rn_select_workspace_bytes(int batch, int capacity) {
  if (batch < 0 || capacity < 0) return 0;
  const size_t n = (size_t)batch * (size_t)capacity;
  size_t bytes = 3 * align256(n * 4);
  if (needs_global_sort(capacity)) bytes += align256((size_t)batch * host_next_pow2((unsigned)capacity) * 12);
  return bytes + 256;
}

extern "C" int rn_decode_select(const rn_pyramid_t* pyr, int batch, int num_classes, const rn_maps_t* scores,
                                int score_format, const rn_maps_t* regs, int regs_are_deltas, float score_thr,
                                int topk_per_level, int capacity, float* d_cand_boxes, float* d_cand_scores,
                                int32_t* d_cand_classes, int32_t* d_cand_anchor, int32_t* d_cand_count,
                                uint32_t* d_status, void* d_workspace, size_t workspace_bytes, void* stream) {
  RN_REQUIRE(pyr && maps_present(scores) && maps_present(regs), RN_ERR_BAD_ARG, "rn_decode_select: NULL input");
  RN_REQUIRE(d_cand_boxes && d_cand_scores && d_cand_classes && d_cand_anchor && d_cand_count && d_workspace,
             RN_ERR_BAD_ARG, "rn_decode_select: NULL output");
  RN_REQUIRE(batch >= 0 && num_classes >= 1 && capacity >= 1, RN_ERR_BAD_ARG, "rn_decode_select: bad sizes");
  RN_REQUIRE(workspace_bytes >= rn_select_workspace_bytes(batch, capacity), RN_ERR_CAPACITY,
             "rn_decode_select: workspace too small");
  RN_REQUIRE(((uintptr_t)d_cand_boxes & 15) == 0, RN_ERR_BAD_ARG, "d_cand_boxes must be 16-byte aligned");
  if (batch == 0) return RN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rn_rows_t rows;
  rn_rows_from_pyramid(pyr, batch, &rows);
  const Segs segs = make_segs(&rows);
  RN_REQUIRE(score_format >= 0 && score_format <= 3, RN_ERR_BAD_ARG, "rn_decode_select: unknown score_format");
  const bool bf16 = (score_format & RN_SCORES_BF16) != 0;
  const int scores_are_logits = score_format & RN_SCORES_LOGITS;
  bool vec4 = num_classes % 4 == 0;
  for (int l = 0; l < pyr->num_levels; ++l) {
    RN_REQUIRE(scores->p[l] && regs->p[l], RN_ERR_BAD_ARG, "rn_decode_select: missing map for level %d", l);
    RN_REQUIRE(((uintptr_t)regs->p[l] & 15) == 0, RN_ERR_BAD_ARG, "regs maps must be 16-byte aligned");
    if ((uintptr_t)scores->p[l] & 15) vec4 = false;
  }
  RN_REQUIRE(!bf16 || (num_classes % 8 == 0 && vec4), RN_ERR_UNSUPPORTED,
             "rn_decode_select: bfloat16 scores need num_classes %% 8 == 0 and 16-byte aligned maps");
  const size_t n = (size_t)batch * capacity;
  char* ws = (char*)d_workspace;
  SelectParams sp;
  sp.scores = make_ptrs(scores);
  sp.num_classes = num_classes;
  sp.units_per_row = bf16 ? num_classes / 8 : (vec4 ? num_classes / 4 : num_classes);
  sp.images = batch;
  sp.capacity = capacity;
  sp.thr = score_thr;
  sp.is_logits = scores_are_logits;
  sp.ws_anchor = (int32_t*)ws; ws += align256(n * 4);
  sp.ws_score = (float*)ws; ws += align256(n * 4);
  sp.ws_class = (int32_t*)ws; ws += align256(n * 4);
  sp.count = d_cand_count;
  sp.status = d_status;
  const int max_upl = 8;   // 16-byte loads per lane and row (bf16x8 units: up to 64 classes per lane)
  int lpr = 1;
  while (lpr < 32 && (sp.units_per_row + lpr - 1) / lpr > max_upl) lpr <<= 1;
  const int upl = (sp.units_per_row + lpr - 1) / lpr;
  RN_REQUIRE(upl <= max_upl, RN_ERR_UNSUPPORTED, "rn_decode_select: more than %d classes", vec4 ? 1024 : 256);
  cudaMemsetAsync(d_cand_count, 0, sizeof(int32_t) * batch, st);
  const long long total_rows = segs.start[RN_MAX_LEVELS];
  long long blocks = (total_rows * lpr + 255) / 256;
  const long long capb = (long long)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  if (blocks < 1) blocks = 1;
  if (bf16) launch_select_lpr<8>(lpr, upl, (unsigned)blocks, st, segs, sp);
  else if (vec4) launch_select_lpr<4>(lpr, upl, (unsigned)blocks, st, segs, sp);
  else launch_select_lpr<1>(lpr, upl, (unsigned)blocks, st, segs, sp);
  RN_CUDA_CHECK_LAUNCH("rn_decode_select(select)");
  if (topk_per_level < 0) return RN_OK;   // select only: the candidates stay unordered in the workspace

  return rn_select_finalize(pyr, batch, regs, regs_are_deltas, topk_per_level, capacity, 0, d_cand_boxes,
                            d_cand_scores, d_cand_classes, d_cand_anchor, d_cand_count, d_workspace, workspace_bytes,
                            stream);
}

extern "C" int rn_select_finalize(const rn_pyramid_t* pyr, int batch, const rn_maps_t* regs, int regs_are_deltas,
                                  int topk_per_level, int capacity, int dedup, float* d_cand_boxes,
                                  float* d_cand_scores, int32_t* d_cand_classes, int32_t* d_cand_anchor,
                                  int32_t* d_cand_count, void* d_workspace, size_t workspace_bytes, void* stream) {
  RN_REQUIRE(pyr && maps_present(regs), RN_ERR_BAD_ARG, "rn_select_finalize: NULL input");
  RN_REQUIRE(d_cand_boxes && d_cand_scores && d_cand_classes && d_cand_anchor && d_cand_count && d_workspace,
             RN_ERR_BAD_ARG, "rn_select_finalize: NULL output");
  RN_REQUIRE(batch >= 0 && capacity >= 1, RN_ERR_BAD_ARG, "rn_select_finalize: bad sizes");
  RN_REQUIRE(workspace_bytes >= rn_select_workspace_bytes(batch, capacity), RN_ERR_CAPACITY,
             "rn_select_finalize: workspace too small");
  RN_REQUIRE(((uintptr_t)d_cand_boxes & 15) == 0, RN_ERR_BAD_ARG, "d_cand_boxes must be 16-byte aligned");
  if (batch == 0) return RN_OK;
  for (int l = 0; l < pyr->num_levels; ++l)
    RN_REQUIRE(regs->p[l] && ((uintptr_t)regs->p[l] & 15) == 0, RN_ERR_BAD_ARG,
               "regs map of level %d missing or not 16-byte aligned", l);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)batch * capacity;
  char* ws = (char*)d_workspace;
  FinalizeParams fp;
  fp.regs = make_ptrs(regs);
  fp.regs_are_deltas = regs_are_deltas;
  fp.capacity = capacity;
  fp.topk = topk_per_level;
  fp.dedup = dedup;
  fp.ws_anchor = (const int32_t*)ws; ws += align256(n * 4);
  fp.ws_score = (const float*)ws; ws += align256(n * 4);
  fp.ws_class = (const int32_t*)ws; ws += align256(n * 4);
  fp.pad = host_next_pow2((unsigned)capacity);
  fp.g_key = nullptr;
  fp.g_pos = nullptr;
  if (needs_global_sort(capacity)) {
    fp.g_key = (unsigned long long*)ws;
    fp.g_pos = (unsigned*)(ws + (size_t)batch * fp.pad * 8);
  }
  fp.out_boxes = (float4*)d_cand_boxes;
  fp.out_scores = d_cand_scores;
  fp.out_classes = d_cand_classes;
  fp.out_anchor = d_cand_anchor;
  fp.count = d_cand_count;
  const size_t smem = sort_smem_bytes(capacity);
  static bool finalize_opted = false;   // once per process, not per call
  if (!finalize_opted) {
    cudaFuncSetAttribute(cand_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemElems * 12);
    finalize_opted = true;
  }
  cand_finalize_kernel<<<batch, kSortThreads, smem, st>>>(*pyr, batch, fp);
  RN_CUDA_CHECK_LAUNCH("rn_select_finalize");
  return RN_OK;
}

extern "C" size_t rn_nms_workspace_bytes(int batch, int num_classes, int capacity) {
  if (batch < 0 || capacity < 0 || num_classes < 0) return 0;
  const size_t n = (size_t)batch * (size_t)capacity;
  size_t bytes = 2 * align256(n * 4) + align256((size_t)batch * (num_classes + 1) * 4) +
                 align256((size_t)batch * num_classes * 4);
  if (needs_global_sort(capacity)) bytes += align256((size_t)batch * host_next_pow2((unsigned)capacity) * 12);
  return bytes + 256;
}

extern "C" int rn_nms_classwise(int batch, int num_classes, int capacity, const float* d_cand_boxes,
                                const float* d_cand_scores, const int32_t* d_cand_classes,
                                const int32_t* d_cand_count, float iou_thr, int max_output_per_class,
                                int out_capacity, float* d_det_boxes, float* d_det_scores, int32_t* d_det_classes,
                                int32_t* d_det_index, int32_t* d_det_count, uint32_t* d_status, void* d_workspace,
                                size_t workspace_bytes, void* stream) {
  RN_REQUIRE(d_cand_boxes && d_cand_scores && d_cand_classes && d_cand_count, RN_ERR_BAD_ARG,
             "rn_nms_classwise: NULL input");
  RN_REQUIRE(d_det_boxes && d_det_scores && d_det_classes && d_det_index && d_det_count && d_workspace,
             RN_ERR_BAD_ARG, "rn_nms_classwise: NULL output");
  RN_REQUIRE(batch >= 0 && num_classes >= 1 && capacity >= 1 && out_capacity >= 1 && max_output_per_class >= 0,
             RN_ERR_BAD_ARG, "rn_nms_classwise: bad sizes");
  RN_REQUIRE(iou_thr >= 0.f && iou_thr <= 1.f, RN_ERR_BAD_ARG, "iou_threshold must be in [0, 1]");
  RN_REQUIRE(workspace_bytes >= rn_nms_workspace_bytes(batch, num_classes, capacity), RN_ERR_CAPACITY,
             "rn_nms_classwise: workspace too small");
  RN_REQUIRE((((uintptr_t)d_cand_boxes | (uintptr_t)d_det_boxes) & 15) == 0, RN_ERR_BAD_ARG,
             "box arrays must be 16-byte aligned");
  if (batch == 0) return RN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)batch * capacity;
  char* ws = (char*)d_workspace;
  NmsParams p;
  p.batch = batch;
  p.num_classes = num_classes;
  p.capacity = capacity;
  p.out_capacity = out_capacity;
  p.max_out = max_output_per_class;
  p.iou_thr = iou_thr;
  p.boxes = (const float4*)d_cand_boxes;
  p.scores = d_cand_scores;
  p.classes = d_cand_classes;
  p.count = d_cand_count;
  p.perm = (unsigned*)ws; ws += align256(n * 4);
  p.kept = (unsigned*)ws; ws += align256(n * 4);
  p.seg = (int32_t*)ws; ws += align256((size_t)batch * (num_classes + 1) * 4);
  p.kept_count = (int32_t*)ws; ws += align256((size_t)batch * num_classes * 4);
  p.pad = host_next_pow2((unsigned)capacity);
  p.g_key = nullptr;
  p.g_pos = nullptr;
  if (needs_global_sort(capacity)) {
    p.g_key = (unsigned long long*)ws;
    p.g_pos = (unsigned*)(ws + (size_t)batch * p.pad * 8);
  }
  p.det_boxes = (float4*)d_det_boxes;
  p.det_scores = d_det_scores;
  p.det_classes = d_det_classes;
  p.det_index = d_det_index;
  p.det_count = d_det_count;
  p.status = d_status;
  p.kept_cap = max_output_per_class < capacity ? max_output_per_class : capacity;
  if (p.kept_cap < 1) p.kept_cap = 1;
  const size_t greedy_smem = (size_t)kNmsWarps * p.kept_cap * sizeof(NmsBox);
  RN_REQUIRE(greedy_smem <= 200 * 1024, RN_ERR_UNSUPPORTED,
             "rn_nms_classwise: max_output_per_class too large for shared memory (limit %d)",
             (int)(200 * 1024 / (kNmsWarps * sizeof(NmsBox))));
  static bool nms_opted = false;   // once per process, not per call
  if (!nms_opted) {
    cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemElems * 12);
    cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    nms_opted = true;
  }
  nms_sort_kernel<<<batch, kSortThreads, sort_smem_bytes(capacity), st>>>(p);
  RN_CUDA_CHECK_LAUNCH("rn_nms_classwise(sort)");
  const long long pairs = (long long)batch * num_classes;
  nms_greedy_kernel<<<(unsigned)((pairs + kNmsWarps - 1) / kNmsWarps), kNmsWarps * 32, greedy_smem, st>>>(p);
  RN_CUDA_CHECK_LAUNCH("rn_nms_classwise(greedy)");
  nms_compact_kernel<<<batch, 256, (size_t)(num_classes + 1) * sizeof(int), st>>>(p);
  RN_CUDA_CHECK_LAUNCH("rn_nms_classwise(compact)");
  return RN_OK;
}

static size_t tail_smem_bytes(int num_classes, int capacity, int kept_smem, bool topk) {
  unsigned pad = host_next_pow2((unsigned)(capacity < kTailMaxCand ? capacity : kTailMaxCand));
  if (pad < 32) pad = 32;
  return (size_t)pad * 8 * (topk ? 2 : 1) + (size_t)(kTailThreads / 32) * kept_smem * sizeof(NmsBox) +
         2 * (size_t)(num_classes + 1) * sizeof(int) + (size_t)pad * 2;
}

constexpr size_t kTailSmemLimit = 200 * 1024;

// kept boxes per warp that fit in shared memory beside the sort buffers (no more than a class can keep)
static int tail_kept_smem(int num_classes, int capacity, int max_out, bool topk) {
  int k = kTailKeptMin;
  const size_t fixed = tail_smem_bytes(num_classes, capacity, 0, topk);
  if (fixed < kTailSmemLimit) {
    const size_t room = (kTailSmemLimit - fixed) / ((kTailThreads / 32) * sizeof(NmsBox));
    if ((size_t)k < room) k = (int)(room < (size_t)kTailKeptMax ? room : (size_t)kTailKeptMax);
  }
  const int most = max_out < capacity ? max_out : capacity;
  if (k > most) k = most > kTailKeptMin ? most : kTailKeptMin;
  return k;
}

extern "C" int rn_detect_tail_supported(int num_classes, int capacity) {
  return num_classes >= 1 && num_classes < (1 << 19) && capacity >= 1 && capacity <= kTailMaxCand &&
         tail_smem_bytes(num_classes, capacity, kTailKeptMin, true) <= kTailSmemLimit;
}

extern "C" int rn_detect_tail(const rn_pyramid_t* pyr, int batch, int num_classes, const rn_maps_t* regs,
                              int regs_are_deltas, int capacity, int dedup, int topk_per_level, float* d_cand_boxes,
                              float* d_cand_scores,
                              int32_t* d_cand_classes, int32_t* d_cand_anchor, int32_t* d_cand_count, float iou_thr,
                              int max_output_per_class, int out_capacity, float* d_det_boxes, float* d_det_scores,
                              int32_t* d_det_classes, int32_t* d_det_index, int32_t* d_det_count, uint32_t* d_status,
                              void* d_select_workspace, size_t select_workspace_bytes, void* stream) {
  RN_REQUIRE(pyr && maps_present(regs), RN_ERR_BAD_ARG, "rn_detect_tail: NULL input");
  RN_REQUIRE(d_cand_boxes && d_cand_scores && d_cand_classes && d_cand_anchor && d_cand_count && d_select_workspace,
             RN_ERR_BAD_ARG, "rn_detect_tail: NULL candidate array");
  RN_REQUIRE(d_det_boxes && d_det_scores && d_det_classes && d_det_index && d_det_count, RN_ERR_BAD_ARG,
             "rn_detect_tail: NULL output");
  RN_REQUIRE(batch >= 0 && num_classes >= 1 && capacity >= 1 && out_capacity >= 1 && max_output_per_class >= 0 &&
                 topk_per_level >= 0,
             RN_ERR_BAD_ARG, "rn_detect_tail: bad sizes");
  RN_REQUIRE(iou_thr >= 0.f && iou_thr <= 1.f, RN_ERR_BAD_ARG, "iou_threshold must be in [0, 1]");
  RN_REQUIRE(rn_detect_tail_supported(num_classes, capacity), RN_ERR_UNSUPPORTED,
             "rn_detect_tail: capacity %d / %d classes exceed the one-launch tail (use rn_select_finalize + "
             "rn_nms_classwise)", capacity, num_classes);
  RN_REQUIRE(select_workspace_bytes >= rn_select_workspace_bytes(batch, capacity), RN_ERR_CAPACITY,
             "rn_detect_tail: workspace too small");
  RN_REQUIRE((((uintptr_t)d_cand_boxes | (uintptr_t)d_det_boxes) & 15) == 0, RN_ERR_BAD_ARG,
             "box arrays must be 16-byte aligned");
  if (batch == 0) return RN_OK;
  for (int l = 0; l < pyr->num_levels; ++l)
    RN_REQUIRE(regs->p[l] && ((uintptr_t)regs->p[l] & 15) == 0, RN_ERR_BAD_ARG,
               "regs map of level %d missing or not 16-byte aligned", l);
  const size_t n = (size_t)batch * capacity;
  char* ws = (char*)d_select_workspace;
  TailParams tp;
  tp.regs = make_ptrs(regs);
  tp.regs_are_deltas = regs_are_deltas;
  tp.capacity = capacity;
  tp.dedup = dedup;
  tp.ws_anchor = (const int32_t*)ws; ws += align256(n * 4);
  tp.ws_score = (const float*)ws; ws += align256(n * 4);
  tp.ws_class = (const int32_t*)ws;
  tp.cand_boxes = (float4*)d_cand_boxes;
  tp.cand_scores = d_cand_scores;
  tp.cand_classes = d_cand_classes;
  tp.cand_anchor = d_cand_anchor;
  tp.cand_count = d_cand_count;
  tp.num_classes = num_classes;
  tp.out_capacity = out_capacity;
  tp.max_out = max_output_per_class;
  tp.topk = topk_per_level;
  tp.kept_smem = tail_kept_smem(num_classes, capacity, max_output_per_class, topk_per_level > 0);
  tp.iou_thr = iou_thr;
  tp.det_boxes = (float4*)d_det_boxes;
  tp.det_scores = d_det_scores;
  tp.det_classes = d_det_classes;
  tp.det_index = d_det_index;
  tp.det_count = d_det_count;
  tp.status = d_status;
  static bool tail_opted = false;   // once per process, not per call
  if (!tail_opted) {
    cudaFuncSetAttribute(detect_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmemLimit);
    tail_opted = true;
  }
  detect_tail_kernel<<<batch, kTailThreads, tail_smem_bytes(num_classes, capacity, tp.kept_smem, topk_per_level > 0),
                       (cudaStream_t)stream>>>(*pyr, tp);
  RN_CUDA_CHECK_LAUNCH("rn_detect_tail");
  return RN_OK;
}
