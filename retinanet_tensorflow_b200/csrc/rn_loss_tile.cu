// K2/K3, product path: classification loss forward + backward (focal | BCE | balanced BCE, +-dice), the
// candidate pre-selection of the detection path and the Huber regression loss in ONE pass over the
// `[rows, C]` logits (losses.py:6-15,50-60,96-110,115-152; utils.py:183-190,249,270).
//
// Why this shape (profiles/r02_*): the round-1 kernel (register loads, static one-wave
// partition, scalar float32 math) sat at 0.86 of the measured copy bandwidth and fell to 0.53 / 0.26 when the
// logits were not all confident negatives.  Three causes, three changes:
//   * loads held registers and were only in flight while a warp was not computing (stall: long_scoreboard);
//     here ONE producer warp per CTA streams row-aligned tiles global -> shared with bulk async copies
//     (cp.async.bulk + mbarrier transaction counts, SASS: UBLKCP) into a ring of stages, so the bytes in flight
//     per SM are the ring, not the register file;
//   * a static equal partition is ~0.5 ms unbalanced on B200 (SMs do not all see the same HBM bandwidth,
//     profiles/r01_microbench_stream.md); tiles are handed out dynamically by an atomic ticket, and the loss sum
//     is accumulated in 64-bit fixed point so the result stays bit-identical for any hand-out order;
//   * 34 instructions per logit; the bulk math is packed float32x2 (fma.rn.f32x2 / mul / add, SASS: FFMA2, FMUL2,
//     FADD2): two logits per FMA-pipe instruction;
//   * the bulk/careful choice was per warp (one hard negative sent 32 lanes through the libm path); here every
//     lane runs the bulk math, and the units that need the careful path (hard negatives, the row's positive
//     class, candidates) are compacted by ballot into a per-warp queue in shared memory that is drained 32 at a
//     time, so the careful path always runs on full warps, whatever the logit distribution.
#include <string.h>

#include <cmath>

#include "rn_loss_math.cuh"
#include <type_traits>
#include "rn_loss_tile.cuh"

namespace rn {

// Tuning constants (overridable with -D for the sweeps recorded in profiles/r02_tile_sweep.md)
#ifndef RN_TILE_CWARPS
#define RN_TILE_CWARPS 8
#endif
#ifndef RN_TILE_STAGES
#define RN_TILE_STAGES 4
#endif
#ifndef RN_TILE_BYTES
#define RN_TILE_BYTES 20480
#endif
#ifndef RN_TILE_MINB
#define RN_TILE_MINB 2
#endif
#ifndef RN_TILE_BOUND_THREADS
#define RN_TILE_BOUND_THREADS (RN_TILE_CWARPS * 32 + 32)
#endif
constexpr int kConsumerWarps = RN_TILE_CWARPS;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kTileThreads = kConsumerThreads + 32;   // + one producer warp
constexpr int kStages = RN_TILE_STAGES;      // ring depth (compact labels); dense labels: half, stages are twice as big
constexpr int kMaxStages = 8;
constexpr int kQueueCap = 64;        // per-warp queue entries (<= 31 left over + <= 32 pushed)
constexpr int kTileBytes = RN_TILE_BYTES;    // target bytes of logits per tile
constexpr int kMaxTileRows = 2048;
constexpr int kUnitsPerThread = (kTileBytes / 16 + kConsumerThreads - 1) / kConsumerThreads;   // units of a tile per thread
constexpr float kFix = 68719476736.0f;   // fixed-point scale of the loss accumulators (2^36)

// ---- device-side parameter block ----------------------------------------------------------------------
struct TileDev {
  Ptrs logits, label, dlogits, pred, target, dpred;
  float alpha, eps;
  const double* stats;
  double* stats_out;
  unsigned* counters;          // [0] finish ticket, [2] tile ticket
  long long* partials;         // [grid][8]: high / low words of the three fixed-point sums
  int num_classes;
  unsigned cv;                 // 16-byte units per row
  unsigned tile_rows;
  unsigned total_tiles;
  unsigned tile_start[RN_MAX_LEVELS + 1];
  unsigned stage_bytes, x_bytes;   // x_bytes: offset of the label area inside a stage
  unsigned stages;                 // ring depth
  unsigned tab_bytes;
  int write_grad, huber;
  float* loss_out;
  int final_mode;
  double* stats_next;
  float plain_bound;           // a unit is "plain" when its max logit is <= this (min(-0.70, select trigger))
  float sel_trigger, sel_thr;
  int sel_images, sel_capacity;
  int32_t* sel_anchor;
  float* sel_score;
  int32_t* sel_class;
  int32_t* sel_count;
  uint32_t* status;
};

// ---- PTX wrappers: mbarrier, bulk async copy ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (16-byte aligned, size % 16 == 0), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// ---- packed float32x2 math (Blackwell: two lanes of an FMA-pipe instruction) ----------------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// Two confident negatives (logit <= -0.70, z = 0) of the focal loss: the packed twin of focal_neg_bulk
// (rn_loss_math.cuh), same operations in the same order.  Returns w^2 log(q + eps); grad is scaled by cg.
__device__ __forceinline__ float2 focal_neg_bulk2(float2 x, float2 eps, float2 cg, float2& grad, float2& pq) {
#ifdef RN_TILE_MUFU2
  // two MUFU per logit: E = e^x, r = 1/(1+E); p = E r, and 1/(q + eps) = d (1 - eps d) with d = 1 + E = 1/q
  const float2 t = mul2(x, f2(1.44269502163e+00f));
  float2 e, p, r;
  e.x = ex2_approx(t.x);
  e.y = ex2_approx(t.y);
  const float2 d = add2(e, f2(1.0f));
  r.x = rcp_approx(d.x);
  r.y = rcp_approx(d.y);
  p = mul2(e, r);
#else
  const float2 t = mul2(x, f2(-1.44269502163e+00f));
  float2 e, p, r;
  e.x = ex2_approx(t.x);
  e.y = ex2_approx(t.y);
  const float2 d = add2(e, f2(1.0f));
  p.x = rcp_approx(d.x);
  p.y = rcp_approx(d.y);
#endif
  const float2 q = fma2(p, f2(-1.0f), f2(1.0f));    // prob_true = 1 - p
  const float2 w = fma2(q, f2(-1.0f), f2(1.0f));    // the reference's naive 1 - prob_true
  const float2 pe = add2(q, eps);
  const float2 f = add2(pe, f2(-1.0f));
  float2 s = f2(4.589447379e-01f);
  s = fma2(s, f, f2(1.008207817e-02f));
  s = fma2(s, f, f2(2.436887920e-01f));
  s = fma2(s, f, f2(-2.447633296e-01f));
  s = fma2(s, f, f2(3.336158395e-01f));
  s = fma2(s, f, f2(-4.999949932e-01f));
  const float2 lg = fma2(mul2(f, f), s, f);
  const float2 u = mul2(w, lg);
#ifdef RN_TILE_MUFU2
  r = fma2(mul2(d, d), mul2(eps, f2(-1.0f)), d);
#else
  r.x = rcp_approx(pe.x);
  r.y = rcp_approx(pe.y);
#endif
  const float2 v = fma2(f2(-2.0f), u, mul2(mul2(w, w), r));
  pq = mul2(p, q);
  grad = mul2(mul2(cg, v), pq);
  return mul2(w, u);
}

// Two confident negatives of BCE: loss = log1p(e^x), p = e^x / (1 + e^x) (packed twin of bce_neg_bulk)
__device__ __forceinline__ float2 bce_neg_bulk2(float2 x, float2& p) {
  const float2 t = mul2(x, f2(1.44269502163e+00f));
  float2 e, r;
  e.x = ex2_approx(t.x);
  e.y = ex2_approx(t.y);
  const float2 d = add2(e, f2(1.0f));
  r.x = rcp_approx(d.x);
  r.y = rcp_approx(d.y);
  p = mul2(e, r);
  float2 s = f2(-2.785309963e-02f);
  s = fma2(s, e, f2(8.751634508e-02f));
  s = fma2(s, e, f2(-1.473636031e-01f));
  s = fma2(s, e, f2(1.960214078e-01f));
  s = fma2(s, e, f2(-2.495450228e-01f));
  s = fma2(s, e, f2(3.333081305e-01f));
  s = fma2(s, e, f2(-4.999995232e-01f));
  return fma2(mul2(e, e), s, e);
}

// ---- per-thread context ---------------------------------------------------------------------------------
struct TileCtx {
  const float* s_tab;   // [4][C] w_neg, w_pos, dice A, dice B (shared memory)
  int C;
  float eps, cg, inv_fg, inv_mc, alpha;
};

struct QuadOut {
  float4 g;
  float neg, pos;   // loss partials: focal negatives' w^2 log(.) | everything else
};

// ---- careful path: one quad (four logits of one row) that is not "four confident negatives" ---------------
// z: the quad's label values (compact labels: 1 at the row's class, else 0).  noinline keeps it out of the
// hot loop's register allocation; it always runs on (nearly) full warps, see the queue below.
template <int MODE, bool DICE, bool SELECT>
__device__ __noinline__ QuadOut careful_quad(const TileCtx cx, const TileDev& prm, const Segs& segs, int s, unsigned row,
                                             unsigned c4, bool ignore, float4 v, float4 z) {
  QuadOut out;
  out.neg = 0.f;
  out.pos = 0.f;
  out.g = make_float4(0.f, 0.f, 0.f, 0.f);
  const float xs[4] = {v.x, v.y, v.z, v.w};
  const float zs[4] = {z.x, z.y, z.z, z.w};
  if (SELECT) {
    float bp = -INFINITY;
    int bc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (xs[k] > prm.sel_trigger) {
        const float pk = sigmoid_f32(xs[k]);
        if (pk > bp) {
          bp = pk;
          bc = (int)c4 + k;
        }
      }
    }
    if (bp > prm.sel_thr) {
      const unsigned per_image = (unsigned)(segs.rows[s] / prm.sel_images);
      const unsigned bimg = row / per_image;
      const int anchor = (int)(segs.start[s] / prm.sel_images) + (int)(row - bimg * per_image);
      const int slot = reserve_slot(prm.sel_count, (int)bimg, prm.sel_capacity);
      if (slot < prm.sel_capacity) {
        const size_t o = (size_t)bimg * prm.sel_capacity + slot;
        prm.sel_anchor[o] = anchor;
        prm.sel_score[o] = bp;
        prm.sel_class[o] = bc;
      } else {
        raise_flag(prm.status, RN_FLAG_CAND_OVERFLOW);
      }
    }
  }
  if (ignore) return out;
  const float* s_wneg = cx.s_tab;
  const float* s_wpos = cx.s_tab + cx.C;
  const float* s_da = cx.s_tab + 2 * cx.C;
  const float* s_db = cx.s_tab + 3 * cx.C;
  float gs[4];
  if constexpr (MODE == RN_CLS_FOCAL) {
    LossScalars sc;
    sc.focus = 2.0f;
    sc.alpha = cx.alpha;
    sc.one_minus_alpha = 1.0f - cx.alpha;
    sc.eps = cx.eps;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float pq;
      const float ln = focal_neg_fast(xs[k], cx.eps, cx.cg, gs[k], pq);
      if (zs[k] == 1.0f) {  // the positive class: redo this element with z = 1 (libm, IEEE division)
        float lp, gp;
        focal_elem<true>(xs[k], true, sc, lp, gp);
        gp *= cx.inv_fg;
        out.pos += lp;
        if (DICE) gp += s_da[c4 + k] * pq;
        gs[k] = gp;
      } else {
        out.neg += ln;
      }
      if (DICE) gs[k] = fmaf(s_db[c4 + k], pq, gs[k]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float loss = 0.f, pk = 0.f, dx = 0.f;
      bce_elem<true>(xs[k], zs[k], loss, pk, dx);
      float gk = 0.f;
      if (MODE == RN_CLS_NONE) {
        loss = 0.f;
      } else {
        float wk = 1.0f;
        if (MODE == RN_CLS_BALANCED_BCE) wk = zs[k] == 1.0f ? s_wpos[c4 + k] : s_wneg[c4 + k];
        loss *= wk;
        gk = dx * (cx.inv_mc * wk);
      }
      if (DICE) gk += (zs[k] * s_da[c4 + k] + s_db[c4 + k]) * (pk * (1.0f - pk));
      out.pos += loss;
      gs[k] = gk;
    }
  }
  out.g = make_float4(gs[0], gs[1], gs[2], gs[3]);
  return out;
}

// ---- bulk path: four confident negatives, packed math -------------------------------------------------------
// live2: {1,1} for a trainable row, {0,0} for an ignored one (its gradient vanishes without a branch; cg2 carries the
// same factor).  Returns the unit's loss pair unmasked: the caller decides whether it counts.
template <int MODE, bool DICE>
__device__ __forceinline__ float2 bulk_quad(const TileCtx& cx, unsigned c4, float4 v, float2 live2, float2 cg2, float4& g) {
  float2 loss = f2(0.f);
  const float2 xa = make_float2(v.x, v.y), xb = make_float2(v.z, v.w);
  float2 ga, gb;
  if constexpr (MODE == RN_CLS_FOCAL) {
    float2 pqa, pqb;
    const float2 la = focal_neg_bulk2(xa, f2(cx.eps), cg2, ga, pqa);
    const float2 lb = focal_neg_bulk2(xb, f2(cx.eps), cg2, gb, pqb);
    if (DICE) {
      const float4 db = *reinterpret_cast<const float4*>(cx.s_tab + 3 * cx.C + c4);
      ga = fma2(mul2(make_float2(db.x, db.y), live2), pqa, ga);
      gb = fma2(mul2(make_float2(db.z, db.w), live2), pqb, gb);
    }
    loss = add2(la, lb);
  } else {
    float2 pa, pb;
    float2 la = bce_neg_bulk2(xa, pa);
    float2 lb = bce_neg_bulk2(xb, pb);
    if (MODE == RN_CLS_NONE) {
      ga = f2(0.f);
      gb = f2(0.f);
    } else {
      // cg2 = live * inv_mc
      float2 wa = cg2, wb = cg2;
      if (MODE == RN_CLS_BALANCED_BCE) {
        const float4 wn = *reinterpret_cast<const float4*>(cx.s_tab + c4);
        wa = mul2(cg2, make_float2(wn.x, wn.y));
        wb = mul2(cg2, make_float2(wn.z, wn.w));
        la = mul2(la, make_float2(wn.x, wn.y));
        lb = mul2(lb, make_float2(wn.z, wn.w));
      }
      ga = mul2(pa, wa);
      gb = mul2(pb, wb);
      loss = add2(la, lb);
    }
    if (DICE) {
      const float4 db = *reinterpret_cast<const float4*>(cx.s_tab + 3 * cx.C + c4);
      const float2 pqa = mul2(pa, fma2(pa, f2(-1.0f), f2(1.0f)));
      const float2 pqb = mul2(pb, fma2(pb, f2(-1.0f), f2(1.0f)));
      ga = fma2(mul2(make_float2(db.x, db.y), live2), pqa, ga);
      gb = fma2(mul2(make_float2(db.z, db.w), live2), pqb, gb);
    }
  }
  g = make_float4(ga.x, ga.y, gb.x, gb.y);
  return loss;
}

__device__ __forceinline__ long long to_fix(float v) { return __float2ll_rn(v * kFix); }

struct __align__(16) QEntry {
  uint4 x;          // the raw 16-byte unit
  uint4 z;          // dense labels of the unit (DENSE only)
  unsigned row;     // row within the segment
  unsigned cseg;    // column (unit index within the row) | segment << 16
  int lab;
  unsigned pad;
};

// MODE: RN_CLS_FOCAL (focus 2) | RN_CLS_BCE | RN_CLS_BALANCED_BCE | RN_CLS_NONE; DICE adds the dice gradient.
// SELECT: fused candidate pre-selection.  BF16: bf16x8 units.  DENSE: float labels [rows, C], all rows trainable.
template <int MODE, bool DICE, bool SELECT, bool BF16, bool DENSE>
__global__ void __launch_bounds__(RN_TILE_BOUND_THREADS, RN_TILE_MINB)
cls_tile_kernel(const __grid_constant__ Segs segs, const __grid_constant__ TileDev prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  // [0,64) full barriers, [64,128) empty barriers, [128,256) tile info, then tables, queues, stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  int4* tile_info = reinterpret_cast<int4*>(smem + 128);
  float* s_tab = reinterpret_cast<float*>(smem + 256);
  constexpr unsigned kEntry = DENSE ? 48u : 32u;
  unsigned char* s_queue = smem + 256 + prm.tab_bytes;
  unsigned char* s_stage = s_queue + kConsumerWarps * kQueueCap * kEntry;
  __shared__ long long s_red[6][kConsumerWarps];
  __shared__ bool s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8);
  constexpr bool kTables = DICE || MODE == RN_CLS_BALANCED_BCE;
  const int C = prm.num_classes;
  const double nfg = prm.stats[RN_STAT_NUM_FG];
  const double M = prm.stats[RN_STAT_TRAINABLE];

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(bar_full + 8 * i, 1 + (DENSE ? 0 : 32));   // expect_tx arrive (+ one per producer lane: labels)
      mbar_init(bar_empty + 8 * i, kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kTables) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double npos = prm.stats[RN_STAT_NPOS(c)];
      float wn = 1.0f, wp = 1.0f, da = 0.0f, db = 0.0f;
      if (MODE == RN_CLS_BALANCED_BCE) {  // losses.py:99-103, axis=0
        wp = (float)((M - npos) / M);
        wn = (float)(npos / M);
      }
      if (DICE) {  // d/dp [1 - 2I/U] / C = z A + B,  A = -2/(C U), B = 2 I/(C U^2)
        const double I = prm.stats[RN_STAT_INTER(C, c)];
        const double U = npos + prm.stats[RN_STAT_PSUM(C, c)];
        da = (float)(-2.0 / ((double)C * U));
        db = (float)(2.0 * I / ((double)C * U * U));
      }
      if (!(M > 0)) wn = wp = da = db = 0.0f;   // every row ignored: 0/0 weights must not leak into the (zero) gradients
      s_tab[c] = wn;
      s_tab[C + c] = wp;
      s_tab[2 * C + c] = da;
      s_tab[3 * C + c] = db;
    }
  }
  __syncthreads();

  const unsigned cv = prm.cv;
  const unsigned nstages = prm.stages;
  long long fx_neg = 0, fx_pos = 0, fx_hub = 0;

  if (warp == kConsumerWarps) {
    // ================= producer warp: tickets + bulk copies =================
    const unsigned row_bytes = cv * 16u;
    unsigned stage = 0, round = 0;   // ring position without a runtime division per tile
    for (unsigned it = 0;; ++it, ++stage) {
      if (stage == nstages) {
        stage = 0;
        ++round;
      }
      if (round) mbar_wait(bar_empty + 8 * stage, (round - 1) & 1);
      unsigned t = 0;
      if (lane == 0) {
        // the first `nstages` tiles of a CTA are static (no ticket latency while the ring fills)
        t = it < nstages ? it * gridDim.x + blockIdx.x : nstages * gridDim.x + atomicAdd(prm.counters + 2, 1u);
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      const uint32_t full = bar_full + 8 * stage;
      if (t >= prm.total_tiles) {
        if (lane == 0) tile_info[stage] = make_int4(-1, 0, 0, 0);
        __syncwarp();
        if (DENSE) {
          if (lane == 0) mbar_arrive(full);
        } else {
          mbar_arrive(full);
          if (lane == 0) mbar_arrive(full);
        }
        break;
      }
      int s = 0;
#pragma unroll
      for (int i = 1; i < RN_MAX_LEVELS; ++i) s += (i < segs.num && t >= prm.tile_start[i]) ? 1 : 0;
      const unsigned row0 = (t - prm.tile_start[s]) * prm.tile_rows;
      const unsigned left = (unsigned)segs.rows[s] - row0;
      const unsigned nrows = left < prm.tile_rows ? left : prm.tile_rows;
      unsigned char* st = s_stage + (size_t)stage * prm.stage_bytes;
      // the tile's compact labels: one more bulk copy when the 4-byte rows allow it (16-byte aligned source, whole
      // 16-byte size -- every full tile of a 16-byte aligned level), else plain loads by the producer lanes
      const int32_t* lab_g = DENSE ? nullptr : reinterpret_cast<const int32_t*>(prm.label.p[s]) + row0;
      const bool lab_bulk = !DENSE && ((reinterpret_cast<uintptr_t>(lab_g) & 15u) == 0) && ((nrows & 3u) == 0);
      if (lane == 0) {
        tile_info[stage] = make_int4(s, (int)row0, (int)nrows, 0);
        const unsigned bytes = nrows * row_bytes;
        mbar_expect_tx(full, DENSE ? 2 * bytes : (lab_bulk ? bytes + nrows * 4u : bytes));
        bulk_g2s(smem_u32(st), reinterpret_cast<const unsigned char*>(prm.logits.p[s]) + (size_t)row0 * row_bytes, bytes,
                 full);
        if (DENSE)
          bulk_g2s(smem_u32(st + prm.x_bytes),
                   reinterpret_cast<const unsigned char*>(prm.label.p[s]) + (size_t)row0 * row_bytes, bytes, full);
        else if (lab_bulk)
          bulk_g2s(smem_u32(st + prm.x_bytes), lab_g, nrows * 4u, full);
      }
      if (!DENSE) {
        if (!lab_bulk) {
          int32_t* lab_s = reinterpret_cast<int32_t*>(st + prm.x_bytes);
          for (unsigned r = lane; r < nrows; r += 32) lab_s[r] = __ldg(lab_g + r);
        }
        mbar_arrive(full);   // one arrival per producer lane (release: orders the plain stores above)
      }
    }
  } else {
    // ================= consumer warps =================
    TileCtx cx;
    cx.s_tab = s_tab;
    cx.C = C;
    cx.eps = prm.eps;
    cx.alpha = prm.alpha;
    cx.inv_fg = (float)(1.0 / fmax(nfg, 1.0));
    cx.inv_mc = M > 0 ? (float)(1.0 / (M * (double)C)) : 0.0f;
    cx.cg = (1.0f - prm.alpha) * cx.inv_fg;
    const float gscale_h = nfg > 0 ? (float)(1.0 / (4.0 * nfg)) : 0.0f;
    const float cgrad = MODE == RN_CLS_FOCAL ? cx.cg : cx.inv_mc;
    const bool write_grad = prm.write_grad != 0;
    const float bound = prm.plain_bound;
    unsigned char* my_queue = s_queue + (size_t)warp * kQueueCap * kEntry;
    unsigned qcount = 0;   // warp-uniform
    const unsigned lane_lt = (1u << lane) - 1u;
    using Unit = uint4;

    // one careful unit (a queue entry): compute, store the gradient, accumulate the loss in fixed point
    auto careful_unit = [&](const Unit& raw, const Unit& zraw, unsigned row, unsigned col, int s, int lab) {
      Unit outv;
      float neg = 0.f, pos = 0.f;
      const bool ignore = !DENSE && lab == RN_LABEL_IGNORE;
      auto zq = [&](unsigned c4, int half) {
        if (DENSE) {
          (void)half;
          return *reinterpret_cast<const float4*>(&zraw);
        }
        const unsigned d = (unsigned)(lab - (int)c4);
        return make_float4(d == 0u ? 1.f : 0.f, d == 1u ? 1.f : 0.f, d == 2u ? 1.f : 0.f, d == 3u ? 1.f : 0.f);
      };
      if constexpr (BF16) {
        float4 lo, hi;
        unpack_bf16x8(raw, lo, hi);
        const QuadOut a = careful_quad<MODE, DICE, SELECT>(cx, prm, segs, s, row, col * 8u, ignore, lo, zq(col * 8u, 0));
        const QuadOut b =
            careful_quad<MODE, DICE, SELECT>(cx, prm, segs, s, row, col * 8u + 4u, ignore, hi, zq(col * 8u + 4u, 1));
        outv = pack_bf16x8(a.g, b.g);
        neg = a.neg + b.neg;
        pos = a.pos + b.pos;
      } else {
        const QuadOut a = careful_quad<MODE, DICE, SELECT>(cx, prm, segs, s, row, col * 4u, ignore,
                                                           *reinterpret_cast<const float4*>(&raw), zq(col * 4u, 0));
        outv = *reinterpret_cast<const Unit*>(&a.g);
        neg = a.neg;
        pos = a.pos;
      }
      if (write_grad) __stcs(reinterpret_cast<Unit*>(prm.dlogits.p[s]) + ((size_t)row * cv + col), outv);
      fx_neg += to_fix(neg);
      fx_pos += to_fix(pos);
    };
    auto drain = [&](unsigned n) {   // the last n (<= 32) entries of the queue
      if ((unsigned)lane < n) {
        const unsigned char* e = my_queue + (size_t)(qcount - n + lane) * kEntry;
        const Unit raw = *reinterpret_cast<const Unit*>(e);
        Unit zraw = make_uint4(0, 0, 0, 0);
        if (DENSE) zraw = *reinterpret_cast<const Unit*>(e + 16);
        const uint4 meta = *reinterpret_cast<const uint4*>(e + (DENSE ? 32 : 16));
        careful_unit(raw, zraw, meta.x, meta.y & 0xffffu, (int)(meta.y >> 16), (int)meta.z);
      }
      qcount -= n;
      __syncwarp();
    };

    // A tile starts on a row boundary, so the (row-in-tile, first class) of this thread's k-th unit is the same for
    // every tile: one division per unit per KERNEL instead of per tile.
    constexpr unsigned kPerUnit = BF16 ? 8u : 4u;
    unsigned geo[kUnitsPerThread];
#pragma unroll
    for (int k = 0; k < kUnitsPerThread; ++k) {
      const unsigned i = threadIdx.x + (unsigned)k * kConsumerThreads;
      const unsigned r = i / cv;   // once per kernel
      geo[k] = r | (((i - r * cv) * kPerUnit) << 16);
      asm volatile("" : "+r"(geo[k]));   // keep it in a register (the compiler would recompute it per tile)
    }

    unsigned stage = 0, round = 0;
    for (;; ++stage) {
      if (stage == nstages) {
        stage = 0;
        ++round;
      }
      mbar_wait(bar_full + 8 * stage, round & 1);
      const int4 info = tile_info[stage];
      if (info.x < 0) break;
      const int s = info.x;
      const unsigned row0 = (unsigned)info.y, nrows = (unsigned)info.z;
      const unsigned char* st = s_stage + (size_t)stage * prm.stage_bytes;
      const Unit* xs = reinterpret_cast<const Unit*>(st);
      const Unit* zs = reinterpret_cast<const Unit*>(st + prm.x_bytes);          // DENSE
      const int32_t* labs = reinterpret_cast<const int32_t*>(st + prm.x_bytes);  // compact
      // this thread's first gradient unit of the tile; its other units sit at fixed distances (immediate offsets)
      Unit* gmine = reinterpret_cast<Unit*>(prm.dlogits.p[s]) + ((size_t)row0 * cv + threadIdx.x);
      asm volatile("" : "+l"(gmine));   // (the compiler would otherwise redo the 64-bit address per unit)
      const unsigned tile_units = nrows * cv;
      float2 acc = f2(0.f);

      // (a full tile -- all but the last tile of a level -- gives every thread all of its units: that copy of the
      //  loop carries no bounds checks and, being the training path, no "gradient wanted?" test either)
      auto units_of_tile = [&](auto full_c) {
      constexpr bool kFull = decltype(full_c)::value;
#pragma unroll
      for (int k = 0; k < kUnitsPerThread; ++k) {
        const unsigned i = threadIdx.x + (unsigned)k * kConsumerThreads;
        if (!kFull && (unsigned)warp * 32u + (unsigned)k * kConsumerThreads >= tile_units) break;   // warp-uniform
        const bool valid = kFull || i < tile_units;
        const unsigned r = geo[k] & 0xffffu, c4 = geo[k] >> 16;
        Unit raw = make_uint4(0, 0, 0, 0);
        if (valid) raw = xs[i];
        int lab = RN_LABEL_BACKGROUND;
        Unit zraw = make_uint4(0, 0, 0, 0);
        bool plain;
        float2 live2 = f2(1.0f);
        if (DENSE) {
          if (valid) zraw = zs[i];
          plain = ((zraw.x | zraw.y) | (zraw.z | zraw.w)) == 0u;   // four exact +0 labels
        } else {
          if (valid) lab = labs[r];
          plain = (unsigned)(lab - (int)c4) >= kPerUnit;   // also true for IGNORE / BACKGROUND (< 0)
          live2 = f2(lab == RN_LABEL_IGNORE ? 0.0f : 1.0f);
        }
        const float2 cg2 = mul2(live2, f2(cgrad));
        Unit outv;
        float2 a2;
        if constexpr (BF16) {
          float4 lo, hi, glo, ghi;
          unpack_bf16x8(raw, lo, hi);
          const float xmax = fmaxf(fmaxf(fmaxf(lo.x, lo.y), fmaxf(lo.z, lo.w)), fmaxf(fmaxf(hi.x, hi.y), fmaxf(hi.z, hi.w)));
          plain = plain && xmax <= bound;
          a2 = bulk_quad<MODE, DICE>(cx, c4, lo, live2, cg2, glo);
          a2 = add2(a2, bulk_quad<MODE, DICE>(cx, c4 + 4u, hi, live2, cg2, ghi));
          outv = pack_bf16x8(glo, ghi);
        } else {
          const float4 v = *reinterpret_cast<const float4*>(&raw);
          const float xmax = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
          plain = plain && xmax <= bound;
          float4 g;
          a2 = bulk_quad<MODE, DICE>(cx, c4, v, live2, cg2, g);
          outv = *reinterpret_cast<const Unit*>(&g);
        }
        plain = plain && valid;
        // the unit's loss only counts when it really was a bulk unit of a trainable row.  A SELECT, not a multiply
        // by zero: the bulk math of a unit that belongs to the careful path may have overflowed (inf * 0 = NaN)
        const bool counts = plain && lab != RN_LABEL_IGNORE;
        a2.x = counts ? a2.x : 0.0f;
        a2.y = counts ? a2.y : 0.0f;
        acc = add2(acc, a2);
        if (plain && (kFull || write_grad)) __stcs(gmine + k * kConsumerThreads, outv);
        const unsigned m = __ballot_sync(0xffffffffu, valid && !plain);
        if (m) {
          if (valid && !plain) {
            unsigned char* e = my_queue + (size_t)(qcount + __popc(m & lane_lt)) * kEntry;
            *reinterpret_cast<Unit*>(e) = raw;
            if (DENSE) *reinterpret_cast<Unit*>(e + 16) = zraw;
            *reinterpret_cast<uint4*>(e + (DENSE ? 32 : 16)) =
                make_uint4(row0 + r, (c4 / kPerUnit) | ((unsigned)s << 16), (unsigned)lab, 0u);
          }
          qcount += __popc(m);
          __syncwarp();
          if (qcount >= 32u) drain(32u);
        }
      }
      };
#ifdef RN_TILE_NO_FULL   // (tuning builds: the one-copy form)
      units_of_tile(std::false_type{});
#else
      if (tile_units == (unsigned)kUnitsPerThread * kConsumerThreads && write_grad) units_of_tile(std::true_type{});
      else units_of_tile(std::false_type{});
#endif
      fx_neg += to_fix(acc.x + acc.y);

      if (!DENSE && prm.huber) {
        // Huber of the tile's rows (losses.py:144-152): label >= 0 <=> foreground
        const float4* pred = reinterpret_cast<const float4*>(prm.pred.p[s]) + row0;
        const float4* targ = reinterpret_cast<const float4*>(prm.target.p[s]) + row0;
        float4* dp = reinterpret_cast<float4*>(prm.dpred.p[s]);
        for (unsigned r = threadIdx.x; r < nrows; r += kConsumerThreads) {
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          if (labs[r] >= 0) {
            const float4 p = __ldg(pred + r);
            const float4 t = __ldg(targ + r);
            const float e[4] = {fsub(p.x, t.x), fsub(p.y, t.y), fsub(p.z, t.z), fsub(p.w, t.w)};
            float ls = 0.f, gg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float ae = fabsf(e[k]);
              const float q = fminf(ae, 1.0f);
              ls += fadd(0.5f * fmul(q, q), fsub(ae, q));
              gg[k] = fminf(fmaxf(e[k], -1.0f), 1.0f) * gscale_h;
            }
            fx_hub += to_fix(ls);
            g = make_float4(gg[0], gg[1], gg[2], gg[3]);
          }
          if (dp) __stcs(dp + row0 + r, g);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
    }
    if (qcount) drain(qcount);
  }

  // ---- order-independent reduction: integer sums.  A thread's total fits 64 bits with a wide margin; sums ACROSS
  // threads and CTAs are kept as separate high / low words so that no realistic or adversarial input can overflow.
  long long acc6[6] = {fx_neg >> 24, fx_neg & 0xffffff, fx_pos >> 24, fx_pos & 0xffffff, fx_hub >> 24, fx_hub & 0xffffff};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc6[k] += __shfl_xor_sync(0xffffffffu, acc6[k], o);
  }
  if (lane == 0 && warp < kConsumerWarps) {
#pragma unroll
    for (int k = 0; k < 6; ++k) s_red[k][warp] = acc6[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    long long a = 0;
    for (int w = 0; w < kConsumerWarps; ++w) a += s_red[threadIdx.x][w];
    prm.partials[8 * (size_t)blockIdx.x + threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(prm.counters, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    long long tot[6] = {0, 0, 0, 0, 0, 0};
    for (unsigned i = lane; i < gridDim.x; i += 32) {
#pragma unroll
      for (int k = 0; k < 6; ++k) tot[k] += __ldcg(prm.partials + 8 * (size_t)i + k);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot[k] += __shfl_xor_sync(0xffffffffu, tot[k], o);
    }
    if (prm.stats_next)   // the next step's statistics vector starts from zero without a launch of its own
      for (int i = lane; i < RN_STATS_HEADER + 3 * C; i += 32) prm.stats_next[i] = 0.0;
    if (lane == 0) {
      const double sc = 1.0 / 68719476736.0;
      const double neg = ((double)tot[0] * 16777216.0 + (double)tot[1]) * sc;
      const double pos = ((double)tot[2] * 16777216.0 + (double)tot[3]) * sc;
      const double hub = ((double)tot[4] * 16777216.0 + (double)tot[5]) * sc;
      // focal: negatives accumulated w^2 log(q+eps), loss = -(1-alpha) * that; BCE modes: plain sum
      prm.stats_out[RN_STAT_CLS_SUM] = MODE == RN_CLS_FOCAL ? pos - (double)(1.0f - prm.alpha) * neg : pos + neg;
      if (prm.huber) prm.stats_out[RN_STAT_HUBER_SUM] = hub;
      if (prm.loss_out) {   // single GPU: nothing to exchange, finish the scalars here (as loss_finalize_kernel)
        const double cls_sum = prm.stats_out[RN_STAT_CLS_SUM];
        const int base = prm.final_mode & 15;
        double cls = 0.0;
        if (base == RN_CLS_FOCAL) cls += cls_sum / fmax(nfg, 1.0);
        if (base == RN_CLS_BCE || base == RN_CLS_BALANCED_BCE) cls += M > 0 ? cls_sum / (M * (double)C) : 0.0;
        if (prm.final_mode & RN_CLS_DICE_FLAG) {
          double dsum = 0.0;
          for (int c = 0; c < C; ++c) {
            const double I = prm.stats[RN_STAT_INTER(C, c)];
            const double U = prm.stats[RN_STAT_NPOS(c)] + prm.stats[RN_STAT_PSUM(C, c)];
            dsum += 1.0 - 2.0 * I / U;
          }
          cls += dsum / (double)C;
        }
        prm.loss_out[0] = (float)cls;
        prm.loss_out[1] = nfg > 0 ? (float)((prm.huber ? hub : prm.stats[RN_STAT_HUBER_SUM]) / (4.0 * nfg)) : 0.0f;
        prm.loss_out[2] = (float)nfg;
        prm.loss_out[3] = (float)M;
      }
      prm.counters[0] = 0;
      prm.counters[2] = 0;
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
bool cls_tile_supported(int num_classes, bool bf16, bool dense) {
  const int per = bf16 ? 8 : 4;
  if (num_classes < per || num_classes % per != 0) return false;
  if (bf16 && dense) return false;
  const size_t row_bytes = (size_t)num_classes * (bf16 ? 2 : 4);
  return row_bytes <= (size_t)kTileBytes && num_classes <= 2048;
}

template <int MODE, bool DICE, bool SELECT, bool BF16, bool DENSE>
static int launch_variant(const Segs& segs, const TileDev& dev, size_t smem, cudaStream_t st) {
  auto kernel = cls_tile_kernel<MODE, DICE, SELECT, BF16, DENSE>;
  static bool configured = false;   // per instantiation
  static int per_sm = 1;
  static size_t configured_smem = 0, occupancy_smem = ~(size_t)0;
  if (!configured || smem > configured_smem) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("tile loss kernel: %zu bytes of shared memory refused", smem);
      cudaGetLastError();
      return RN_ERR_CUDA;
    }
    configured_smem = smem;
    configured = true;
  }
  if (occupancy_smem != smem) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTileThreads, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    occupancy_smem = smem;
  }
  long long grid = (long long)num_sms() * per_sm;
  if (grid > dev.total_tiles) grid = dev.total_tiles;
  if (grid < 1) grid = 1;
  kernel<<<(unsigned)grid, kTileThreads, smem, st>>>(segs, dev);
  RN_CUDA_CHECK_LAUNCH("tile classification-loss kernel");
  return RN_OK;
}

template <int MODE, bool DICE>
static int launch_mode(const Segs& segs, const TileDev& dev, bool select, bool bf16, bool dense, size_t smem,
                       cudaStream_t st) {
#ifdef RN_TILE_SWEEP   // tuning builds: only the headline instantiation
  if (MODE == RN_CLS_FOCAL && !DICE && select && !bf16 && !dense)
    return launch_variant<RN_CLS_FOCAL, false, true, false, false>(segs, dev, smem, st);
  return RN_ERR_UNSUPPORTED;
#else
  if (dense) return launch_variant<MODE, DICE, false, false, true>(segs, dev, smem, st);
  if (bf16)
    return select ? launch_variant<MODE, DICE, true, true, false>(segs, dev, smem, st)
                  : launch_variant<MODE, DICE, false, true, false>(segs, dev, smem, st);
  return select ? launch_variant<MODE, DICE, true, false, false>(segs, dev, smem, st)
                : launch_variant<MODE, DICE, false, false, false>(segs, dev, smem, st);
#endif
}

int launch_cls_tile(const Segs& segs, const TileParams& tp, int mode, bool select, cudaStream_t st) {
  const bool bf16 = tp.bf16 != 0, dense = tp.dense != 0;
  RN_REQUIRE(cls_tile_supported(tp.num_classes, bf16, dense), RN_ERR_UNSUPPORTED, "tile loss kernel: unsupported shape");
  TileDev dev;
  memset(&dev, 0, sizeof(dev));
  dev.logits = tp.logits;
  dev.label = tp.label;
  dev.dlogits = tp.dlogits;
  dev.pred = tp.pred;
  dev.target = tp.target;
  dev.dpred = tp.dpred;
  dev.alpha = tp.alpha;
  dev.eps = tp.eps;
  dev.stats = tp.stats;
  dev.stats_out = tp.stats_out;
  dev.counters = reinterpret_cast<unsigned*>(tp.workspace);
  dev.partials = reinterpret_cast<long long*>(reinterpret_cast<char*>(tp.workspace) + 256);
  dev.num_classes = tp.num_classes;
  dev.cv = (unsigned)(tp.num_classes / (bf16 ? 8 : 4));
  const unsigned row_bytes = dev.cv * 16u;
  unsigned tr = (unsigned)kTileBytes / row_bytes;   // dense labels: the stage also holds a label tile of this size
  if (tr < 1) tr = 1;
  if (tr > (unsigned)kMaxTileRows) tr = kMaxTileRows;
  if (tr >= 4) tr &= ~3u;
  // prefer a tile whose units divide evenly among the consumer threads (every thread then does the same number of
  // units per tile): look a little below the byte budget for such a row count
  for (unsigned r = tr; r >= 4 && r * 8u >= tr * 7u; r -= 4)
    if ((r * dev.cv) % (unsigned)kConsumerThreads == 0) {
      tr = r;
      break;
    }
  dev.tile_rows = tr;
  unsigned long long tiles = 0;
  for (int i = 0; i < RN_MAX_LEVELS; ++i) {
    dev.tile_start[i] = (unsigned)tiles;
    if (i < segs.num) {
      RN_REQUIRE(segs.rows[i] < (1ll << 32) / (long long)dev.cv - (1ll << 16), RN_ERR_UNSUPPORTED,
                 "tile loss kernel: a level with >= 2^32 16-byte units");
      tiles += (unsigned long long)((segs.rows[i] + tr - 1) / tr);
    }
  }
  dev.tile_start[RN_MAX_LEVELS] = (unsigned)tiles;
  RN_REQUIRE(tiles < 0xfff00000ull, RN_ERR_UNSUPPORTED, "tile loss kernel: too many tiles");
  dev.total_tiles = (unsigned)tiles;
  dev.x_bytes = tr * row_bytes;
  const unsigned lab_bytes = dense ? tr * row_bytes : ((tr * 4u + 15u) & ~15u);
  dev.stage_bytes = (dev.x_bytes + lab_bytes + 127u) & ~127u;
  const int base = mode & 15;
  const bool dice = (mode & RN_CLS_DICE_FLAG) != 0;
  dev.tab_bytes = (dice || base == RN_CLS_BALANCED_BCE) ? (unsigned)((4 * tp.num_classes * sizeof(float) + 127) & ~127u) : 0u;
  dev.write_grad = tp.write_grad;
  dev.loss_out = tp.loss_out;
  dev.final_mode = tp.final_mode;
  dev.stats_next = tp.stats_next;
  dev.huber = (!dense && tp.pred.p[0] != nullptr) ? 1 : 0;
  dev.sel_trigger = tp.sel_trigger;
  dev.sel_thr = tp.sel_thr;
  dev.plain_bound = select ? fminf(-0.70f, tp.sel_trigger) : -0.70f;
  dev.sel_images = tp.sel_images;
  dev.sel_capacity = tp.sel_capacity;
  dev.sel_anchor = tp.sel_anchor;
  dev.sel_score = tp.sel_score;
  dev.sel_class = tp.sel_class;
  dev.sel_count = tp.sel_count;
  dev.status = tp.status;
  dev.stages = dense ? (kStages >= 4 ? kStages / 2 : 2) : kStages;
  const size_t smem = 256 + dev.tab_bytes + (size_t)kConsumerWarps * kQueueCap * (dense ? 48 : 32) +
                      (size_t)dev.stages * dev.stage_bytes;
  if (dev.total_tiles == 0) {   // nothing to stream: the sums are zero
    // (kernel launch with grid 1 handles it: the producer immediately publishes the end marker)
  }
#define RN_TILE_MODE(M)                                                            \
  (dice ? launch_mode<M, true>(segs, dev, select, bf16, dense, smem, st)           \
        : launch_mode<M, false>(segs, dev, select, bf16, dense, smem, st))
  switch (base) {
    case RN_CLS_FOCAL: return RN_TILE_MODE(RN_CLS_FOCAL);
    case RN_CLS_BCE: return RN_TILE_MODE(RN_CLS_BCE);
    case RN_CLS_BALANCED_BCE: return RN_TILE_MODE(RN_CLS_BALANCED_BCE);
    default: return launch_mode<RN_CLS_NONE, true>(segs, dev, select, bf16, dense, smem, st);
  }
#undef RN_TILE_MODE
}

}  // namespace rn
