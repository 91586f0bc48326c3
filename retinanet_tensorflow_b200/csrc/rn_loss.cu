// K2/K3 entry points (classification + regression losses, forward and backward: losses.py:6-15,50-60,96-110,
// 115-152; utils.process_labels_and_logits' sigmoid and trainable masking, utils.py:249,270) and the kernels that are
// NOT the product's streaming kernel.
//
// Normalisers (M, num_fg, per-class sums) live in a small double-precision stats vector in device memory, so nothing
// is synchronised with the host and the multi-GPU exchange is one all-reduce of that vector.
//
// Layout:
//   rn_loss_tile.cu       (other file) cls_tile_kernel -- the product path: TMA-bulk ring, dynamic tiles, packed
//                         f32x2 math, careful queue; compact or dense labels, float32 or bf16, +-dice, fused Huber /
//                         candidate pre-selection / final scalars.  The entry points below route to it.
//   cls_loss_kernel       generic kernel (any mode, dense or compact labels, any C, any focus): libm + IEEE division,
//                         ~110 instructions per logit -- the fallback for C % 4 != 0 / focus != 2 and, with
//                         RN_CLS_GENERIC_FLAG, the parity tests' second implementation
//   cls_sums_*            dice phase 1 (per-class sums); huber_*; loss_finalize; elementwise / column-sum / softmax
//                         forms for the reference's own signatures
#include <string.h>

#include <cmath>

#include <cuda_bf16.h>

#include "rn_common.cuh"
#include "rn_loss_math.cuh"
#include "rn_loss_tile.cuh"

namespace rn {

constexpr int kLossThreads = 256;
constexpr int kMaxLossBlocks = 2048;


struct ClsParams {
  Ptrs logits, label, dlogits;
  LossScalars sc;
  const double* stats;
  double* stats_out;     // stats (written: RN_STAT_CLS_SUM)
  double* partials;      // [gridDim.x]
  unsigned* ticket;
  int num_classes;
  unsigned cv;           // units (of VEC elements) per row
  int write_grad;
};

// MODE: RN_CLS_NONE/FOCAL/BCE/BALANCED_BCE;  DICE: add dice gradient;  DENSE: float labels
// [rows, C] instead of compact int32;  VEC: elements per thread-unit (4 -> 128-bit accesses).
template <int MODE, bool DICE, bool DENSE, int VEC>
__global__ void __launch_bounds__(kLossThreads)
cls_loss_kernel(const __grid_constant__ Segs segs, const __grid_constant__ ClsParams prm) {
  extern __shared__ float s_tab[];  // [4][C]: w_pos, w_neg, dice A, dice B
  __shared__ double s_red[32];
  const int C = prm.num_classes;
  float* s_wpos = s_tab;
  float* s_wneg = s_tab + C;
  float* s_da = s_tab + 2 * C;
  float* s_db = s_tab + 3 * C;

  const double M = prm.stats[RN_STAT_TRAINABLE];
  const double nfg = prm.stats[RN_STAT_NUM_FG];
  const float inv_fg = (float)(1.0 / fmax(nfg, 1.0));                  // losses.py:121
  const float inv_mc = M > 0 ? (float)(1.0 / (M * (double)C)) : 0.0f;  // reduce_mean over [M, C]
  if (MODE == RN_CLS_BALANCED_BCE || DICE) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double npos = prm.stats[RN_STAT_NPOS(c)];
      if (MODE == RN_CLS_BALANCED_BCE) {  // losses.py:99-103, axis=0
        s_wpos[c] = (float)((M - npos) / M);
        s_wneg[c] = (float)(npos / M);
      }
      if (DICE) {  // d/dp [1 - 2I/U] / C = (z * A + B),  A = -2/(C U), B = 2 I/(C U^2)
        const double I = prm.stats[RN_STAT_INTER(C, c)];
        const double U = npos + prm.stats[RN_STAT_PSUM(C, c)];
        s_da[c] = (float)(-2.0 / ((double)C * U));
        s_db[c] = (float)(2.0 * I / ((double)C * U * U));
      }
    }
    __syncthreads();
  }

  const long long total_units = segs.start[RN_MAX_LEVELS] * (long long)prm.cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;

  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < total_units; u += stride) {
    const long long r = u / prm.cv;
    const int c0 = (int)(u - r * prm.cv) * VEC;
    const int s = seg_of(segs, r);
    const long long rl = r - segs.start[s];
    const long long off = rl * C + c0;
    const float* xp = reinterpret_cast<const float*>(prm.logits.p[s]) + off;
    float x[VEC], z[VEC], g[VEC];
    bool trainable = true;
    if constexpr (VEC == 4) {
      const float4 v = ld_stream4(xp);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
      x[0] = __ldcs(xp);
    }
    if (DENSE) {
      const float* zp = reinterpret_cast<const float*>(prm.label.p[s]) + off;
      if constexpr (VEC == 4) {
        const float4 v = ld_stream4(zp);
        z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
      } else {
        z[0] = __ldcs(zp);
      }
    } else {
      const int lab = __ldg(reinterpret_cast<const int32_t*>(prm.label.p[s]) + rl);
      trainable = lab != RN_LABEL_IGNORE;
#pragma unroll
      for (int k = 0; k < VEC; ++k) z[k] = (lab == c0 + k) ? 1.0f : 0.0f;
    }
    float lsum = 0.0f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      g[k] = 0.0f;
      if (!trainable) continue;
      float p = 0.0f, loss = 0.0f;
      bool have_p = false;
      if (MODE == RN_CLS_FOCAL) {
        float gr = 0.0f;
        focal_elem<true>(x[k], z[k] == 1.0f, prm.sc, loss, gr);
        g[k] = gr * inv_fg;
      } else if (MODE == RN_CLS_BCE || MODE == RN_CLS_BALANCED_BCE) {
        float dx = 0.0f;
        bce_elem<true>(x[k], z[k], loss, p, dx);
        have_p = true;
        float wgt = inv_mc;
        if (MODE == RN_CLS_BALANCED_BCE) {
          const float wc = (z[k] == 1.0f) ? s_wpos[c0 + k] : s_wneg[c0 + k];
          loss *= wc;
          wgt *= wc;
        }
        g[k] = dx * wgt;
      }
      if (DICE) {
        if (!have_p) p = sigmoid_f32(x[k]);
        g[k] += (z[k] * s_da[c0 + k] + s_db[c0 + k]) * (p * (1.0f - p));
      }
      lsum += loss;
    }
    acc += (double)lsum;
    if (prm.write_grad) {
      float* gp = reinterpret_cast<float*>(prm.dlogits.p[s]) + off;
      if constexpr (VEC == 4) {
        st_stream4(gp, make_float4(g[0], g[1], g[2], g[3]));
      } else {
        __stcs(gp, g[0]);
      }
    }
  }

  // ---- deterministic reduction of the loss sum
  acc = block_sum(acc, s_red);
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    prm.partials[blockIdx.x] = acc;
    __threadfence();
    const unsigned t = atomicAdd(prm.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double v = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(prm.partials + i);
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) {
      prm.stats_out[RN_STAT_CLS_SUM] = v;
      *prm.ticket = 0;
    }
  }
}

// Arguments of the streaming classification-loss kernel as the entry points below assemble them
struct FastParams {
  Ptrs logits, label, dlogits;
  float alpha, eps;
  const double* stats;
  double* stats_out;
  double* partials;
  unsigned* ticket;
  int num_classes;
  int cv;           // 16-byte units per row (C/4 float32, C/8 bfloat16)
  int write_grad;
  int bf16;         // logits / dlogits are bfloat16
  // fused candidate pre-selection (rn_cls_loss_select): the same pass also finds the anchors whose
  // max class probability exceeds `thr` (utils.boxes_decode, utils.py:183-190)
  float sel_trigger;   // logit below which sigmoid(x) <= thr for sure
  float sel_thr;
  int sel_images, sel_capacity;
  int32_t* sel_anchor;   // [B, capacity] unordered, one entry per firing UNIT (deduplicated later)
  float* sel_score;
  int32_t* sel_class;
  int32_t* sel_count;    // [B]
  uint32_t* status;
};

// ---- per-class sums for dice: I_c = sum z p, P_c = sum p (losses.py:53-56) ---------------------
struct SumsParams {
  Ptrs logits, label;
  double* stats_out;
  double* partials;  // [gridDim.x][2C]
  unsigned* ticket;
  int num_classes;
  int cv;   // units per row
  int ry;   // rows per block iteration
};

template <bool DENSE, int VEC>
__global__ void __launch_bounds__(1024) cls_sums_kernel(const __grid_constant__ Segs segs,
                                                        const __grid_constant__ SumsParams prm) {
  extern __shared__ double s_acc[];  // [ry][2C]
  const int C = prm.num_classes;
  const int tx = threadIdx.x % prm.cv, ty = threadIdx.x / prm.cv;
  const long long total_rows = segs.start[RN_MAX_LEVELS];
  double ai[VEC], ap[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) ai[k] = ap[k] = 0.0;
  const int c0 = tx * VEC;
  if (ty < prm.ry) {
    for (long long r = (long long)blockIdx.x * prm.ry + ty; r < total_rows; r += (long long)gridDim.x * prm.ry) {
      const int s = seg_of(segs, r);
      const long long rl = r - segs.start[s];
      const long long off = rl * C + c0;
      const float* xp = reinterpret_cast<const float*>(prm.logits.p[s]) + off;
      float x[VEC], z[VEC];
      if constexpr (VEC == 4) {
        const float4 v = ld_stream4(xp);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      } else {
        x[0] = __ldcs(xp);
      }
      if (DENSE) {
        const float* zp = reinterpret_cast<const float*>(prm.label.p[s]) + off;
        if constexpr (VEC == 4) {
          const float4 v = ld_stream4(zp);
          z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
        } else {
          z[0] = __ldcs(zp);
        }
      } else {
        const int lab = __ldg(reinterpret_cast<const int32_t*>(prm.label.p[s]) + rl);
        if (lab == RN_LABEL_IGNORE) continue;
#pragma unroll
        for (int k = 0; k < VEC; ++k) z[k] = (lab == c0 + k) ? 1.0f : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float p;
        if (x[k] <= -0.70f) {  // same bulk formula as cls_fast_kernel: p = e / (1 + e), e = exp(x)
          const float e = ex2_approx(x[k] * 1.44269502163e+00f);
          p = e * rcp_approx(1.0f + e);
        } else {
          p = sigmoid_f32(x[k]);
        }
        ai[k] += (double)(z[k] * p);
        ap[k] += (double)p;
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s_acc[(size_t)ty * 2 * C + c0 + k] = ai[k];
      s_acc[(size_t)ty * 2 * C + C + c0 + k] = ap[k];
    }
  }
  __syncthreads();
  double* mine = prm.partials + (size_t)blockIdx.x * 2 * C;
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
    double v = 0.0;
    for (int y = 0; y < prm.ry; ++y) v += s_acc[(size_t)y * 2 * C + j];
    mine[j] = v;
  }
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(prm.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(prm.partials + (size_t)b * 2 * C + j);
      prm.stats_out[RN_STATS_HEADER + C + j] = v;  // INTER block then PSUM block
    }
    if (threadIdx.x == 0) *prm.ticket = 0;
  }
}

// Fast variant for compact labels and C % 4 == 0: thread (tx, ty) owns the float4 column tx of the
// rows ty, ty + RY, ...; four rows per iteration with all loads issued first; float partials per
// class flushed to double every iteration.
template <int UNROLL>
__global__ void __launch_bounds__(512) cls_sums_fast_kernel(const __grid_constant__ Segs segs,
                                                           const __grid_constant__ SumsParams prm) {
  extern __shared__ double s_acc[];  // [ry][2C]
  const int C = prm.num_classes;
  const unsigned cv = (unsigned)prm.cv;
  const unsigned tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  const unsigned ry = (unsigned)prm.ry;
  const int c0 = (int)tx * 4;
  double ai[4] = {0, 0, 0, 0}, ap[4] = {0, 0, 0, 0};
  if (ty < ry) {
    const unsigned rstride = gridDim.x * ry;
    for (int s = 0; s < segs.num; ++s) {
      const unsigned rows = (unsigned)segs.rows[s];
      const float4* __restrict__ xin = reinterpret_cast<const float4*>(prm.logits.p[s]);
      const int32_t* __restrict__ lab_p = reinterpret_cast<const int32_t*>(prm.label.p[s]);
      for (unsigned r0 = blockIdx.x * ry + ty; r0 < rows; r0 += UNROLL * rstride) {
        float4 v[UNROLL];
        int lab[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
          const unsigned r = r0 + j * rstride;
          lab[j] = RN_LABEL_IGNORE;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) {
            v[j] = __ldcs(xin + (size_t)r * cv + tx);
            lab[j] = __ldg(lab_p + r);
          }
        }
        float fp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
          if (lab[j] == RN_LABEL_IGNORE) continue;
          const float xs[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float p;
            if (xs[k] <= -0.70f) {
              const float e = ex2_approx(xs[k] * 1.44269502163e+00f);
              p = e * rcp_approx(1.0f + e);
            } else {
              p = sigmoid_f32(xs[k]);
            }
            fp[k] += p;
            if (lab[j] == c0 + k) ai[k] += (double)p;   // at most one class per row
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) ap[k] += (double)fp[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_acc[(size_t)ty * 2 * C + c0 + k] = ai[k];
      s_acc[(size_t)ty * 2 * C + C + c0 + k] = ap[k];
    }
  }
  __syncthreads();
  double* mine = prm.partials + (size_t)blockIdx.x * 2 * C;
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
    double v = 0.0;
    for (unsigned y = 0; y < ry; ++y) v += s_acc[(size_t)y * 2 * C + j];
    mine[j] = v;
  }
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(prm.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(prm.partials + (size_t)b * 2 * C + j);
      prm.stats_out[RN_STATS_HEADER + C + j] = v;  // INTER block then PSUM block
    }
    if (threadIdx.x == 0) *prm.ticket = 0;
  }
}

// Flat variant (the one the pipeline uses): same streaming skeleton as cls_fast_kernel.  The grid
// stride is a multiple of the row length, so a thread's float4 column -- its four classes -- is
// constant and the per-class sums live in registers for the whole kernel; four units per
// iteration with all loads first.  The in-block reduction is ordered (thread order), so the result
// is run-to-run deterministic.
template <int U, int MINB, bool BF16>   // U 16-byte loads in flight per thread, MINB resident CTAs per SM
__global__ void __launch_bounds__(kLossThreads, MINB)
cls_sums_flat_kernel(const __grid_constant__ Segs segs, const __grid_constant__ SumsParams prm) {
  constexpr int K = BF16 ? 8 : 4;      // classes per 16-byte unit (bf16x8 or float4)
  extern __shared__ double s_part[];   // [blockDim][2K]: I[K], P[K] of every thread
  const int C = prm.num_classes;
  const unsigned cv = (unsigned)prm.cv;
  const unsigned gstride = gridDim.x * blockDim.x;   // multiple of cv (host)
  const unsigned drow = gstride / cv;
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned col = gtid % cv;
  const int c0 = (int)col * K;
  // sum of p per class in double; sum of z*p gets a handful of terms per thread (a positive row reaches exactly
  // one of the cv threads that share it), so a float is exact enough and keeps the registers for the loads
  double ap[K];
  float ai[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ap[k] = 0.0;
    ai[k] = 0.f;
  }

  auto prob = [](float x) {
    if (x <= -0.70f) {  // same bulk formula as cls_fast_kernel: p = e / (1 + e), e = exp(x)
      const float e = ex2_approx(x * 1.44269502163e+00f);
      return e * rcp_approx(1.0f + e);
    }
    return sigmoid_f32(x);
  };
  // four sigmoids; confident negatives share ONE reciprocal (1/(abcd) times the other factors; 1 + e is in
  // [1, 1.5], so nothing over- or underflows).  The kernel was limited by the MUFU pipe (exp2 + rcp per logit):
  // this takes it from 8 to 5 MUFU per four logits.
  auto prob4 = [&](const float4& v, float* p) {
    if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) <= -0.70f) {
      const float e0 = ex2_approx(v.x * 1.44269502163e+00f), e1 = ex2_approx(v.y * 1.44269502163e+00f);
      const float e2 = ex2_approx(v.z * 1.44269502163e+00f), e3 = ex2_approx(v.w * 1.44269502163e+00f);
      float i0, i1, i2, i3;
      rcp4(1.0f + e0, 1.0f + e1, 1.0f + e2, 1.0f + e3, i0, i1, i2, i3);
      p[0] = e0 * i0;
      p[1] = e1 * i1;
      p[2] = e2 * i2;
      p[3] = e3 * i3;
    } else {
      p[0] = prob(v.x);
      p[1] = prob(v.y);
      p[2] = prob(v.z);
      p[3] = prob(v.w);
    }
  };

  for (int s = 0; s < segs.num; ++s) {
    const unsigned units = (unsigned)segs.rows[s] * cv;
    const uint4* __restrict__ xin = reinterpret_cast<const uint4*>(prm.logits.p[s]);
    const int32_t* __restrict__ lab_p = reinterpret_cast<const int32_t*>(prm.label.p[s]);
    unsigned u = gtid;
    unsigned row = u / cv;
    float fp[K];   // float partials of sum p, flushed to the doubles once per iteration
#pragma unroll
    for (int k = 0; k < K; ++k) fp[k] = 0.f;

    auto do_unit = [&](const uint4& raw, int lab) {
      if (lab == RN_LABEL_IGNORE) return;
      float p[K];
      if constexpr (BF16) {
        float4 lo, hi;
        unpack_bf16x8(raw, lo, hi);
        prob4(lo, p);
        prob4(hi, p + 4);
      } else {
        prob4(*reinterpret_cast<const float4*>(&raw), p);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) fp[k] += p[k];
      const unsigned d = (unsigned)(lab - c0);
      if (d < (unsigned)K) {  // the row's positive class is one of mine (rare)
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (d == (unsigned)k) ai[k] += p[k];
      }
    };
    auto flush = [&]() {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        ap[k] += (double)fp[k];
        fp[k] = 0.f;
      }
    };

    while (u + (unsigned)(U - 1) * gstride < units) {
      uint4 v[U];
      int l[U];
#pragma unroll
      for (int k = 0; k < U; ++k) v[k] = __ldcs(xin + u + (unsigned)k * gstride);
#pragma unroll
      for (int k = 0; k < U; ++k) l[k] = __ldg(lab_p + row + (unsigned)k * drow);
#pragma unroll
      for (int k = 0; k < U; ++k) do_unit(v[k], l[k]);
      u += (unsigned)U * gstride;
      row += (unsigned)U * drow;
      flush();
    }
    for (; u < units; u += gstride, row += drow) do_unit(__ldcs(xin + u), __ldg(lab_p + row));
    flush();
  }

  // ordered in-block reduction: output j = (kind, class) sums the threads that own that class
#pragma unroll
  for (int k = 0; k < K; ++k) {
    s_part[(size_t)threadIdx.x * 2 * K + k] = (double)ai[k];
    s_part[(size_t)threadIdx.x * 2 * K + K + k] = ap[k];
  }
  __syncthreads();
  double* mine = prm.partials + (size_t)blockIdx.x * 2 * C;
  const unsigned col0 = (blockIdx.x * blockDim.x) % cv;   // column of thread 0
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
    const int kind = j / C, c = j % C;
    const unsigned want = (unsigned)c / (unsigned)K;              // owning column
    unsigned t = (want + cv - col0) % cv;                         // first thread with that column
    double v = 0.0;
    for (; t < blockDim.x; t += cv) v += s_part[(size_t)t * 2 * K + kind * K + (c % K)];
    mine[j] = v;
  }
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(prm.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(prm.partials + (size_t)b * 2 * C + j);
      prm.stats_out[RN_STATS_HEADER + C + j] = v;  // INTER block then PSUM block
    }
    if (threadIdx.x == 0) *prm.ticket = 0;
  }
}

// ---- dense-label statistics (compat API): M = rows, num_fg = #rows with max z > 0.5, n_pos ------
__global__ void dense_label_stats_kernel(long long rows, int C, const float* __restrict__ labels, double* stats) {
  int n_fg = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float z = labels[r * C + c];
      mx = fmaxf(mx, z);
      if (z != 0.0f) atomicAdd(stats + RN_STAT_NPOS(c), (double)z);
    }
    n_fg += mx > 0.5f;  // utils.classmap_decode, utils.py:173-174
  }
  n_fg = warp_sum(n_fg);
  if ((threadIdx.x & 31) == 0 && n_fg) atomicAdd(stats + RN_STAT_NUM_FG, (double)n_fg);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + RN_STAT_TRAINABLE, (double)rows);
}

// ---- Huber (tf.losses.huber_loss, delta = 1, SUM_BY_NONZERO_WEIGHTS; losses.py:144-152) ---------
struct HuberParams {
  Ptrs pred, target, label, fg_mask, dpred;
  const double* stats;
  double* stats_out;
  double* partials;
  unsigned* ticket;
};

__global__ void __launch_bounds__(kLossThreads)
huber_kernel(const __grid_constant__ Segs segs, const __grid_constant__ HuberParams prm) {
  __shared__ double s_red[32];
  const double nfg = prm.stats[RN_STAT_NUM_FG];
  const float gscale = nfg > 0 ? (float)(1.0 / (4.0 * nfg)) : 0.0f;
  const bool by_label = prm.label.p[0] != nullptr;
  const bool write = prm.dpred.p[0] != nullptr;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;

  for (int s = 0; s < segs.num; ++s) {
    const long long rows = segs.rows[s];
    const int32_t* __restrict__ lab_p = reinterpret_cast<const int32_t*>(prm.label.p[s]);
    const unsigned char* __restrict__ fg_p = reinterpret_cast<const unsigned char*>(prm.fg_mask.p[s]);
    const float4* __restrict__ pred = reinterpret_cast<const float4*>(prm.pred.p[s]);
    const float4* __restrict__ targ = reinterpret_cast<const float4*>(prm.target.p[s]);
    float4* __restrict__ out = reinterpret_cast<float4*>(prm.dpred.p[s]);

    auto one_row = [&](long long r, bool fg) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fg) {  // ~1 % of the rows
        const float4 p = __ldg(pred + r);
        const float4 t = __ldg(targ + r);
        const float e[4] = {fsub(p.x, t.x), fsub(p.y, t.y), fsub(p.z, t.z), fsub(p.w, t.w)};
        float ls = 0.f, gg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float ae = fabsf(e[k]);
          const float q = fminf(ae, 1.0f);
          ls += fadd(0.5f * fmul(q, q), fsub(ae, q));  // 0.5 q^2 + delta (|e| - q)
          gg[k] = fminf(fmaxf(e[k], -1.0f), 1.0f) * gscale;
        }
        acc += (double)ls;
        g = make_float4(gg[0], gg[1], gg[2], gg[3]);
      }
      if (write) __stcs(out + r, g);
    };
    auto is_fg = [&](long long r) {
      return by_label ? (__ldg(lab_p + r) >= 0) : (fg_p[r] != 0);
    };

    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // four rows per iteration: the four label loads are issued before the four 16-byte stores
    for (; r + 3 * gstride < rows; r += 4 * gstride) {
      const bool f0 = is_fg(r), f1 = is_fg(r + gstride), f2 = is_fg(r + 2 * gstride), f3 = is_fg(r + 3 * gstride);
      one_row(r, f0);
      one_row(r + gstride, f1);
      one_row(r + 2 * gstride, f2);
      one_row(r + 3 * gstride, f3);
    }
    for (; r < rows; r += gstride) one_row(r, is_fg(r));
  }

  acc = block_sum(acc, s_red);
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    prm.partials[blockIdx.x] = acc;
    __threadfence();
    const unsigned t = atomicAdd(prm.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double v = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(prm.partials + i);
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) {
      prm.stats_out[RN_STAT_HUBER_SUM] = v;
      *prm.ticket = 0;
    }
  }
}

// Same loss for [rows, width] arrays of any width with a u8 foreground mask (the reference's
// regression_loss signature, losses.py:144-152; its own test uses width 1, losses_test.py:17-27).
__global__ void __launch_bounds__(kLossThreads)
huber_rows_kernel(long long rows, int width, const float* __restrict__ pred, const float* __restrict__ target,
                  const unsigned char* __restrict__ fg_mask, float* __restrict__ dpred, const double* stats,
                  float* loss_out, double* partials, unsigned* ticket) {
  __shared__ double s_red[32];
  const long long n = rows * width;
  const double nfg = stats[RN_STAT_NUM_FG];
  const float gscale = nfg > 0 ? (float)(1.0 / ((double)width * nfg)) : 0.0f;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (fg_mask[i / width]) {
      const float e = fsub(pred[i], target[i]);
      const float ae = fabsf(e);
      const float q = fminf(ae, 1.0f);
      acc += (double)fadd(0.5f * fmul(q, q), fsub(ae, q));
      g = fminf(fmaxf(e, -1.0f), 1.0f) * gscale;
    }
    if (dpred) dpred[i] = g;
  }
  acc = block_sum(acc, s_red);
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double v = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(partials + i);
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) {
      *loss_out = nfg > 0 ? (float)(v / ((double)width * nfg)) : 0.0f;  // SUM_BY_NONZERO_WEIGHTS
      *ticket = 0;
    }
  }
}

__global__ void loss_finalize_kernel(int C, int mode, const double* __restrict__ stats, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double M = stats[RN_STAT_TRAINABLE], nfg = stats[RN_STAT_NUM_FG];
  const int base = mode & 15;
  double cls = 0.0;
  if (base == RN_CLS_FOCAL) cls += stats[RN_STAT_CLS_SUM] / fmax(nfg, 1.0);
  if (base == RN_CLS_BCE || base == RN_CLS_BALANCED_BCE) cls += M > 0 ? stats[RN_STAT_CLS_SUM] / (M * (double)C) : 0.0;
  if (mode & RN_CLS_DICE_FLAG) {
    double d = 0.0;
    for (int c = 0; c < C; ++c) {
      const double I = stats[RN_STAT_INTER(C, c)];
      const double U = stats[RN_STAT_NPOS(c)] + stats[RN_STAT_PSUM(C, c)];
      d += 1.0 - 2.0 * I / U;
    }
    cls += d / (double)C;
  }
  out[0] = (float)cls;
  out[1] = nfg > 0 ? (float)(stats[RN_STAT_HUBER_SUM] / (4.0 * nfg)) : 0.0f;
  out[2] = (float)nfg;
  out[3] = (float)M;
}

// ---- elementwise, unreduced (the reference functions' own return shape) -------------------------
__global__ void cls_elementwise_kernel(long long n, int C, const float* __restrict__ labels,
                                       const float* __restrict__ logits, int mode, LossScalars sc,
                                       const float* __restrict__ w_pos, const float* __restrict__ w_neg,
                                       float* __restrict__ loss_out, float* __restrict__ grad_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = logits[i], z = labels[i];
    float loss = 0.f, grad = 0.f;
    if (mode == RN_CLS_FOCAL) {
      focal_elem<true>(x, z == 1.0f, sc, loss, grad);
    } else {
      float p;
      bce_elem<true>(x, z, loss, p, grad);
      if (mode == RN_CLS_BALANCED_BCE && w_pos) {
        const int c = (int)(i % C);
        const float w = (z == 1.0f) ? w_pos[c] : w_neg[c];
        loss *= w;
        grad *= w;
      }
    }
    if (loss_out) loss_out[i] = loss;
    if (grad_out) grad_out[i] = grad;
  }
}

// Backward of any loss of the per-class sums I_c = sum z p, P_c = sum p (dice / jaccard /
// fixed-IoU, losses.py:37-73):  d/dx[m,c] = (z * a[c] + b[c]) * p (1 - p).
__global__ void colsum_backward_kernel(long long n, int C, const float* __restrict__ labels,
                                       const float* __restrict__ logits, const float* __restrict__ a,
                                       const float* __restrict__ b, float* __restrict__ grad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float p = sigmoid_f32(logits[i]);
    grad[i] = (labels[i] * a[c] + b[c]) * (p * (1.0f - p));
  }
}

// losses.focal_softmax_cross_entropy_with_logits (losses.py:19-34): one warp per row.
//   prob = softmax(x);  a = z==1 ? alpha : 1-alpha;  pt = z==1 ? prob : 1-prob
//   loss[row] = -sum_c a * (1-pt)^focus * z * log(prob + eps)
// grad[row, j] = p_j (D_j - sum_c D_c p_c) with D_c = d loss / d prob_c.
__global__ void focal_softmax_kernel(long long rows, int C, const float* __restrict__ labels,
                                     const float* __restrict__ logits, float focus, float alpha, float eps,
                                     float* __restrict__ loss_out, float* __restrict__ grad_out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* x = logits + r * C;
    const float* z = labels + r * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
    se = warp_sum(se);
    float loss = 0.f, dot = 0.f;   // dot = sum_c D_c p_c
    for (int c = lane; c < C; c += 32) {
      const float p = fdiv(expf(x[c] - mx), se);
      const float zc = z[c];
      const bool is1 = zc == 1.0f;
      const float a = is1 ? alpha : 1.0f - alpha;
      const float pt = is1 ? p : 1.0f - p;
      const float w = 1.0f - pt;
      const float pe = p + eps;
      const float lg = logf(pe);
      const float mod = powf(w, focus);
      loss -= a * mod * zc * lg;
      const float dmod = focus * powf(w, focus - 1.0f) * (is1 ? -1.0f : 1.0f);   // d (1-pt)^f / d p
      const float D = -a * zc * (dmod * lg + mod / pe);
      dot += D * p;
    }
    loss = warp_sum(loss);
    dot = warp_sum(dot);
    if (lane == 0 && loss_out) loss_out[r] = loss;
    if (grad_out) {
      for (int c = lane; c < C; c += 32) {
        const float p = fdiv(expf(x[c] - mx), se);
        const float zc = z[c];
        const bool is1 = zc == 1.0f;
        const float a = is1 ? alpha : 1.0f - alpha;
        const float pt = is1 ? p : 1.0f - p;
        const float w = 1.0f - pt;
        const float pe = p + eps;
        const float dmod = focus * powf(w, focus - 1.0f) * (is1 ? -1.0f : 1.0f);
        const float D = -a * zc * (dmod * logf(pe) + powf(w, focus) / pe);
        grad_out[r * C + c] = p * (D - dot);
      }
    }
  }
}

__global__ void sigmoid_kernel(long long n, const float* __restrict__ in, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = sigmoid_f32(in[i]);
}


// Persistent grid: exactly the number of CTAs that are resident at once (one full wave), so a
// grid-stride loop never runs a partially filled second wave.
template <typename K>
static unsigned persistent_grid(K kernel, int threads, size_t smem, long long work_items) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  long long blocks = (long long)num_sms() * per_sm;
  if (blocks > kMaxLossBlocks) blocks = kMaxLossBlocks;
  const long long need = (work_items + threads - 1) / threads;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static int launch_fast(const Segs& segs, const FastParams& fp, int mode, bool select, long long units,
                             cudaStream_t st);

static unsigned gcd_u(unsigned a, unsigned b) {
  while (b) {
    const unsigned t = a % b;
    a = b;
    b = t;
  }
  return a;
}

static unsigned loss_grid(long long units, int threads, int per_sm) {
  long long blocks = (units + threads - 1) / threads;
  long long cap = (long long)num_sms() * per_sm;
  if (cap > kMaxLossBlocks) cap = kMaxLossBlocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static bool aligned16(const rn_maps_t* m, int n) {
  for (int i = 0; i < n; ++i)
    if (((uintptr_t)m->p[i]) & 15) return false;
  return true;
}

template <int MODE, bool DICE, bool DENSE>
static void launch_cls(bool vec4, unsigned grid, size_t smem, cudaStream_t st, const Segs& segs, const ClsParams& prm) {
  if (vec4)
    cls_loss_kernel<MODE, DICE, DENSE, 4><<<grid, kLossThreads, smem, st>>>(segs, prm);
  else
    cls_loss_kernel<MODE, DICE, DENSE, 1><<<grid, kLossThreads, smem, st>>>(segs, prm);
}

template <int MODE, bool DICE>
static void launch_cls2(bool dense, bool vec4, unsigned grid, size_t smem, cudaStream_t st, const Segs& segs,
                        const ClsParams& prm) {
  if (dense)
    launch_cls<MODE, DICE, true>(vec4, grid, smem, st, segs, prm);
  else
    launch_cls<MODE, DICE, false>(vec4, grid, smem, st, segs, prm);
}

// The streaming (tile) kernel for compact labels; `units` is unused (kept for the call sites' symmetry).
static int launch_fast(const Segs& segs, const FastParams& fp, int mode, bool select, long long units,
                             cudaStream_t st) {
  (void)units;
  TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.logits = fp.logits;
  tp.label = fp.label;
  tp.dlogits = fp.dlogits;
  tp.alpha = fp.alpha;
  tp.eps = fp.eps;
  tp.stats = fp.stats;
  tp.stats_out = fp.stats_out;
  tp.workspace = fp.ticket;
  tp.num_classes = fp.num_classes;
  tp.bf16 = fp.bf16;
  tp.write_grad = fp.write_grad;
  tp.sel_trigger = fp.sel_trigger;
  tp.sel_thr = fp.sel_thr;
  tp.sel_images = fp.sel_images;
  tp.sel_capacity = fp.sel_capacity;
  tp.sel_anchor = fp.sel_anchor;
  tp.sel_score = fp.sel_score;
  tp.sel_class = fp.sel_class;
  tp.sel_count = fp.sel_count;
  tp.status = fp.status;
  return launch_cls_tile(segs, tp, mode & ~RN_CLS_GENERIC_FLAG, select, st);
}

}  // namespace rn

using namespace rn;

extern "C" int rn_set_option(const char* name, int value) {
  RN_REQUIRE(name, RN_ERR_BAD_ARG, "rn_set_option: NULL name");
  if (strcmp(name, "exchange_timeout_ms") == 0) {
    rn::set_exchange_timeout_ms(value);
    return RN_OK;
  }
  rn::set_error("rn_set_option: unknown option '%s'", name);
  return RN_ERR_BAD_ARG;
}

extern "C" size_t rn_loss_workspace_bytes(int num_classes) {
  const size_t per_block = (size_t)(2 * (num_classes > 0 ? num_classes : 1)) * sizeof(double);
  return 256 + (size_t)kMaxLossBlocks * per_block;
}

static int check_rows(const rn_rows_t* rows, const char* who) {
  RN_REQUIRE(rows, RN_ERR_BAD_ARG, "%s: NULL rows", who);
  RN_REQUIRE(rows->num_segments >= 1 && rows->num_segments <= RN_MAX_LEVELS, RN_ERR_BAD_SHAPE,
             "%s: num_segments must be in 1..%d", who, RN_MAX_LEVELS);
  for (int i = 0; i < rows->num_segments; ++i) RN_REQUIRE(rows->rows[i] >= 0, RN_ERR_BAD_SHAPE, "%s: negative rows", who);
  return RN_OK;
}

extern "C" int rn_cls_sums(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits, const rn_maps_t* label,
                           int dense_labels, int logits_dtype, double* d_stats, void* d_workspace,
                           size_t workspace_bytes, void* stream) {
  int rc = check_rows(rows, "rn_cls_sums");
  if (rc) return rc;
  RN_REQUIRE(maps_present(logits) && maps_present(label) && d_stats && d_workspace, RN_ERR_BAD_ARG,
             "rn_cls_sums: NULL input");
  RN_REQUIRE(num_classes >= 1, RN_ERR_BAD_ARG, "rn_cls_sums: num_classes must be >= 1");
  RN_REQUIRE(workspace_bytes >= rn_loss_workspace_bytes(num_classes), RN_ERR_CAPACITY, "rn_cls_sums: workspace too small");
  const Segs segs = make_segs(rows);
  cudaStream_t st = (cudaStream_t)stream;
  RN_REQUIRE(logits_dtype == RN_DTYPE_F32 || logits_dtype == RN_DTYPE_BF16, RN_ERR_BAD_ARG,
             "rn_cls_sums: unknown logits_dtype");
  const bool bf16 = logits_dtype == RN_DTYPE_BF16;
  const bool vec4 = (num_classes % 4 == 0) && aligned16(logits, rows->num_segments) &&
                    (!dense_labels || aligned16(label, rows->num_segments));
  const int vec = bf16 ? 8 : (vec4 ? 4 : 1);
  RN_REQUIRE(!bf16 || (!dense_labels && num_classes % 8 == 0 && aligned16(logits, rows->num_segments)),
             RN_ERR_UNSUPPORTED,
             "rn_cls_sums: bf16 logits need compact labels, num_classes %% 8 == 0 and 16-byte aligned maps");
  SumsParams prm;
  prm.logits = make_ptrs(logits);
  prm.label = make_ptrs(label);
  prm.stats_out = d_stats;
  prm.ticket = reinterpret_cast<unsigned*>(d_workspace);
  prm.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + 256);
  prm.num_classes = num_classes;
  prm.cv = num_classes / vec;
  RN_REQUIRE(prm.cv <= 1024, RN_ERR_UNSUPPORTED, "rn_cls_sums: too many classes");
  prm.ry = 256 / prm.cv > 0 ? 256 / prm.cv : 1;
  const int threads = prm.cv * prm.ry;
  const size_t smem = (size_t)prm.ry * 2 * num_classes * sizeof(double);
  RN_REQUIRE(smem <= 200 * 1024, RN_ERR_UNSUPPORTED, "rn_cls_sums: too many classes for shared memory");
  const long long total_rows = segs.start[RN_MAX_LEVELS];
  cudaMemsetAsync(prm.ticket, 0, sizeof(unsigned), st);
  bool rows32 = true;
  for (int i = 0; i < rows->num_segments; ++i)
    if (rows->rows[i] >= (1ll << 31) - (1ll << 26)) rows32 = false;
  bool units32 = true;
  for (int i = 0; i < rows->num_segments; ++i)
    if ((unsigned long long)rows->rows[i] * (unsigned)prm.cv + (1ull << 24) >= (1ull << 32)) units32 = false;
  if (!dense_labels && (vec4 || bf16) && units32 && prm.cv <= kLossThreads) {
    // six float4 loads in flight per thread at 4 resident CTAs per SM measured best (4: 0.90 ms, 6: 0.85 ms,
    // 8 with 3 CTAs: 0.95 ms, 8 with 2 CTAs: 1.24 ms at COCO batch 64): the kernel is bound by resident warps, not
    // by bytes in flight.  bf16x8 units carry twice the classes (and accumulators): four loads in flight.
#define RN_SUMS_FLAT(U, BF)                                                                                      \
    do {                                                                                                           \
      const size_t fsmem = (size_t)kLossThreads * (BF ? 16 : 8) * sizeof(double);                                  \
      unsigned g = persistent_grid(cls_sums_flat_kernel<U, 4, BF>, kLossThreads, fsmem,                            \
                                   (total_rows * prm.cv + 3) / 4);                                                 \
      const unsigned q = (unsigned)prm.cv / gcd_u((unsigned)prm.cv, (unsigned)kLossThreads);                       \
      unsigned gf = g / q * q;                                                                                     \
      if (BF && gf == 0) gf = q;   /* bf16 has no other kernel: a few idle CTAs instead of a fallback */            \
      if (gf > 0 && (BF || (g - gf) * 25u <= g)) {                                                                 \
        cls_sums_flat_kernel<U, 4, BF><<<gf, kLossThreads, fsmem, st>>>(segs, prm);                                \
        RN_CUDA_CHECK_LAUNCH("rn_cls_sums(flat)");                                                                 \
        return RN_OK;                                                                                              \
      }                                                                                                            \
    } while (0)
    if (bf16) RN_SUMS_FLAT(4, true); else RN_SUMS_FLAT(6, false);
#undef RN_SUMS_FLAT
  }
  RN_REQUIRE(!bf16, RN_ERR_UNSUPPORTED, "rn_cls_sums: bf16 logits: no launch configuration for this shape");
  if (!dense_labels && vec4 && rows32 && prm.cv <= 256) {
    prm.ry = 256 / prm.cv > 0 ? 256 / prm.cv : 1;
    const int fthreads = prm.cv * prm.ry;
    const size_t fsmem = (size_t)prm.ry * 2 * num_classes * sizeof(double);
    if (fsmem <= 48 * 1024) {
      int per_sm = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cls_sums_fast_kernel<4>, fthreads, fsmem);
      long long fgrid = (long long)num_sms() * (per_sm > 0 ? per_sm : 1);
      const long long need = (total_rows + prm.ry - 1) / prm.ry;
      if (fgrid > need) fgrid = need;
      if (fgrid > kMaxLossBlocks) fgrid = kMaxLossBlocks;
      if (fgrid < 1) fgrid = 1;
      cls_sums_fast_kernel<4><<<(unsigned)fgrid, fthreads, fsmem, st>>>(segs, prm);
      RN_CUDA_CHECK_LAUNCH("rn_cls_sums(fast)");
      return RN_OK;
    }
  }
  unsigned grid = loss_grid((total_rows + prm.ry - 1) / prm.ry, 1, 4);
#define RN_LAUNCH_SUMS(DENSE, VEC)                                                                             \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      cudaFuncSetAttribute(cls_sums_kernel<DENSE, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    cls_sums_kernel<DENSE, VEC><<<grid, threads, smem, st>>>(segs, prm);                                       \
  } while (0)
  if (dense_labels) {
    if (vec4) RN_LAUNCH_SUMS(true, 4); else RN_LAUNCH_SUMS(true, 1);
  } else {
    if (vec4) RN_LAUNCH_SUMS(false, 4); else RN_LAUNCH_SUMS(false, 1);
  }
#undef RN_LAUNCH_SUMS
  RN_CUDA_CHECK_LAUNCH("rn_cls_sums");
  return RN_OK;
}

extern "C" int rn_cls_loss_fwd_bwd(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits,
                                   const rn_maps_t* label, int dense_labels, int mode, float focus, float alpha,
                                   float eps, const rn_maps_t* dlogits, int logits_dtype, double* d_stats,
                                   void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_rows(rows, "rn_cls_loss_fwd_bwd");
  if (rc) return rc;
  RN_REQUIRE(maps_present(logits) && maps_present(label) && d_stats && d_workspace, RN_ERR_BAD_ARG,
             "rn_cls_loss_fwd_bwd: NULL input");
  RN_REQUIRE(num_classes >= 1, RN_ERR_BAD_ARG, "rn_cls_loss_fwd_bwd: num_classes must be >= 1");
  RN_REQUIRE(workspace_bytes >= rn_loss_workspace_bytes(num_classes), RN_ERR_CAPACITY,
             "rn_cls_loss_fwd_bwd: workspace too small");
  const int base = mode & 15;
  const bool dice = (mode & RN_CLS_DICE_FLAG) != 0;
  RN_REQUIRE(base >= RN_CLS_NONE && base <= RN_CLS_BALANCED_BCE && (mode & ~(15 | RN_CLS_DICE_FLAG | RN_CLS_GENERIC_FLAG)) == 0,
             RN_ERR_BAD_ARG, "rn_cls_loss_fwd_bwd: unknown mode %d", mode);
  const bool generic = (mode & RN_CLS_GENERIC_FLAG) != 0;
  RN_REQUIRE(base != RN_CLS_NONE || dice, RN_ERR_BAD_ARG, "rn_cls_loss_fwd_bwd: empty mode");
  const Segs segs = make_segs(rows);
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = maps_present(dlogits);
  const bool vec4 = (num_classes % 4 == 0) && aligned16(logits, rows->num_segments) &&
                    (!grad || aligned16(dlogits, rows->num_segments)) &&
                    (!dense_labels || aligned16(label, rows->num_segments));
  ClsParams prm;
  prm.logits = make_ptrs(logits);
  prm.label = make_ptrs(label);
  prm.dlogits = make_ptrs(grad ? dlogits : nullptr);
  prm.sc.focus = focus;
  prm.sc.alpha = alpha;
  prm.sc.one_minus_alpha = 1.0f - alpha;
  prm.sc.eps = eps;
  prm.stats = d_stats;
  prm.stats_out = d_stats;
  prm.ticket = reinterpret_cast<unsigned*>(d_workspace);
  prm.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + 256);
  prm.num_classes = num_classes;
  prm.cv = (unsigned)(num_classes / (vec4 ? 4 : 1));
  prm.write_grad = grad ? 1 : 0;
  const long long units = segs.start[RN_MAX_LEVELS] * (long long)prm.cv;
  const unsigned grid = loss_grid(units, kLossThreads, 8);
  const size_t smem = 4 * (size_t)num_classes * sizeof(float);
  RN_REQUIRE(smem <= 48 * 1024, RN_ERR_UNSUPPORTED, "rn_cls_loss_fwd_bwd: too many classes");
  cudaMemsetAsync(prm.ticket, 0, 4 * sizeof(unsigned), st);
  const bool dense = dense_labels != 0;
  RN_REQUIRE(logits_dtype == RN_DTYPE_F32 || logits_dtype == RN_DTYPE_BF16, RN_ERR_BAD_ARG, "unknown logits_dtype");
  const bool bf16 = logits_dtype == RN_DTYPE_BF16;
  bool fits32 = true;  // the fast kernel indexes float4 units with 32-bit integers
  for (int i = 0; i < rows->num_segments; ++i)
    if ((unsigned long long)rows->rows[i] * (num_classes / 4 > 0 ? num_classes / 4 : 1) + (1ull << 24) >= (1ull << 32)) fits32 = false;
  const bool fast_ok = (base != RN_CLS_FOCAL || focus == 2.0f) && !dense && vec4 && !generic && fits32 &&
                       num_classes <= 2048;
  if (dense && !bf16 && vec4 && !generic && (base != RN_CLS_FOCAL || focus == 2.0f) &&
      cls_tile_supported(num_classes, false, true)) {
    // dense float labels [rows, C] (the reference's own loss signatures, losses.py:115-141): the streaming
    // kernel with a second tile of labels per stage
    TileParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.logits = prm.logits;
    tp.label = prm.label;
    tp.dlogits = prm.dlogits;
    tp.alpha = alpha;
    tp.eps = eps;
    tp.stats = d_stats;
    tp.stats_out = d_stats;
    tp.workspace = d_workspace;
    tp.num_classes = num_classes;
    tp.dense = 1;
    tp.write_grad = prm.write_grad;
    return launch_cls_tile(segs, tp, mode, false, st);
  }
  RN_REQUIRE(!bf16 || (fast_ok && (dice || base != RN_CLS_NONE) && num_classes % 8 == 0), RN_ERR_UNSUPPORTED,
             "bfloat16 logits need the fast kernel: compact labels, C %% 8 == 0, focal (focus 2) / BCE / balanced BCE "
             "and/or dice");
  if (fast_ok) {
    FastParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.logits = prm.logits;
    fp.label = prm.label;
    fp.dlogits = prm.dlogits;
    fp.alpha = alpha;
    fp.eps = eps;
    fp.stats = d_stats;
    fp.stats_out = d_stats;
    fp.partials = prm.partials;
    fp.ticket = prm.ticket;
    fp.num_classes = num_classes;
    fp.cv = num_classes / (bf16 ? 8 : 4);
    fp.bf16 = bf16 ? 1 : 0;
    fp.write_grad = prm.write_grad;
    return launch_fast(segs, fp, mode, false, segs.start[RN_MAX_LEVELS] * (long long)fp.cv, st);
  }
  switch (base) {
    case RN_CLS_FOCAL:
      if (dice) launch_cls2<RN_CLS_FOCAL, true>(dense, vec4, grid, smem, st, segs, prm);
      else launch_cls2<RN_CLS_FOCAL, false>(dense, vec4, grid, smem, st, segs, prm);
      break;
    case RN_CLS_BCE:
      if (dice) launch_cls2<RN_CLS_BCE, true>(dense, vec4, grid, smem, st, segs, prm);
      else launch_cls2<RN_CLS_BCE, false>(dense, vec4, grid, smem, st, segs, prm);
      break;
    case RN_CLS_BALANCED_BCE:
      if (dice) launch_cls2<RN_CLS_BALANCED_BCE, true>(dense, vec4, grid, smem, st, segs, prm);
      else launch_cls2<RN_CLS_BALANCED_BCE, false>(dense, vec4, grid, smem, st, segs, prm);
      break;
    default:
      launch_cls2<RN_CLS_NONE, true>(dense, vec4, grid, smem, st, segs, prm);
      break;
  }
  RN_CUDA_CHECK_LAUNCH("rn_cls_loss_fwd_bwd");
  return RN_OK;
}

extern "C" int rn_fused_select_supported(int num_classes, int mode, float focus, const rn_rows_t* rows) {
  const int base = mode & 15;
  if (!rows || num_classes % 4 != 0 || num_classes > 2048 || (base == RN_CLS_FOCAL && focus != 2.0f) ||
      (base == RN_CLS_NONE && !(mode & RN_CLS_DICE_FLAG)) || (mode & RN_CLS_GENERIC_FLAG))
    return 0;
  for (int i = 0; i < rows->num_segments; ++i)
    if ((unsigned long long)rows->rows[i] * (num_classes / 4) + (1ull << 24) >= (1ull << 32)) return 0;
  return 1;
}

extern "C" int rn_cls_loss_select(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits,
                                  const rn_maps_t* label, int mode, float focus, float alpha, float eps,
                                  const rn_maps_t* dlogits, int logits_dtype, double* d_stats, void* d_loss_workspace,
                                  size_t loss_workspace_bytes, float score_thr, int capacity, int32_t* d_cand_count,
                                  uint32_t* d_status, void* d_select_workspace, size_t select_workspace_bytes,
                                  void* stream) {
  int rc = check_rows(rows, "rn_cls_loss_select");
  if (rc) return rc;
  RN_REQUIRE(maps_present(logits) && maps_present(label) && d_stats && d_loss_workspace && d_select_workspace &&
                 d_cand_count,
             RN_ERR_BAD_ARG, "rn_cls_loss_select: NULL input");
  RN_REQUIRE(rn_fused_select_supported(num_classes, mode, focus, rows), RN_ERR_UNSUPPORTED,
             "rn_cls_loss_select: needs focus == 2 (focal), C %% 4 == 0 and < 2^32 float4 per level; "
             "use rn_cls_loss_fwd_bwd + rn_decode_select");
  RN_REQUIRE(rows->images >= 1 && capacity >= 1, RN_ERR_BAD_ARG, "rn_cls_loss_select: bad images/capacity");
  RN_REQUIRE(score_thr > 0.f && score_thr < 1.f, RN_ERR_BAD_ARG, "rn_cls_loss_select: score_thr must be in (0, 1)");
  RN_REQUIRE(loss_workspace_bytes >= rn_loss_workspace_bytes(num_classes), RN_ERR_CAPACITY,
             "rn_cls_loss_select: loss workspace too small");
  RN_REQUIRE(select_workspace_bytes >= rn_select_workspace_bytes(rows->images, capacity), RN_ERR_CAPACITY,
             "rn_cls_loss_select: select workspace too small");
  const bool grad = maps_present(dlogits);
  RN_REQUIRE(aligned16(logits, rows->num_segments) && (!grad || aligned16(dlogits, rows->num_segments)),
             RN_ERR_BAD_ARG, "rn_cls_loss_select: maps must be 16-byte aligned");
  for (int i = 0; i < rows->num_segments; ++i)
    RN_REQUIRE(rows->rows[i] % rows->images == 0, RN_ERR_BAD_SHAPE, "rows[%d] is not a multiple of images", i);
  const Segs segs = make_segs(rows);
  cudaStream_t st = (cudaStream_t)stream;
  FastParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.logits = make_ptrs(logits);
  fp.label = make_ptrs(label);
  fp.dlogits = make_ptrs(grad ? dlogits : nullptr);
  fp.alpha = alpha;
  fp.eps = eps;
  fp.stats = d_stats;
  fp.stats_out = d_stats;
  fp.ticket = reinterpret_cast<unsigned*>(d_loss_workspace);
  fp.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_loss_workspace) + 256);
  RN_REQUIRE(logits_dtype == RN_DTYPE_F32 || logits_dtype == RN_DTYPE_BF16, RN_ERR_BAD_ARG, "unknown logits_dtype");
  const bool bf16 = logits_dtype == RN_DTYPE_BF16;
  RN_REQUIRE(!bf16 || num_classes % 8 == 0,
             RN_ERR_UNSUPPORTED, "bfloat16 logits need C %% 8 == 0");
  fp.num_classes = num_classes;
  fp.cv = num_classes / (bf16 ? 8 : 4);
  fp.bf16 = bf16 ? 1 : 0;
  fp.write_grad = grad ? 1 : 0;
  // sigmoid(x) > thr needs x > logit(thr); leave a margin for the float32 sigmoid's rounding
  const double lg = std::log((double)score_thr / (1.0 - (double)score_thr));
  fp.sel_trigger = (float)(lg - 1e-3 * (1.0 + std::fabs(lg)));
  fp.sel_thr = score_thr;
  fp.sel_images = rows->images;
  fp.sel_capacity = capacity;
  const size_t n = (size_t)rows->images * capacity;
  const size_t chunk = (n * 4 + 255) & ~(size_t)255;  // same carving as rn_decode_select
  char* ws = (char*)d_select_workspace;
  fp.sel_anchor = (int32_t*)ws;
  fp.sel_score = (float*)(ws + chunk);
  fp.sel_class = (int32_t*)(ws + 2 * chunk);
  fp.sel_count = d_cand_count;
  fp.status = d_status;
  cudaMemsetAsync(fp.ticket, 0, 4 * sizeof(unsigned), st);
  cudaMemsetAsync(d_cand_count, 0, sizeof(int32_t) * rows->images, st);
  const long long units = segs.start[RN_MAX_LEVELS] * (long long)fp.cv;
  return launch_fast(segs, fp, mode, true, units, st);
}

// One launch for the whole loss stage of a training step: classification loss forward + backward, Huber
// regression loss forward + backward, optional candidate pre-selection (capacity > 0) and -- when d_loss_out is
// given (single GPU: no statistics exchange in between) -- the final scalars.
extern "C" int rn_head_loss(const rn_rows_t* rows, int num_classes, const rn_maps_t* logits, const rn_maps_t* label,
                            int mode, float focus, float alpha, float eps, const rn_maps_t* dlogits, int logits_dtype,
                            const rn_maps_t* reg_pred, const rn_maps_t* reg_target, const rn_maps_t* dreg,
                            double* d_stats, double* d_stats_next, void* d_loss_workspace, size_t loss_workspace_bytes,
                            float* d_loss_out, float score_thr, int capacity, int32_t* d_cand_count, uint32_t* d_status,
                            void* d_select_workspace, size_t select_workspace_bytes, void* stream) {
  int rc = check_rows(rows, "rn_head_loss");
  if (rc) return rc;
  RN_REQUIRE(maps_present(logits) && maps_present(label) && maps_present(reg_pred) && maps_present(reg_target) &&
                 d_stats && d_loss_workspace,
             RN_ERR_BAD_ARG, "rn_head_loss: NULL input");
  RN_REQUIRE(num_classes >= 1, RN_ERR_BAD_ARG, "rn_head_loss: num_classes must be >= 1");
  RN_REQUIRE(loss_workspace_bytes >= rn_loss_workspace_bytes(num_classes), RN_ERR_CAPACITY,
             "rn_head_loss: loss workspace too small");
  const int base = mode & 15;
  const bool dice = (mode & RN_CLS_DICE_FLAG) != 0;
  RN_REQUIRE(base >= RN_CLS_NONE && base <= RN_CLS_BALANCED_BCE && (mode & ~(15 | RN_CLS_DICE_FLAG | RN_CLS_GENERIC_FLAG)) == 0 &&
                 (base != RN_CLS_NONE || dice),
             RN_ERR_BAD_ARG, "rn_head_loss: unknown mode %d", mode);
  RN_REQUIRE(logits_dtype == RN_DTYPE_F32 || logits_dtype == RN_DTYPE_BF16, RN_ERR_BAD_ARG, "unknown logits_dtype");
  const bool bf16 = logits_dtype == RN_DTYPE_BF16;
  const bool select = capacity > 0;
  const bool grad = maps_present(dlogits);
  const bool aligned = aligned16(logits, rows->num_segments) && (!grad || aligned16(dlogits, rows->num_segments)) &&
                       aligned16(reg_pred, rows->num_segments) && aligned16(reg_target, rows->num_segments) &&
                       (!maps_present(dreg) || aligned16(dreg, rows->num_segments));
  const bool tile_ok = !(mode & RN_CLS_GENERIC_FLAG) && aligned && (base != RN_CLS_FOCAL || focus == 2.0f) &&
                       cls_tile_supported(num_classes, bf16, false);
  if (!tile_ok) {   // separate launches (generic kernels); the fused selection needs the tile kernel
    RN_REQUIRE(!select, RN_ERR_UNSUPPORTED,
               "rn_head_loss: the fused candidate selection needs focus == 2 (focal), C %% 4 == 0 and 16-byte aligned "
               "maps; use rn_decode_select");
    rc = rn_cls_loss_fwd_bwd(rows, num_classes, logits, label, 0, mode, focus, alpha, eps, dlogits, logits_dtype, d_stats,
                             d_loss_workspace, loss_workspace_bytes, stream);
    if (rc) return rc;
    rc = rn_huber_fwd_bwd(rows, reg_pred, reg_target, label, nullptr, dreg, d_stats, d_loss_workspace,
                          loss_workspace_bytes, stream);
    if (rc) return rc;
    if (d_stats_next)
      cudaMemsetAsync(d_stats_next, 0, sizeof(double) * (RN_STATS_HEADER + 3 * (size_t)num_classes), (cudaStream_t)stream);
    return d_loss_out ? rn_loss_finalize(num_classes, mode, d_stats, d_loss_out, stream) : RN_OK;
  }
  const Segs segs = make_segs(rows);
  cudaStream_t st = (cudaStream_t)stream;
  TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.logits = make_ptrs(logits);
  tp.label = make_ptrs(label);
  tp.dlogits = make_ptrs(grad ? dlogits : nullptr);
  tp.pred = make_ptrs(reg_pred);
  tp.target = make_ptrs(reg_target);
  tp.dpred = make_ptrs(maps_present(dreg) ? dreg : nullptr);
  tp.alpha = alpha;
  tp.eps = eps;
  tp.stats = d_stats;
  tp.stats_out = d_stats;
  tp.workspace = d_loss_workspace;   // counters are left at zero by every kernel that uses them
  tp.num_classes = num_classes;
  tp.bf16 = bf16 ? 1 : 0;
  tp.write_grad = grad ? 1 : 0;
  tp.loss_out = d_loss_out;
  tp.final_mode = mode;
  tp.stats_next = d_stats_next;
  if (select) {
    RN_REQUIRE(d_select_workspace && d_cand_count && rows->images >= 1, RN_ERR_BAD_ARG, "rn_head_loss: NULL select output");
    RN_REQUIRE(score_thr > 0.f && score_thr < 1.f, RN_ERR_BAD_ARG, "rn_head_loss: score_thr must be in (0, 1)");
    RN_REQUIRE(select_workspace_bytes >= rn_select_workspace_bytes(rows->images, capacity), RN_ERR_CAPACITY,
               "rn_head_loss: select workspace too small");
    for (int i = 0; i < rows->num_segments; ++i)
      RN_REQUIRE(rows->rows[i] % rows->images == 0, RN_ERR_BAD_SHAPE, "rows[%d] is not a multiple of images", i);
    const double lg = std::log((double)score_thr / (1.0 - (double)score_thr));
    tp.sel_trigger = (float)(lg - 1e-3 * (1.0 + std::fabs(lg)));
    tp.sel_thr = score_thr;
    tp.sel_images = rows->images;
    tp.sel_capacity = capacity;
    const size_t n = (size_t)rows->images * capacity;
    const size_t chunk = (n * 4 + 255) & ~(size_t)255;  // same carving as rn_decode_select
    char* ws = (char*)d_select_workspace;
    tp.sel_anchor = (int32_t*)ws;
    tp.sel_score = (float*)(ws + chunk);
    tp.sel_class = (int32_t*)(ws + 2 * chunk);
    tp.sel_count = d_cand_count;
    tp.status = d_status;
    cudaMemsetAsync(d_cand_count, 0, sizeof(int32_t) * rows->images, st);
  }
  return launch_cls_tile(segs, tp, mode, select, st);
}

extern "C" int rn_huber_fwd_bwd(const rn_rows_t* rows, const rn_maps_t* pred, const rn_maps_t* target,
                                const rn_maps_t* label, const rn_maps_t* fg_mask, const rn_maps_t* dpred,
                                double* d_stats, void* d_workspace, size_t workspace_bytes, void* stream) {
  int rc = check_rows(rows, "rn_huber_fwd_bwd");
  if (rc) return rc;
  RN_REQUIRE(maps_present(pred) && maps_present(target) && d_stats && d_workspace, RN_ERR_BAD_ARG,
             "rn_huber_fwd_bwd: NULL input");
  RN_REQUIRE(maps_present(label) || maps_present(fg_mask), RN_ERR_BAD_ARG, "rn_huber_fwd_bwd: need label or fg_mask");
  RN_REQUIRE(workspace_bytes >= rn_loss_workspace_bytes(1), RN_ERR_CAPACITY, "rn_huber_fwd_bwd: workspace too small");
  RN_REQUIRE(aligned16(pred, rows->num_segments) && aligned16(target, rows->num_segments) &&
                 (!maps_present(dpred) || aligned16(dpred, rows->num_segments)),
             RN_ERR_BAD_ARG, "rn_huber_fwd_bwd: maps must be 16-byte aligned");
  const Segs segs = make_segs(rows);
  cudaStream_t st = (cudaStream_t)stream;
  HuberParams prm;
  prm.pred = make_ptrs(pred);
  prm.target = make_ptrs(target);
  prm.label = make_ptrs(maps_present(label) ? label : nullptr);
  prm.fg_mask = make_ptrs(maps_present(fg_mask) ? fg_mask : nullptr);
  prm.dpred = make_ptrs(maps_present(dpred) ? dpred : nullptr);
  prm.stats = d_stats;
  prm.stats_out = d_stats;
  prm.ticket = reinterpret_cast<unsigned*>(d_workspace);
  prm.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + 256);
  cudaMemsetAsync(prm.ticket, 0, sizeof(unsigned), st);
  const unsigned grid = persistent_grid(huber_kernel, kLossThreads, 0, (segs.start[RN_MAX_LEVELS] + 3) / 4);
  huber_kernel<<<grid, kLossThreads, 0, st>>>(segs, prm);
  RN_CUDA_CHECK_LAUNCH("rn_huber_fwd_bwd");
  return RN_OK;
}

extern "C" int rn_huber_rows(int64_t rows, int width, const float* d_pred, const float* d_target,
                             const uint8_t* d_fg_mask, float* d_dpred, const double* d_stats, float* d_loss,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
  RN_REQUIRE(rows >= 0 && width >= 1, RN_ERR_BAD_ARG, "rn_huber_rows: bad shape");
  RN_REQUIRE(d_stats && d_loss && d_workspace, RN_ERR_BAD_ARG, "rn_huber_rows: NULL pointer");
  RN_REQUIRE(workspace_bytes >= rn_loss_workspace_bytes(1), RN_ERR_CAPACITY, "rn_huber_rows: workspace too small");
  RN_REQUIRE(rows == 0 || (d_pred && d_target && d_fg_mask), RN_ERR_BAD_ARG, "rn_huber_rows: NULL input");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* ticket = reinterpret_cast<unsigned*>(d_workspace);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + 256);
  cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);
  huber_rows_kernel<<<loss_grid((long long)rows * width, kLossThreads, 8), kLossThreads, 0, st>>>(
      rows, width, d_pred, d_target, d_fg_mask, d_dpred, d_stats, d_loss, partials, ticket);
  RN_CUDA_CHECK_LAUNCH("rn_huber_rows");
  return RN_OK;
}

extern "C" int rn_loss_finalize(int num_classes, int mode, const double* d_stats, float* d_out, void* stream) {
  RN_REQUIRE(d_stats && d_out && num_classes >= 1, RN_ERR_BAD_ARG, "rn_loss_finalize: bad argument");
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(num_classes, mode, d_stats, d_out);
  RN_CUDA_CHECK_LAUNCH("rn_loss_finalize");
  return RN_OK;
}

extern "C" int rn_label_stats_dense(int64_t rows, int num_classes, const float* d_labels, double* d_stats,
                                    void* stream) {
  RN_REQUIRE(rows >= 0 && num_classes >= 1 && d_stats, RN_ERR_BAD_ARG, "rn_label_stats_dense: bad argument");
  if (rows == 0) return RN_OK;
  RN_REQUIRE(d_labels, RN_ERR_BAD_ARG, "rn_label_stats_dense: NULL labels");
  dense_label_stats_kernel<<<loss_grid(rows, 256, 8), 256, 0, (cudaStream_t)stream>>>(rows, num_classes, d_labels,
                                                                                     d_stats);
  RN_CUDA_CHECK_LAUNCH("rn_label_stats_dense");
  return RN_OK;
}

extern "C" int rn_cls_loss_elementwise(int64_t rows, int num_classes, const float* d_labels, const float* d_logits,
                                       int mode, float focus, float alpha, float eps, const float* d_w_pos,
                                       const float* d_w_neg, float* d_loss, float* d_grad, void* stream) {
  RN_REQUIRE(rows >= 0 && num_classes >= 1, RN_ERR_BAD_ARG, "rn_cls_loss_elementwise: bad shape");
  RN_REQUIRE(mode == RN_CLS_FOCAL || mode == RN_CLS_BCE || mode == RN_CLS_BALANCED_BCE, RN_ERR_BAD_ARG,
             "rn_cls_loss_elementwise: unknown mode %d", mode);
  RN_REQUIRE((d_w_pos == nullptr) == (d_w_neg == nullptr), RN_ERR_BAD_ARG, "w_pos and w_neg go together");
  const long long n = (long long)rows * num_classes;
  if (n == 0) return RN_OK;
  RN_REQUIRE(d_labels && d_logits && (d_loss || d_grad), RN_ERR_BAD_ARG, "rn_cls_loss_elementwise: NULL pointer");
  LossScalars sc{focus, alpha, 1.0f - alpha, eps};
  cls_elementwise_kernel<<<loss_grid(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(n, num_classes, d_labels, d_logits,
                                                                                 mode, sc, d_w_pos, d_w_neg, d_loss,
                                                                                 d_grad);
  RN_CUDA_CHECK_LAUNCH("rn_cls_loss_elementwise");
  return RN_OK;
}

extern "C" int rn_colsum_backward(int64_t rows, int num_classes, const float* d_labels, const float* d_logits,
                                  const float* d_a, const float* d_b, float* d_grad, void* stream) {
  RN_REQUIRE(rows >= 0 && num_classes >= 1, RN_ERR_BAD_ARG, "rn_colsum_backward: bad shape");
  const long long n = (long long)rows * num_classes;
  if (n == 0) return RN_OK;
  RN_REQUIRE(d_labels && d_logits && d_a && d_b && d_grad, RN_ERR_BAD_ARG, "rn_colsum_backward: NULL pointer");
  colsum_backward_kernel<<<loss_grid(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(n, num_classes, d_labels, d_logits,
                                                                                 d_a, d_b, d_grad);
  RN_CUDA_CHECK_LAUNCH("rn_colsum_backward");
  return RN_OK;
}

extern "C" int rn_focal_softmax(int64_t rows, int num_classes, const float* d_labels, const float* d_logits,
                                float focus, float alpha, float eps, float* d_loss, float* d_grad, void* stream) {
  RN_REQUIRE(rows >= 0 && num_classes >= 1, RN_ERR_BAD_ARG, "rn_focal_softmax: bad shape");
  if (rows == 0) return RN_OK;
  RN_REQUIRE(d_labels && d_logits && (d_loss || d_grad), RN_ERR_BAD_ARG, "rn_focal_softmax: NULL pointer");
  focal_softmax_kernel<<<loss_grid(rows * 32, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      rows, num_classes, d_labels, d_logits, focus, alpha, eps, d_loss, d_grad);
  RN_CUDA_CHECK_LAUNCH("rn_focal_softmax");
  return RN_OK;
}

extern "C" int rn_sigmoid(int64_t n, const float* d_in, float* d_out, void* stream) {
  RN_REQUIRE(n >= 0, RN_ERR_BAD_ARG, "rn_sigmoid: negative n");
  if (n == 0) return RN_OK;
  RN_REQUIRE(d_in && d_out, RN_ERR_BAD_ARG, "rn_sigmoid: NULL pointer");
  sigmoid_kernel<<<loss_grid(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(n, d_in, d_out);
  RN_CUDA_CHECK_LAUNCH("rn_sigmoid");
  return RN_OK;
}
