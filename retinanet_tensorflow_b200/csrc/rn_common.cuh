// Shared device/host helpers for librndense (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "rndense.h"

#ifndef __CUDA_ARCH__
#else
#if __CUDA_ARCH__ < 1000
#error "librndense is written for sm_100a (B200) only"
#endif
#endif

namespace rn {

// ---- host-side status plumbing -------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
void set_exchange_timeout_ms(long long ms);

#define RN_REQUIRE(cond, status, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      rn::set_error(__VA_ARGS__);      \
      return (status);                 \
    }                                  \
  } while (0)

#define RN_CUDA_CHECK_LAUNCH(what)                                          \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      rn::set_error("%s: %s", (what), cudaGetErrorString(e__));             \
      return RN_ERR_CUDA;                                                   \
    }                                                                       \
    rn::count_launch();                                                     \
  } while (0)

static inline bool maps_present(const rn_maps_t* m) { return m != nullptr && m->p[0] != nullptr; }

// Kernel-side copies of the small host structs (passed by value as __grid_constant__).
struct Segs {
  int num;
  long long rows[RN_MAX_LEVELS];    // rows per segment
  long long start[RN_MAX_LEVELS + 1];  // exclusive prefix over rows
};

static inline Segs make_segs(const rn_rows_t* r) {
  Segs s;
  s.num = r->num_segments;
  long long acc = 0;
  for (int i = 0; i < RN_MAX_LEVELS; ++i) {
    s.rows[i] = i < s.num ? r->rows[i] : 0;
    s.start[i] = acc;
    acc += s.rows[i];
  }
  s.start[RN_MAX_LEVELS] = acc;
  for (int i = s.num; i < RN_MAX_LEVELS; ++i) s.start[i] = acc;
  return s;
}

struct Ptrs {
  void* p[RN_MAX_LEVELS];
};

static inline Ptrs make_ptrs(const rn_maps_t* m) {
  Ptrs q;
  for (int i = 0; i < RN_MAX_LEVELS; ++i) q.p[i] = m ? m->p[i] : nullptr;
  return q;
}

#ifdef __CUDACC__
// ---- exact float32 arithmetic (never contracted into FMA) -----------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// tf.nn.sigmoid as the oracle states it: 1 / (1 + exp(-x)), every op rounded to float32.
__device__ __forceinline__ float sigmoid_f32(float x) { return fdiv(1.0f, fadd(1.0f, expf(-x))); }

// ---- cheap ~1 ulp primitives shared by the fast kernels ---------------------------------------
__device__ __forceinline__ float ex2_approx(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}
__device__ __forceinline__ float rcp_approx(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}
__device__ __forceinline__ float rcp_newton(float d) {
  const float r = rcp_approx(d);
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}
// log(a) for positive normal a (any positive normal float); no I2F (XU pipe): the
// exponent becomes a float through the 1.5*2^23 magic constant.
__device__ __forceinline__ float log_fast(float a) {
  const int ia = __float_as_int(a);
  const int e = (ia - 0x3f2aaaab) & 0xff800000;
  const float f = __int_as_float(ia - e) - 1.0f;  // mantissa in [2/3, 4/3] minus 1
  const float kf = __int_as_float((e >> 23) + 0x4b400000) - 12582912.0f;
  float q = 1.402632743e-01f;
  q = fmaf(q, f, -1.538854539e-01f);
  q = fmaf(q, f, 1.399581283e-01f);
  q = fmaf(q, f, -1.641595960e-01f);
  q = fmaf(q, f, 2.001116872e-01f);
  q = fmaf(q, f, -2.500782013e-01f);
  q = fmaf(q, f, 3.333319724e-01f);
  q = fmaf(q, f, -4.999993742e-01f);
  const float r = fmaf(f * f, q, f);
  return fmaf(kf, 0.693147182f, r);
}

// ---- streaming loads / stores ----------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// ---- reductions --------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum `v` over the block; result valid in thread 0.  `scratch` holds >= 32 T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
  if (warp == 0) v = warp_sum(v);
  return v;
}

// Segment lookup: which segment does global row `r` fall in.
// Reserve one slot of the per-image candidate list `count[b]` for every calling lane.  Lanes that are
// converged here and append to the same list share one atomic (opportunistic warp aggregation), and a
// list that is already full is not touched again -- so a dense firing pattern (most anchors above the
// threshold) costs one L2 read per warp instead of an atomic per candidate.  Returns the slot; a slot
// >= capacity means "dropped" (the caller raises RN_FLAG_CAND_OVERFLOW).
__device__ __forceinline__ int reserve_slot(int32_t* count, int b, int capacity) {
  const unsigned act = __activemask();
  const unsigned same = __match_any_sync(act, b);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(same) - 1;
  int base = capacity;
  if (lane == leader) {
    if (__ldcg(count + b) < capacity) base = atomicAdd(count + b, __popc(same));
  }
  base = __shfl_sync(same, base, leader);
  return base + __popc(same & ((1u << lane) - 1u));
}

// Raise a status bit at most a few times per grid (not once per event).
__device__ __forceinline__ void raise_flag(uint32_t* status, uint32_t flag) {
  if (status && !(__ldcg(status) & flag)) atomicOr(status, flag);
}

__device__ __forceinline__ int seg_of(const Segs& s, long long r) {
  int k = 0;
#pragma unroll
  for (int i = 1; i < RN_MAX_LEVELS; ++i) k += (i < s.num && r >= s.start[i]) ? 1 : 0;
  return k;
}
#endif  // __CUDACC__

}  // namespace rn
