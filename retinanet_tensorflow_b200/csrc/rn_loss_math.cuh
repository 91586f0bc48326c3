// Element math shared by the classification-loss kernels (rn_loss.cu, rn_loss_tile.cu):
// the reference-order generic forms (libm, IEEE division) and the cheap ~1 ulp forms of the
// streaming kernels.  losses.py:6-15 (focal), tf.nn.sigmoid_cross_entropy_with_logits (BCE).
#pragma once

#include <cmath>

#include <cuda_bf16.h>

#include "rn_common.cuh"

namespace rn {

struct LossScalars {
  float focus, alpha, one_minus_alpha, eps;
};

// ---- per-element math --------------------------------------------------------------------------
// Focal (losses.py:9-13): loss = -alpha_t * (1 - p_t)^focus * log(p_t + eps), with the naive
// 1 - p and eps inside the log; gradient follows the autodiff chain through the same
// float32 intermediates.
template <bool GRAD>
__device__ __forceinline__ void focal_elem(float x, bool z, const LossScalars& sc, float& loss, float& grad) {
  const float p = sigmoid_f32(x);
  const float q = fsub(1.0f, p);
  const float pt = z ? p : q;
  const float at = z ? sc.alpha : sc.one_minus_alpha;
  const float w = fsub(1.0f, pt);
  const float pe = fadd(pt, sc.eps);
  const float lg = logf(pe);
  float mod, dmod;  // w^focus and focus * w^(focus-1)
  if (sc.focus == 2.0f) {
    mod = w * w;
    dmod = 2.0f * w;
  } else {
    mod = powf(w, sc.focus);
    dmod = sc.focus * powf(w, sc.focus - 1.0f);
  }
  loss = (-at * mod) * lg;
  if (GRAD) {
    const float d_pt = -at * (mod / pe - dmod * lg);
    grad = (z ? d_pt : -d_pt) * (p * q);
  }
}

// tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|)).  Its autodiff
// gradient is relu'(x) - z - sign(x) e/(1+e) with e = exp(-|x|) (== p - z, minus the
// cancellation of `p - 1`).  Also returns p = sigmoid(x) for the dice term.
template <bool GRAD>
__device__ __forceinline__ void bce_elem(float x, float z, float& loss, float& p_out, float& dx) {
  const float e = expf(-fabsf(x));
  loss = fadd(fsub(fmaxf(x, 0.0f), fmul(x, z)), log1pf(e));
  if (GRAD) {
    const float r = fdiv(1.0f, fadd(1.0f, e));
    const float s = e * r;
    p_out = x >= 0.0f ? r : s;
    dx = x >= 0.0f ? (1.0f - z) - s : s - z;
  }
}

// ---- fast path: compact labels, C % 4 == 0 ------------------------------------------------------------
// The generic kernel above spends ~110 instructions per logit (IEEE divisions, libm expf/logf with
// their special-case branches, a 64-bit division per unit) and is issue-bound at a third of the
// HBM roofline.  The fast kernel keeps the reference's op sequence (p, naive 1-p, eps inside the
// log) but evaluates each primitive with the cheapest form that stays within ~1 ulp:
//   exp  : ex2.approx (on an exactly split x*log2(e) where p -> 1 matters)   (rel. err <= 2.4e-7)
//   1/d  : rcp.approx (+ one Newton step where p -> 1 matters)               (<= 1 ulp)
//   log  : exponent/mantissa split + minimax log1p, no branches              (rel. err <= 1.2e-7)
// and treats every element as a negative (z = 0); the at-most-one positive element of a row is
// recomputed on a rare path.  See DESIGN.md section 4 for the loop structure and its history.
// exp(-x), clamped so that 1 + e stays finite.  The product x*log2(e) is split into its rounded
// value and its exact rounding error (one FFMA), so the result carries only ex2.approx's error.
__device__ __forceinline__ float exp_neg_fast(float x) {
  const float kL2E = 1.44269502163e+00f;
  const float t = x * -kL2E;
  const float r = fmaf(x, -kL2E, -t);
  const float e = ex2_approx(fminf(t, 126.0f));
  return fmaf(e * r, 0.693147182f, e);
}

// One NEGATIVE element (z = 0): returns w^2 log(q + eps) (the loss without its -(1-alpha) factor)
// and the gradient scaled by cg = (1-alpha)/max(num_fg,1).
__device__ __forceinline__ float focal_neg_fast(float x, float eps, float cg, float& grad, float& pq) {
  const float e = exp_neg_fast(x);
  const float p = rcp_newton(1.0f + e);
  const float q = 1.0f - p;     // prob_true
  const float w = 1.0f - q;     // naive 1 - prob_true, as the reference computes it
  const float pe = q + eps;
  const float lg = log_fast(pe);
  const float u = w * lg;
  const float mod = w * w;
  // w^2/(q+eps) - 2 w log(q+eps): both terms positive, so rcp.approx's 1e-7 is not amplified
  const float v = fmaf(-2.0f, u, mod * rcp_approx(pe));
  pq = p * q;
  grad = (cg * v) * pq;
  return w * u;
}

// Same for a logit known to be <= -0.70 (p < 1/3, q + eps in [2/3, 4/3]): log needs no exponent
// split and rcp.approx alone gives p to 1 ulp.  This is the path nearly every negative takes.
__device__ __forceinline__ float focal_neg_bulk(float x, float eps, float cg, float& grad, float& pq) {
  // x <= -0.70: the rounding error of x*log2(e) adds at most |x| * 6e-8 relative to e (3e-7 at
  // the prior logit), so the exact-product correction of exp_neg_fast is not needed here.
  // (no clamp: for very negative x the exponential overflows to +inf, rcp.approx(inf) = 0, and p = 0 is exact)
  const float e = ex2_approx(x * -1.44269502163e+00f);
  const float p = rcp_approx(1.0f + e);
  const float q = 1.0f - p;
  const float w = 1.0f - q;
  const float pe = q + eps;
  const float f = pe - 1.0f;   // in [-1/3, 1.2e-7]: one-sided degree-7 minimax log1p, 9.2e-8 relative
  float t = 4.589447379e-01f;
  t = fmaf(t, f, 1.008207817e-02f);
  t = fmaf(t, f, 2.436887920e-01f);
  t = fmaf(t, f, -2.447633296e-01f);
  t = fmaf(t, f, 3.336158395e-01f);
  t = fmaf(t, f, -4.999949932e-01f);
  const float lg = fmaf(f * f, t, f);
  const float u = w * lg;
  const float v = fmaf(-2.0f, u, (w * w) * rcp_approx(pe));
  pq = p * q;
  grad = (cg * v) * pq;
  return w * u;
}

// 1/a, 1/b, 1/c, 1/d with ONE rcp.approx (the reciprocal of the product times the other factors); the
// arguments are O(1), so the product neither over- nor underflows.  Three extra roundings (~2e-7 relative).
// Pays where the MUFU pipe (a quarter of the FMA rate) is the limit: the dice sums pass (0.85 -> 0.69 ms); the
// focal fwd+bwd kernel is HBM/issue-limited and did not gain from it (measured), so it keeps one rcp per use.
__device__ __forceinline__ void rcp4(float a, float b, float c, float d, float& ia, float& ib, float& ic, float& id) {
  const float ab = a * b, cd = c * d;
  const float r = rcp_approx(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  ia = rab * b;
  ib = rab * a;
  ic = rcd * d;
  id = rcd * c;
}

// BCE of a confident negative (x <= -0.70, z = 0): loss = log1p(e^x) by a one-sided minimax
// polynomial on e in [0, 0.5] (7.9e-8 relative), p = e / (1 + e) = d loss / dx.
__device__ __forceinline__ float bce_neg_bulk(float x, float& p) {
  const float e = ex2_approx(x * 1.44269502163e+00f);   // exp(x) < 0.5
  p = e * rcp_approx(1.0f + e);
  float t = -2.785309963e-02f;
  t = fmaf(t, e, 8.751634508e-02f);
  t = fmaf(t, e, -1.473636031e-01f);
  t = fmaf(t, e, 1.960214078e-01f);
  t = fmaf(t, e, -2.495450228e-01f);
  t = fmaf(t, e, 3.333081305e-01f);
  t = fmaf(t, e, -4.999995232e-01f);
  return fmaf(e * e, t, e);
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float4& lo, float4& hi) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  const float2 c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
  lo = make_float4(a.x, a.y, b.x, b.y);
  hi = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float4& lo, const float4& hi) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(lo.x, lo.y);
  h[1] = __floats2bfloat162_rn(lo.z, lo.w);
  h[2] = __floats2bfloat162_rn(hi.x, hi.y);
  h[3] = __floats2bfloat162_rn(hi.z, hi.w);
  return r;
}

}  // namespace rn
