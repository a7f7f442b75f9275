// Device-side helpers shared by the kernel translation units.
#pragma once
#include "dv3_common.h"

namespace dv3 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// programmatic dependent launch (see launch_pdl in dv3_common.h): let the next kernel of the stream start its prologue now,
// and do not touch global memory before the previous kernel has completed
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// fast variants for the latency-critical persistent scan kernels (ex2.approx based, rel. error ~1e-6: far inside the
// 1e-4 parity budget, and nothing integer-valued depends on them)
__device__ __forceinline__ float ex2_ftz(float x) {   // single MUFU.EX2, no denormal fix-up branches
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_ftz(-2.8853900817779268f * fabsf(x));
  const float t = __fdividef(1.f - e, 1.f + e);
  return x < 0.f ? -t : t;
}

// sign(x) * log(|x| + 1) (utils/symlog.py:9-12): the add is fp32 (as TF does), the log is evaluated in
// fp64 and rounded once, so two-hot bucket indices derived from it are bit-identical to the oracle's.
__device__ __forceinline__ float symlog_dev(float x) {
  const float t = __fadd_rn(fabsf(x), 1.0f);
  const float l = (float)log((double)t);
  return x > 0.f ? l : (x < 0.f ? -l : 0.f);
}
__device__ __forceinline__ float inv_symlog_dev(float y) {
  const float m = expf(fabsf(y)) - 1.f;
  return y > 0.f ? m : (y < 0.f ? -m : 0.f);
}

// utils/two_hot.py:39-66 for 255 buckets in [-20, 20]; explicit round-to-nearest intrinsics keep the
// compiler from contracting into FMAs, so (k, kp1) are bit-exact against the NumPy restatement.
struct TwoHot { int k, kp1; float wk, wkp1; };
__device__ __forceinline__ TwoHot two_hot_dev(float value) {
  const float lo = DV3_BUCKET_LO, hi = DV3_BUCKET_HI;
  const float delta = (float)((20.0 - (-20.0)) / (double)(DV3_NB - 1));
  const float v = fminf(fmaxf(value, lo), hi);
  const float idx = __fdiv_rn(__fadd_rn(-lo, v), delta);
  float k = floorf(idx), kp1 = ceilf(idx);
  if (k == kp1) kp1 += 1.0f;
  if (kp1 == (float)DV3_NB) kp1 -= 2.0f;
  const float vk = __fadd_rn(lo, __fmul_rn(k, delta));
  const float vkp1 = __fadd_rn(lo, __fmul_rn(kp1, delta));
  TwoHot t;
  t.wk = __fdiv_rn(__fsub_rn(v, vkp1), __fsub_rn(vk, vkp1));
  t.wkp1 = __fsub_rn(1.0f, t.wk);
  t.k = (int)k; t.kp1 = (int)kp1;
  return t;
}
__device__ __forceinline__ float bucket_value(int k) {
  const float delta = (float)((20.0 - (-20.0)) / (double)(DV3_NB - 1));
  return DV3_BUCKET_LO + (float)k * delta;
}

// Inverse-CDF draw over one warp (lane = class, K <= 32 active lanes): fp64 inclusive scan of the fp32
// probabilities (exact, hence order independent), index = #(cdf <= u * total), clamped.
__device__ __forceinline__ int categorical_pick(float p, bool act, int K, float u, int lane) {
  double cdf = act ? (double)p : 0.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, cdf, o);
    if (lane >= o) cdf += t;
  }
  const double tot = __shfl_sync(0xffffffffu, cdf, K - 1);
  const double thr = (double)u * tot;
  const unsigned below = __ballot_sync(0xffffffffu, act && (cdf <= thr));
  return min(__popc(below), K - 1);
}

}  // namespace dv3
