// The action of one dreamed step and the input layer of the sequence model in ONE launch, one warp per dream row
// (actor_network.py:63-118 followed by sequence_model.py:56-66 inside DreamerModel.dream_trajectory, dreamer_model.py:189-208):
//
//   act   = SiLU(LN(pre))                 last hidden layer of the actor (and of its std net for Box spaces)
//   out   = act . W_out + b               D x A output layer(s), A <= 8: warp-reduced dot products
//   a     = one_hot(sample(unimix(softmax(out))))   or   tanh(mu) + (0.9 sigmoid(s + 2) + 0.1) eps
//   x_pre = z . W_pre[:Z] (already in x_pre, written by the contraction that ran beside the actor) + a . W_pre[Z:]
//   u     = SiLU(LN(x_pre))               (+ bf16 twin: the operand of u . K)
//
// Per dreamed step this replaces five dependent launches of the per-op path (LN, D x A GEMV, sample, A x D GEMV, LN): the rollout
// is a chain of ~5 us launches, so its length, not its arithmetic, is what a step costs.
#include <cuda_bf16.h>

#include "dv3_device.cuh"
#include "dv3_common.h"

namespace dv3 {
namespace {

__device__ __forceinline__ void st_bf16x4(void* base16, size_t idx, const float4& v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base16) + idx) = o;
}

// LayerNorm (eps 1e-3, biased variance, two passes) + SiLU of a row held as NPL float4 per lane (columns 4 (lane + 32 i) ..);
// the same operation order as ln_silu_fwd_v4_kernel (dv3_elem.cu), so both give the same bits
template <int NPL>
__device__ __forceinline__ void ln_silu_row(float4 (&v)[NPL], const float* __restrict__ g, const float* __restrict__ be, float& mean,
                                            float& rstd, int lane) {
  constexpr int D = NPL * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) / (float)D + DV3_LN_EPS);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = 4 * (lane + 32 * i);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + c)), b4 = __ldg(reinterpret_cast<const float4*>(be + c));
    float4 o;
    float n;
    n = (v[i].x - mean) * rstd * g4.x + b4.x; o.x = n * sigmoidf_(n);
    n = (v[i].y - mean) * rstd * g4.y + b4.y; o.y = n * sigmoidf_(n);
    n = (v[i].z - mean) * rstd * g4.z + b4.z; o.z = n * sigmoidf_(n);
    n = (v[i].w - mean) * rstd * g4.w + b4.w; o.w = n * sigmoidf_(n);
    v[i] = o;
  }
}

template <int NPL>
__global__ void __launch_bounds__(256) actor_step_kernel(const ActorStepArgs a) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = NPL * 128;
  constexpr int MA = kActorStepMaxA;
  static_assert(MA == 8, "the softmax sum below spells out the 8-lane butterfly");
  const int lane = threadIdx.x & 31, A = a.A;
  const bool box = !a.discrete;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < a.rows; r += nwarps) {
    // every load of the row is issued before the first use
    float4 pm[NPL], ps[NPL], x[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const size_t c = (size_t)r * D + 4 * (lane + 32 * i);
      pm[i] = *reinterpret_cast<const float4*>(a.net[0].pre + c);
      ps[i] = box ? *reinterpret_cast<const float4*>(a.net[1].pre + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[i] = a.do_pre ? *reinterpret_cast<const float4*>(a.x_pre + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float noise[MA];
#pragma unroll
    for (int j = 0; j < MA; ++j) noise[j] = (j < (box ? A : 1)) ? __ldg(a.noise + (size_t)r * (box ? A : 1) + j) : 0.f;

    // ---- hidden layer and the D x A output layer of each net ----
    float out[2][MA];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      if (net == 1 && !box) break;
      const ActorStepNet& m = a.net[net];
      float4 (&v)[NPL] = net == 0 ? pm : ps;
      float mean, rstd;
      ln_silu_row<NPL>(v, m.g, m.b, mean, rstd, lane);
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const size_t c = (size_t)r * D + 4 * (lane + 32 * i);
        *reinterpret_cast<float4*>(m.act + c) = v[i];
        if (m.act16 != nullptr) st_bf16x4(m.act16, c, v[i]);
      }
      if (lane == 0) { m.mean[r] = mean; m.rstd[r] = rstd; }
#pragma unroll
      for (int j = 0; j < MA; ++j) {
        float s = 0.f;
        if (j < A) {
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            const int c = 4 * (lane + 32 * i);
            s += v[i].x * __ldg(m.out_w + (size_t)c * A + j) + v[i].y * __ldg(m.out_w + (size_t)(c + 1) * A + j) +
                 v[i].z * __ldg(m.out_w + (size_t)(c + 2) * A + j) + v[i].w * __ldg(m.out_w + (size_t)(c + 3) * A + j);
          }
          s = warp_sum(s) + __ldg(m.out_b + j);
        }
        out[net][j] = s;
      }
    }

    // ---- the action ----
    float act[MA];
    if (!box) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < MA; ++j) if (j < A) mx = fmaxf(mx, out[0][j]);
      float e[MA];
#pragma unroll
      for (int j = 0; j < MA; ++j) e[j] = j < A ? expf(out[0][j] - mx) : 0.f;
      // the butterfly order of warp_sum over lanes 0 .. 7 (repr_fwd_kernel keeps one class per lane): same bits as the per-op path
      const float s = ((e[0] + e[4]) + (e[2] + e[6])) + ((e[1] + e[5]) + (e[3] + e[7]));
      float pu[MA];
      double tot = 0.0;
#pragma unroll
      for (int j = 0; j < MA; ++j) { pu[j] = j < A ? 0.99f * (e[j] / s) + 0.01f * (1.0f / (float)A) : 0.f; tot += (double)pu[j]; }
      const double thr = (double)noise[0] * tot;   // fp64 sums of fp32 values are exact: order independent
      double cdf = 0.0;
      int pick = 0;
#pragma unroll
      for (int j = 0; j < MA; ++j) if (j < A) { cdf += (double)pu[j]; if (cdf <= thr) ++pick; }
      pick = min(pick, A - 1);
#pragma unroll
      for (int j = 0; j < MA; ++j) act[j] = j == pick ? 1.f : 0.f;
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < MA; ++j)
          if (j < A) { a.logits[(size_t)r * A + j] = out[0][j]; a.probs[(size_t)r * A + j] = pu[j]; a.act[(size_t)r * A + j] = act[j]; }
        a.a_idx[r] = pick;
      }
    } else {
#pragma unroll
      for (int j = 0; j < MA; ++j) {
        act[j] = 0.f;
        if (j < A) {
          const float sd = 0.9f * sigmoidf_(out[1][j] + 2.0f) + 0.1f;
          act[j] = tanhf(out[0][j]) + sd * noise[j];
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < MA; ++j)
          if (j < A) { a.logits[(size_t)r * A + j] = out[0][j]; a.s_raw[(size_t)r * A + j] = out[1][j]; a.act[(size_t)r * A + j] = act[j]; }
      }
    }
    if (!a.do_pre) continue;

    // ---- x_pre += a . W_pre[Z:];  u = SiLU(LN(x_pre)) ----
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = 4 * (lane + 32 * i);
#pragma unroll
      for (int j = 0; j < MA; ++j) {
        if (j < A) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(a.W_pre_a + (size_t)j * D + c));
          x[i].x += act[j] * w.x; x[i].y += act[j] * w.y; x[i].z += act[j] * w.z; x[i].w += act[j] * w.w;
        }
      }
      *reinterpret_cast<float4*>(a.x_pre + (size_t)r * D + c) = x[i];
    }
    float mean, rstd;
    ln_silu_row<NPL>(x, a.pre_g, a.pre_b, mean, rstd, lane);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const size_t c = (size_t)r * D + 4 * (lane + 32 * i);
      *reinterpret_cast<float4*>(a.u + c) = x[i];
      if (a.u16 != nullptr) st_bf16x4(a.u16, c, x[i]);
    }
    if (lane == 0) { a.x_mean[r] = mean; a.x_rstd[r] = rstd; }
  }
}

}  // namespace

bool actor_step_supported(int D, int A) { return D % 128 == 0 && D >= 128 && D <= 1024 && A >= 1 && A <= kActorStepMaxA; }

cudaError_t actor_step(cudaStream_t s, const ActorStepArgs& a) {
  if (a.rows <= 0) return cudaSuccess;
  long long blocks = (a.rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaError_t e = cudaErrorInvalidValue;
#define DV3_ACTOR_STEP(V_) e = launch_pdl(actor_step_kernel<V_>, dim3((unsigned)blocks), dim3(256), 0, s, a)
  switch (a.D / 128) {
    case 1: DV3_ACTOR_STEP(1); break; case 2: DV3_ACTOR_STEP(2); break; case 3: DV3_ACTOR_STEP(3); break; case 4: DV3_ACTOR_STEP(4); break;
    case 5: DV3_ACTOR_STEP(5); break; case 6: DV3_ACTOR_STEP(6); break; case 7: DV3_ACTOR_STEP(7); break; case 8: DV3_ACTOR_STEP(8); break;
    default: break;
  }
#undef DV3_ACTOR_STEP
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
