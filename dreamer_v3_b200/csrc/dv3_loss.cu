// Fused loss kernels: symlog / two-hot cross-entropies, Bernoulli continue loss, decoder SSE, KL balance with
// free bits (world_model_losses.py), reverse lambda-return scan (critic_loss.py:92-139), critic and actor
// losses (critic_loss.py:15-89, actor_loss.py:12-138) and percentile-EMA return normalisation.
// All kernels also emit the gradients w.r.t. the network outputs they consume.
#include "dv3_device.cuh"

namespace dv3 {
namespace {

constexpr int kNB = DV3_NB;
constexpr int kPerLane = 8;  // ceil(255 / 32)

// log-softmax statistics of one 255-bucket row held as 8 values per lane.
struct Row255 {
  float l[kPerLane];
  float mx, lse;  // lse = log(sum(exp(l - mx))) + mx
  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      l[i] = (c < kNB) ? row[c] : -INFINITY;
      m = fmaxf(m, l[i]);
    }
    mx = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) s += (lane + 32 * i < kNB) ? expf(l[i] - mx) : 0.f;
    lse = logf(warp_sum(s)) + mx;
  }
  __device__ __forceinline__ float prob(int i) const { return expf(l[i] - lse); }
  // log-prob of bucket k (k is warp-uniform)
  __device__ __forceinline__ float logp_at(int k, int lane) const {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i)
      if (lane + 32 * i == k) v = l[i] - lse;
    return warp_sum(v);
  }
};

__global__ void two_hot_kernel(const float* __restrict__ v, long long n, int32_t* k, int32_t* kp1, float* wk,
                               float* wkp1, float* dense) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const TwoHot t = two_hot_dev(v[i]);
    if (k) k[i] = t.k;
    if (kp1) kp1[i] = t.kp1;
    if (wk) wk[i] = t.wk;
    if (wkp1) wkp1[i] = t.wkp1;
    if (dense) {
      float* row = dense + (size_t)i * kNB;
      for (int c = 0; c < kNB; ++c) row[c] = 0.f;
      row[t.k] += t.wk;
      row[t.kp1] += t.wkp1;
    }
  }
}

__global__ void bucket_expectation_kernel(const float* __restrict__ logits, int ldl, float* __restrict__ out,
                                          long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    Row255 R;
    R.load(logits + (size_t)r * ldl, lane);
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < kNB) e += R.prob(i) * bucket_value(c);
    }
    e = warp_sum(e);
    if (lane == 0) out[r] = e;
  }
}

__global__ void bucket_expectation_bwd_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ dE,
                                              float* __restrict__ dlogits, int lddl, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    Row255 R;
    R.load(logits + (size_t)r * ldl, lane);
    float e = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < kNB) e += R.prob(i) * bucket_value(c);
    }
    e = warp_sum(e);
    const float g = dE[r];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < kNB) dlogits[(size_t)r * lddl + c] = g * R.prob(i) * (bucket_value(c) - e);
    }
  }
}

__global__ void sample_categorical_kernel(const float* __restrict__ probs, const float* __restrict__ u, long long rows,
                                          int K, int32_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const bool act = lane < K;
    const float p = act ? probs[(size_t)r * K + lane] : 0.f;
    const int pick = categorical_pick(p, act, K, u[r], lane);
    if (lane == 0) idx[r] = pick;
  }
}

// ---- world-model losses: one 256-thread block per (b,t) row ----------------------------------------------------
__global__ void __launch_bounds__(256) wm_losses_kernel(WmLossArgs a) {
  __shared__ float red[8];
  __shared__ float s_kl;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // decoder: sum of squared differences over the flattened observation (world_model_losses.py:36)
  float sse = 0.f;
  {
    const float* o = a.obs + (size_t)n * a.obs_flat;
    const float* l = a.loc + (size_t)n * a.obs_flat;
    float* d = a.dloc + (size_t)n * a.obs_flat;
    const float sc = 2.f * a.inv_count;
    if ((a.obs_flat & 3) == 0 && (((uintptr_t)o | (uintptr_t)l | (uintptr_t)d) & 15) == 0) {
      // image observations (12 288 values per row, 150 MB per update over the three tensors): 128-bit accesses, four
      // independent load pairs in flight per thread
      const float4* o4 = reinterpret_cast<const float4*>(o);
      const float4* l4 = reinterpret_cast<const float4*>(l);
      float4* d4 = reinterpret_cast<float4*>(d);
      const int n4 = a.obs_flat >> 2;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int i = tid;
      for (; i + 3 * 256 < n4; i += 4 * 256) {
        float4 lv[4], ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { lv[j] = __ldg(l4 + i + j * 256); ov[j] = __ldg(o4 + i + j * 256); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 df = make_float4(lv[j].x - ov[j].x, lv[j].y - ov[j].y, lv[j].z - ov[j].z, lv[j].w - ov[j].w);
          s0 += df.x * df.x; s1 += df.y * df.y; s2 += df.z * df.z; s3 += df.w * df.w;
          d4[i + j * 256] = make_float4(sc * df.x, sc * df.y, sc * df.z, sc * df.w);   // (stays in the L2 for the decoder backward)
        }
      }
      for (; i < n4; i += 256) {
        const float4 lv = l4[i], ov = o4[i];
        const float4 df = make_float4(lv.x - ov.x, lv.y - ov.y, lv.z - ov.z, lv.w - ov.w);
        s0 += df.x * df.x; s1 += df.y * df.y; s2 += df.z * df.z; s3 += df.w * df.w;
        d4[i] = make_float4(sc * df.x, sc * df.y, sc * df.z, sc * df.w);
      }
      sse = (s0 + s1) + (s2 + s3);
    } else {
      for (int i = tid; i < a.obs_flat; i += 256) {
        const float df = l[i] - o[i];
        sse += df * df;
        d[i] = sc * df;
      }
    }
  }
  sse = warp_sum(sse);
  if (lane == 0) red[warp] = sse;
  __syncthreads();
  float L_dec = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) L_dec += red[i];
  }
  __syncthreads();

  // KL(post || prior) summed over all categoricals; each warp takes categoricals warp, warp+8, ...
  float klsum = 0.f;
  for (int c = warp; c < a.Zc; c += 8) {
    const bool act = lane < a.Zk;
    const size_t off = (size_t)n * a.Z + c * a.Zk + lane;
    const float p = act ? a.post[off] : 0.f, q = act ? a.prior[off] : 0.f;
    const float sp = warp_sum(p), sq = warp_sum(q);
    const float ph = p / sp;
    const float term = act ? ph * ((logf(p) - logf(sp)) - (logf(q) - logf(sq))) : 0.f;
    klsum += warp_sum(term);
  }
  if (lane == 0) red[warp] = klsum;
  __syncthreads();
  if (tid == 0) {
    float k = 0.f;
    for (int i = 0; i < 8; ++i) k += red[i];
    s_kl = k;
  }
  __syncthreads();
  const float kl = s_kl;
  const float gate = kl > 1.0f ? 1.f : 0.f;  // free bits: tf.maximum(1.0, kl) passes gradient only where kl > 1
  for (int c = warp; c < a.Zc; c += 8) {
    const bool act = lane < a.Zk;
    const size_t off = (size_t)n * a.Z + c * a.Zk + lane;
    const float p = act ? a.post[off] : 0.f, q = act ? a.prior[off] : 1.f;
    const float sp = warp_sum(p), sq = warp_sum(act ? q : 0.f);
    const float ph = p / sp;
    const float aj = act ? ((logf(p) - logf(sp)) - (logf(q) - logf(sq))) : 0.f;
    const float klc = warp_sum(act ? ph * aj : 0.f);
    if (act) {
      // L_dyn = KL(sg(post) || prior), weight 0.5 ; L_rep = KL(post || sg(prior)), weight 0.1
      a.dprior[off] = 0.5f * a.inv_count * gate * (-ph / q + 1.f / sq);
      a.dpost[off] = 0.1f * a.inv_count * gate * ((aj - klc) / sp);
    }
  }

  // reward: two-hot cross-entropy of symlog(reward) against the 255 logits (world_model_losses.py:45-59)
  if (warp == 0) {
    Row255 R;
    R.load(a.rew_logits + (size_t)n * kNB, lane);
    const TwoHot th = two_hot_dev(symlog_dev(a.rewards[n]));
    const float L_rew = -(th.wk * R.logp_at(th.k, lane) + th.wkp1 * R.logp_at(th.kp1, lane));
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < kNB) {
        float t = 0.f;
        if (c == th.k) t += th.wk;
        if (c == th.kp1) t += th.wkp1;
        a.drew_logits[(size_t)n * kNB + c] = a.inv_count * (R.prob(i) - t);
      }
    }
    if (lane == 0) {
      // continue: -Bernoulli(logits).log_prob(1 - is_terminated) (world_model_losses.py:73-74)
      const float x = 1.f - a.is_terminated[n];
      const float lg = a.cont_logit[n];
      const float sp_neg = fmaxf(-lg, 0.f) + log1pf(expf(-fabsf(lg)));  // softplus(-l) = -log sigmoid(l)
      const float sp_pos = fmaxf(lg, 0.f) + log1pf(expf(-fabsf(lg)));   // softplus(l)  = -log sigmoid(-l)
      const float L_con = x * sp_neg + (1.f - x) * sp_pos;
      a.dcont_logit[n] = a.inv_count * (sigmoidf_(lg) - x);
      const float L_kl = fmaxf(1.0f, kl);
      a.L_dec[n] = L_dec; a.L_rew[n] = L_rew; a.L_con[n] = L_con;
      a.L_dyn[n] = L_kl; a.L_rep[n] = L_kl;
      a.L_tot[n] = (L_dec + L_rew + L_con) + 0.5f * L_kl + 0.1f * L_kl;
      if (a.rew_k) a.rew_k[n] = th.k;
    }
  }
}

// means of up to 16 arrays: block b reduces array b (deterministic tree)
struct MeanArgs { const float* in[16]; float* out[16]; long long n[16]; float scale[16]; int count; };
__global__ void __launch_bounds__(256) multi_mean_kernel(MeanArgs m) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  if (b >= m.count) return;
  double s = 0.0;
  for (long long i = threadIdx.x; i < m.n[b]; i += 256) s += (double)m.in[b][i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *m.out[b] = (float)(red[0] * (double)m.scale[b]);
}

// ---- actor-critic: real-space rewards/continues/values, loss weights, lambda returns; one thread per column ----
__global__ void ac_value_targets_kernel(AcArgs a) {
  const int H = a.H, N = a.N;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float cum = 1.f;
    for (int t = 0; t <= H; ++t) {
      const size_t i = (size_t)t * N + n;
      const float c = (t == 0) ? (1.f - a.is_terminated[n]) : (a.cont_logit[i] > 0.f ? 1.f : 0.f);
      a.r[i] = inv_symlog_dev(a.rew_E[i]);
      a.v[i] = inv_symlog_dev(a.v_E[i]);
      a.c[i] = c;
      cum = cum * (a.gamma * c);            // tf.math.cumprod(gamma * c, axis=0)
      a.w[i] = cum / a.gamma;               // ... / gamma (dreamer_model.py:261)
    }
    float R = a.v[(size_t)H * N + n];
    for (int t = H - 1; t >= 0; --t) {
      const size_t j = (size_t)(t + 1) * N + n;
      const float disc = a.c[j] * a.gamma;
      const float rj = a.r_intr ? a.r[j] + a.r_intr[j] : a.r[j];   // critic_loss.py:118-120
      const float inter = rj + disc * (1.f - a.lambda_) * a.v[j];
      R = inter + disc * a.lambda_ * R;
      a.targets[(size_t)t * N + n] = R;
    }
  }
}


// isolated lambda-return op on real-space inputs [H+1, N] (critic_loss.py:92-139)
__global__ void value_targets_raw_kernel(const float* __restrict__ r, const float* __restrict__ c, const float* __restrict__ v,
                                         int H, int N, float gamma, float lambda_, float* __restrict__ targets) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float R = v[(size_t)H * N + n];
    for (int t = H - 1; t >= 0; --t) {
      const size_t j = (size_t)(t + 1) * N + n;
      const float disc = c[j] * gamma;
      const float inter = r[j] + disc * (1.f - lambda_) * v[j];
      R = inter + disc * lambda_ * R;
      targets[(size_t)t * N + n] = R;
    }
  }
}

__global__ void ac_returns_bwd_kernel(AcArgs a, const float* __restrict__ dR, const float* __restrict__ dV,
                                      float* __restrict__ drew_E, float* __restrict__ dv_E) {
  const int H = a.H, N = a.N;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float g = dR[n];  // dL/dR_0
    float dv_next_direct;  // direct dL/dv_t term
    // t = 0: r_0 unused, v_0 only direct
    drew_E[n] = 0.f;
    dv_E[n] = dV[n] * expf(fabsf(a.v_E[n]));
    for (int t = 0; t < H; ++t) {
      const size_t j = (size_t)(t + 1) * N + n;
      const float disc = a.c[j] * a.gamma;
      const float dr = g;
      float dv = g * disc * (1.f - a.lambda_);
      float gn = g * disc * a.lambda_;
      if (t + 1 < H) { gn += dR[j]; dv_next_direct = dV[j]; } else { dv_next_direct = 0.f; dv += gn; }
      dv += dv_next_direct;
      drew_E[j] = dr * expf(fabsf(a.rew_E[j]));  // d inverse_symlog(y)/dy = exp(|y|)
      dv_E[j] = dv * expf(fabsf(a.v_E[j]));
      g = gn;
    }
  }
}

// ---- percentile (tfp 'nearest') by exact radix select + EMA update; single block ----------------------------------
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(1024) percentile_ema_kernel(const float* __restrict__ x, long long n, long long rank_lo,
                                                             long long rank_hi, float decay, float* ema,
                                                             float* raw_out) {
  __shared__ unsigned int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ long long s_rank;
  float result[2];
  for (int which = 0; which < 2; ++which) {
    if (threadIdx.x == 0) { s_prefix = 0u; s_rank = which == 0 ? rank_lo : rank_hi; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t k = float_key(x[i]);
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        long long r = s_rank;
        int b = 0;
        for (; b < 256; ++b) {
          if (r < (long long)hist[b]) break;
          r -= hist[b];
        }
        s_rank = r;
        s_prefix = prefix | ((uint32_t)b << shift);
      }
      __syncthreads();
    }
    result[which] = key_float(s_prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    raw_out[0] = result[0]; raw_out[1] = result[1];
    // actor_loss.py:115-129: first call assigns, later calls EMA
    if (isnan(ema[0])) { ema[0] = result[0]; ema[1] = result[1]; }
    else {
      const float om = (float)(1.0 - (double)decay);
      ema[0] = decay * ema[0] + om * result[0];
      ema[1] = decay * ema[1] + om * result[1];
    }
  }
}

// ---- critic + actor losses: one warp per (t < H, n) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) ac_losses_kernel(AcLossArgs a) {
  const AcArgs& b = a.base;
  const int H = b.H, N = b.N, A = b.A;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)(H + 1) * N;
  const float off = a.ema[0];
  const float invscale = fmaxf(1e-8f, a.ema[1] - a.ema[0]);
  for (long long e = warp0; e < total; e += nwarps) {
    const int t = (int)(e / N);
    if (t == H) {  // last dreamed step: no loss; zero its gradient rows
      for (int c = lane; c < kNB; c += 32) a.dcritic_logits[(size_t)e * kNB + c] = 0.f;
      if (b.discrete) { for (int c = lane; c < A; c += 32) a.dactor_logits[(size_t)e * A + c] = 0.f; }
      else { for (int c = lane; c < A; c += 32) { a.dmu[(size_t)e * A + c] = 0.f; a.ds[(size_t)e * A + c] = 0.f; } if (lane == 0) a.dV[e] = 0.f; }
      continue;
    }
    const float w = b.w[e];
    const float Rt = b.targets[e];
    // ---- critic: two-hot CE against symlog(targets) and against the EMA critic's expectation ----
    Row255 R;
    R.load(a.critic_logits + (size_t)e * kNB, lane);
    const float rs = symlog_dev(Rt);
    const TwoHot t1 = two_hot_dev(rs);
    const TwoHot t2 = two_hot_dev(b.ema_E[e]);
    const float L_val = -(t1.wk * R.logp_at(t1.k, lane) + t1.wkp1 * R.logp_at(t1.kp1, lane));
    const float L_reg = -(t2.wk * R.logp_at(t2.k, lane) + t2.wkp1 * R.logp_at(t2.kp1, lane));
    const float cs = b.inv_count * w;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < kNB) {
        float tt = 0.f;
        if (c == t1.k) tt += t1.wk;
        if (c == t1.kp1) tt += t1.wkp1;
        if (c == t2.k) tt += t2.wk;
        if (c == t2.kp1) tt += t2.wkp1;
        a.dcritic_logits[(size_t)e * kNB + c] = cs * (2.f * R.prob(i) - tt);
      }
    }
    // ---- actor ----
    const float scaled = (Rt - off) / invscale - (b.v[e] - off) / invscale;
    float logp = 0.f, ent = 0.f, L_reinf;
    if (b.discrete) {
      const int ai = a.action_idx[e];
      // softmax over A logits, strided over lanes
      float mx = -INFINITY;
      for (int c = lane; c < A; c += 32) mx = fmaxf(mx, a.actor_logits[(size_t)e * A + c]);
      mx = warp_max(mx);
      float s = 0.f;
      for (int c = lane; c < A; c += 32) s += expf(a.actor_logits[(size_t)e * A + c] - mx);
      s = warp_sum(s);
      float S = 0.f;
      for (int c = lane; c < A; c += 32) {
        const float p = expf(a.actor_logits[(size_t)e * A + c] - mx) / s;
        S += 0.99f * p + 0.01f * (1.0f / (float)A);
      }
      S = warp_sum(S);
      const float logS = logf(S);
      float lp = 0.f, en = 0.f;
      for (int c = lane; c < A; c += 32) {
        const float p = expf(a.actor_logits[(size_t)e * A + c] - mx) / s;
        const float pu = 0.99f * p + 0.01f * (1.0f / (float)A);
        const float lpu = logf(pu);
        if (c == ai) lp = lpu;  // dist.logits = log(unimixed probs) (actor_loss.py:34-43)
        en -= (pu / S) * (lpu - logS);
      }
      logp = warp_sum(lp);
      ent = warp_sum(en);
      L_reinf = -(logp * scaled);
      // gradient w.r.t. unimixed probs, then through unimix (0.99) and the softmax Jacobian
      float dot = 0.f;
      for (int c = lane; c < A; c += 32) {
        const float p = expf(a.actor_logits[(size_t)e * A + c] - mx) / s;
        const float pu = 0.99f * p + 0.01f * (1.0f / (float)A);
        float g = b.entropy_scale * ((logf(pu) - logS) + ent) / S;
        if (c == ai) g -= scaled / pu;
        dot += (0.99f * cs * g) * p;
      }
      dot = warp_sum(dot);
      for (int c = lane; c < A; c += 32) {
        const float p = expf(a.actor_logits[(size_t)e * A + c] - mx) / s;
        const float pu = 0.99f * p + 0.01f * (1.0f / (float)A);
        float g = b.entropy_scale * ((logf(pu) - logS) + ent) / S;
        if (c == ai) g -= scaled / pu;
        a.dactor_logits[(size_t)e * A + c] = p * (0.99f * cs * g - dot);
      }
    } else {
      float lp = 0.f, en = 0.f;
      for (int c = lane; c < A; c += 32) {
        const float mu = a.actor_mu[(size_t)e * A + c], sr = a.actor_s[(size_t)e * A + c];
        const float sg = sigmoidf_(sr + 2.0f);
        const float sd = 0.9f * sg + 0.1f;
        const float loc = tanhf(mu);
        const float z = (a.actions[(size_t)e * A + c] - loc) / sd;
        lp += -0.5f * z * z - logf(sd) - 0.9189385332046727f;
        en += 1.4189385332046727f + logf(sd);
        // d(-eta * H * w / count)/ds through std = 0.9 sigmoid(s+2) + 0.1
        a.dmu[(size_t)e * A + c] = 0.f;
        a.ds[(size_t)e * A + c] = -b.entropy_scale * cs / sd * 0.9f * sg * (1.f - sg);
      }
      logp = warp_sum(lp);
      ent = warp_sum(en);
      L_reinf = -scaled;  // actor_loss.py:57; gradient flows through the dream (ac_returns_bwd)
      if (lane == 0) {
        a.dR[e] = -cs / invscale;
        a.dV[e] = cs / invscale;
      }
    }
    if (lane == 0) {
      const float L_ent = -b.entropy_scale * ent;
      a.L_val_HB[e] = L_val; a.L_reg_HB[e] = L_reg; a.L_critic_HB[e] = (L_val + L_reg) * w;
      a.targets_symlog[e] = rs;
      a.logp_HB[e] = logp; a.scaled_HB[e] = scaled; a.ent_HB[e] = ent;
      a.L_reinf_HB[e] = L_reinf; a.L_ent_HB[e] = L_ent;
      a.L_actor_HB[e] = (L_reinf + L_ent) * w;
    }
  }
}

__global__ void box_actor_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sraw,
                                     const float* __restrict__ eps, float* __restrict__ act, int lda, long long rows, int A) {
  const long long n = rows * A;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / A; const int c = (int)(i % A);
    const float sd = 0.9f * sigmoidf_(sraw[i] + 2.0f) + 0.1f;
    act[(size_t)r * lda + c] = tanhf(mu[i]) + sd * eps[i];
  }
}
__global__ void box_actor_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ sraw,
                                     const float* __restrict__ eps, const float* __restrict__ da, int ldda,
                                     float* __restrict__ dmu, float* __restrict__ ds, long long rows, int A) {
  const long long n = rows * A;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / A; const int c = (int)(i % A);
    const float g = da[(size_t)r * ldda + c];
    const float th = tanhf(mu[i]);
    const float sg = sigmoidf_(sraw[i] + 2.0f);
    dmu[i] += g * (1.f - th * th);
    ds[i] += g * eps[i] * 0.9f * sg * (1.f - sg);
  }
}

inline int grid_warps(long long rows, int cap = 148 * 8) {
  long long b = (rows + 7) / 8;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}
inline int grid_threads(long long n, int cap = 148 * 8) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

}  // namespace

void two_hot_op(cudaStream_t s, const float* v, long long n, int32_t* k, int32_t* kp1, float* wk, float* wkp1,
                float* dense) {
  if (n <= 0) return;
  two_hot_kernel<<<grid_threads(n), 256, 0, s>>>(v, n, k, kp1, wk, wkp1, dense); count_launch();
}
void bucket_expectation(cudaStream_t s, const float* logits, int ldl, float* out, long long rows) {
  if (rows <= 0) return;
  bucket_expectation_kernel<<<grid_warps(rows), 256, 0, s>>>(logits, ldl, out, rows); count_launch();
}
void bucket_expectation_bwd(cudaStream_t s, const float* logits, int ldl, const float* dE, float* dlogits, int lddl,
                            long long rows) {
  if (rows <= 0) return;
  bucket_expectation_bwd_kernel<<<grid_warps(rows), 256, 0, s>>>(logits, ldl, dE, dlogits, lddl, rows); count_launch();
}
void sample_categorical_op(cudaStream_t s, const float* probs, const float* u, long long rows, int K, int32_t* idx) {
  if (rows <= 0) return;
  sample_categorical_kernel<<<grid_warps(rows), 256, 0, s>>>(probs, u, rows, K, idx); count_launch();
}

void wm_losses(cudaStream_t s, const WmLossArgs& a) {
  const int N = a.B * a.T;
  wm_losses_kernel<<<N, 256, 0, s>>>(a); count_launch();
  MeanArgs m{};
  const float* ins[6] = {a.L_dec, a.L_rew, a.L_con, a.L_dyn, a.L_rep, a.L_tot};
  const int outs[6] = {0, 1, 2, 4, 5, 6};
  for (int i = 0; i < 6; ++i) { m.in[i] = ins[i]; m.out[i] = a.scalars + outs[i]; m.n[i] = N; m.scale[i] = 1.0f / (float)N; }
  m.count = 6;
  multi_mean_kernel<<<6, 256, 0, s>>>(m); count_launch();
}

void ac_value_targets(cudaStream_t s, const AcArgs& a) {
  ac_value_targets_kernel<<<grid_threads(a.N), 256, 0, s>>>(a); count_launch();
}
void value_targets_raw(cudaStream_t s, const float* r, const float* c, const float* v, int H, int N, float gamma,
                       float lambda_, float* targets) {
  value_targets_raw_kernel<<<grid_threads(N), 256, 0, s>>>(r, c, v, H, N, gamma, lambda_, targets); count_launch();
}
void ac_returns_bwd(cudaStream_t s, const AcArgs& a, const float* dR, const float* dV, float* drew_E, float* dv_E) {
  ac_returns_bwd_kernel<<<grid_threads(a.N), 256, 0, s>>>(a, dR, dV, drew_E, dv_E); count_launch();
}

static long long desc_rank(long long n, double q) {
  // tfp: index round((n-1) * (1 - q/100)) into the DESCENDING sort, half-to-even
  const double x = (double)(n - 1) * (1.0 - q / 100.0);
  return (long long)nearbyint(x);
}
void percentile_ema(cudaStream_t s, const float* x, long long n, float decay, float* ema, float* raw_out,
                    uint32_t* /*scratch*/) {
  const long long lo = (n - 1) - desc_rank(n, 5.0);
  const long long hi = (n - 1) - desc_rank(n, 95.0);
  percentile_ema_kernel<<<1, 1024, 0, s>>>(x, n, lo, hi, decay, ema, raw_out); count_launch();
}

void ac_losses(cudaStream_t s, const AcLossArgs& a) {
  const long long total = (long long)(a.base.H + 1) * a.base.N;
  ac_losses_kernel<<<grid_warps(total), 256, 0, s>>>(a); count_launch();
  const long long hn = (long long)a.base.H * a.base.N;
  MeanArgs m{};
  // scalars: 0 CRITIC_L_total, 1 CRITIC_L_neg_logp, 2 CRITIC_L_slow_reg, 3 ACTOR_L_total, 4 ACTOR_action_entropy,
  //          5 ACTOR_L_neglogp_reinforce_term, 6 ACTOR_L_neg_entropy_term
  const float* ins[7] = {a.L_critic_HB, a.L_val_HB, a.L_reg_HB, a.L_actor_HB, a.ent_HB, a.L_reinf_HB, a.L_ent_HB};
  for (int i = 0; i < 7; ++i) { m.in[i] = ins[i]; m.out[i] = a.scalars + i; m.n[i] = hn; m.scale[i] = 1.0f / (float)hn; }
  m.count = 7;
  multi_mean_kernel<<<7, 256, 0, s>>>(m); count_launch();
}

void box_actor_fwd(cudaStream_t s, const float* mu, const float* sraw, const float* eps, float* a, int lda,
                   long long rows, int A) {
  box_actor_fwd_kernel<<<grid_threads(rows * A), 256, 0, s>>>(mu, sraw, eps, a, lda, rows, A); count_launch();
}
void box_actor_bwd(cudaStream_t s, const float* mu, const float* sraw, const float* eps, const float* da, int ldda,
                   float* dmu, float* ds, long long rows, int A) {
  box_actor_bwd_kernel<<<grid_threads(rows * A), 256, 0, s>>>(mu, sraw, eps, da, ldda, dmu, ds, rows, A); count_launch();
}
void disc_actor_fwd(cudaStream_t s, const float* logits, const float* u, float* probs, int32_t* idx, float* a,
                    int lda, long long rows, int A) {
  repr_fwd(s, logits, A, probs, A, u, 1, idx, 1, a, lda, (int)rows, 1, A, 1);
}

}  // namespace dv3
