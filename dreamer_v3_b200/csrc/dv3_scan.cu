// Persistent RSSM observe-scan kernels (world_model.py:244-273 and its BPTT): ONE cooperative launch walks all T
// timesteps; the grid synchronises between the dependent stages of a step instead of returning to the host.
//
// Forward, per step t (4 grid barriers):
//   A  x_pre = actW[t] + sum_c W_pre[c*Zk + idx_in[c]]        one-hot z -> row gather, no 1024-deep contraction
//      (+ materialise the masked inputs h_in / z_in of world_model.py:246-250 for the backward pass)
//   B  gx    = SiLU(LN(x_pre)) . K + b0                        LayerNorm applied on the fly by the consumer
//   C  h     = GRU gates(gx, gh, h_in)  (on the fly)           p_pre += h . W_post[E:]      gh(t+1) = h_in(t+1) . U + b1
//   D  logits= SiLU(LN(p_pre)) . W_zq + b -> softmax, 1% unimix, fp64 inverse-CDF sample -> z_t
// Every contraction is "B (<=16) rows x 32 output columns" per task; a task's K range is split over the 8 warps
// of a CTA and reduced through shared memory, tasks are dealt round-robin to the CTAs of the grid. Weights are
// streamed from L2 (they are re-read every step by the same CTAs and stay L2 resident up to size L); activations
// of the step are staged in shared memory in [k][row] order so one LDS.128 feeds four FMAs.
#include <cstdlib>

#include "dv3_scan_common.cuh"

namespace dv3 {
namespace {

using namespace scan;

// Block-wide LayerNorm statistics of one row held as up to 4 values per thread (d = tid + 256 i < D).
__device__ __forceinline__ void row_stats(Smem& sm, const float (&x)[4], int D, float& mean, float& rstd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* red = &sm.red[0][0][0];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (threadIdx.x + kThreads * i < D) ? x[i] : 0.f;
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) tot += red[w];
  mean = tot / (float)D;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float d = (threadIdx.x + kThreads * i < D) ? x[i] - mean : 0.f; q += d * d; }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) var += red[w];
  rstd = rsqrtf(var / (float)D + DV3_LN_EPS);
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) observe_scan_fwd_kernel(ScanFwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  unsigned int bar_target = 0;
  const int B = p.B, T = p.T, G = p.G, D = p.D, Z = p.Z, Zc = p.Zc, Zk = p.Zk, E = p.E;
  const int W = G + Z;
  const int nblk = gridDim.x, blk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tasks_unit = (G + 7) / 8;                 // stage B: 8 GRU units x 3 gates = 24 columns per task
  const int tasks_gate = (3 * G + 31) / 32, tasks_D = (D + 31) / 32, tasks_Z = (Z + 31) / 32;
  const float* W_post_h = p.W_post + (size_t)E * D;

  // ---- weights resident in shared memory (launcher sets cache_weights when every CTA owns <= 1 task per stage
  // and the three tiles fit): copied once, reused for all T steps ----
  float* wc_B = reinterpret_cast<float*>(smem_raw + sizeof(Smem));           // stage B tile [D][32] (24 columns used)
  float* wc_C = wc_B + (size_t)D * 32;                                       // stage C tile [G][32]
  float* wc_D = wc_C + (size_t)G * 32;                                       // stage D tile [D][32]
  const bool cached = p.cache_weights != 0;
  // stage-B column of this lane: gate (lane / 8), unit 8*task + lane % 8
  auto unit_col = [&](int task) {
    const int u = task * 8 + (lane & 7);
    return (lane < 24 && u < G) ? (lane >> 3) * G + u : -1;
  };
  if (cached) {
    if (blk < tasks_unit)
      for (int e = threadIdx.x; e < D * 32; e += kThreads) {
        const int k = e >> 5, j = e & 31, u = blk * 8 + (j & 7);
        wc_B[e] = (j < 24 && u < G) ? p.gru_K[(size_t)k * 3 * G + (j >> 3) * G + u] : 0.f;
      }
    if (blk < tasks_D)
      for (int e = threadIdx.x; e < G * 32; e += kThreads) { const int k = e >> 5, j = e & 31, c = blk * 32 + j; wc_C[e] = c < D ? W_post_h[(size_t)k * D + c] : 0.f; }
    else if (blk < tasks_D + tasks_gate)
      for (int e = threadIdx.x; e < G * 32; e += kThreads) { const int k = e >> 5, j = e & 31, c = (blk - tasks_D) * 32 + j; wc_C[e] = c < 3 * G ? p.gru_U[(size_t)k * 3 * G + c] : 0.f; }
    if (blk < tasks_Z)
      for (int e = threadIdx.x; e < D * 32; e += kThreads) { const int k = e >> 5, j = e & 31, c = blk * 32 + j; wc_D[e] = c < Z ? p.W_zq[(size_t)k * Z + c] : 0.f; }
    __syncthreads();
  }

  // ---- prologue: gh(0) = h0 . U + b1 (every row starts from the learned initial state at t = 0) ----
  for (int task = blk; task < tasks_gate; task += nblk) {
    const int c0 = task * 32;
    task_gemv(sm, p.gru_U, 3 * G, c0, 3 * G, G, B, [&](int, int k) { return p.h0[k]; });
    for (int e = threadIdx.x; e < B * 32; e += kThreads) {
      const int b = e >> 5, j = e & 31;
      if (c0 + j < 3 * G) p.gh[((size_t)b * T + 0) * 3 * G + c0 + j] = sm.tile[b][j] + p.gru_b[3 * G + c0 + j];
    }
    __syncthreads();
  }
  grid_bar(p.bar, bar_target);
  auto stamp = [&](int i) {
    if (p.stamps != nullptr && blk == 0 && threadIdx.x == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  stamp(0);

  for (int t = 0; t < T; ++t) {
    // ===== stage A (row owners): x_pre = actW + gather-sum of W_pre rows, LayerNorm + SiLU -> u; masked inputs =====
    {
      int* idx_s = reinterpret_cast<int*>(&sm.tile[0][0]);
      for (int b = blk; b < B; b += nblk) {
        const size_t bt = (size_t)b * T + t;
        const bool first = (t == 0) || (p.is_first[bt] != 0.f);
        for (int c = threadIdx.x; c < Zc; c += kThreads) idx_s[c] = first ? p.z0_idx[c] : __ldcg(p.z_idx + (bt - 1) * Zc + c);
        __syncthreads();
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = threadIdx.x + kThreads * i;
          float s = 0.f;
          if (d < D) {
            s = p.actW[bt * D + d];
            int c = 0;
            for (; c + 8 <= Zc; c += 8) {
              float w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = p.W_pre[(size_t)((c + j) * Zk + idx_s[c + j]) * D + d];
#pragma unroll
              for (int j = 0; j < 8; ++j) s += w8[j];
            }
            for (; c < Zc; ++c) s += p.W_pre[(size_t)(c * Zk + idx_s[c]) * D + d];
          }
          x[i] = s;
        }
        float mean, rstd;
        row_stats(sm, x, D, mean, rstd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int d = threadIdx.x + kThreads * i;
          if (d < D) {
            const float n = (x[i] - mean) * rstd * p.pre_g[d] + p.pre_b[d];
            p.x_pre[bt * D + d] = x[i];
            p.u_act[bt * D + d] = n * sigmoid_fast(n);
          }
        }
        if (threadIdx.x == 0) { p.x_mean[bt] = mean; p.x_rstd[bt] = rstd; }
        // masked recurrent inputs of this row, kept for the backward pass (world_model.py:246-250)
        for (int j = threadIdx.x; j < Z; j += kThreads) p.z_in[bt * Z + j] = (idx_s[j / Zk] == (j % Zk)) ? 1.f : 0.f;
        for (int j = threadIdx.x; j < G; j += kThreads) p.h_in[bt * G + j] = first ? p.h0[j] : __ldcg(p.hz + (bt - 1) * W + j);
        __syncthreads();
      }
    }
    stamp(t * 8 + 1);
    grid_bar(p.bar, bar_target);
    stamp(t * 8 + 2);

    // ===== stage B: gx = u . K + b0 for 8 units x 3 gates per task, then the GRU cell -> h_t =====
    for (int task = blk; task < tasks_unit; task += nblk) {
      const int col = unit_col(task);
      if (cached) task_gemv_col(sm, wc_B, 32, lane < 24 ? lane : -1, D, B, [&](int b, int k) { return __ldcg(p.u_act + ((size_t)b * T + t) * D + k); });
      else task_gemv_col(sm, p.gru_K, 3 * G, col, D, B, [&](int b, int k) { return __ldcg(p.u_act + ((size_t)b * T + t) * D + k); });
      for (int e = threadIdx.x; e < B * 8; e += kThreads) {
        const int b = e >> 3, j = e & 7, u = task * 8 + j;
        if (u < G) {
          const size_t bt = (size_t)b * T + t;
          const float* ghr = p.gh + bt * 3 * G;
          const float z = sigmoid_fast(sm.tile[b][j] + p.gru_b[u] + __ldcg(ghr + u));
          const float r = sigmoid_fast(sm.tile[b][8 + j] + p.gru_b[G + u] + __ldcg(ghr + G + u));
          const float c = tanh_fast(sm.tile[b][16 + j] + p.gru_b[2 * G + u] + r * __ldcg(ghr + 2 * G + u));
          p.hz[bt * W + u] = z * __ldcg(p.h_in + bt * G + u) + (1.f - z) * c;
          float* gt = p.gates + bt * 3 * G;
          gt[u] = z; gt[G + u] = r; gt[2 * G + u] = c;
        }
      }
      __syncthreads();
    }
    stamp(t * 8 + 3);
    grid_bar(p.bar, bar_target);
    stamp(t * 8 + 4);

    // ===== stage C: p_pre += h . W_post[E:];  gh(t+1) = h_in(t+1) . U + b1 =====
    {
      const int ntask = tasks_D + ((t + 1 < T) ? tasks_gate : 0);
      for (int task = blk; task < ntask; task += nblk) {
        if (task < tasks_D) {
          const int c0 = task * 32;
          auto h_of = [&](int b, int k) { return __ldcg(p.hz + ((size_t)b * T + t) * W + k); };
          if (cached) task_gemv(sm, wc_C, 32, 0, 32, G, B, h_of);
          else task_gemv(sm, W_post_h, D, c0, D, G, B, h_of);
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < D) { float* q = p.p_pre + ((size_t)b * T + t) * D + c0 + j; *q = __ldcg(q) + sm.tile[b][j]; }
          }
          __syncthreads();
        } else {
          const int c0 = (task - tasks_D) * 32;
          auto hn_of = [&](int b, int k) {
            const bool first = p.is_first[(size_t)b * T + t + 1] != 0.f;
            return first ? p.h0[k] : __ldcg(p.hz + ((size_t)b * T + t) * W + k);
          };
          if (cached) task_gemv(sm, wc_C, 32, 0, 32, G, B, hn_of);
          else task_gemv(sm, p.gru_U, 3 * G, c0, 3 * G, G, B, hn_of);
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < 3 * G) p.gh[((size_t)b * T + t + 1) * 3 * G + c0 + j] = sm.tile[b][j] + p.gru_b[3 * G + c0 + j];
          }
          __syncthreads();
        }
      }
    }
    stamp(t * 8 + 5);
    grid_bar(p.bar, bar_target);

    stamp(t * 8 + 6);

    // ===== stage D: p_act = SiLU(LN(p_pre)) by every task CTA itself (one warp per row, staged straight into shared
    // memory; the first CTA also writes p_act / mean / rstd out for the backward pass), logits = p_act . W_zq + b
    // -> unimix probs, sample z_t =====
    for (int task = blk; task < tasks_Z; task += nblk) {
      const int c0 = task * 32;
      float* in_s = &sm.in_chunk[0][0][0];
      const bool writer = task == 0;
      cta_ln_silu_stage(in_s, p.p_pre + (size_t)t * D, (size_t)T * D, p.post_g, p.post_b, B, D,
                        writer ? p.p_act + (size_t)t * D : nullptr, (size_t)T * D, p.p_mean + t, p.p_rstd + t, (size_t)T);
      float acc[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
      if (cached) gemv_acc<32>(acc, in_s, wc_D, 32, lane, D);
      else gemv_acc<32>(acc, in_s, p.W_zq, Z, (c0 + lane) < Z ? c0 + lane : -1, D);
      gemv_finish<32>(sm, acc);
      for (int e = threadIdx.x; e < B * 32; e += kThreads) {
        const int b = e >> 5, j = e & 31;
        if (c0 + j < Z) {
          const float v = sm.tile[b][j] + p.b_zq[c0 + j];
          sm.tile[b][j] = v;
          p.post_logits[((size_t)b * T + t) * Z + c0 + j] = v;
        }
      }
      __syncthreads();
      // softmax / unimix / sample: one warp per (row, categorical) of this 32-column block
      const int cats = 32 / Zk;  // Zk divides 32
      for (int pr = warp; pr < B * cats; pr += kWarps) {
        const int b = pr / cats, cl = pr % cats;
        const int cat = c0 / Zk + cl;
        if (cat >= Zc) continue;
        const bool act = lane < Zk;
        const float l = act ? sm.tile[b][cl * Zk + lane] : -INFINITY;
        const float mx = warp_max(l);
        const float ex = act ? expf(l - mx) : 0.f;
        const float s = warp_sum(ex);
        const float pu = 0.99f * (ex / s) + 0.01f * (1.0f / (float)Zk);
        const size_t bt = (size_t)b * T + t;
        if (act) p.post_probs[bt * Z + cat * Zk + lane] = pu;
        const int pick = categorical_pick(pu, act, Zk, p.u_post[((size_t)t * B + b) * Zc + cat], lane);
        if (lane == 0) p.z_idx[bt * Zc + cat] = pick;
        if (act) p.hz[bt * W + G + cat * Zk + lane] = (lane == pick) ? 1.f : 0.f;
      }
      __syncthreads();
    }
    stamp(t * 8 + 7);
    grid_bar(p.bar, bar_target);
    stamp(t * 8 + 8);
  }
}

}  // namespace

size_t scan_fwd_smem_bytes() { return sizeof(Smem); }

int scan_fwd_max_ctas(int device) {
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaFuncSetAttribute(observe_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, observe_scan_fwd_kernel, kThreads, sizeof(Smem));
  return per_sm > 0 ? sms : 0;
}

cudaError_t launch_observe_scan_fwd(cudaStream_t s, const ScanFwdParams& p, int ctas) {
  ScanFwdParams pp = p;
  // smem-resident weights need one task per CTA per stage and 3 tiles of [K x 32] fp32 next to the staging buffers
  const int tg = (3 * p.G + 31) / 32, td = (p.D + 31) / 32, tz = (p.Z + 31) / 32, tu = (p.G + 7) / 8;
  const size_t cache = ((size_t)2 * p.D + p.G) * 32 * sizeof(float);
  const bool fits = p.D <= 4 * kThreads && sizeof(Smem) + cache <= 227 * 1024 && ctas >= tg + td && ctas >= tz && ctas >= tu;
  pp.cache_weights = fits ? 1 : 0;
  const size_t smem = sizeof(Smem) + (fits ? cache : 0);
  // CTAs beyond the widest stage only lengthen the grid barrier (every arrival is an atomic on one address)
  int need = tg + td;
  if (tz > need) need = tz;
  if (tu > need) need = tu;
  if (p.B > need) need = p.B;
  if (const char* f = std::getenv("DV3_SCAN_CTAS")) need = std::atoi(f);
  if (need < ctas && need > 0) ctas = need;
  cudaFuncSetAttribute(observe_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(pp.bar, 0, sizeof(unsigned int), s);
  void* args[] = {&pp};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_fwd_kernel, dim3(ctas), dim3(kThreads), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
