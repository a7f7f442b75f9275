// Persistent RSSM observe-scan kernels (world_model.py:244-273 and its BPTT): ONE cooperative launch walks all T
// timesteps; the grid synchronises between the dependent stages of a step instead of returning to the host.
//
// Forward, per step t (4 grid syncs):
//   A  x_pre = actW[t] + sum_c W_pre[c*Zk + idx_in[c]]        one-hot z -> row gather, no 1024-deep contraction
//      (+ materialise the masked inputs h_in / z_in of world_model.py:246-250 for the backward pass)
//   B  gx    = SiLU(LN(x_pre)) . K + b0                        LayerNorm applied on the fly by the consumer
//   C  h     = GRU gates(gx, gh, h_in)  (on the fly)           p_pre += h . W_post[E:]      gh(t+1) = h_in(t+1) . U + b1
//   D  logits= SiLU(LN(p_pre)) . W_zq + b -> softmax, 1% unimix, fp64 inverse-CDF sample -> z_t
// Every contraction is "B (<=16) rows x 32 output columns" per task; a task's K range is split over the 8 warps
// of a CTA and reduced through shared memory, tasks are dealt round-robin to the CTAs of the grid. Weights are
// streamed from L2 (they are re-read every step by the same CTAs and stay L2 resident up to size L); activations
// of the step are staged in shared memory in [k][row] order so one LDS.128 feeds four FMAs.
#include "dv3_scan_common.cuh"

namespace cg = cooperative_groups;

namespace dv3 {
namespace {

using namespace scan;

__global__ void __launch_bounds__(kThreads, 1) observe_scan_fwd_kernel(ScanFwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  cg::grid_group grid = cg::this_grid();
  const int B = p.B, T = p.T, G = p.G, D = p.D, Z = p.Z, Zc = p.Zc, Zk = p.Zk, E = p.E;
  const int W = G + Z;
  const int nblk = gridDim.x, blk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gtid = blk * kThreads + threadIdx.x, gthreads = nblk * kThreads;
  const int tasks_gate = (3 * G + 31) / 32, tasks_D = (D + 31) / 32, tasks_Z = (Z + 31) / 32;
  const float* W_post_h = p.W_post + (size_t)E * D;

  // ---- weights resident in shared memory: when every CTA owns at most one task per stage (XS and smaller), its
  // three [K x 32] weight tiles are copied once and stay for all T steps ----
  float* wc_B = reinterpret_cast<float*>(smem_raw + sizeof(Smem));           // stage B tile  [D][32]
  float* wc_C = wc_B + (size_t)D * 32;                                       // stage C tile  [G][32]
  float* wc_D = wc_C + (size_t)G * 32;                                       // stage D tile  [D][32]
  const bool cached = p.cache_weights != 0;
  if (cached) {
    if (blk < tasks_gate)
      for (int e = threadIdx.x; e < D * 32; e += kThreads) { const int k = e >> 5, j = e & 31, c = blk * 32 + j; wc_B[e] = c < 3 * G ? p.gru_K[(size_t)k * 3 * G + c] : 0.f; }
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
  grid.sync();
  auto stamp = [&](int i) {
    if (p.stamps != nullptr && blk == 0 && threadIdx.x == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  stamp(0);

  for (int t = 0; t < T; ++t) {
    // ================= stage A: masked inputs + gather-sum pre-GRU dense =================
    {
      // indices of the (masked) previous z for every row: one L2 latency, then Zc independent row loads per output
      int* idx_s = reinterpret_cast<int*>(&sm.red[0][0][0]);
      for (int e = threadIdx.x; e < B * Zc; e += kThreads) {
        const int b = e / Zc, c = e % Zc;
        const size_t bt = (size_t)b * T + t;
        const bool first = (t == 0) || (p.is_first[bt] != 0.f);
        idx_s[e] = first ? p.z0_idx[c] : __ldcg(p.z_idx + (bt - 1) * Zc + c);
      }
      __syncthreads();
      const int nA = B * D;
      for (int i = gtid; i < nA; i += gthreads) {
        const int b = i / D, d = i % D;
        const size_t bt = (size_t)b * T + t;
        const int* idx = idx_s + b * Zc;
        float s = p.actW[bt * D + d];
        int c = 0;
        for (; c + 8 <= Zc; c += 8) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = p.W_pre[(size_t)((c + j) * Zk + idx[c + j]) * D + d];
#pragma unroll
          for (int j = 0; j < 8; ++j) s += w8[j];
        }
        for (; c < Zc; ++c) s += p.W_pre[(size_t)(c * Zk + idx[c]) * D + d];
        p.x_pre[bt * D + d] = s;
      }
      const int nZ = B * Z;
      for (int i = gtid; i < nZ; i += gthreads) {
        const int b = i / Z, j = i % Z;
        p.z_in[((size_t)b * T + t) * Z + j] = (idx_s[b * Zc + j / Zk] == (j % Zk)) ? 1.f : 0.f;
      }
      const int nH = B * G;
      for (int i = gtid; i < nH; i += gthreads) {
        const int b = i / G, j = i % G;
        const size_t bt = (size_t)b * T + t;
        const bool first = (t == 0) || (p.is_first[bt] != 0.f);
        p.h_in[bt * G + j] = first ? p.h0[j] : __ldcg(p.hz + (bt - 1) * W + j);
      }
    }
    stamp(t * 8 + 1);
    grid.sync();
    stamp(t * 8 + 2);

    // ================= stage B: gx = SiLU(LN(x_pre)) . K + b0 =================
    {
      const float* xp = p.x_pre + (size_t)t * D;  // row b at stride T*D
      if (t == 10) stamp(T * 8 + 1);
      cta_ln_stats(sm, xp, T * D, B, D);
      if (t == 10) stamp(T * 8 + 2);
      for (int e = gtid; e < B * D; e += gthreads) {  // every CTA publishes a slice of u = SiLU(LN(x_pre)) for the backward pass
        const int b = e / D, k = e % D;
        const float n = (__ldcg(xp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.pre_g[k] + p.pre_b[k];
        p.u_act[((size_t)b * T + t) * D + k] = n * sigmoid_fast(n);
      }
      if (blk == 0 && threadIdx.x < B) { p.x_mean[(size_t)threadIdx.x * T + t] = sm.mean[threadIdx.x]; p.x_rstd[(size_t)threadIdx.x * T + t] = sm.rstd[threadIdx.x]; }
      if (t == 10) stamp(T * 8 + 3);
      for (int task = blk; task < tasks_gate; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, cached ? wc_B : p.gru_K, cached ? 32 : 3 * G, cached ? 0 : c0, cached ? 32 : 3 * G, D, B, [&](int b, int k) {
          const float n = (__ldcg(xp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.pre_g[k] + p.pre_b[k];
          return n * sigmoid_fast(n);
        });
        for (int e = threadIdx.x; e < B * 32; e += kThreads) {
          const int b = e >> 5, j = e & 31;
          if (c0 + j < 3 * G) p.gx_raw[(size_t)b * 3 * G + c0 + j] = sm.tile[b][j] + p.gru_b[c0 + j];
        }
        __syncthreads();
      }
      if (t == 10) stamp(T * 8 + 4);
    }
    stamp(t * 8 + 3);
    grid.sync();
    stamp(t * 8 + 4);

    // ================= stage C: h = GRU(gx, gh, h_in);  p_pre += h . W_post[E:];  gh(t+1) = h_in(t+1) . U + b1 ==========
    {
      const int ntask = tasks_D + ((t + 1 < T) ? tasks_gate : 0);
      auto h_of = [&](int b, int k) {
        const size_t bt = (size_t)b * T + t;
        const float* gxr = p.gx_raw + (size_t)b * 3 * G;
        const float* ghr = p.gh + bt * 3 * G;
        const float z = sigmoid_fast(__ldcg(gxr + k) + __ldcg(ghr + k));
        const float r = sigmoid_fast(__ldcg(gxr + G + k) + __ldcg(ghr + G + k));
        const float c = tanh_fast(__ldcg(gxr + 2 * G + k) + r * __ldcg(ghr + 2 * G + k));
        return z * __ldcg(p.h_in + bt * G + k) + (1.f - z) * c;
      };
      // every CTA publishes a slice of h_t and of the gate activations (z, r, c) for the backward pass
      for (int e = gtid; e < B * G; e += gthreads) {
        const int b = e / G, k = e % G;
        const size_t bt = (size_t)b * T + t;
        const float* gxr = p.gx_raw + (size_t)b * 3 * G;
        const float* ghr = p.gh + bt * 3 * G;
        const float z = sigmoid_fast(__ldcg(gxr + k) + __ldcg(ghr + k));
        const float r = sigmoid_fast(__ldcg(gxr + G + k) + __ldcg(ghr + G + k));
        const float c = tanh_fast(__ldcg(gxr + 2 * G + k) + r * __ldcg(ghr + 2 * G + k));
        p.hz[bt * W + k] = z * __ldcg(p.h_in + bt * G + k) + (1.f - z) * c;
        float* gt = p.gates + bt * 3 * G;
        gt[k] = z; gt[G + k] = r; gt[2 * G + k] = c;
      }
      for (int task = blk; task < ntask; task += nblk) {
        if (task < tasks_D) {
          const int c0 = task * 32;
          task_gemv(sm, cached ? wc_C : W_post_h, cached ? 32 : D, cached ? 0 : c0, cached ? 32 : D, G, B, h_of);
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < D) { float* q = p.p_pre + ((size_t)b * T + t) * D + c0 + j; *q = __ldcg(q) + sm.tile[b][j]; }
          }
          __syncthreads();
        } else {
          const int c0 = (task - tasks_D) * 32;
          task_gemv(sm, cached ? wc_C : p.gru_U, cached ? 32 : 3 * G, cached ? 0 : c0, cached ? 32 : 3 * G, G, B, [&](int b, int k) {
            const bool first = p.is_first[(size_t)b * T + t + 1] != 0.f;
            return first ? p.h0[k] : h_of(b, k);
          });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < 3 * G) p.gh[((size_t)b * T + t + 1) * 3 * G + c0 + j] = sm.tile[b][j] + p.gru_b[3 * G + c0 + j];
          }
          __syncthreads();
        }
      }
    }
    stamp(t * 8 + 5);
    grid.sync();
    stamp(t * 8 + 6);

    // ================= stage D: logits = SiLU(LN(p_pre)) . W_zq + b -> unimix probs, sample z_t =================
    {
      const float* pp = p.p_pre + (size_t)t * D;
      cta_ln_stats(sm, pp, T * D, B, D);
      for (int e = gtid; e < B * D; e += gthreads) {
        const int b = e / D, k = e % D;
        const float n = (__ldcg(pp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.post_g[k] + p.post_b[k];
        p.p_act[((size_t)b * T + t) * D + k] = n * sigmoid_fast(n);
      }
      if (blk == 0 && threadIdx.x < B) { p.p_mean[(size_t)threadIdx.x * T + t] = sm.mean[threadIdx.x]; p.p_rstd[(size_t)threadIdx.x * T + t] = sm.rstd[threadIdx.x]; }
      for (int task = blk; task < tasks_Z; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, cached ? wc_D : p.W_zq, cached ? 32 : Z, cached ? 0 : c0, cached ? 32 : Z, D, B, [&](int b, int k) {
          const float n = (__ldcg(pp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.post_g[k] + p.post_b[k];
          return n * sigmoid_fast(n);
        });
        // bias + publish logits
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
    }
    stamp(t * 8 + 7);
    grid.sync();
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
  const int tg = (3 * p.G + 31) / 32, td = (p.D + 31) / 32, tz = (p.Z + 31) / 32;
  const size_t cache = ((size_t)2 * p.D + p.G) * 32 * sizeof(float);
  const bool fits = sizeof(Smem) + cache <= 227 * 1024 && ctas >= tg + td && ctas >= tz;
  pp.cache_weights = fits ? 1 : 0;
  const size_t smem = sizeof(Smem) + (fits ? cache : 0);
  cudaFuncSetAttribute(observe_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void* args[] = {&pp};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_fwd_kernel, dim3(ctas), dim3(kThreads), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
