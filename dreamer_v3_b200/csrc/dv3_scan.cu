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
#include <cooperative_groups.h>

#include "dv3_device.cuh"
#include "dv3_scan.h"

namespace cg = cooperative_groups;

namespace dv3 {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRows = 16;      // max rows (replay sequences) handled per launch
constexpr int kKC = 128;       // k-chunk staged per warp
constexpr int kRedLd = 33;
constexpr int kInLd = 20;     // row stride of the staged activations (16 rows + pad): conflict-free float4 stores

struct Smem {
  float in_chunk[kWarps][kKC][kInLd];   // 80 KB
  float red[kWarps][kRows][kRedLd];     // 16.5 KB
  float tile[kRows][kRedLd];            // reduced output tile
  float mean[kRows], rstd[kRows];
};

// One task: out_tile[b][j] = sum_k in(b,k) * W[k*ldw + c0 + j], j < 32, k < K. `in_fn(b, k)` produces the input
// on the fly (LayerNorm / GRU gates are applied here by the consumer). Result is left in sm.tile (all threads
// synchronised on return).
template <class InFn>
__device__ __forceinline__ void task_gemv(Smem& sm, const float* __restrict__ W, int ldw, int c0, int ncols, int K,
                                          int B, InFn in_fn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = (K + kWarps - 1) / kWarps;
  const int k0 = warp * ks, k1 = min(K, k0 + ks);
  const bool col_ok = (c0 + lane) < ncols;
  float acc[kRows];
#pragma unroll
  for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
  for (int kc = k0; kc < k1; kc += kKC) {
    const int kn = min(kKC, k1 - kc);
    // stage in[b][kc .. kc+kn) -> in_chunk[warp][kk][b]; lanes run along k (coalesced global reads), each lane
    // gathers the 16 rows of its k in registers and stores them as four float4
    for (int kb = 0; kb < kn; kb += 32) {
      const int kk = kb + lane;
      float v[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) v[b] = (b < B && kk < kn) ? in_fn(b, kc + kk) : 0.f;
      if (kk < kn) {
        float4* dst = reinterpret_cast<float4*>(&sm.in_chunk[warp][kk][0]);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        dst[2] = make_float4(v[8], v[9], v[10], v[11]);
        dst[3] = make_float4(v[12], v[13], v[14], v[15]);
      }
    }
    __syncwarp();
    const float* wp = W + (size_t)kc * ldw + c0 + lane;
    for (int kk0 = 0; kk0 < kn; kk0 += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = (col_ok && kk0 + j < kn) ? wp[(size_t)(kk0 + j) * ldw] : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = min(kk0 + j, kn - 1);
        const float4* iv = reinterpret_cast<const float4*>(&sm.in_chunk[warp][kk][0]);
        const float4 a0 = iv[0], a1 = iv[1], a2 = iv[2], a3 = iv[3];
        const float w_ = wv[j];
        acc[0] = fmaf(a0.x, w_, acc[0]); acc[1] = fmaf(a0.y, w_, acc[1]); acc[2] = fmaf(a0.z, w_, acc[2]); acc[3] = fmaf(a0.w, w_, acc[3]);
        acc[4] = fmaf(a1.x, w_, acc[4]); acc[5] = fmaf(a1.y, w_, acc[5]); acc[6] = fmaf(a1.z, w_, acc[6]); acc[7] = fmaf(a1.w, w_, acc[7]);
        acc[8] = fmaf(a2.x, w_, acc[8]); acc[9] = fmaf(a2.y, w_, acc[9]); acc[10] = fmaf(a2.z, w_, acc[10]); acc[11] = fmaf(a2.w, w_, acc[11]);
        acc[12] = fmaf(a3.x, w_, acc[12]); acc[13] = fmaf(a3.y, w_, acc[13]); acc[14] = fmaf(a3.z, w_, acc[14]); acc[15] = fmaf(a3.w, w_, acc[15]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int b = 0; b < kRows; ++b) sm.red[warp][b][lane] = acc[b];
  __syncthreads();
  for (int e = threadIdx.x; e < kRows * 32; e += kThreads) {
    const int b = e >> 5, j = e & 31;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sm.red[w][b][j];
    sm.tile[b][j] = s;
  }
  __syncthreads();
}

// LayerNorm statistics of rows x[b*ldx .. +D) for b < B, by the whole CTA -> sm.mean / sm.rstd. Each warp keeps its
// row in registers (D <= 1024): all loads are issued together, so the cost is one L2 latency, not a chain of them.
__device__ __forceinline__ void cta_ln_stats(Smem& sm, const float* x, int ldx, int B, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += kWarps) {
    const float* xr = x + (size_t)b * ldx;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < D) ? __ldcg(xr + c) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = (lane + 32 * i < D) ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float var = warp_sum(q) / (float)D;
    if (lane == 0) { sm.mean[b] = mean; sm.rstd[b] = rsqrtf(var + DV3_LN_EPS); }
  }
  __syncthreads();
}

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
    grid.sync();

    // ================= stage B: gx = SiLU(LN(x_pre)) . K + b0 =================
    {
      const float* xp = p.x_pre + (size_t)t * D;  // row b at stride T*D
      cta_ln_stats(sm, xp, T * D, B, D);
      for (int e = gtid; e < B * D; e += gthreads) {  // every CTA publishes a slice of u = SiLU(LN(x_pre)) for the backward pass
        const int b = e / D, k = e % D;
        const float n = (__ldcg(xp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.pre_g[k] + p.pre_b[k];
        p.u_act[((size_t)b * T + t) * D + k] = n * sigmoidf_(n);
      }
      if (blk == 0 && threadIdx.x < B) { p.x_mean[(size_t)threadIdx.x * T + t] = sm.mean[threadIdx.x]; p.x_rstd[(size_t)threadIdx.x * T + t] = sm.rstd[threadIdx.x]; }
      for (int task = blk; task < tasks_gate; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, p.gru_K, 3 * G, c0, 3 * G, D, B, [&](int b, int k) {
          const float n = (__ldcg(xp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.pre_g[k] + p.pre_b[k];
          return n * sigmoidf_(n);
        });
        for (int e = threadIdx.x; e < B * 32; e += kThreads) {
          const int b = e >> 5, j = e & 31;
          if (c0 + j < 3 * G) p.gx_raw[(size_t)b * 3 * G + c0 + j] = sm.tile[b][j] + p.gru_b[c0 + j];
        }
        __syncthreads();
      }
    }
    grid.sync();

    // ================= stage C: h = GRU(gx, gh, h_in);  p_pre += h . W_post[E:];  gh(t+1) = h_in(t+1) . U + b1 ==========
    {
      const int ntask = tasks_D + ((t + 1 < T) ? tasks_gate : 0);
      auto h_of = [&](int b, int k) {
        const size_t bt = (size_t)b * T + t;
        const float* gxr = p.gx_raw + (size_t)b * 3 * G;
        const float* ghr = p.gh + bt * 3 * G;
        const float z = sigmoidf_(__ldcg(gxr + k) + __ldcg(ghr + k));
        const float r = sigmoidf_(__ldcg(gxr + G + k) + __ldcg(ghr + G + k));
        const float c = tanhf(__ldcg(gxr + 2 * G + k) + r * __ldcg(ghr + 2 * G + k));
        return z * __ldcg(p.h_in + bt * G + k) + (1.f - z) * c;
      };
      // every CTA publishes a slice of h_t and of the gate activations (z, r, c) for the backward pass
      for (int e = gtid; e < B * G; e += gthreads) {
        const int b = e / G, k = e % G;
        const size_t bt = (size_t)b * T + t;
        const float* gxr = p.gx_raw + (size_t)b * 3 * G;
        const float* ghr = p.gh + bt * 3 * G;
        const float z = sigmoidf_(__ldcg(gxr + k) + __ldcg(ghr + k));
        const float r = sigmoidf_(__ldcg(gxr + G + k) + __ldcg(ghr + G + k));
        const float c = tanhf(__ldcg(gxr + 2 * G + k) + r * __ldcg(ghr + 2 * G + k));
        p.hz[bt * W + k] = z * __ldcg(p.h_in + bt * G + k) + (1.f - z) * c;
        float* gt = p.gates + bt * 3 * G;
        gt[k] = z; gt[G + k] = r; gt[2 * G + k] = c;
      }
      for (int task = blk; task < ntask; task += nblk) {
        if (task < tasks_D) {
          const int c0 = task * 32;
          task_gemv(sm, W_post_h, D, c0, D, G, B, h_of);
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < D) { float* q = p.p_pre + ((size_t)b * T + t) * D + c0 + j; *q = __ldcg(q) + sm.tile[b][j]; }
          }
          __syncthreads();
        } else {
          const int c0 = (task - tasks_D) * 32;
          task_gemv(sm, p.gru_U, 3 * G, c0, 3 * G, G, B, [&](int b, int k) {
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
    grid.sync();

    // ================= stage D: logits = SiLU(LN(p_pre)) . W_zq + b -> unimix probs, sample z_t =================
    {
      const float* pp = p.p_pre + (size_t)t * D;
      cta_ln_stats(sm, pp, T * D, B, D);
      for (int e = gtid; e < B * D; e += gthreads) {
        const int b = e / D, k = e % D;
        const float n = (__ldcg(pp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.post_g[k] + p.post_b[k];
        p.p_act[((size_t)b * T + t) * D + k] = n * sigmoidf_(n);
      }
      if (blk == 0 && threadIdx.x < B) { p.p_mean[(size_t)threadIdx.x * T + t] = sm.mean[threadIdx.x]; p.p_rstd[(size_t)threadIdx.x * T + t] = sm.rstd[threadIdx.x]; }
      for (int task = blk; task < tasks_Z; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, p.W_zq, Z, c0, Z, D, B, [&](int b, int k) {
          const float n = (__ldcg(pp + (size_t)b * T * D + k) - sm.mean[b]) * sm.rstd[b] * p.post_g[k] + p.post_b[k];
          return n * sigmoidf_(n);
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
    grid.sync();
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
  cudaFuncSetAttribute(observe_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  ScanFwdParams pp = p;
  void* args[] = {&pp};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_fwd_kernel, dim3(ctas), dim3(kThreads), args, sizeof(Smem), s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
