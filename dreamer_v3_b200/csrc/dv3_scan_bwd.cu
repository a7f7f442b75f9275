// Reverse-time BPTT through the RSSM observe scan as ONE cooperative launch (replaces the ~700 per-op launches of
// the reverse loop in Engine::wm_loss_backward). Only input gradients are on the recurrence; every weight gradient
// is a single N-row contraction after the kernel. Per step t (T-1 .. 0), 4 grid-synchronised stages:
//   R2  dp_act = dlogits[t] . W_zq^T
//   R3  dp_pre = LN'SiLU'(dp_act)  (applied by the consumer)      dh_tot = d_hz[t].h + dp_pre . W_post[E:]^T
//   R4  GRU gate gradients from dh_tot (applied by the consumer)   dh_in  = dh_tot*z + dgh . U^T ;  du = dgx . K^T
//   R5  dx_pre = LN'SiLU'(du)                                      dz_in  = dx_pre . W_pre[:Z]^T
//       epilogue: d_hz[t-1] += (1 - is_first[t]) * (dh_in, dz_in), then the straight-through / KL backward of the
//       posterior categoricals of step t-1 right away (dlogits[t-1]) -- so no separate stage for it.
// The contractions read TRANSPOSED weight copies (refreshed once per update) so that weight reads stay coalesced.
#include "dv3_scan_common.cuh"

namespace cg = cooperative_groups;

namespace dv3 {
namespace {

using namespace scan;

// d(SiLU(LN(x)))/dx helpers: n = xhat*g + b ; dn = dy * silu'(n)
__device__ __forceinline__ float silu_grad(float n) {
  const float sg = sigmoidf_(n);
  return sg * (1.f + n * (1.f - sg));
}

// Row statistics of the LayerNorm backward for rows b < B: s1 = mean(dxhat), s2 = mean(dxhat * xhat), plus mean/rstd
// copied to smem. dy rows at dy + b*lddy, x rows at x + b*ldx. One warp per row, row cached in registers.
__device__ __forceinline__ void cta_lnbwd_stats(Smem& sm, const float* dy, int lddy, const float* x, int ldx,
                                                const float* mean, const float* rstd, int sstride,
                                                const float* __restrict__ gamma, const float* __restrict__ beta, int B, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += kWarps) {
    const float m = mean[(size_t)b * sstride], r = rstd[(size_t)b * sstride];
    float xv[32], dv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      xv[i] = (c < D) ? __ldcg(x + (size_t)b * ldx + c) : 0.f;
      dv[i] = (c < D) ? __ldcg(dy + (size_t)b * lddy + c) : 0.f;
    }
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float xh = (xv[i] - m) * r;
        const float g = gamma[c];
        const float dxh = dv[i] * silu_grad(xh * g + beta[c]) * g;
        a1 += dxh; a2 += dxh * xh;
      }
    }
    a1 = warp_sum(a1) / (float)D; a2 = warp_sum(a2) / (float)D;
    if (lane == 0) { sm.mean[b] = m; sm.rstd[b] = r; sm.s1[b] = a1; sm.s2[b] = a2; }
  }
  __syncthreads();
}
// dx element of the LayerNorm+SiLU backward given the row statistics in smem
__device__ __forceinline__ float lnbwd_elem(const Smem& sm, float dy, float x, float g, float bt, int b) {
  const float xh = (x - sm.mean[b]) * sm.rstd[b];
  const float dxh = dy * silu_grad(xh * g + bt) * g;
  return sm.rstd[b] * (dxh - sm.s1[b] - xh * sm.s2[b]);
}

// Straight-through + KL backward of the posterior categoricals (representation_layer.py:73-113) for the (row, cat)
// pairs `pair0 .. ` handled by this warp: dlogits = J_softmax^T (0.99 (dz + dprobs)). Lanes = classes (Zk | 32).
__device__ __forceinline__ void repr_bwd_pair(const ScanBwdParams& p, int b, int t, int cat, int lane) {
  const int Zk = p.Zk, Z = p.Z, W = p.G + p.Z;
  const size_t bt = (size_t)b * p.T + t;
  const bool act = lane < Zk;
  const int col = cat * Zk + lane;
  const float l = act ? p.post_logits[bt * Z + col] : -INFINITY;
  const float mx = warp_max(l);
  const float e = act ? expf(l - mx) : 0.f;
  const float s = warp_sum(e);
  const float pr = e / s;
  const float g = act ? 0.99f * (__ldcg(p.d_hz + bt * W + p.G + col) + p.dpost[bt * Z + col]) : 0.f;
  const float dot = warp_sum(g * pr);
  if (act) p.dlogits[bt * Z + col] = pr * (g - dot);
}

__global__ void __launch_bounds__(kThreads, 1) observe_scan_bwd_kernel(ScanBwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  cg::grid_group grid = cg::this_grid();
  const int B = p.B, T = p.T, G = p.G, D = p.D, Z = p.Z, Zc = p.Zc, Zk = p.Zk;
  const int W = G + Z;
  const int nblk = gridDim.x, blk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gtid = blk * kThreads + threadIdx.x, gthreads = nblk * kThreads;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
  const int tasks_G = (G + 31) / 32, tasks_D = (D + 31) / 32, tasks_Z = (Z + 31) / 32;

  // prologue: dlogits of the last step (no recurrent contribution yet)
  for (int pr = gwarp; pr < B * Zc; pr += gwarps) repr_bwd_pair(p, pr / Zc, T - 1, pr % Zc, lane);
  grid.sync();

  for (int t = T - 1; t >= 0; --t) {
    // ================= R2: dp_act = dlogits[t] . W_zq^T   (Wt_zq [Z, D]) =================
    for (int task = blk; task < tasks_D; task += nblk) {
      const int c0 = task * 32;
      task_gemv(sm, p.Wt_zq, D, c0, D, Z, B, [&](int b, int k) { return __ldcg(p.dlogits + ((size_t)b * T + t) * Z + k); });
      for (int e = threadIdx.x; e < B * 32; e += kThreads) {
        const int b = e >> 5, j = e & 31;
        if (c0 + j < D) p.dp_act[((size_t)b * T + t) * D + c0 + j] = sm.tile[b][j];
      }
      __syncthreads();
    }
    grid.sync();

    // ================= R3: dp_pre = LN'SiLU'(dp_act);  dh_tot = d_hz[t].h + dp_pre . W_post_h^T  (Wt_post_h [D, G]) ==========
    {
      const float* dy = p.dp_act + (size_t)t * D;
      const float* xx = p.p_pre + (size_t)t * D;
      cta_lnbwd_stats(sm, dy, T * D, xx, T * D, p.p_mean + t, p.p_rstd + t, T, p.post_g, p.post_b, B, D);
      for (int e = gtid; e < B * D; e += gthreads) {  // publish dp_pre[t] for the weight gradients
        const int b = e / D, k = e % D;
        p.dp_pre[((size_t)b * T + t) * D + k] =
            lnbwd_elem(sm, __ldcg(dy + (size_t)b * T * D + k), __ldcg(xx + (size_t)b * T * D + k), p.post_g[k], p.post_b[k], b);
      }
      for (int task = blk; task < tasks_G; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, p.Wt_post_h, G, c0, G, D, B, [&](int b, int k) {
          return lnbwd_elem(sm, __ldcg(dy + (size_t)b * T * D + k), __ldcg(xx + (size_t)b * T * D + k), p.post_g[k], p.post_b[k], b);
        });
        for (int e = threadIdx.x; e < B * 32; e += kThreads) {
          const int b = e >> 5, j = e & 31;
          if (c0 + j < G) p.dh_tot[(size_t)b * G + c0 + j] = sm.tile[b][j] + __ldcg(p.d_hz + ((size_t)b * T + t) * W + c0 + j);
        }
        __syncthreads();
      }
    }
    grid.sync();

    // ================= R4: GRU backward (gates applied by the consumer) =================
    {
      // gate gradient k in [0, 3G) of row b; for_gh selects dgh (third block scaled by r) vs dgx
      auto dgate = [&](int b, int k, bool for_gh) {
        const size_t bt = (size_t)b * T + t;
        const int j = k % G, blk3 = k / G;
        const float* gt = p.gates + bt * 3 * G;
        const float z = gt[j], r = gt[G + j], c = gt[2 * G + j];
        const float d = __ldcg(p.dh_tot + (size_t)b * G + j);
        const float dc = d * (1.f - z) * (1.f - c * c);
        if (blk3 == 0) return d * (p.h_in[bt * G + j] - c) * z * (1.f - z);
        if (blk3 == 1) return dc * p.gh[bt * 3 * G + 2 * G + j] * r * (1.f - r);
        return for_gh ? dc * r : dc;
      };
      for (int e = gtid; e < B * 3 * G; e += gthreads) {  // publish dgx[t], dgh[t] for the weight gradients
        const int b = e / (3 * G), k = e % (3 * G);
        const size_t o = ((size_t)b * T + t) * 3 * G + k;
        const float vx = dgate(b, k, false);
        p.dgx[o] = vx;
        p.dgh[o] = (k >= 2 * G) ? vx * p.gates[((size_t)b * T + t) * 3 * G + G + (k - 2 * G)] : vx;
      }
      const int ntask = tasks_G + tasks_D;
      for (int task = blk; task < ntask; task += nblk) {
        if (task < tasks_G) {  // dh_in = dh_tot * z + dgh . U^T   (Ut [3G, G])
          const int c0 = task * 32;
          task_gemv(sm, p.Ut, G, c0, G, 3 * G, B, [&](int b, int k) { return dgate(b, k, true); });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < G) {
              const size_t bt = (size_t)b * T + t;
              p.dh_in[bt * G + c0 + j] = sm.tile[b][j] + __ldcg(p.dh_tot + (size_t)b * G + c0 + j) * p.gates[bt * 3 * G + c0 + j];
            }
          }
          __syncthreads();
        } else {               // du = dgx . K^T   (Kt [3G, D])
          const int c0 = (task - tasks_G) * 32;
          task_gemv(sm, p.Kt, D, c0, D, 3 * G, B, [&](int b, int k) { return dgate(b, k, false); });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < D) p.du[((size_t)b * T + t) * D + c0 + j] = sm.tile[b][j];
          }
          __syncthreads();
        }
      }
    }
    grid.sync();

    // ================= R5: dx_pre = LN'SiLU'(du);  dz_in = dx_pre . W_pre[:Z]^T  (Wt_pre_z [D, Z]); pass to step t-1 ==========
    {
      const float* dy = p.du + (size_t)t * D;
      const float* xx = p.x_pre + (size_t)t * D;
      cta_lnbwd_stats(sm, dy, T * D, xx, T * D, p.x_mean + t, p.x_rstd + t, T, p.pre_g, p.pre_b, B, D);
      for (int e = gtid; e < B * D; e += gthreads) {
        const int b = e / D, k = e % D;
        p.dx_pre[((size_t)b * T + t) * D + k] =
            lnbwd_elem(sm, __ldcg(dy + (size_t)b * T * D + k), __ldcg(xx + (size_t)b * T * D + k), p.pre_g[k], p.pre_b[k], b);
      }
      if (t > 0) {
        // h part of the recurrent gradient: d_hz[t-1].h += (1 - f_t) * dh_in[t]
        for (int e = gtid; e < B * G; e += gthreads) {
          const int b = e / G, j = e % G;
          const size_t bt = (size_t)b * T + t;
          const float keep = 1.f - p.is_first[bt];
          float* dst = p.d_hz + (bt - 1) * W + j;
          *dst = __ldcg(dst) + keep * __ldcg(p.dh_in + bt * G + j);
        }
        for (int task = blk; task < tasks_Z; task += nblk) {
          const int c0 = task * 32;
          task_gemv(sm, p.Wt_pre_z, Z, c0, Z, D, B, [&](int b, int k) {
            return lnbwd_elem(sm, __ldcg(dy + (size_t)b * T * D + k), __ldcg(xx + (size_t)b * T * D + k), p.pre_g[k], p.pre_b[k], b);
          });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < Z) {
              const size_t bt = (size_t)b * T + t;
              float* dst = p.d_hz + (bt - 1) * W + G + c0 + j;
              *dst = __ldcg(dst) + (1.f - p.is_first[bt]) * sm.tile[b][j];
            }
          }
          __syncthreads();
          // the categoricals of this 32-column block are final for step t-1: run their repr backward now
          const int cats = 32 / Zk;
          for (int pr = warp; pr < B * cats; pr += kWarps) {
            const int b = pr / cats, cat = c0 / Zk + pr % cats;
            if (cat < Zc) repr_bwd_pair(p, b, t - 1, cat, lane);
          }
          __syncthreads();
        }
      }
    }
    grid.sync();
  }
}

// out[c, r] = in[r, c]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < rows && c2 < cols) out[(size_t)c2 * rows + r2] = tile[threadIdx.x][i];
  }
}

}  // namespace

void transpose_f32(cudaStream_t s, const float* in, float* out, int rows, int cols) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
  count_launch();
}

int scan_bwd_max_ctas(int device) {
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaFuncSetAttribute(observe_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, observe_scan_bwd_kernel, kThreads, sizeof(Smem));
  return per_sm > 0 ? sms : 0;
}

cudaError_t launch_observe_scan_bwd(cudaStream_t s, const ScanBwdParams& p, int ctas) {
  cudaFuncSetAttribute(observe_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
  ScanBwdParams pp = p;
  void* args[] = {&pp};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_bwd_kernel, dim3(ctas), dim3(kThreads), args, sizeof(Smem), s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
