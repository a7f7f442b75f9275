// Reverse-time BPTT through the RSSM observe scan as ONE cooperative launch (replaces the ~700 per-op launches of
// the reverse loop in Engine::wm_loss_backward). Only input gradients are on the recurrence; every weight gradient
// is a single N-row contraction after the kernel. Per step t (T-1 .. 0), 6 short grid-synchronised stages:
//   R2   dp_act = dlogits[t] . W_zq^T                                       (32-column tasks, K split over 8 warps)
//   R3a  dp_pre = LN'SiLU'(dp_act)                                          (row owners: one CTA per sequence)
//   R3   dh_tot = d_hz[t].h + dp_pre . W_post[E:]^T, GRU gate gradients dgx / dgh of the task's 32 units
//   R4   dh_in  = dh_tot*z + dgh . U^T ;  du = dgx . K^T
//   R5a  dx_pre = LN'SiLU'(du);  d_hz[t-1].h += (1 - is_first[t]) dh_in     (row owners)
//   R5   dz_in  = dx_pre . W_pre[:Z]^T;  d_hz[t-1].z += (1 - is_first[t]) dz_in, then the straight-through / KL
//        backward of the posterior categoricals of step t-1 right away (dlogits[t-1]) -- no separate stage for it.
// Stage inputs are published to global memory by their producer and staged by plain copies, so no CTA recomputes
// another CTA's non-linearities. The contractions read TRANSPOSED weight copies (refreshed once per update) so that
// weight reads stay coalesced.
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

// Row-owner LayerNorm+SiLU backward of one row: dy/x held as up to 4 values per thread (d = tid + 256 i < D).
__device__ __forceinline__ void row_lnbwd(Smem& sm, const float (&dy)[4], const float (&x)[4], float mean, float rstd,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, int D, float (&dx)[4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* red = &sm.red[0][0][0];
  float xh[4], dxh[4];
  float a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = threadIdx.x + kThreads * i;
    if (d < D) {
      xh[i] = (x[i] - mean) * rstd;
      const float g = gamma[d];
      dxh[i] = dy[i] * silu_grad(xh[i] * g + beta[d]) * g;
      a1 += dxh[i]; a2 += dxh[i] * xh[i];
    } else { xh[i] = 0.f; dxh[i] = 0.f; }
  }
  a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) { red[warp] = a1; red[kWarps + warp] = a2; }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) { s1 += red[w]; s2 += red[kWarps + w]; }
  s1 /= (float)D; s2 /= (float)D;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) dx[i] = rstd * (dxh[i] - s1 - xh[i] * s2);
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
    // ===== R2: dp_act = dlogits[t] . W_zq^T   (Wt_zq [Z, D]) =====
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

    // ===== R3a (row owners): dp_pre = LN'SiLU'(dp_act) =====
    for (int b = blk; b < B; b += nblk) {
      const size_t bt = (size_t)b * T + t;
      float dy[4], x[4], dx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = threadIdx.x + kThreads * i;
        dy[i] = d < D ? __ldcg(p.dp_act + bt * D + d) : 0.f;
        x[i] = d < D ? p.p_pre[bt * D + d] : 0.f;
      }
      row_lnbwd(sm, dy, x, p.p_mean[bt], p.p_rstd[bt], p.post_g, p.post_b, D, dx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = threadIdx.x + kThreads * i;
        if (d < D) p.dp_pre[bt * D + d] = dx[i];
      }
    }
    grid.sync();

    // ===== R3: dh_tot = d_hz[t].h + dp_pre . W_post_h^T (Wt_post_h [D, G]); GRU gate gradients for the task's 32 units =====
    for (int task = blk; task < tasks_G; task += nblk) {
      const int c0 = task * 32;
      task_gemv(sm, p.Wt_post_h, G, c0, G, D, B, [&](int b, int k) { return __ldcg(p.dp_pre + ((size_t)b * T + t) * D + k); });
      for (int e = threadIdx.x; e < B * 32; e += kThreads) {
        const int b = e >> 5, jj = e & 31, j = c0 + jj;
        if (j < G) {
          const size_t bt = (size_t)b * T + t;
          const float d = sm.tile[b][jj] + __ldcg(p.d_hz + bt * W + j);
          const float* gt = p.gates + bt * 3 * G;
          const float z = gt[j], r = gt[G + j], c = gt[2 * G + j];
          const float dz = d * (p.h_in[bt * G + j] - c) * z * (1.f - z);
          const float dc = d * (1.f - z) * (1.f - c * c);
          const float dr = dc * p.gh[bt * 3 * G + 2 * G + j] * r * (1.f - r);
          p.dh_tot[(size_t)b * G + j] = d;
          float* gxo = p.dgx + bt * 3 * G;
          float* gho = p.dgh + bt * 3 * G;
          gxo[j] = dz; gxo[G + j] = dr; gxo[2 * G + j] = dc;
          gho[j] = dz; gho[G + j] = dr; gho[2 * G + j] = dc * r;
        }
      }
      __syncthreads();
    }
    grid.sync();

    // ===== R4: dh_in = dh_tot * z + dgh . U^T (Ut [3G, G]);  du = dgx . K^T (Kt [3G, D]) =====
    {
      const int ntask = tasks_G + tasks_D;
      for (int task = blk; task < ntask; task += nblk) {
        if (task < tasks_G) {
          const int c0 = task * 32;
          task_gemv(sm, p.Ut, G, c0, G, 3 * G, B, [&](int b, int k) { return __ldcg(p.dgh + ((size_t)b * T + t) * 3 * G + k); });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < G) {
              const size_t bt = (size_t)b * T + t;
              p.dh_in[bt * G + c0 + j] = sm.tile[b][j] + __ldcg(p.dh_tot + (size_t)b * G + c0 + j) * p.gates[bt * 3 * G + c0 + j];
            }
          }
          __syncthreads();
        } else {
          const int c0 = (task - tasks_G) * 32;
          task_gemv(sm, p.Kt, D, c0, D, 3 * G, B, [&](int b, int k) { return __ldcg(p.dgx + ((size_t)b * T + t) * 3 * G + k); });
          for (int e = threadIdx.x; e < B * 32; e += kThreads) {
            const int b = e >> 5, j = e & 31;
            if (c0 + j < D) p.du[((size_t)b * T + t) * D + c0 + j] = sm.tile[b][j];
          }
          __syncthreads();
        }
      }
    }
    grid.sync();

    // ===== R5a (row owners): dx_pre = LN'SiLU'(du); h part of the recurrent gradient: d_hz[t-1].h += (1 - f_t) dh_in[t] =====
    for (int b = blk; b < B; b += nblk) {
      const size_t bt = (size_t)b * T + t;
      float dy[4], x[4], dx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = threadIdx.x + kThreads * i;
        dy[i] = d < D ? __ldcg(p.du + bt * D + d) : 0.f;
        x[i] = d < D ? p.x_pre[bt * D + d] : 0.f;
      }
      row_lnbwd(sm, dy, x, p.x_mean[bt], p.x_rstd[bt], p.pre_g, p.pre_b, D, dx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = threadIdx.x + kThreads * i;
        if (d < D) p.dx_pre[bt * D + d] = dx[i];
      }
      if (t > 0) {
        const float keep = 1.f - p.is_first[bt];
        for (int j = threadIdx.x; j < G; j += kThreads) {
          float* dst = p.d_hz + (bt - 1) * W + j;
          *dst = __ldcg(dst) + keep * __ldcg(p.dh_in + bt * G + j);
        }
      }
    }
    grid.sync();

    // ===== R5: dz_in = dx_pre . W_pre[:Z]^T (Wt_pre_z [D, Z]); pass to step t-1 and run its posterior repr backward =====
    if (t > 0) {
      for (int task = blk; task < tasks_Z; task += nblk) {
        const int c0 = task * 32;
        task_gemv(sm, p.Wt_pre_z, Z, c0, Z, D, B, [&](int b, int k) { return __ldcg(p.dx_pre + ((size_t)b * T + t) * D + k); });
        for (int e = threadIdx.x; e < B * 32; e += kThreads) {
          const int b = e >> 5, j = e & 31;
          if (c0 + j < Z) {
            const size_t bt = (size_t)b * T + t;
            float* dst = p.d_hz + (bt - 1) * W + G + c0 + j;
            *dst = __ldcg(dst) + (1.f - p.is_first[bt]) * sm.tile[b][j];
          }
        }
        __syncthreads();
        const int cats = 32 / Zk;
        for (int pr = warp; pr < B * cats; pr += kWarps) {
          const int b = pr / cats, cat = c0 / Zk + pr % cats;
          if (cat < Zc) repr_bwd_pair(p, b, t - 1, cat, lane);
        }
        __syncthreads();
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
