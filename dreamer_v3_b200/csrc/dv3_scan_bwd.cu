// Reverse-time BPTT through the RSSM observe scan as ONE cooperative launch (replaces the ~700 per-op launches of
// the reverse loop in Engine::wm_loss_backward). Only input gradients are on the recurrence; every weight gradient
// is a single N-row contraction after the kernel. Two kernels (launch_observe_scan_bwd picks one by layer width):
// observe_scan_bwd_kernel, described here, and the two-stage split-K observe_scan_bwd2_kernel further down (small
// layer sizes). Per step t (T-1 .. 0), 4 stages separated by grid barriers:
//   S1  dp_act = dlogits[t] . W_zq^T                                 (narrow column tasks: the K = Z contraction is the
//                                                                      longest of the step, so it is spread over many CTAs)
//   S2  dp_pre = LN'SiLU'(dp_act) recomputed by every task CTA (one warp per row, staged in shared memory);
//       dh_tot = d_hz[t].h + dp_pre . W_post[E:]^T, GRU gate gradients dgx / dgh of the task's units
//   S3  dh_in  = dh_tot*z + dgh . U^T  (and d_hz[t-1].h += (1 - is_first[t]) dh_in);  du = dgx . K^T
//   S4  dx_pre = LN'SiLU'(du) recomputed per task CTA; dz_in = dx_pre . W_pre[:Z]^T; d_hz[t-1].z += (1 - is_first[t]) dz_in,
//       then the straight-through / KL backward of the posterior categoricals of step t-1 right away (dlogits[t-1]).
// The LayerNorm backward needs whole-row statistics; instead of a separate row-owner stage (one more barrier) every
// consumer CTA redoes it for the 16 rows from data it has to stage anyway. The contractions read TRANSPOSED weight
// copies (refreshed once per update) so that weight reads stay coalesced.
#include <algorithm>
#include <cstdlib>

#include "dv3_scan_common.cuh"

namespace dv3 {
namespace {

using namespace scan;

// d(SiLU(LN(x)))/dx helpers: n = xhat*g + b ; dn = dy * silu'(n)
__device__ __forceinline__ float silu_grad(float n) {
  const float sg = sigmoid_fast(n);
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

// LayerNorm+SiLU backward of rows b < B by the whole CTA (one warp per row, NR rows per warp at once): dy row b at
// dy + b*ld (read through L2), x likewise; the result is staged as in_s[d][b] for the contraction that follows and
// optionally written out. All global loads are issued before the first reduction (the grid barrier invalidates L1).
template <int EPL, int NR>
__device__ __forceinline__ void cta_lnbwd_stage_t(float* in_s, const float* dy, const float* x, size_t ld, const float* mean,
                                                  const float* rstd, size_t lds, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, int B, int D, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int bb = warp; bb < kRows; bb += kWarps * NR) {
    float xh[NR][EPL], dxh[NR][EPL], g[EPL], be[EPL], mu[NR], rs[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int b = bb + r * kWarps;
      mu[r] = b < B ? mean[(size_t)b * lds] : 0.f;
      rs[r] = b < B ? rstd[(size_t)b * lds] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = lane + 32 * i;
      const bool ok = c < D;
      g[i] = ok ? gamma[c] : 0.f;
      be[i] = ok ? beta[c] : 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int b = bb + r * kWarps;
        xh[r][i] = (ok && b < B) ? x[(size_t)b * ld + c] : 0.f;
        dxh[r][i] = (ok && b < B) ? __ldcg(dy + (size_t)b * ld + c) : 0.f;
      }
    }
    asm volatile("" ::: "memory");   // keep every load of the pass ahead of the reductions
    float s1[NR], s2[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const float xhat = (xh[r][i] - mu[r]) * rs[r];
        const float d = dxh[r][i] * silu_grad(xhat * g[i] + be[i]) * g[i];   // 0 beyond D (g = 0) and beyond B (dy = 0)
        xh[r][i] = xhat; dxh[r][i] = d;
        a1 += d; a2 += d * xhat;
      }
      s1[r] = a1; s2[r] = a2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
        s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int b = bb + r * kWarps;
      if (b >= kRows) continue;
      const float m1 = s1[r] / (float)D, m2 = s2[r] / (float)D;
      const bool row_ok = b < B, wr = out != nullptr && row_ok;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const int c = lane + 32 * i;
        const float o = row_ok ? rs[r] * (dxh[r][i] - m1 - xh[r][i] * m2) : 0.f;
        if (c < D) in_s[(size_t)c * kInLd + b] = o;
        if (wr && c < D) out[(size_t)b * ld + c] = o;
      }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void cta_lnbwd_stage(float* in_s, const float* dy, const float* x, size_t ld, const float* mean,
                                                const float* rstd, size_t lds, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, int B, int D, float* out) {
  if (D <= 256) cta_lnbwd_stage_t<8, 2>(in_s, dy, x, ld, mean, rstd, lds, gamma, beta, B, D, out);
  else if (D <= 512) cta_lnbwd_stage_t<16, 2>(in_s, dy, x, ld, mean, rstd, lds, gamma, beta, B, D, out);
  else cta_lnbwd_stage_t<32, 1>(in_s, dy, x, ld, mean, rstd, lds, gamma, beta, B, D, out);
}

// columns per task of a contraction with `ncols` outputs: the narrowest of 32 / 16 / 8 that still gives every CTA work
__device__ __forceinline__ int pick_cw(int ncols, int nblk) {
  if ((ncols + 31) / 32 * 2 > nblk) return 32;
  if ((ncols + 15) / 16 * 2 > nblk) return 16;
  return 8;
}

// out_tile = in(b, 0..K) . Wt[:, c0 .. c0+CW): K staged in chunks of kStageK
// wc != nullptr: this CTA's slice [K][CW] of Wt is resident in shared memory
template <int CW, class InFn>
__device__ __forceinline__ void task_contract(Smem& sm, const float* __restrict__ Wt, int ldw, const float* wc, int c0, int ncols, int K, int B, InFn in_fn) {
  const int lane = threadIdx.x & 31;
  float* in_s = &sm.in_chunk[0][0][0];
  const int c = c0 + (lane % CW);
  float acc[kRows];
#pragma unroll
  for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
  for (int kc = 0; kc < K; kc += kStageK) {
    const int kn = min(kStageK, K - kc);
    stage_rows(in_s, kn, B, [&](int b, int k) { return in_fn(b, kc + k); });
    if (wc != nullptr) gemv_acc<CW>(acc, in_s, wc + (size_t)kc * CW, CW, c < ncols ? (lane % CW) : -1, kn);
    else gemv_acc<CW>(acc, in_s, Wt + (size_t)kc * ldw, ldw, c < ncols ? c : -1, kn);
    __syncthreads();
  }
  gemv_finish<CW>(sm, acc);
}
template <class InFn>
__device__ __forceinline__ void task_contract_cw(Smem& sm, int cw, const float* __restrict__ Wt, int ldw, const float* wc, int c0, int ncols, int K, int B, InFn in_fn) {
  if (cw == 32) task_contract<32>(sm, Wt, ldw, wc, c0, ncols, K, B, in_fn);
  else if (cw == 16) task_contract<16>(sm, Wt, ldw, wc, c0, ncols, K, B, in_fn);
  else task_contract<8>(sm, Wt, ldw, wc, c0, ncols, K, B, in_fn);
}
// copies columns [c0, c0 + cw) of Wt [K, ldw] into wc [K][cw] (zero beyond ncols)
__device__ __forceinline__ void cache_slice(float* wc, const float* __restrict__ Wt, int ldw, int c0, int cw, int ncols, int K) {
  for (int e = threadIdx.x; e < K * cw; e += kThreads) {
    const int k = e / cw, j = e % cw;
    wc[e] = (c0 + j < ncols) ? Wt[(size_t)k * ldw + c0 + j] : 0.f;
  }
}

__global__ void __launch_bounds__(kThreads, 1) observe_scan_bwd_kernel(ScanBwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  unsigned int bar_target = 0;
  const int B = p.B, T = p.T, G = p.G, D = p.D, Z = p.Z, Zc = p.Zc, Zk = p.Zk;
  const int W = G + Z;
  const int nblk = gridDim.x, blk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gtid = blk * kThreads + threadIdx.x, gthreads = nblk * kThreads;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
  const int tasks_G = (G + 31) / 32, tasks_Z = (Z + 31) / 32;
  const int cw1 = pick_cw(D, nblk), ntask1 = (D + cw1 - 1) / cw1;                 // S1: D output columns
  const int cw3 = pick_cw(G + D, nblk), nt3g = (G + cw3 - 1) / cw3, nt3d = (D + cw3 - 1) / cw3;   // S3: G + D output columns
  float* in_s = &sm.in_chunk[0][0][0];

  // ---- weight slices resident in shared memory for all T steps (launcher: one task per CTA per stage and they fit) ----
  const bool cached = p.cache_weights != 0;
  float* wc1 = reinterpret_cast<float*>(smem_raw + sizeof(Smem));   // [Z][cw1]
  float* wc2 = wc1 + (size_t)Z * cw1;                               // [D][32]
  float* wc3 = wc2 + (size_t)D * 32;                                // [3G][cw3]
  float* wc4 = wc3 + (size_t)3 * G * cw3;                           // [D][32]
  if (cached) {
    if (blk < ntask1) cache_slice(wc1, p.Wt_zq, D, blk * cw1, cw1, D, Z);
    if (blk < tasks_G) cache_slice(wc2, p.Wt_post_h, G, blk * 32, 32, G, D);
    if (blk < nt3g) cache_slice(wc3, p.Ut, G, blk * cw3, cw3, G, 3 * G);
    else if (blk < nt3g + nt3d) cache_slice(wc3, p.Kt, D, (blk - nt3g) * cw3, cw3, D, 3 * G);
    if (blk < tasks_Z) cache_slice(wc4, p.Wt_pre_z, Z, blk * 32, 32, Z, D);
    __syncthreads();
  }

  // prologue: dlogits of the last step (no recurrent contribution yet)
  for (int pr = gwarp; pr < B * Zc; pr += gwarps) repr_bwd_pair(p, pr / Zc, T - 1, pr % Zc, lane);
  grid_bar(p.bar, bar_target);

  auto stamp = [&](int i) {
    if (p.stamps != nullptr && blk == 0 && threadIdx.x == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  stamp(0);
  for (int t = T - 1; t >= 0; --t) {
    const int si = (T - 1 - t) * 8;
    // ===== S1: dp_act = dlogits[t] . W_zq^T   (Wt_zq [Z, D]) =====
    for (int task = blk; task < ntask1; task += nblk) {
      const int c0 = task * cw1;
      task_contract_cw(sm, cw1, p.Wt_zq, D, cached ? wc1 : nullptr, c0, D, Z, B, [&](int b, int k) { return __ldcg(p.dlogits + ((size_t)b * T + t) * Z + k); });
      for (int e = threadIdx.x; e < B * cw1; e += kThreads) {
        const int b = e / cw1, j = e % cw1;
        if (c0 + j < D) p.dp_act[((size_t)b * T + t) * D + c0 + j] = sm.tile[b][j];
      }
      __syncthreads();
    }
    stamp(si + 1);
    grid_bar(p.bar, bar_target);
    stamp(si + 2);

    // ===== S2: dp_pre = LN'SiLU'(dp_act) (per task CTA); dh_tot = d_hz[t].h + dp_pre . W_post_h^T (Wt_post_h [D, G]);
    //       GRU gate gradients for the task's 32 units =====
    for (int task = blk; task < tasks_G; task += nblk) {
      const int c0 = task * 32;
      // operands of the epilogue do not depend on the contraction: issue their loads now, one L2 round trip earlier
      float e_d[2], e_z[2], e_r[2], e_c[2], e_h[2], e_g[2];
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads, b = e >> 5, j = c0 + (e & 31);
        const bool ok = b < B && j < G;
        const size_t bt = (size_t)(ok ? b : 0) * T + t;
        const int jj = ok ? j : 0;
        e_d[it] = ok ? __ldcg(p.d_hz + bt * W + jj) : 0.f;
        e_z[it] = ok ? p.gates[bt * 3 * G + jj] : 0.f;
        e_r[it] = ok ? p.gates[bt * 3 * G + G + jj] : 0.f;
        e_c[it] = ok ? p.gates[bt * 3 * G + 2 * G + jj] : 0.f;
        e_h[it] = ok ? p.h_in[bt * G + jj] : 0.f;
        e_g[it] = ok ? p.gh[bt * 3 * G + 2 * G + jj] : 0.f;
      }
      cta_lnbwd_stage(in_s, p.dp_act + (size_t)t * D, p.p_pre + (size_t)t * D, (size_t)T * D, p.p_mean + t, p.p_rstd + t, (size_t)T,
                      p.post_g, p.post_b, B, D, task == 0 ? p.dp_pre + (size_t)t * D : nullptr);
      float acc[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
      if (cached) gemv_acc<32>(acc, in_s, wc2, 32, (c0 + lane) < G ? lane : -1, D);
      else gemv_acc<32>(acc, in_s, p.Wt_post_h, G, (c0 + lane) < G ? c0 + lane : -1, D);
      gemv_finish<32>(sm, acc);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads;
        const int b = e >> 5, jj = e & 31, j = c0 + jj;
        if (b < B && j < G) {
          const size_t bt = (size_t)b * T + t;
          const float d = sm.tile[b][jj] + e_d[it];
          const float z = e_z[it], r = e_r[it], c = e_c[it];
          const float dz = d * (e_h[it] - c) * z * (1.f - z);
          const float dc = d * (1.f - z) * (1.f - c * c);
          const float dr = dc * e_g[it] * r * (1.f - r);
          p.dh_tot[(size_t)b * G + j] = d;
          float* gxo = p.dgx + bt * 3 * G;
          float* gho = p.dgh + bt * 3 * G;
          gxo[j] = dz; gxo[G + j] = dr; gxo[2 * G + j] = dc;
          gho[j] = dz; gho[G + j] = dr; gho[2 * G + j] = dc * r;
        }
      }
      __syncthreads();
    }
    stamp(si + 3);
    grid_bar(p.bar, bar_target);
    stamp(si + 4);

    // ===== S3: dh_in = dh_tot * z + dgh . U^T (Ut [3G, G]), passed on to d_hz[t-1].h;  du = dgx . K^T (Kt [3G, D]) =====
    for (int task = blk; task < nt3g + nt3d; task += nblk) {
      if (task < nt3g) {
        const int c0 = task * cw3;
        task_contract_cw(sm, cw3, p.Ut, G, cached ? wc3 : nullptr, c0, G, 3 * G, B, [&](int b, int k) { return __ldcg(p.dgh + ((size_t)b * T + t) * 3 * G + k); });
        for (int e = threadIdx.x; e < B * cw3; e += kThreads) {
          const int b = e / cw3, j = c0 + e % cw3;
          if (j < G) {
            const size_t bt = (size_t)b * T + t;
            const float v = sm.tile[b][e % cw3] + __ldcg(p.dh_tot + (size_t)b * G + j) * p.gates[bt * 3 * G + j];
            p.dh_in[bt * G + j] = v;
            if (t > 0) {
              float* dst = p.d_hz + (bt - 1) * W + j;
              *dst = __ldcg(dst) + (1.f - p.is_first[bt]) * v;
            }
          }
        }
        __syncthreads();
      } else {
        const int c0 = (task - nt3g) * cw3;
        task_contract_cw(sm, cw3, p.Kt, D, cached ? wc3 : nullptr, c0, D, 3 * G, B, [&](int b, int k) { return __ldcg(p.dgx + ((size_t)b * T + t) * 3 * G + k); });
        for (int e = threadIdx.x; e < B * cw3; e += kThreads) {
          const int b = e / cw3, j = c0 + e % cw3;
          if (j < D) p.du[((size_t)b * T + t) * D + j] = sm.tile[b][e % cw3];
        }
        __syncthreads();
      }
    }
    stamp(si + 5);
    grid_bar(p.bar, bar_target);
    stamp(si + 6);

    // ===== S4: dx_pre = LN'SiLU'(du) (per task CTA); dz_in = dx_pre . W_pre[:Z]^T (Wt_pre_z [D, Z]); pass to step t-1 and
    //       run its posterior repr backward. At t = 0 only dx_pre is needed (weight gradients). =====
    for (int task = blk; task < (t > 0 ? tasks_Z : 1); task += nblk) {
      const int c0 = task * 32;
      // epilogue / categorical-backward operands (independent of the contraction): loads issued up front
      float e_old[2], e_keep[2], r_l[2], r_dp[2];
      const int cats = 32 / Zk;
      if (t > 0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int e = threadIdx.x + it * kThreads, b = e >> 5, j = e & 31;
          const bool ok = b < B && c0 + j < Z;
          const size_t bt = (size_t)(ok ? b : 0) * T + t;
          e_old[it] = ok ? __ldcg(p.d_hz + (bt - 1) * W + G + c0 + (ok ? j : 0)) : 0.f;
          e_keep[it] = ok ? 1.f - p.is_first[bt] : 0.f;
          const int pr = warp + it * kWarps, rb = pr / cats, cat = c0 / Zk + pr % cats;
          const bool rok = pr < B * cats && cat < Zc && lane < Zk;
          const size_t rbt = (size_t)(rok ? rb : 0) * T + (t - 1);
          const int col = rok ? cat * Zk + lane : 0;
          r_l[it] = rok ? p.post_logits[rbt * Z + col] : -INFINITY;
          r_dp[it] = rok ? p.dpost[rbt * Z + col] : 0.f;
        }
      }
      cta_lnbwd_stage(in_s, p.du + (size_t)t * D, p.x_pre + (size_t)t * D, (size_t)T * D, p.x_mean + t, p.x_rstd + t, (size_t)T,
                      p.pre_g, p.pre_b, B, D, task == 0 ? p.dx_pre + (size_t)t * D : nullptr);
      if (t == 0) break;
      float acc[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
      if (cached) gemv_acc<32>(acc, in_s, wc4, 32, (c0 + lane) < Z ? lane : -1, D);
      else gemv_acc<32>(acc, in_s, p.Wt_pre_z, Z, (c0 + lane) < Z ? c0 + lane : -1, D);
      gemv_finish<32>(sm, acc);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads;
        const int b = e >> 5, j = e & 31;
        if (b < B && c0 + j < Z) {
          const size_t bt = (size_t)b * T + t;
          const float v = e_old[it] + e_keep[it] * sm.tile[b][j];
          p.d_hz[(bt - 1) * W + G + c0 + j] = v;
          sm.tile[b][j] = v;      // the categorical backward below takes the updated gradient from here, not from L2
        }
      }
      __syncthreads();
      // straight-through / KL backward of the step t-1 categoricals of this 32-column block (B * cats <= 16 pairs, 2 per warp)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int pr = warp + it * kWarps;
        if (pr < B * cats) {
          const int b = pr / cats, cl_ = pr % cats, cat = c0 / Zk + cl_;
          if (cat < Zc) {
            const bool act = lane < Zk;
            const int col = cat * Zk + lane;
            const float l = r_l[it];
            const float mx = warp_max(l);
            const float ex = act ? expf(l - mx) : 0.f;
            const float sden = warp_sum(ex);
            const float prb = ex / sden;
            const float g = act ? 0.99f * (sm.tile[b][cl_ * Zk + lane] + r_dp[it]) : 0.f;
            const float dot = warp_sum(g * prb);
            if (act) p.dlogits[((size_t)b * T + (t - 1)) * Z + col] = prb * (g - dot);
          }
        }
      }
      for (int pr = warp + 2 * kWarps; pr < B * cats; pr += kWarps) {   // narrow categoricals (Zk < 32): more than 16 pairs per block
        const int b = pr / cats, cat = c0 / Zk + pr % cats;
        if (cat < Zc) repr_bwd_pair(p, b, t - 1, cat, lane);
      }
      __syncthreads();
    }
    stamp(si + 7);
    if (t > 0) grid_bar(p.bar, bar_target);
    stamp(si + 8);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-stage version: every contraction whose K runs over what the previous stage has just produced is done SPLIT-K by
// the producer itself - the task that owns 32 dlogit columns (or cwx GRU units) multiplies its own K-slice with the
// matching weight rows and adds the [B x all-columns] partial product to the consumer's input with fp32 atomics (RED
// to L2). That removes the S1 and S3 stages of the kernel above together with their grid barriers: two barriers and two
// latency chains per step instead of four.
//   X(t)  dp_pre = LN'SiLU'(dp_act[t]); dh_tot = d_hz[t].h (+ recurrent part of step t+1) + dp_pre . W_post_h^T; GRU gate
//         gradients of the task's units; then dh_in[t] += dgh_task . Ut[task rows, :],  du[t] += dgx_task . Kt[task rows, :]
//   Y(t)  dx_pre = LN'SiLU'(du[t]); dz_in = dx_pre . W_pre[:Z]^T for the task's 32 columns -> d_hz[t-1].z, categorical
//         backward of step t-1 -> dlogits[t-1]; then dp_act[t-1] += dlogits_task . Wt_zq[task rows, :]
// dp_act, du and dh_in are accumulators: the launcher zeroes them. Summation order inside them is not fixed (atomics);
// the values agree with the four-stage kernel to fp32 rounding.

// out[b*ldo + c] += sum_{k < K} in_s[k][b] * wrow(k)[c], c < ncols: thread = output column (coalesced weight rows),
// the 16 rows of a k are one broadcast LDS.128 quad
template <class RowFn>
__device__ __forceinline__ void partial_accumulate(const float* in_s, int K, RowFn wrow, int ncols, float* out, size_t ldo, int B) {
  for (int c = threadIdx.x; c < ncols; c += kThreads) {
    float acc[kRows];
#pragma unroll
    for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float w_ = wrow(k)[c];
      const float4* iv = reinterpret_cast<const float4*>(in_s + (size_t)k * kInLd);
      const float4 a0 = iv[0], a1 = iv[1], a2 = iv[2], a3 = iv[3];
      acc[0] = fmaf(a0.x, w_, acc[0]); acc[1] = fmaf(a0.y, w_, acc[1]); acc[2] = fmaf(a0.z, w_, acc[2]); acc[3] = fmaf(a0.w, w_, acc[3]);
      acc[4] = fmaf(a1.x, w_, acc[4]); acc[5] = fmaf(a1.y, w_, acc[5]); acc[6] = fmaf(a1.z, w_, acc[6]); acc[7] = fmaf(a1.w, w_, acc[7]);
      acc[8] = fmaf(a2.x, w_, acc[8]); acc[9] = fmaf(a2.y, w_, acc[9]); acc[10] = fmaf(a2.z, w_, acc[10]); acc[11] = fmaf(a2.w, w_, acc[11]);
      acc[12] = fmaf(a3.x, w_, acc[12]); acc[13] = fmaf(a3.y, w_, acc[13]); acc[14] = fmaf(a3.z, w_, acc[14]); acc[15] = fmaf(a3.w, w_, acc[15]);
    }
#pragma unroll
    for (int b = 0; b < kRows; ++b)
      if (b < B) atomicAdd(out + (size_t)b * ldo + c, acc[b]);
  }
}

// Categorical backward of the (row, categorical) pairs of the 32-column block c0 at step tt, gradient w.r.t. z taken from
// dz(b, col) ; dlogits -> global and, transposed, into dl_s[col - c0][b] for the split-K contraction that follows.
// The first two pairs of every warp (all of them when a categorical has 32 classes) take their logits / KL gradient from
// `pre` (repr_prefetch, issued before the contraction: one L2 round trip off the critical path).
struct ReprPre { float l[2], dp[2]; };
__device__ __forceinline__ ReprPre repr_prefetch(const ScanBwdParams& p, int tt, int c0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Zk = p.Zk, Z = p.Z, cats = 32 / Zk;
  ReprPre r;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int pr = warp + it * kWarps, b = pr / cats, cat = c0 / Zk + pr % cats;
    const bool act = pr < kRows * cats && b < p.B && cat < p.Zc && lane < Zk;
    const size_t bt = (size_t)(act ? b : 0) * p.T + tt;
    const int col = act ? cat * Zk + lane : 0;
    r.l[it] = act ? p.post_logits[bt * Z + col] : -INFINITY;
    r.dp[it] = act ? p.dpost[bt * Z + col] : 0.f;
  }
  return r;
}
template <class DzFn>
__device__ __forceinline__ void repr_bwd_block(const ScanBwdParams& p, int tt, int c0, float* dl_s, const ReprPre& pre, DzFn dz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Zk = p.Zk, Z = p.Z, cats = 32 / Zk;
  int it = 0;
  for (int pr = warp; pr < kRows * cats; pr += kWarps, ++it) {
    const int b = pr / cats, cl_ = pr % cats, cat = c0 / Zk + cl_;
    const bool rok = b < p.B && cat < p.Zc;
    const bool act = rok && lane < Zk;
    const int col = cat * Zk + lane;
    const size_t bt = (size_t)(rok ? b : 0) * p.T + tt;
    const float l = it < 2 ? (it == 0 ? pre.l[0] : pre.l[1]) : (act ? p.post_logits[bt * Z + col] : -INFINITY);
    const float dpv = it < 2 ? (it == 0 ? pre.dp[0] : pre.dp[1]) : (act ? p.dpost[bt * Z + col] : 0.f);
    const float mx = warp_max(l);
    const float ex = act ? expf(l - mx) : 0.f;
    const float sden = warp_sum(ex);
    const float prb = rok ? ex / sden : 0.f;
    const float g = act ? 0.99f * (dz(b, col) + dpv) : 0.f;
    const float dot = warp_sum(g * prb);
    const float dl = act ? prb * (g - dot) : 0.f;
    if (act) p.dlogits[bt * Z + col] = dl;
    if (lane < Zk && cl_ * Zk + lane < 32) dl_s[(size_t)(cl_ * Zk + lane) * kInLd + b] = dl;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) observe_scan_bwd2_kernel(ScanBwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  unsigned int bar_target = 0;
  const int B = p.B, T = p.T, G = p.G, D = p.D, Z = p.Z;
  const int W = G + Z;
  const int nblk = gridDim.x, blk = blockIdx.x;
  const int tasks_Z = (Z + 31) / 32;
  const int cwx = p.cwx, ntx = (G + cwx - 1) / cwx;   // X: GRU units per task
  float* in_s = &sm.in_chunk[0][0][0];

  // ---- weight slices resident in shared memory for all T steps (launcher: one task per CTA per stage and they fit) ----
  const bool cached = p.cache_weights != 0;
  float* wcP = reinterpret_cast<float*>(smem_raw + sizeof(Smem));   // Wt_post_h[:, my cwx units]   [D][cwx]
  float* wcU = wcP + (size_t)D * cwx;                               // Ut rows of my units           [3 cwx][G]
  float* wcK = wcU + (size_t)3 * cwx * G;                           // Kt rows of my units           [3 cwx][D]
  float* wcX = wcK + (size_t)3 * cwx * D;                           // Wt_pre_z[:, my 32 columns]    [D][32]
  float* wcQ = wcX + (size_t)D * 32;                                // Wt_zq rows of my 32 columns   [32][D]
  if (cached) {
    if (blk < ntx) {
      cache_slice(wcP, p.Wt_post_h, G, blk * cwx, cwx, G, D);
      for (int e = threadIdx.x; e < 3 * cwx * G; e += kThreads) {
        const int k = e / G, c = e % G, u = blk * cwx + k % cwx;
        wcU[e] = u < G ? p.Ut[(size_t)((k / cwx) * G + u) * G + c] : 0.f;
      }
      for (int e = threadIdx.x; e < 3 * cwx * D; e += kThreads) {
        const int k = e / D, c = e % D, u = blk * cwx + k % cwx;
        wcK[e] = u < G ? p.Kt[(size_t)((k / cwx) * G + u) * D + c] : 0.f;
      }
    }
    if (blk < tasks_Z) {
      cache_slice(wcX, p.Wt_pre_z, Z, blk * 32, 32, Z, D);
      for (int e = threadIdx.x; e < 32 * D; e += kThreads) {
        const int k = e / D, c = e % D;
        wcQ[e] = blk * 32 + k < Z ? p.Wt_zq[(size_t)(blk * 32 + k) * D + c] : 0.f;
      }
    }
    __syncthreads();
  }
  auto stamp = [&](int i) {
    if (p.stamps != nullptr && blk == 0 && threadIdx.x == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  // dlogits[tt] of a 32-column block are in in_s rows [0, 32): add their share of dp_act[tt]
  auto dp_act_partial = [&](int tt, int c0) {
    const int kq = min(32, Z - c0);
    if (cached) partial_accumulate(in_s, kq, [&](int k) { return wcQ + (size_t)k * D; }, D, p.dp_act + (size_t)tt * D, (size_t)T * D, B);
    else partial_accumulate(in_s, kq, [&](int k) { return p.Wt_zq + (size_t)(c0 + k) * D; }, D, p.dp_act + (size_t)tt * D, (size_t)T * D, B);
    __syncthreads();
  };

  // prologue: dlogits of the last step (no recurrent contribution yet) and their dp_act
  for (int task = blk; task < tasks_Z; task += nblk) {
    const int c0 = task * 32;
    repr_bwd_block(p, T - 1, c0, in_s, repr_prefetch(p, T - 1, c0), [&](int b, int col) { return __ldcg(p.d_hz + ((size_t)b * T + (T - 1)) * W + G + col); });
    dp_act_partial(T - 1, c0);
  }
  grid_bar(p.bar, bar_target);
  stamp(0);

  for (int t = T - 1; t >= 0; --t) {
    const int si = (T - 1 - t) * 8;
    // ===== X(t) =====
    for (int task = blk; task < ntx; task += nblk) {
      const int c0 = task * cwx;
      // epilogue operands (independent of the contraction): loads issued up front. thread -> (row b, unit jj), 2 passes
      float e_d[2], e_z[2], e_r[2], e_c[2], e_h[2], e_g[2];
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads, b = e / cwx, j = c0 + e % cwx;
        const bool ok = e < kRows * cwx && b < B && j < G;
        const size_t bt = (size_t)(ok ? b : 0) * T + t;
        const int jj = ok ? j : 0;
        float d = ok ? __ldcg(p.d_hz + bt * W + jj) : 0.f;
        if (ok && t + 1 < T) d += (1.f - p.is_first[bt + 1]) * __ldcg(p.dh_in + (bt + 1) * G + jj);   // recurrent part (X(t+1), complete)
        e_d[it] = d;
        e_z[it] = ok ? p.gates[bt * 3 * G + jj] : 0.f;
        e_r[it] = ok ? p.gates[bt * 3 * G + G + jj] : 0.f;
        e_c[it] = ok ? p.gates[bt * 3 * G + 2 * G + jj] : 0.f;
        e_h[it] = ok ? p.h_in[bt * G + jj] : 0.f;
        e_g[it] = ok ? p.gh[bt * 3 * G + 2 * G + jj] : 0.f;
      }
      cta_lnbwd_stage(in_s, p.dp_act + (size_t)t * D, p.p_pre + (size_t)t * D, (size_t)T * D, p.p_mean + t, p.p_rstd + t, (size_t)T,
                      p.post_g, p.post_b, B, D, task == 0 ? p.dp_pre + (size_t)t * D : nullptr);
      float acc[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
      const int lane = threadIdx.x & 31;
      if (cwx == 32) {
        if (cached) gemv_acc<32>(acc, in_s, wcP, 32, (c0 + lane) < G ? lane : -1, D);
        else gemv_acc<32>(acc, in_s, p.Wt_post_h, G, (c0 + lane) < G ? c0 + lane : -1, D);
        gemv_finish<32>(sm, acc);
      } else if (cwx == 16) {
        if (cached) gemv_acc<16>(acc, in_s, wcP, 16, (c0 + lane % 16) < G ? lane % 16 : -1, D);
        else gemv_acc<16>(acc, in_s, p.Wt_post_h, G, (c0 + lane % 16) < G ? c0 + lane % 16 : -1, D);
        gemv_finish<16>(sm, acc);
      } else {
        if (cached) gemv_acc<8>(acc, in_s, wcP, 8, (c0 + lane % 8) < G ? lane % 8 : -1, D);
        else gemv_acc<8>(acc, in_s, p.Wt_post_h, G, (c0 + lane % 8) < G ? c0 + lane % 8 : -1, D);
        gemv_finish<8>(sm, acc);
      }
      // gate gradients -> global (weight gradients after the kernel) and, as [k][row], into the staging buffer:
      // rows [0, 3 cwx) = dgh (z, r, c blocks of cwx), rows [3 cwx, 6 cwx) = dgx   (dp_pre in in_s is dead now)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads;
        if (e < kRows * cwx) {
          const int b = e / cwx, jj = e % cwx, j = c0 + jj;
          float dz = 0.f, dr = 0.f, dc = 0.f, dcr = 0.f;
          if (b < B && j < G) {
            const size_t bt = (size_t)b * T + t;
            const float d = sm.tile[b][jj] + e_d[it];
            const float z = e_z[it], r = e_r[it], c = e_c[it];
            dz = d * (e_h[it] - c) * z * (1.f - z);
            dc = d * (1.f - z) * (1.f - c * c);
            dr = dc * e_g[it] * r * (1.f - r);
            dcr = dc * r;
            p.d_hz[bt * W + j] = e_d[it];
            float* gxo = p.dgx + bt * 3 * G;
            float* gho = p.dgh + bt * 3 * G;
            gxo[j] = dz; gxo[G + j] = dr; gxo[2 * G + j] = dc;
            gho[j] = dz; gho[G + j] = dr; gho[2 * G + j] = dcr;
            atomicAdd(p.dh_in + bt * G + j, d * z);   // the direct path h_in -> h of the GRU cell
          }
          in_s[(size_t)(0 * cwx + jj) * kInLd + b] = dz;
          in_s[(size_t)(1 * cwx + jj) * kInLd + b] = dr;
          in_s[(size_t)(2 * cwx + jj) * kInLd + b] = dcr;
          in_s[(size_t)(3 * cwx + jj) * kInLd + b] = dz;
          in_s[(size_t)(4 * cwx + jj) * kInLd + b] = dr;
          in_s[(size_t)(5 * cwx + jj) * kInLd + b] = dc;
        }
      }
      __syncthreads();
      // split-K: this task's 3 cwx gate rows of U^T / K^T
      auto urow = [&](int k) { return p.Ut + (size_t)((k / cwx) * G + min(c0 + k % cwx, G - 1)) * G; };
      auto krow = [&](int k) { return p.Kt + (size_t)((k / cwx) * G + min(c0 + k % cwx, G - 1)) * D; };
      if (cached) {
        partial_accumulate(in_s, 3 * cwx, [&](int k) { return wcU + (size_t)k * G; }, G, p.dh_in + (size_t)t * G, (size_t)T * G, B);
        partial_accumulate(in_s + (size_t)3 * cwx * kInLd, 3 * cwx, [&](int k) { return wcK + (size_t)k * D; }, D, p.du + (size_t)t * D, (size_t)T * D, B);
      } else {
        partial_accumulate(in_s, 3 * cwx, urow, G, p.dh_in + (size_t)t * G, (size_t)T * G, B);
        partial_accumulate(in_s + (size_t)3 * cwx * kInLd, 3 * cwx, krow, D, p.du + (size_t)t * D, (size_t)T * D, B);
      }
      __syncthreads();
    }
    stamp(si + 1); stamp(si + 2);
    stamp(si + 3);
    grid_bar(p.bar, bar_target);
    stamp(si + 4);

    // ===== Y(t) (at t = 0 only dx_pre is needed, for the weight gradients) =====
    for (int task = blk; task < (t > 0 ? tasks_Z : 1); task += nblk) {
      const int c0 = task * 32;
      float e_old[2], e_keep[2];
      if (t > 0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int e = threadIdx.x + it * kThreads, b = e >> 5, j = e & 31;
          const bool ok = b < B && c0 + j < Z;
          const size_t bt = (size_t)(ok ? b : 0) * T + t;
          e_old[it] = ok ? __ldcg(p.d_hz + (bt - 1) * W + G + c0 + (ok ? j : 0)) : 0.f;
          e_keep[it] = ok ? 1.f - p.is_first[bt] : 0.f;
        }
      }
      const ReprPre rpre = repr_prefetch(p, t > 0 ? t - 1 : 0, c0);
      cta_lnbwd_stage(in_s, p.du + (size_t)t * D, p.x_pre + (size_t)t * D, (size_t)T * D, p.x_mean + t, p.x_rstd + t, (size_t)T,
                      p.pre_g, p.pre_b, B, D, task == 0 ? p.dx_pre + (size_t)t * D : nullptr);
      if (t == 0) break;
      float acc[kRows];
#pragma unroll
      for (int b = 0; b < kRows; ++b) acc[b] = 0.f;
      const int lane = threadIdx.x & 31;
      if (cached) gemv_acc<32>(acc, in_s, wcX, 32, (c0 + lane) < Z ? lane : -1, D);
      else gemv_acc<32>(acc, in_s, p.Wt_pre_z, Z, (c0 + lane) < Z ? c0 + lane : -1, D);
      gemv_finish<32>(sm, acc);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int e = threadIdx.x + it * kThreads;
        const int b = e >> 5, j = e & 31;
        if (b < B && c0 + j < Z) {
          const float v = e_old[it] + e_keep[it] * sm.tile[b][j];
          p.d_hz[((size_t)b * T + t - 1) * W + G + c0 + j] = v;
          sm.tile[b][j] = v;      // the categorical backward below takes the updated gradient from here, not from L2
        }
      }
      __syncthreads();
      repr_bwd_block(p, t - 1, c0, in_s, rpre, [&](int b, int col) { return sm.tile[b][col - c0]; });
      dp_act_partial(t - 1, c0);
    }
    stamp(si + 5); stamp(si + 6);
    stamp(si + 7);
    if (t > 0) grid_bar(p.bar, bar_target);
    stamp(si + 8);
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
  ScanBwdParams pp = p;
  if (const char* f = std::getenv("DV3_SCAN_BWD_CTAS")) { const int n = std::atoi(f); if (n > 0 && n < ctas) ctas = n; }
  // Two-stage split-K kernel for the small layer sizes (XS: 17.9 vs 27.3 us per step). Its atomics grow with
  // tasks x 16 x (G + D): at S (1 M per step) it ties with the four-stage kernel, at M (3.4 M) it loses (C4: 34.1 vs 31.9 ms
  // per update), so the larger sizes stay on the four-stage kernel. DV3_SCAN_BWD_V1 / _V2 force one of them (tests, A / B).
  const bool v1 = std::getenv("DV3_SCAN_BWD_V1") != nullptr || ((p.G > 256 || p.D > 256) && std::getenv("DV3_SCAN_BWD_V2") == nullptr);
  auto cw_of = [&](int ncols) { return ((ncols + 31) / 32 * 2 > ctas) ? 32 : (((ncols + 15) / 16 * 2 > ctas) ? 16 : 8); };
  void* args[] = {&pp};
  cudaMemsetAsync(pp.bar, 0, sizeof(unsigned int), s);
  if (!v1) {
    // accumulators of the split-K contractions
    const size_t n = (size_t)p.B * p.T;
    cudaMemsetAsync(pp.dp_act, 0, n * p.D * sizeof(float), s);
    cudaMemsetAsync(pp.du, 0, n * p.D * sizeof(float), s);
    cudaMemsetAsync(pp.dh_in, 0, n * p.G * sizeof(float), s);
    const int tz = (p.Z + 31) / 32;
    // X task width: the narrowest of 32 / 16 / 8 units that still gives every CTA work (DV3_SCAN_BWD_CW overrides: probing aid)
    int cwx = cw_of(p.G);
    if (const char* f = std::getenv("DV3_SCAN_BWD_CW")) { const int c = std::atoi(f); if (c == 8 || c == 16 || c == 32) cwx = c; }
    const int ntx = (p.G + cwx - 1) / cwx;
    pp.cwx = cwx;
    // CTAs beyond the widest stage only lengthen the grid barrier
    if (std::max(ntx, tz) < ctas) ctas = std::max(ntx, tz);
    const size_t cache = ((size_t)p.D * cwx + (size_t)3 * cwx * p.G + (size_t)3 * cwx * p.D + (size_t)p.D * 32 + (size_t)32 * p.D) * sizeof(float);
    const bool fits = sizeof(Smem) + cache <= 227 * 1024 && ntx <= ctas && tz <= ctas && std::getenv("DV3_SCAN_NO_CACHE") == nullptr;
    pp.cache_weights = fits ? 1 : 0;
    const size_t smem = sizeof(Smem) + (fits ? cache : 0);
    cudaFuncSetAttribute(observe_scan_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_bwd2_kernel, dim3(ctas), dim3(kThreads), args, smem, s);
    if (e == cudaSuccess) count_launch();
    return e;
  }
  // same task-width choice as the kernel (pick_cw), to size the shared-memory weight cache
  const int cw1 = cw_of(p.D), cw3 = cw_of(p.G + p.D);
  const int nt1 = (p.D + cw1 - 1) / cw1, nt3 = (p.G + cw3 - 1) / cw3 + (p.D + cw3 - 1) / cw3;
  const size_t cache = ((size_t)p.Z * cw1 + (size_t)p.D * 32 * 2 + (size_t)3 * p.G * cw3) * sizeof(float);
  const bool fits = sizeof(Smem) + cache <= 227 * 1024 && nt1 <= ctas && nt3 <= ctas && (p.G + 31) / 32 <= ctas && (p.Z + 31) / 32 <= ctas &&
                    std::getenv("DV3_SCAN_NO_CACHE") == nullptr;
  pp.cache_weights = fits ? 1 : 0;
  const size_t smem = sizeof(Smem) + (fits ? cache : 0);
  cudaFuncSetAttribute(observe_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_bwd_kernel, dim3(ctas), dim3(kThreads), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
