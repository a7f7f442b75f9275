// Shared device pieces of the persistent scan kernels (forward: dv3_scan.cu, backward: dv3_scan_bwd.cu).
#pragma once
#include "dv3_device.cuh"
#include "dv3_scan.h"

namespace dv3 {
namespace scan {

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
};

// One task: out_tile[b][j] = sum_k in(b,k) * W[k*ldw + c0 + j], j < 32, k < K. `in_fn(b, k)` produces the input
// on the fly (LayerNorm / GRU gates are applied here by the consumer). Result is left in sm.tile (all threads
// synchronised on return).
template <class InFn>
__device__ __forceinline__ void task_gemv_col(Smem& sm, const float* __restrict__ W, int ldw, int col, int K, int B, InFn in_fn);

template <class InFn>
__device__ __forceinline__ void task_gemv(Smem& sm, const float* __restrict__ W, int ldw, int c0, int ncols, int K,
                                          int B, InFn in_fn) {
  const int lane = threadIdx.x & 31;
  task_gemv_col(sm, W, ldw, (c0 + lane) < ncols ? c0 + lane : -1, K, B, in_fn);
}

// Same, with the output column of each lane given explicitly (col < 0: lane idle): out_tile[b][lane] = sum_k in(b,k) W[k*ldw + col]
template <class InFn>
__device__ __forceinline__ void task_gemv_col(Smem& sm, const float* __restrict__ W, int ldw, int col, int K, int B, InFn in_fn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = (K + kWarps - 1) / kWarps;
  const int k0 = warp * ks, k1 = min(K, k0 + ks);
  const bool col_ok = col >= 0;
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
    const float* wp = W + (size_t)kc * ldw + (col_ok ? col : 0);
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

// ---- grid-wide barrier for the cooperative (co-resident) scan kernels --------------------------------------------
// One monotonically increasing counter in global memory (zeroed by the host before the launch): every CTA adds 1 and
// spins until the count reaches its next multiple of gridDim.x. Cheaper than cooperative_groups' grid.sync() (one
// fence + one atomic + an acquire-load spin by a single thread per CTA). Data exchanged across the barrier is read
// with __ldcg (L2), never through L1.
__device__ __forceinline__ void grid_bar(unsigned int* ctr, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (++spins > (1u << 24)) { printf("dv3 scan: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    } while (v < target);
  }
  __syncthreads();
}

// ---- contraction on activations already staged CTA-wide as in_s[k][kInLd] (k < K <= kStageK) -----------------------
// CW output columns per task (lane % CW), the 32 / CW lane groups and the 8 warps split K. Accumulates into acc[16].
constexpr int kStageK = kWarps * kKC;   // k lines that fit the staging buffer (in_chunk viewed as [kStageK][kInLd])

template <int CW>
__device__ __forceinline__ void gemv_acc(float (&acc)[kRows], const float* in_s, const float* __restrict__ W, int ldw, int col, int K) {
  constexpr int NPH = 32 / CW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = (K + kWarps - 1) / kWarps;
  const int k0 = warp * ks + lane / CW, k1 = min(K, (warp + 1) * ks);
  const bool col_ok = col >= 0;
  const float* wp = W + (col_ok ? col : 0);
  for (int kk0 = k0; kk0 < k1; kk0 += 8 * NPH) {
    float wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int k = kk0 + j * NPH; wv[j] = (col_ok && k < k1) ? wp[(size_t)k * ldw] : 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = min(kk0 + j * NPH, k1 - 1);
      const float4* iv = reinterpret_cast<const float4*>(in_s + (size_t)k * kInLd);
      const float4 a0 = iv[0], a1 = iv[1], a2 = iv[2], a3 = iv[3];
      const float w_ = wv[j];
      acc[0] = fmaf(a0.x, w_, acc[0]); acc[1] = fmaf(a0.y, w_, acc[1]); acc[2] = fmaf(a0.z, w_, acc[2]); acc[3] = fmaf(a0.w, w_, acc[3]);
      acc[4] = fmaf(a1.x, w_, acc[4]); acc[5] = fmaf(a1.y, w_, acc[5]); acc[6] = fmaf(a1.z, w_, acc[6]); acc[7] = fmaf(a1.w, w_, acc[7]);
      acc[8] = fmaf(a2.x, w_, acc[8]); acc[9] = fmaf(a2.y, w_, acc[9]); acc[10] = fmaf(a2.z, w_, acc[10]); acc[11] = fmaf(a2.w, w_, acc[11]);
      acc[12] = fmaf(a3.x, w_, acc[12]); acc[13] = fmaf(a3.y, w_, acc[13]); acc[14] = fmaf(a3.z, w_, acc[14]); acc[15] = fmaf(a3.w, w_, acc[15]);
    }
  }
}
// Reduces the lane groups and the 8 warps: sm.tile[b][j], j < CW (all threads synchronised on return).
template <int CW>
__device__ __forceinline__ void gemv_finish(Smem& sm, float (&acc)[kRows]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= CW; off >>= 1) {
#pragma unroll
    for (int b = 0; b < kRows; ++b) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], off);
  }
  if (lane < CW) {
#pragma unroll
    for (int b = 0; b < kRows; ++b) sm.red[warp][b][lane] = acc[b];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kRows * CW; e += kThreads) {
    const int b = e / CW, j = e % CW;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sm.red[w][b][j];
    sm.tile[b][j] = s;
  }
  __syncthreads();
}
// CTA-wide staging of in(b, k), k < K <= kStageK, into in_s[k][b] (rows >= B zero); lanes run along k (coalesced reads).
// All loads of the pass are issued before the first shared-memory store: one L2 round trip instead of K / 256.
template <int NIT, class InFn>
__device__ __forceinline__ void stage_rows_t(float* in_s, int K, int B, InFn in_fn) {
  float v[NIT][kRows];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int k = threadIdx.x + it * kThreads;
#pragma unroll
    for (int b = 0; b < kRows; ++b) v[it][b] = (b < B && k < K) ? in_fn(b, k) : 0.f;
  }
  asm volatile("" ::: "memory");
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int k = threadIdx.x + it * kThreads;
    if (k < K) {
      float4* dst = reinterpret_cast<float4*>(in_s + (size_t)k * kInLd);
      dst[0] = make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
      dst[1] = make_float4(v[it][4], v[it][5], v[it][6], v[it][7]);
      dst[2] = make_float4(v[it][8], v[it][9], v[it][10], v[it][11]);
      dst[3] = make_float4(v[it][12], v[it][13], v[it][14], v[it][15]);
    }
  }
  __syncthreads();
}
template <class InFn>
__device__ __forceinline__ void stage_rows(float* in_s, int K, int B, InFn in_fn) {
  if (K <= kThreads) stage_rows_t<1>(in_s, K, B, in_fn);
  else if (K <= 2 * kThreads) stage_rows_t<2>(in_s, K, B, in_fn);
  else if (K <= 3 * kThreads) stage_rows_t<3>(in_s, K, B, in_fn);
  else stage_rows_t<4>(in_s, K, B, in_fn);
}

// Forward LayerNorm + SiLU of rows b < B (x row b at x + b*ldx, read through L2), one warp per row, result staged as
// in_s[d][b] for the contraction that follows; optionally also written out (act / mean / rstd, row b at + b*ldo / b*lds).
// Every global load of the pass (NR rows per warp at once, gain and offset) is issued before the first reduction: the
// grid barrier invalidates L1, so each dependent load would cost a full L2 round trip.
template <int EPL, int NR>
__device__ __forceinline__ void cta_ln_silu_stage_t(float* in_s, const float* x, size_t ldx, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, int B, int D, float* act_out, size_t ldo,
                                                    float* mean_out, float* rstd_out, size_t lds) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int bb = warp; bb < kRows; bb += kWarps * NR) {
    float v[NR][EPL], g[EPL], be[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = lane + 32 * i;
      const bool ok = c < D;
      g[i] = ok ? gamma[c] : 0.f;
      be[i] = ok ? beta[c] : 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int b = bb + r * kWarps;
        v[r][i] = (ok && b < B) ? __ldcg(x + (size_t)b * ldx + c) : 0.f;
      }
    }
    asm volatile("" ::: "memory");   // keep every load of the pass ahead of the reductions
    float mean[NR], rstd[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) s += v[r][i];
      mean[r] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      mean[r] /= (float)D;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) { const float d = (lane + 32 * i < D) ? v[r][i] - mean[r] : 0.f; q += d * d; }
      rstd[r] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int b = bb + r * kWarps;
      if (b >= kRows) continue;
      rstd[r] = rsqrtf(rstd[r] / (float)D + DV3_LN_EPS);
      const bool row_ok = b < B, wr = act_out != nullptr && row_ok;
#pragma unroll
      for (int i = 0; i < EPL; ++i) {   // branch-free: columns beyond D have g = be = 0 -> o = 0, only the stores are predicated
        const int c = lane + 32 * i;
        const float n = (v[r][i] - mean[r]) * rstd[r] * g[i] + be[i];
        const float o = row_ok ? n * sigmoid_fast(n) : 0.f;
        if (c < D) in_s[(size_t)c * kInLd + b] = o;
        if (wr && c < D) act_out[(size_t)b * ldo + c] = o;
      }
      if (act_out != nullptr && lane == 0 && b < B) { mean_out[(size_t)b * lds] = mean[r]; rstd_out[(size_t)b * lds] = rstd[r]; }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void cta_ln_silu_stage(float* in_s, const float* x, size_t ldx, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, int B, int D, float* act_out, size_t ldo,
                                                  float* mean_out, float* rstd_out, size_t lds) {
  if (D <= 256) cta_ln_silu_stage_t<8, 2>(in_s, x, ldx, gamma, beta, B, D, act_out, ldo, mean_out, rstd_out, lds);
  else if (D <= 512) cta_ln_silu_stage_t<16, 2>(in_s, x, ldx, gamma, beta, B, D, act_out, ldo, mean_out, rstd_out, lds);
  else cta_ln_silu_stage_t<32, 2>(in_s, x, ldx, gamma, beta, B, D, act_out, ldo, mean_out, rstd_out, lds);
}

}  // namespace scan
}  // namespace dv3
