// fp32 SIMT GEMM for sm_100a: the exact-precision path (1e-4 parity mode) and the small-M scan steps.
// C[M,N] (+)= op(A) op(B) (+ bias). Shared-memory tiled, register-blocked, float4 global loads where
// the operand is 16-byte aligned, split-K with fp32 atomics when the tile grid is smaller than the GPU.
#include <cstdlib>

#include <cstdio>
#include <cstdlib>

#include "dv3_common.h"
#include "dv3_device.cuh"

namespace dv3 {

long long g_launches = 0;

namespace {

constexpr int kNumSMs = 148;

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(GemmArgs g, int k_per_split, int vecA, int vecB, int use_atomic) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  pdl_trigger();
  pdl_wait();

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile into As[k][m] ----
    if (!TA) {
      // A[m*lda + k], k contiguous: float4 along k
      for (int e = tid; e < BM * (BK / 4); e += NT) {
        const int m = e / (BK / 4), kq = (e % (BK / 4)) * 4;
        const int gm = m0 + m, gk = k0 + kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < g.M) {
          const float* p = g.A + (size_t)gm * g.lda + gk;
          if (vecA && gk + 3 < kend) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (gk + 0 < kend) v.x = p[0];
            if (gk + 1 < kend) v.y = p[1];
            if (gk + 2 < kend) v.z = p[2];
            if (gk + 3 < kend) v.w = p[3];
          }
        }
        As[kq + 0][m] = v.x; As[kq + 1][m] = v.y; As[kq + 2][m] = v.z; As[kq + 3][m] = v.w;
      }
    } else {
      // A[k*lda + m], m contiguous: float4 along m
      for (int e = tid; e < BK * (BM / 4); e += NT) {
        const int k = e / (BM / 4), mq = (e % (BM / 4)) * 4;
        const int gk = k0 + k, gm = m0 + mq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gk < kend) {
          const float* p = g.A + (size_t)gk * g.lda + gm;
          if (vecA && gm + 3 < g.M) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (gm + 0 < g.M) v.x = p[0];
            if (gm + 1 < g.M) v.y = p[1];
            if (gm + 2 < g.M) v.z = p[2];
            if (gm + 3 < g.M) v.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&As[k][mq]) = v;
      }
    }
    // ---- load B tile into Bs[k][n] ----
    if (!TB) {
      // B[k*ldb + n], n contiguous
      for (int e = tid; e < BK * (BN / 4); e += NT) {
        const int k = e / (BN / 4), nq = (e % (BN / 4)) * 4;
        const int gk = k0 + k, gn = n0 + nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gk < kend) {
          const float* p = g.B + (size_t)gk * g.ldb + gn;
          if (vecB && gn + 3 < g.N) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (gn + 0 < g.N) v.x = p[0];
            if (gn + 1 < g.N) v.y = p[1];
            if (gn + 2 < g.N) v.z = p[2];
            if (gn + 3 < g.N) v.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[k][nq]) = v;
      }
    } else {
      // B[n*ldb + k], k contiguous
      for (int e = tid; e < BN * (BK / 4); e += NT) {
        const int n = e / (BK / 4), kq = (e % (BK / 4)) * 4;
        const int gn = n0 + n, gk = k0 + kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gn < g.N) {
          const float* p = g.B + (size_t)gn * g.ldb + gk;
          if (vecB && gk + 3 < kend) {
            v = *reinterpret_cast<const float4*>(p);
          } else {
            if (gk + 0 < kend) v.x = p[0];
            if (gk + 1 < kend) v.y = p[1];
            if (gk + 2 < kend) v.z = p[2];
            if (gk + 3 < kend) v.w = p[3];
          }
        }
        Bs[kq + 0][n] = v.x; Bs[kq + 1][n] = v.y; Bs[kq + 2][n] = v.z; Bs[kq + 3][n] = v.w;
      }
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      float c = acc[i][j];
      if (g.bias != nullptr && blockIdx.z == 0) c += g.bias[gn];
      float* dst = g.C + (size_t)gm * g.ldc + gn;
      if (use_atomic) atomicAdd(dst, c);
      else if (g.beta != 0.f) *dst += c;
      else *dst = c;
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
void launch_tile(cudaStream_t s, const GemmArgs& g) {
  const int gx = cdiv(g.N, BN), gy = cdiv(g.M, BM);
  const long long tiles = (long long)gx * gy;
  int splits = 1;
  if (tiles < kNumSMs && g.K >= 8 * BK) {
    splits = (int)min((long long)(2 * kNumSMs / tiles), (long long)(g.K / (4 * BK)));
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
  }
  int kps = cdiv(cdiv(g.K, splits), BK) * BK;
  splits = cdiv(g.K, kps);
  const int use_atomic = splits > 1;
  if (use_atomic && g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
  const bool a16 = ((uintptr_t)g.A % 16 == 0) && (g.lda % 4 == 0);
  const bool b16 = ((uintptr_t)g.B % 16 == 0) && (g.ldb % 4 == 0);
  dim3 grid(gx, gy, splits), block((BM / TM) * (BN / TN));
  const int vA = a16, vB = b16;
  if (!g.ta && !g.tb) launch_pdl(gemm_f32_kernel<BM, BN, BK, TM, TN, false, false>, grid, block, 0, s, g, kps, vA, vB, use_atomic);
  else if (!g.ta && g.tb) launch_pdl(gemm_f32_kernel<BM, BN, BK, TM, TN, false, true>, grid, block, 0, s, g, kps, vA, vB, use_atomic);
  else if (g.ta && !g.tb) launch_pdl(gemm_f32_kernel<BM, BN, BK, TM, TN, true, false>, grid, block, 0, s, g, kps, vA, vB, use_atomic);
  else launch_pdl(gemm_f32_kernel<BM, BN, BK, TM, TN, true, true>, grid, block, 0, s, g, kps, vA, vB, use_atomic);
  count_launch();
}

// ---- degenerate shapes of the hot path ---------------------------------------------------------------------------
// Heads with a handful of outputs (actor mean / std / logits, continue logit: N <= 32) and rank-few updates (the action
// part of the pre-GRU layer, input gradients of those heads: K <= 32) are latency-bound launches in a tiled GEMM
// (10 us per node in the replayed graph for [1024 x 256] x [256 x 1]); these two kernels do them in one pass.

// C[M, N] (+)= A[M, K] * op(B)[K, N] (+ bias), N <= NMAX <= 32: one warp per row, lanes stride K (128-bit loads when the
// row is aligned), N shuffle reductions; two rows per warp in flight
template <bool TB, int NMAX>
__global__ void __launch_bounds__(256) gemm_skinny_n_kernel(GemmArgs g, int vec) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = warp0 * 2; r0 < g.M; r0 += nwarps * 2) {
    float acc[2][NMAX];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) acc[q][n] = 0.f;
    const bool two = r0 + 1 < g.M;
    const float* a0 = g.A + (size_t)r0 * g.lda;
    const float* a1 = g.A + (size_t)(two ? r0 + 1 : r0) * g.lda;
    if (vec) {
      for (int k = 4 * lane; k < g.K; k += 128) {
        const float4 x0 = *reinterpret_cast<const float4*>(a0 + k);
        const float4 x1 = *reinterpret_cast<const float4*>(a1 + k);
        const float xs0[4] = {x0.x, x0.y, x0.z, x0.w}, xs1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int n = 0; n < NMAX; ++n) {
            if (n < g.N) {
              const float w = TB ? g.B[(size_t)n * g.ldb + k + j] : g.B[(size_t)(k + j) * g.ldb + n];
              acc[0][n] = fmaf(xs0[j], w, acc[0][n]);
              acc[1][n] = fmaf(xs1[j], w, acc[1][n]);
            }
          }
        }
      }
    } else {
      for (int k = lane; k < g.K; k += 32) {
        const float x0 = a0[k], x1 = a1[k];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
          if (n < g.N) {
            const float w = TB ? g.B[(size_t)n * g.ldb + k] : g.B[(size_t)k * g.ldb + n];
            acc[0][n] = fmaf(x0, w, acc[0][n]);
            acc[1][n] = fmaf(x1, w, acc[1][n]);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < g.N) {
        float v0 = acc[0][n], v1 = acc[1][n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v0 += __shfl_xor_sync(0xffffffffu, v0, o);
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        if (lane < 2 && (lane == 0 || two)) {
          float v = lane == 0 ? v0 : v1;
          if (g.bias != nullptr) v += g.bias[n];
          float* dst = g.C + (size_t)(r0 + lane) * g.ldc + n;
          *dst = (g.beta != 0.f) ? *dst + v : v;
        }
      }
    }
  }
}

// C[M, N] (+)= A[M, K] * op(B)[K, N] (+ bias), K <= 32: one thread per output element
template <bool TB>
__global__ void __launch_bounds__(256) gemm_skinny_k_kernel(GemmArgs g) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)g.M * g.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / g.N;
    const int c = (int)(i - r * g.N);
    const float* a = g.A + (size_t)r * g.lda;
    float v = g.bias != nullptr ? g.bias[c] : 0.f;
    for (int k = 0; k < g.K; ++k) v = fmaf(a[k], TB ? g.B[(size_t)c * g.ldb + k] : g.B[(size_t)k * g.ldb + c], v);
    float* dst = g.C + (size_t)r * g.ldc + c;
    *dst = (g.beta != 0.f) ? *dst + v : v;
  }
}

// The single-pass kernels are for the latency-bound launches over the B*T or (H+1)*B*T rows of the heads. The conv
// contractions over ~10^6 pixel rows have the same "few outputs / short K" shapes but are throughput problems: they stay
// on the tiled / tensor-core kernels (4 ms instead of 0.2 ms per launch otherwise).
constexpr long long kSkinnyMaxRows = 32768;
bool skinny_shape(const GemmArgs& g) {
  return !g.ta && g.M >= 64 && g.M <= kSkinnyMaxRows && ((g.N <= 32 && g.K >= 32) || (g.K <= 32 && g.N >= 32));
}

bool gemm_skinny(cudaStream_t s, const GemmArgs& g) {
  if (!skinny_shape(g)) return false;
  if (g.N <= 32 && g.K >= 32) {
    const int grid = (int)min((long long)148 * 8, ((long long)g.M + 15) / 16);
    const int vec = ((uintptr_t)g.A % 16 == 0) && (g.lda % 4 == 0) && (g.K % 4 == 0);
#define DV3_SKINNY_N(TB_, NM_) launch_pdl(gemm_skinny_n_kernel<TB_, NM_>, dim3(grid), dim3(256), 0, s, g, vec)
    if (g.N == 1) { if (g.tb) DV3_SKINNY_N(true, 1); else DV3_SKINNY_N(false, 1); }
    else if (g.N <= 8) { if (g.tb) DV3_SKINNY_N(true, 8); else DV3_SKINNY_N(false, 8); }
    else { if (g.tb) DV3_SKINNY_N(true, 32); else DV3_SKINNY_N(false, 32); }
#undef DV3_SKINNY_N
    count_launch();
    return true;
  }
  if (g.K <= 32 && g.N >= 32) {
    const int grid = (int)min((long long)148 * 16, ((long long)g.M * g.N + 255) / 256);
    if (g.tb) launch_pdl(gemm_skinny_k_kernel<true>, dim3(grid), dim3(256), 0, s, g);
    else launch_pdl(gemm_skinny_k_kernel<false>, dim3(grid), dim3(256), 0, s, g);
    count_launch();
    return true;
  }
  return false;
}

}  // namespace

void gemm_f32(cudaStream_t s, const GemmArgs& g) {
  if (g.M <= 0 || g.N <= 0) return;
  if (g.K <= 0) {
    if (g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
    return;
  }
  if (gemm_skinny(s, g)) return;
  if (g.M <= 16) {
    launch_tile<16, 128, 16, 1, 8>(s, g);
  } else if (g.M <= 64 || g.N <= 64 || (long long)cdiv(g.M, 128) * cdiv(g.N, 128) < kNumSMs) {
    launch_tile<64, 64, 16, 4, 4>(s, g);
  } else {
    launch_tile<128, 128, 16, 8, 8>(s, g);
  }
}


bool pdl_enabled() {
  // measured on B200 inside the replayed graphs: no gain (6.739 vs 6.740 ms per update), so it is opt-in
  static const bool on = std::getenv("DV3_PDL") != nullptr;
  return on;
}

// true when gemm_dispatch would run this contraction on the TMA-fed tcgen05 kernel, which reads only the bf16 operand copies
// (so a producer may skip writing the fp32 original of an operand whose every consumer says yes)
bool gemm_uses_tma(const GemmArgs& g, int mode) {
  if (!(g.M > 0 && g.N > 0 && g.K > 0) || mode != 1) return false;
  if (gemv16_eligible(g) || skinny_shape(g)) return false;
  if (gemm_tc_eligible(g) && gemm_small_preferred(g)) return true;   // (reads the bf16 copies only, like the TMA kernel)
  return gemm_tma_eligible(g) && gemm_tc_eligible(g);
}

void gemm_dispatch(cudaStream_t s, const GemmArgs& g, int mode) {
  if (g.M > 0 && g.N > 0 && g.K > 0 && mode != 2 && gemv16_eligible(g)) {   // the 16 scan rows against a large weight matrix
    gemv16(s, g);
    return;
  }
  if (g.M > 0 && g.N > 0 && g.K > 0 && skinny_shape(g) && mode != 2) {
    gemm_f32(s, g);   // degenerate shapes: skinny kernels, exact fp32 in every mode
    return;
  }
  // small contractions on a dependency chain (the per-step products of the XS dream): the fixed cost of the TMA kernel is their cost
  if (mode == 1 && g.M > 0 && g.N > 0 && g.K > 0 && gemm_tc_eligible(g) && gemm_small_preferred(g) && gemm_small(s, g)) return;
  if (mode >= 1 && g.M > 0 && g.N > 0 && g.K > 0 && gemm_tma_eligible(g) && (mode == 2 || gemm_tc_eligible(g)) && gemm_tma(s, g)) return;
  if (mode == 1 && g.M > 0 && g.N > 0 && g.K > 0 && gemm_tc_eligible(g)) {
    static const bool trace = std::getenv("DV3_TRACE_FALLBACK") != nullptr;   // which contractions miss the TMA kernel, and why
    if (trace)
      fprintf(stderr, "gemm_tc fallback: M %d N %d K %d ta %d tb %d lda %d ldb %d ldc %d beta %g Ah %d(ld %d) Bh %d(ld %d) Bmn %d(ld %d) C%%16 %d\n", g.M, g.N, g.K,
              g.ta, g.tb, g.lda, g.ldb, g.ldc, (double)g.beta, g.Ah != nullptr, g.ldah, g.Bh != nullptr, g.ldbh, g.Bmn != nullptr, g.ldbmn,
              (int)((uintptr_t)g.C % 16));
    gemm_tc(s, g);
  }
  else if (mode == 2 && g.M > 0 && g.N > 0 && g.K > 0) gemm_tc(s, g);  // forced (tests)
  else gemm_f32(s, g);
}

}  // namespace dv3
