// Small-shape bf16 contraction for the latency chains of the dream rollout and its BPTT (XS layer sizes: [1024 x 256..1280] x
// [.. x 256..1024] per dreamed step, 8-9 dependent contractions per step):
//
//   C[M,N] (+)= A16[M,K] * B16[N,K]^T (+ bias[N])          both operands bf16, K-major, fp32 accumulate, fp32 output
//
// The TMA / tcgen05 kernel (dv3_gemm_tma.cu) has a fixed cost of ~6 us per launch (tensor maps, mbarrier and TMEM set-up,
// pipeline fill, TMEM -> shared -> TMA-store epilogue, split-K reduction): right for contractions that take tens of
// microseconds, but these take 0.1 GFLOP each and sit on a dependency chain, where the fixed cost is the cost. Here: 64 x 64
// (or 64 x 32) tiles, 4 warps, cp.async 16-byte copies into padded shared memory (4 stages, all issued up front for K <= 256),
// mma.sync.m16n8k16 with fragments read by plain conflict-free 32-bit loads, results stored from registers. No tensor map, no
// mbarrier, no TMEM, no split-K.
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>

#include "dv3_common.h"
#include "dv3_device.cuh"

namespace dv3 {
namespace {

constexpr int kBM = 64, kBK = 64, kLdS = kBK + 8, kStages = 4, kThreadsS = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {   // bytes beyond src_bytes are zero-filled
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct SmallArgs {
  const __nv_bfloat16* A; int lda;
  const __nv_bfloat16* B; int ldb;
  float* C; int ldc;
  int M, N, K;
  const float* bias; float beta;
};

template <int BN>
__global__ void __launch_bounds__(kThreadsS) gemm_mma_small_kernel(const SmallArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16(*As)[kBM][kLdS] = reinterpret_cast<__nv_bfloat16(*)[kBM][kLdS]>(smem_raw);
  __nv_bfloat16(*Bs)[BN][kLdS] = reinterpret_cast<__nv_bfloat16(*)[BN][kLdS]>(smem_raw + (size_t)kStages * kBM * kLdS * 2);
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, tq = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;             // warp -> 32 rows x BN/2 columns of the tile
  constexpr int NT = BN / 16;                          // n-tiles (8 columns) per warp
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int nchunks = (g.K + kBK - 1) / kBK;

  auto load_chunk = [&](int ch) {
    const int st = ch % kStages, k0 = ch * kBK;
#pragma unroll
    for (int i = 0; i < kBM * 8 / kThreadsS; ++i) {
      const int p = tid + kThreadsS * i, row = p >> 3, c8 = (p & 7) * 8;
      const int rem = (g.K - (k0 + c8)) * 2;
      const bool ok = m0 + row < g.M && rem > 0;
      cp_async16(&As[st][row][c8], g.A + (size_t)(ok ? m0 + row : 0) * g.lda + (ok ? k0 + c8 : 0), ok ? min(rem, 16) : 0);
    }
#pragma unroll
    for (int i = 0; i < BN * 8 / kThreadsS; ++i) {
      const int p = tid + kThreadsS * i, row = p >> 3, c8 = (p & 7) * 8;
      const int rem = (g.K - (k0 + c8)) * 2;
      const bool ok = n0 + row < g.N && rem > 0;
      cp_async16(&Bs[st][row][c8], g.B + (size_t)(ok ? n0 + row : 0) * g.ldb + (ok ? k0 + c8 : 0), ok ? min(rem, 16) : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float c[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { c[mt][nt][0] = 0.f; c[mt][nt][1] = 0.f; c[mt][nt][2] = 0.f; c[mt][nt][3] = 0.f; }

  const int pre = min(nchunks, kStages - 1);
  for (int ch = 0; ch < pre; ++ch) load_chunk(ch);
  for (int ch = 0; ch < nchunks; ++ch) {
    // chunk ch has landed when at most (groups issued after it) are still pending
    const int newer = min(nchunks, ch + kStages - 1) - (ch + 1);
    if (newer >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (newer == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // chunk ch visible to every warp; every warp has finished chunk ch - 1 (its stage is refilled next)
    if (ch + kStages - 1 < nchunks) load_chunk(ch + kStages - 1);
    const int st = ch % kStages;
#pragma unroll
    for (int ks = 0; ks < kBK / 16; ++ks) {
      const int kb = ks * 16 + 2 * tq;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = wm * 32 + mt * 16 + g8;
        a[mt][0] = *reinterpret_cast<const uint32_t*>(&As[st][r][kb]);
        a[mt][1] = *reinterpret_cast<const uint32_t*>(&As[st][r + 8][kb]);
        a[mt][2] = *reinterpret_cast<const uint32_t*>(&As[st][r][kb + 8]);
        a[mt][3] = *reinterpret_cast<const uint32_t*>(&As[st][r + 8][kb + 8]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = wn * (BN / 2) + nt * 8 + g8;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Bs[st][n][kb]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Bs[st][n][kb + 8]);
        mma16816(c[0][nt], a[0], b0, b1);
        mma16816(c[1][nt], a[1], b0, b1);
      }
    }
  }

  // epilogue: thread holds rows g8 / g8 + 8 of each m-tile, column pairs 2 tq of each n-tile
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = m0 + wm * 32 + mt * 16 + g8 + 8 * h;
      if (row >= g.M) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = n0 + wn * (BN / 2) + nt * 8 + 2 * tq;
        if (col >= g.N) continue;
        float v0 = c[mt][nt][2 * h], v1 = c[mt][nt][2 * h + 1];
        float* dst = g.C + (size_t)row * g.ldc + col;
        if (g.bias != nullptr) { v0 += g.bias[col]; if (col + 1 < g.N) v1 += g.bias[col + 1]; }
        if (col + 1 < g.N) {
          float2 o = make_float2(v0, v1);
          if (g.beta != 0.f) { const float2 old = *reinterpret_cast<const float2*>(dst); o.x += g.beta * old.x; o.y += g.beta * old.y; }
          *reinterpret_cast<float2*>(dst) = o;
        } else {
          dst[0] = g.beta != 0.f ? v0 + g.beta * dst[0] : v0;
        }
      }
    }
  }
}

template <int BN>
cudaError_t launch_small(cudaStream_t s, const SmallArgs& a) {
  constexpr size_t smem = (size_t)kStages * (kBM + BN) * kLdS * 2;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mma_small_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(gemm_mma_small_kernel<BN>, dim3((unsigned)cdiv(a.N, BN), (unsigned)cdiv(a.M, kBM)), dim3(kThreadsS), smem, s, a);
}

}  // namespace

// What the kernel can run: both operands in bf16 (K-major), M in [64, 4096], <= 1.4 GFLOP
bool gemm_small_eligible(const GemmArgs& g) {
  if (g.ta || g.Ah == nullptr || g.Bh == nullptr) return false;
  if (g.M < 64 || g.M > 4096 || g.N < 16 || g.K < 32) return false;
  if ((long long)g.M * g.N * g.K > 1024LL * 1024 * 1280) return false;
  return g.K % 8 == 0 && g.ldah % 8 == 0 && g.ldbh % 8 == 0 && (uintptr_t)g.Ah % 16 == 0 && (uintptr_t)g.Bh % 16 == 0 && g.ldc % 2 == 0 &&
         (uintptr_t)g.C % 8 == 0;
}

// Where it is the faster of the two bf16 kernels as a node of a replayed graph (tools/gemm_small_probe.py, us per dependent
// launch, TMA kernel vs this one): [1024 x 256] x [256 x 256] 4.7 vs 3.2, [2048 x 256] x [256 x 256] 4.9 vs 4.2, 512^3 6.1 vs 4.8 -
// but 5.0 vs 5.4 at N = 768, 6.4 vs 7.7 at K = 1024 (the TMA kernel splits K over the idle SMs), and accumulating into C costs this
// kernel a dependent read (4.7 vs 5.0): narrow outputs, short K, plain stores only.
bool gemm_small_preferred(const GemmArgs& g) {
  static const bool off = std::getenv("DV3_NO_GEMM_SMALL") != nullptr;
  return !off && gemm_small_eligible(g) && g.beta == 0.f && g.N <= 512 && g.K <= 512 && (long long)g.M * g.N <= 1024LL * 512;
}

bool gemm_small(cudaStream_t s, const GemmArgs& g) {
  SmallArgs a;
  a.A = reinterpret_cast<const __nv_bfloat16*>(g.Ah); a.lda = g.ldah;
  a.B = reinterpret_cast<const __nv_bfloat16*>(g.Bh); a.ldb = g.ldbh;
  a.C = g.C; a.ldc = g.ldc; a.M = g.M; a.N = g.N; a.K = g.K; a.bias = g.bias; a.beta = g.beta;
  // 64-column tiles once they fill the GPU, 32-column tiles for the narrow outputs (N = 256 at 1024 rows: 128 CTAs instead of 64)
  const long long tiles64 = (long long)cdiv(g.M, kBM) * cdiv(g.N, 64);
  const cudaError_t e = tiles64 >= 148 ? launch_small<64>(s, a) : launch_small<32>(s, a);
  if (e != cudaSuccess) return false;
  count_launch();
  return true;
}

}  // namespace dv3
