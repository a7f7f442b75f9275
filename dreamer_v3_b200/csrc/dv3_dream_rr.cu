// Row-resident imagination rollout (DreamerModel.dream_trajectory, dreamer_model.py:189-208) for the XS layer sizes in the
// tensor-core mode: ONE launch walks all H dreamed steps, and a CTA never talks to another one - dream rows are independent
// through the whole rollout, so CTA c owns rows [16 c, 16 c + 16) from the start state to step H and keeps their state
// (h in fp32 and bf16, the 32 class indices of z, the hidden activations of the step) in shared memory. Per dreamed step:
//
//   gathers   the one-hot products z . W[z rows] of the actor's first layer(s) and of the pre-GRU layer are sums of 32 rows of the
//             plain bf16 weight shadows (one 16-byte load per lane and row: a warp reads a whole 512-byte row), no contraction
//   actor     h . W_actor[:G] on mma.sync + the gathered part, LayerNorm+SiLU, D x A output layer, unimix + inverse-CDF draw or the
//             reparameterised Normal (same for the std net of Box spaces)
//   pre-GRU   x = gathered part + a . W_pre[Z:], LayerNorm+SiLU -> u                    (a warp owns two whole rows: no contraction)
//   GRU       h . U and u . K (24 weight tiles), cell -> h'                             (a warp owns all three gates of its 32 units)
//   prior     h' . W_dyn, LayerNorm+SiLU, q . W_z (16 tiles), softmax / 1 % unimix / fp64 inverse-CDF draw of 32 categoricals
//             (a warp owns 4 whole categoricals of all 16 rows: reductions stay inside a quad)
//
// The 16 rows are exactly one m16 tile: contractions are mma.sync.m16n8k16 (bf16, fp32 accumulate), A fragments from padded
// [row][k] bf16 buffers written by the previous layer's epilogue, B fragments by conflict-free 32-bit loads from [256 rows x 64 k]
// weight tiles that a producer warp streams from the transposed bf16 shadows with TMA (SWIZZLE_128B, 4-deep ring, ~1.6 MB per
// step and CTA out of L2): the producer runs ahead across layer and step boundaries, so the weight stream never waits for an
// epilogue. Everything the heads, the losses and the BPTT read afterwards is written on the side in the layout of the per-op path.
//
// Measured (tools/dream_rr_prof.py, XS, us per dreamed step): gathers 6 - 9, weight-tile phases 25 - 28, epilogues 21 - 24 -> 52
// (Discrete) / 61 (Box) in total against ~64 on the per-kernel path inside the replayed graph: opt-in (DV3_DREAM_RR=1), parity-tested.
// The tile phases with the multiplications switched off: 21 us (the stream alone: 68 GB/s into one SM with 96 KB in flight; the same
// with 16 instead of 64 CTAs, and the same with tile-major weights moved by one contiguous 32 KB bulk copy per tile); with the stream
// switched off: 14.7 us (5632 HMMA). A CTA that owns its rows for a whole step ingests the step's 1.6 MB of weights whatever its row
// count: ~30 us per step with the gathered rows is this design's floor.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "dv3_device.cuh"
#include "dv3_dream_tc.h"

namespace dv3 {
namespace {

constexpr int kRowsR = 16;                 // dream rows per CTA
constexpr int kCW = 8;                     // warps: all consume; lane 0 of warp 0 also feeds the TMA ring
constexpr int kThreadsR = 32 * kCW;        // (a 9th warp would cost registers: allocation rounds the CTA up to 12 warps -> 168 per thread)
constexpr int kDR = 256;                   // G = D (XS)
constexpr int kZR = 1024, kZcR = 32;
constexpr int kRingR = 4;
constexpr int kTileBytes = 256 * 128;      // [256 weight rows][64 k] bf16
constexpr int kLdA = kDR + 8;              // bf16 row stride of the A operands (fragment loads hit 32 distinct banks)
constexpr int kLdF = kDR + 8;              // fp32 row stride of hf / zacc
constexpr int MA = kDreamTcMaxA;

enum { M_ACT = 0, M_STD, M_U, M_K, M_DYN, M_ZP, M_COUNT };
struct RMaps { CUtensorMap w[M_COUNT]; };

struct RParams {
  DreamTcArgs a;
  const __nv_bfloat16* Wn_pre; int ld_n_pre;        // plain shadow of W_pre [Z + A, D]
  const __nv_bfloat16* Wn_act[2]; int ld_n_act[2];  // plain shadows of the actor / std first layers [G + Z, D]
  const int32_t* z0_idx;                            // class indices of the start states [N, Zc]
  int nets;                                         // 1 Discrete, 2 Box (mean + std net)
};

struct RSmem {
  unsigned char ring[kRingR][kTileBytes];
  __nv_bfloat16 hA[kRowsR][kLdA];          // h_{i-1} (h_i after the GRU), u, q: A operands
  __nv_bfloat16 uA[kRowsR][kLdA];
  __nv_bfloat16 qA[kRowsR][kLdA];
  float hf[kRowsR][kLdF];                  // h in fp32 (the GRU cell's h_{t-1})
  float zacc[2][kRowsR][kLdF];             // gathered z-part of the actor / std first layer
  float red[kRowsR][kCW][4];               // per-warp LayerNorm partial sums (sum, sum of squares) of up to two nets
  float headp[kRowsR][kCW][2 * MA];        // per-warp partial output-layer products
  float act[kRowsR][MA];                   // the sampled action
  int zidx[kRowsR][kZcR];                  // class indices of z_{i-1}
  int seq[52][3];                          // the weight tiles of one step in consumption order: (matrix, 256-row block, k-block)
  uint64_t full[kRingR], empty[kRingR];
};
static_assert(sizeof(RSmem) + 1024 <= 227 * 1024, "dream_rr shared memory");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) { printf("dv3 dream_rr: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cbar() { __syncthreads(); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// One weight tile [256 rows][64 k] (TMA, SWIZZLE_128B) against the A operand at k offset k0: acc[j] += A . W^T for this warp's four
// n-tiles (tile rows 32 w + 8 j + g). Lane: g = lane / 4 (A rows g, g + 8; weight row g of the n-tile), tq = lane % 4 (k pairs).
__device__ __forceinline__ void mma_tile(const unsigned char* tile, const __nv_bfloat16 (*A)[kLdA], int k0, int w, int g, int tq,
                                         float (&acc)[4][4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {   // 16 k per step
    const int ka = k0 + 16 * s + 2 * tq;
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&A[g][ka]);
    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(&A[g + 8][ka]);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(&A[g][ka + 8]);
    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(&A[g + 8][ka + 8]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 32 * w + 8 * j + g;   // (n & 7) == g
      const unsigned char* row = tile + n * 128 + 4 * tq;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(row + (((2 * s) ^ g) << 4));
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(row + (((2 * s + 1) ^ g) << 4));
      mma16816(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }
}

__global__ void __launch_bounds__(kThreadsR, 1) dream_rr_kernel(const __grid_constant__ RMaps maps, const RParams P) {
  extern __shared__ unsigned char smem_raw[];
  // (offset added to the shared-space pointer itself: an integer round trip would turn every access into a generic load)
  RSmem& sm = *reinterpret_cast<RSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const DreamTcArgs& a = P.a;
  const int N = a.N, H = a.H, A = a.A, nets = P.nets;
  constexpr int G = kDR, D = kDR, Z = kZR, W = kDR + kZR, Zc = kZcR;
  const int row0 = blockIdx.x * kRowsR;

  if (tid == 0) {
    for (int i = 0; i < M_COUNT; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[i]) : "memory");
    for (int s = 0; s < kRingR; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the tile sequence of a step (the same every step)
  const int per_step = 4 * nets + 12 + 12 + 4 + 16;
  if (tid == 0) {
    int n = 0;
    auto put = [&](int m, int nb, int kb) { sm.seq[n][0] = m; sm.seq[n][1] = nb; sm.seq[n][2] = kb; ++n; };
    for (int net = 0; net < nets; ++net) for (int kb = 0; kb < 4; ++kb) put(net == 0 ? M_ACT : M_STD, 0, kb);
    for (int nb = 0; nb < 3; ++nb) for (int kb = 0; kb < 4; ++kb) put(M_U, nb, kb);
    for (int nb = 0; nb < 3; ++nb) for (int kb = 0; kb < 4; ++kb) put(M_K, nb, kb);
    for (int kb = 0; kb < 4; ++kb) put(M_DYN, 0, kb);
    for (int nb = 0; nb < 4; ++nb) for (int kb = 0; kb < 4; ++kb) put(M_ZP, nb, kb);
  }
  __syncthreads();

  const int w = warp, g = lane >> 2, tq = lane & 3, ct = tid;
  const int r0 = 2 * w, c8 = 8 * lane;                   // row-oriented phases: this warp's two rows, this lane's 8 columns
  const bool box = !a.discrete;
  const float* b0g = a.gru_b;                            // [2, 3G]: input-kernel bias, recurrent bias
  const float* b1g = a.gru_b + 3 * G;

  // start state: h (fp32 + bf16 A operand) and the class indices of z
  for (int e = ct; e < kRowsR * D; e += 32 * kCW) {
    const int r = e >> 8, c = e & 255;
    const float v = a.hz[(size_t)(row0 + r) * W + c];
    sm.hf[r][c] = v;
    sm.hA[r][c] = __float2bfloat16_rn(v);
  }
  for (int e = ct; e < kRowsR * Zc; e += 32 * kCW) sm.zidx[e >> 5][e & 31] = P.z0_idx[(size_t)(row0 + (e >> 5)) * Zc + (e & 31)];
  cbar();

  // TMA ring: tile t lives in slot t % kRingR. Lane 0 of warp 0 issues tile t + kRingR - 1 just before the CTA consumes tile t (its
  // slot was released by all 8 warps when they finished tile t - 1), so three tiles are in flight while one is consumed, across
  // layer and step boundaries.
  const uint32_t total_tiles = (uint32_t)H * (uint32_t)per_step;
  auto issue = [&](uint32_t t) {
    const int slot = t % kRingR, e = t % per_step;
    mbar_wait(&sm.empty[slot], ((t / kRingR) & 1) ^ 1);
    mbar_expect_tx(&sm.full[slot], kTileBytes);
    tma_load_2d(smem_u32(sm.ring[slot]), &maps.w[sm.seq[e][0]], sm.seq[e][2] * 64, sm.seq[e][1] * 256, &sm.full[slot]);
  };
  if (tid == 0) for (uint32_t t = 0; t + 1 < kRingR && t < total_tiles; ++t) issue(t);
  uint32_t it = 0;   // tile counter
  // optional profile (a.prof): cycles this CTA's first consumer warp spends in each phase of a step, summed over the steps
  const bool prof = a.prof != nullptr && w == 0 && lane == 0;
  long long pt[12], plast = prof ? clock64() : 0;
#pragma unroll
  for (int k = 0; k < 12; ++k) pt[k] = 0;
  auto lap = [&](int k) { if (prof) { const long long now = clock64(); pt[k] += now - plast; plast = now; } };
  auto tiles = [&](const __nv_bfloat16 (*Aop)[kLdA], float (&acc)[4][4]) {   // the four k-blocks of one 256-row weight block
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb, ++it) {
      const int slot = it % kRingR;
      if (tid == 0 && it + kRingR - 1 < total_tiles) issue(it + kRingR - 1);
      mbar_wait(&sm.full[slot], (it / kRingR) & 1);
      mma_tile(sm.ring[slot], Aop, kb * 64, w, g, tq, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
    }
  };
  auto zero = [](float (&acc)[4][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
  };

  for (int i = 1; i <= H; ++i) {
    const size_t ra0 = (size_t)(i - 1) * N + row0;   // rows of state i-1 (actor buffers, hz) = rows of the per-step buffers of step i
    const size_t rc0 = (size_t)i * N + row0;         // rows of state i in hz

    // ===== gathers: z-part of the actor's first layer(s) -> zacc, z-part of the pre-GRU layer -> registers =====
    float xacc[2][8];
    {
      float zac[2][2][8];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int k = 0; k < 8; ++k) { xacc[rr][k] = 0.f; zac[0][rr][k] = 0.f; zac[1][rr][k] = 0.f; }
      const __nv_bfloat16* wp = P.Wn_pre + c8;
      const __nv_bfloat16* wa0 = P.Wn_act[0] + (size_t)G * P.ld_n_act[0] + c8;
      const __nv_bfloat16* wa1 = box ? P.Wn_act[1] + (size_t)G * P.ld_n_act[1] + c8 : wa0;
#pragma unroll 1
      for (int cb = 0; cb < Zc; cb += 4) {
        uint4 vp[2][4], v0[2][4], v1[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const size_t zr = (size_t)((cb + q) * 32 + sm.zidx[r0 + rr][cb + q]);
            vp[rr][q] = __ldg(reinterpret_cast<const uint4*>(wp + zr * P.ld_n_pre));
            v0[rr][q] = __ldg(reinterpret_cast<const uint4*>(wa0 + zr * P.ld_n_act[0]));
            if (box) v1[rr][q] = __ldg(reinterpret_cast<const uint4*>(wa1 + zr * P.ld_n_act[1]));
          }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t* up = reinterpret_cast<const uint32_t*>(&vp[rr][q]);
            const uint32_t* u0 = reinterpret_cast<const uint32_t*>(&v0[rr][q]);
            const uint32_t* u1 = reinterpret_cast<const uint32_t*>(&v1[rr][q]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              xacc[rr][2 * k] += __uint_as_float(up[k] << 16); xacc[rr][2 * k + 1] += __uint_as_float(up[k] & 0xffff0000u);
              zac[0][rr][2 * k] += __uint_as_float(u0[k] << 16); zac[0][rr][2 * k + 1] += __uint_as_float(u0[k] & 0xffff0000u);
              if (box) { zac[1][rr][2 * k] += __uint_as_float(u1[k] << 16); zac[1][rr][2 * k + 1] += __uint_as_float(u1[k] & 0xffff0000u); }
            }
          }
      }
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        if (net >= nets) break;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          float4* dst = reinterpret_cast<float4*>(&sm.zacc[net][r0 + rr][c8]);
          dst[0] = make_float4(zac[net][rr][0], zac[net][rr][1], zac[net][rr][2], zac[net][rr][3]);
          dst[1] = make_float4(zac[net][rr][4], zac[net][rr][5], zac[net][rr][6], zac[net][rr][7]);
        }
      }
    }
    lap(0);
    cbar();
    lap(1);

    // ===== actor: h . W[:G] + gathered part, LayerNorm + SiLU, output layer =====
    {
      float pre[2][4][4];
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        if (net >= nets) break;
        zero(pre[net]);
        tiles(sm.hA, pre[net]);
        lap(2);
        const DreamTcMlp& m = net == 0 ? a.actor : a.actor_std;
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 32 * w + 8 * j + 2 * tq;
          const float2 z0 = *reinterpret_cast<const float2*>(&sm.zacc[net][g][col]);
          const float2 z1 = *reinterpret_cast<const float2*>(&sm.zacc[net][g + 8][col]);
          pre[net][j][0] += z0.x; pre[net][j][1] += z0.y; pre[net][j][2] += z1.x; pre[net][j][3] += z1.y;
          *reinterpret_cast<float2*>(m.pre[0] + (ra0 + g) * D + col) = make_float2(pre[net][j][0], pre[net][j][1]);
          *reinterpret_cast<float2*>(m.pre[0] + (ra0 + g + 8) * D + col) = make_float2(pre[net][j][2], pre[net][j][3]);
          s0 += pre[net][j][0] + pre[net][j][1]; q0 += pre[net][j][0] * pre[net][j][0] + pre[net][j][1] * pre[net][j][1];
          s1 += pre[net][j][2] + pre[net][j][3]; q1 += pre[net][j][2] * pre[net][j][2] + pre[net][j][3] * pre[net][j][3];
        }
        s0 = quad_sum(s0); q0 = quad_sum(q0); s1 = quad_sum(s1); q1 = quad_sum(q1);
        if (tq == 0) {
          sm.red[g][w][2 * net] = s0; sm.red[g][w][2 * net + 1] = q0;
          sm.red[g + 8][w][2 * net] = s1; sm.red[g + 8][w][2 * net + 1] = q1;
        }
      }
      cbar();
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        if (net >= nets) break;
        const DreamTcMlp& m = net == 0 ? a.actor : a.actor_std;
        float mean[2], rstd[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int ww = 0; ww < kCW; ++ww) { s += sm.red[g + 8 * hh][ww][2 * net]; q += sm.red[g + 8 * hh][ww][2 * net + 1]; }
          mean[hh] = s / (float)D;
          rstd[hh] = rsqrtf(fmaxf(q / (float)D - mean[hh] * mean[hh], 0.f) + DV3_LN_EPS);
        }
        if (w == 0 && tq == 0) {
          m.mean[0][ra0 + g] = mean[0]; m.rstd[0][ra0 + g] = rstd[0];
          m.mean[0][ra0 + g + 8] = mean[1]; m.rstd[0][ra0 + g + 8] = rstd[1];
        }
        const float* wo = net == 0 ? a.actor_out_w : a.std_out_w;   // [D, A]
        float hp[2][MA];
#pragma unroll
        for (int k = 0; k < MA; ++k) { hp[0][k] = 0.f; hp[1][k] = 0.f; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 32 * w + 8 * j + 2 * tq;
          const float2 gm = __ldg(reinterpret_cast<const float2*>(m.g[0] + col)), bt = __ldg(reinterpret_cast<const float2*>(m.b[0] + col));
          float v[4];
          v[0] = (pre[net][j][0] - mean[0]) * rstd[0] * gm.x + bt.x; v[1] = (pre[net][j][1] - mean[0]) * rstd[0] * gm.y + bt.y;
          v[2] = (pre[net][j][2] - mean[1]) * rstd[1] * gm.x + bt.x; v[3] = (pre[net][j][3] - mean[1]) * rstd[1] * gm.y + bt.y;
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] *= sigmoid_fast(v[k]);
          *reinterpret_cast<float2*>(m.act[0] + (ra0 + g) * D + col) = make_float2(v[0], v[1]);
          *reinterpret_cast<float2*>(m.act[0] + (ra0 + g + 8) * D + col) = make_float2(v[2], v[3]);
          if (m.act16[0] != nullptr) {
            unsigned short* a16 = reinterpret_cast<unsigned short*>(m.act16[0]);
            *reinterpret_cast<uint32_t*>(a16 + (ra0 + g) * D + col) = pack_bf16(v[0], v[1]);
            *reinterpret_cast<uint32_t*>(a16 + (ra0 + g + 8) * D + col) = pack_bf16(v[2], v[3]);
          }
#pragma unroll
          for (int k = 0; k < MA; ++k) {
            if (k < A) {
              const float w0 = __ldg(wo + (size_t)col * A + k), w1 = __ldg(wo + (size_t)(col + 1) * A + k);
              hp[0][k] += v[0] * w0 + v[1] * w1;
              hp[1][k] += v[2] * w0 + v[3] * w1;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < MA; ++k) {
          if (k < A) {
            const float p0 = quad_sum(hp[0][k]), p1 = quad_sum(hp[1][k]);
            if (tq == 0) { sm.headp[g][w][net * MA + k] = p0; sm.headp[g + 8][w][net * MA + k] = p1; }
          }
        }
      }
      cbar();
      lap(3);
    }

    // ===== the action (actor_network.py:63-118): lanes 0 / 1 of warp w finish rows 2 w / 2 w + 1 =====
    if (lane < 2) {
      const int r = r0 + lane;
      const size_t ra = ra0 + r;
      float om[MA], os[MA], av[MA];
#pragma unroll
      for (int k = 0; k < MA; ++k) {
        om[k] = 0.f; os[k] = 0.f; av[k] = 0.f;
        if (k < A) {
          float s = 0.f, t = 0.f;
#pragma unroll
          for (int ww = 0; ww < kCW; ++ww) { s += sm.headp[r][ww][k]; t += sm.headp[r][ww][MA + k]; }
          om[k] = s + __ldg(a.actor_out_b + k);
          if (box) os[k] = t + __ldg(a.std_out_b + k);
        }
      }
      if (!box) {
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < MA; ++k) if (k < A) mx = fmaxf(mx, om[k]);
        float e[MA];
#pragma unroll
        for (int k = 0; k < MA; ++k) e[k] = k < A ? expf(om[k] - mx) : 0.f;
        const float s = ((e[0] + e[4]) + (e[2] + e[6])) + ((e[1] + e[5]) + (e[3] + e[7]));   // warp_sum order of the per-op kernel
        float pu[MA];
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < MA; ++k) { pu[k] = k < A ? 0.99f * (e[k] / s) + 0.01f * (1.0f / (float)A) : 0.f; tot += (double)pu[k]; }
        const double thr = (double)__ldg(a.act_noise + ra) * tot;
        double cdf = 0.0;
        int pick = 0;
#pragma unroll
        for (int k = 0; k < MA; ++k) if (k < A) { cdf += (double)pu[k]; if (cdf <= thr) ++pick; }
        pick = min(pick, A - 1);
#pragma unroll
        for (int k = 0; k < MA; ++k) {
          av[k] = k == pick ? 1.f : 0.f;
          if (k < A) { a.actor_logits[ra * A + k] = om[k]; a.actor_probs[ra * A + k] = pu[k]; a.dr_act[ra * A + k] = av[k]; }
        }
        a.a_idx[ra] = pick;
      } else {
#pragma unroll
        for (int k = 0; k < MA; ++k) {
          if (k < A) {
            const float sd = 0.9f * sigmoidf_(os[k] + 2.0f) + 0.1f;
            av[k] = tanhf(om[k]) + sd * __ldg(a.act_noise + ra * A + k);
            a.actor_logits[ra * A + k] = om[k]; a.actor_s[ra * A + k] = os[k]; a.dr_act[ra * A + k] = av[k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < MA; ++k) sm.act[r][k] = av[k];
    }
    __syncwarp();

    // ===== pre-GRU layer (row-oriented): x = gathered part + a . W_pre[Z:]; u = SiLU(LN(x)) =====
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = r0 + rr;
      const size_t rs = ra0 + r;
      float x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = xacc[rr][k];
#pragma unroll
      for (int j = 0; j < MA; ++j) {
        if (j < A) {
          const float aj = sm.act[r][j];
          const float4 wa = __ldg(reinterpret_cast<const float4*>(a.W_pre_a + (size_t)j * D + c8));
          const float4 wb = __ldg(reinterpret_cast<const float4*>(a.W_pre_a + (size_t)j * D + c8 + 4));
          x[0] += aj * wa.x; x[1] += aj * wa.y; x[2] += aj * wa.z; x[3] += aj * wa.w;
          x[4] += aj * wb.x; x[5] += aj * wb.y; x[6] += aj * wb.z; x[7] += aj * wb.w;
        }
      }
      float4* xo = reinterpret_cast<float4*>(a.x_pre + rs * D + c8);
      xo[0] = make_float4(x[0], x[1], x[2], x[3]); xo[1] = make_float4(x[4], x[5], x[6], x[7]);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += x[k];
      const float mean = warp_sum(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = x[k] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / (float)D + DV3_LN_EPS);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.pre_g + c8)), g1 = __ldg(reinterpret_cast<const float4*>(a.pre_g + c8 + 4));
      const float4 e0 = __ldg(reinterpret_cast<const float4*>(a.pre_b + c8)), e1 = __ldg(reinterpret_cast<const float4*>(a.pre_b + c8 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float n = (x[k] - mean) * rstd * gm[k] + bt[k]; x[k] = n * sigmoid_fast(n); }
      float4* uo = reinterpret_cast<float4*>(a.u + rs * D + c8);
      uo[0] = make_float4(x[0], x[1], x[2], x[3]); uo[1] = make_float4(x[4], x[5], x[6], x[7]);
      uint4 ub;
      ub.x = pack_bf16(x[0], x[1]); ub.y = pack_bf16(x[2], x[3]); ub.z = pack_bf16(x[4], x[5]); ub.w = pack_bf16(x[6], x[7]);
      if (a.u16 != nullptr) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(a.u16) + rs * D + c8) = ub;
      *reinterpret_cast<uint4*>(&sm.uA[r][c8]) = ub;
      if (lane == 0) { a.x_mean[rs] = mean; a.x_rstd[rs] = rstd; }
    }
    lap(4);
    cbar();
    lap(1);

    // ===== GRU: gh = h . U, gx = u . K (this warp: all three gates of units 32 w .. 32 w + 31), cell -> h' =====
    {
      float aU[3][4][4], aK[3][4][4];
#pragma unroll
      for (int gt = 0; gt < 3; ++gt) { zero(aU[gt]); tiles(sm.hA, aU[gt]); }
#pragma unroll
      for (int gt = 0; gt < 3; ++gt) { zero(aK[gt]); tiles(sm.uA, aK[gt]); }
      lap(5);
      cbar();   // every warp has read h_{i-1} (A operand) for the last time
      lap(1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 32 * w + 8 * j + 2 * tq;
        float2 bx[3], bh[3];
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) {
          bx[gt] = __ldg(reinterpret_cast<const float2*>(b0g + gt * G + col));
          bh[gt] = __ldg(reinterpret_cast<const float2*>(b1g + gt * G + col));
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = g + 8 * hh;
          const size_t rs = ra0 + r, rc = rc0 + r;
          const float2 hp = *reinterpret_cast<const float2*>(&sm.hf[r][col]);
          float gxv[3][2], ghv[3][2];
#pragma unroll
          for (int gt = 0; gt < 3; ++gt) {
            gxv[gt][0] = aK[gt][j][2 * hh] + bx[gt].x; gxv[gt][1] = aK[gt][j][2 * hh + 1] + bx[gt].y;
            ghv[gt][0] = aU[gt][j][2 * hh] + bh[gt].x; ghv[gt][1] = aU[gt][j][2 * hh + 1] + bh[gt].y;
          }
          float zz[2], rg[2], cc[2], hn[2];
          const float hpv[2] = {hp.x, hp.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            zz[e] = sigmoid_fast(gxv[0][e] + ghv[0][e]);
            rg[e] = sigmoid_fast(gxv[1][e] + ghv[1][e]);
            cc[e] = tanh_fast(gxv[2][e] + rg[e] * ghv[2][e]);
            hn[e] = zz[e] * hpv[e] + (1.f - zz[e]) * cc[e];
          }
          float* gt_o = a.gx + rs * 3 * G + col;   // the gates (what gru_fwd leaves in gx for the backward pass)
          *reinterpret_cast<float2*>(gt_o) = make_float2(zz[0], zz[1]);
          *reinterpret_cast<float2*>(gt_o + G) = make_float2(rg[0], rg[1]);
          *reinterpret_cast<float2*>(gt_o + 2 * G) = make_float2(cc[0], cc[1]);
          float* gh_o = a.gh + rs * 3 * G + col;
          *reinterpret_cast<float2*>(gh_o) = make_float2(ghv[0][0], ghv[0][1]);
          *reinterpret_cast<float2*>(gh_o + G) = make_float2(ghv[1][0], ghv[1][1]);
          *reinterpret_cast<float2*>(gh_o + 2 * G) = make_float2(ghv[2][0], ghv[2][1]);
          *reinterpret_cast<float2*>(a.hz + rc * W + col) = make_float2(hn[0], hn[1]);
          const uint32_t hb = pack_bf16(hn[0], hn[1]);
          if (a.hz16 != nullptr) *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(a.hz16) + rc * W + col) = hb;
          *reinterpret_cast<float2*>(&sm.hf[r][col]) = make_float2(hn[0], hn[1]);
          *reinterpret_cast<uint32_t*>(&sm.hA[r][col]) = hb;
        }
      }
      lap(6);
      cbar();
      lap(1);
    }

    // ===== prior net: q = SiLU(LN(h' . W_dyn)) =====
    {
      float qa[4][4];
      zero(qa);
      tiles(sm.hA, qa);
      lap(7);
      float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 32 * w + 8 * j + 2 * tq;
        *reinterpret_cast<float2*>(a.q_pre + (ra0 + g) * D + col) = make_float2(qa[j][0], qa[j][1]);
        *reinterpret_cast<float2*>(a.q_pre + (ra0 + g + 8) * D + col) = make_float2(qa[j][2], qa[j][3]);
        s0 += qa[j][0] + qa[j][1]; q0 += qa[j][0] * qa[j][0] + qa[j][1] * qa[j][1];
        s1 += qa[j][2] + qa[j][3]; q1 += qa[j][2] * qa[j][2] + qa[j][3] * qa[j][3];
      }
      s0 = quad_sum(s0); q0 = quad_sum(q0); s1 = quad_sum(s1); q1 = quad_sum(q1);
      if (tq == 0) { sm.red[g][w][0] = s0; sm.red[g][w][1] = q0; sm.red[g + 8][w][0] = s1; sm.red[g + 8][w][1] = q1; }
      cbar();
      float mean[2], rstd[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int ww = 0; ww < kCW; ++ww) { s += sm.red[g + 8 * hh][ww][0]; q += sm.red[g + 8 * hh][ww][1]; }
        mean[hh] = s / (float)D;
        rstd[hh] = rsqrtf(fmaxf(q / (float)D - mean[hh] * mean[hh], 0.f) + DV3_LN_EPS);
      }
      if (w == 0 && tq == 0) {
        a.q_mean[ra0 + g] = mean[0]; a.q_rstd[ra0 + g] = rstd[0];
        a.q_mean[ra0 + g + 8] = mean[1]; a.q_rstd[ra0 + g + 8] = rstd[1];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 32 * w + 8 * j + 2 * tq;
        const float2 gm = __ldg(reinterpret_cast<const float2*>(a.dyn_g + col)), bt = __ldg(reinterpret_cast<const float2*>(a.dyn_b + col));
        float v[4];
        v[0] = (qa[j][0] - mean[0]) * rstd[0] * gm.x + bt.x; v[1] = (qa[j][1] - mean[0]) * rstd[0] * gm.y + bt.y;
        v[2] = (qa[j][2] - mean[1]) * rstd[1] * gm.x + bt.x; v[3] = (qa[j][3] - mean[1]) * rstd[1] * gm.y + bt.y;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] *= sigmoid_fast(v[k]);
        *reinterpret_cast<float2*>(a.q_act + (ra0 + g) * D + col) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(a.q_act + (ra0 + g + 8) * D + col) = make_float2(v[2], v[3]);
        const uint32_t p0 = pack_bf16(v[0], v[1]), p1 = pack_bf16(v[2], v[3]);
        if (a.q_act16 != nullptr) {
          unsigned short* q16 = reinterpret_cast<unsigned short*>(a.q_act16);
          *reinterpret_cast<uint32_t*>(q16 + (ra0 + g) * D + col) = p0;
          *reinterpret_cast<uint32_t*>(q16 + (ra0 + g + 8) * D + col) = p1;
        }
        *reinterpret_cast<uint32_t*>(&sm.qA[g][col]) = p0;
        *reinterpret_cast<uint32_t*>(&sm.qA[g + 8][col]) = p1;
      }
      cbar();
      lap(8);
    }

    // ===== prior logits + straight-through sample (representation_layer.py:73-113): this warp owns categoricals nb * 8 + w =====
    {
      float lg[4][4][4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) { zero(lg[nb]); tiles(sm.qA, lg[nb]); }
      lap(9);
      // The four categoricals of this warp go through every reduction together (their shuffle / MUFU / fp64 chains interleave
      // instead of running back to back: with 2 warps per scheduler a single dependent chain leaves the SM idle most of the time).
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = g + 8 * hh;
        const size_t rs = ra0 + r, rc = rc0 + r;
        float pu[4][4][2], mx[4], sden[4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const int cat = nb * 8 + w;
          mx[nb] = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 bz = __ldg(reinterpret_cast<const float2*>(a.zp_bias + cat * 32 + 8 * j + 2 * tq));
            pu[nb][j][0] = lg[nb][j][2 * hh] + bz.x; pu[nb][j][1] = lg[nb][j][2 * hh + 1] + bz.y;
            mx[nb] = fmaxf(mx[nb], fmaxf(pu[nb][j][0], pu[nb][j][1]));
            *reinterpret_cast<float2*>(a.prior_logits + rs * Z + cat * 32 + 8 * j + 2 * tq) = make_float2(pu[nb][j][0], pu[nb][j][1]);
          }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) mx[nb] = fmaxf(mx[nb], __shfl_xor_sync(0xffffffffu, mx[nb], o));
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          sden[nb] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pu[nb][j][0] = __expf(pu[nb][j][0] - mx[nb]); pu[nb][j][1] = __expf(pu[nb][j][1] - mx[nb]);
            sden[nb] += pu[nb][j][0] + pu[nb][j][1];
          }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) sden[nb] += __shfl_xor_sync(0xffffffffu, sden[nb], o);
        double incl[4][4];   // fp64 sums of fp32 probabilities are exact: any grouping gives the same cdf
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const int cat = nb * 8 + w;
          const float inv = __fdividef(0.99f, sden[nb]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pu[nb][j][0] = pu[nb][j][0] * inv + 0.01f * (1.0f / 32.0f); pu[nb][j][1] = pu[nb][j][1] * inv + 0.01f * (1.0f / 32.0f);
            *reinterpret_cast<float2*>(a.prior_probs + rs * Z + cat * 32 + 8 * j + 2 * tq) = make_float2(pu[nb][j][0], pu[nb][j][1]);
            incl[nb][j] = (double)pu[nb][j][0] + (double)pu[nb][j][1];
          }
        }
        // inclusive scan over the quad (classes 8 j + 2 tq + e: lanes with a smaller tq come first)
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double t = __shfl_up_sync(0xffffffffu, incl[nb][j], o, 4);
              if (tq >= o) incl[nb][j] += t;
            }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const int cat = nb * 8 + w;
          double blk[4], tot = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) { blk[j] = __shfl_sync(0xffffffffu, incl[nb][j], 3, 4); tot += blk[j]; }   // the 8 classes 8 j .. 8 j + 7
          const double thr = (double)__ldg(a.u_dyn + rs * Zc + cat) * tot;
          int cnt = 0;
          double base = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double c1 = base + incl[nb][j], c0 = c1 - (double)pu[nb][j][1];
            cnt += (c0 <= thr) + (c1 <= thr);
            base += blk[j];
          }
          cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
          cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
          const int pick = min(cnt, 31);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cls = 8 * j + 2 * tq;
            const float o0 = cls == pick ? 1.f : 0.f, o1 = cls + 1 == pick ? 1.f : 0.f;
            *reinterpret_cast<float2*>(a.hz + rc * W + G + cat * 32 + cls) = make_float2(o0, o1);
            if (a.hz16 != nullptr)
              *reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned short*>(a.hz16) + rc * W + G + cat * 32 + cls) =
                  (cls == pick ? 0x3F80u : 0u) | (cls + 1 == pick ? 0x3F800000u : 0u);
          }
          if (tq == 0) { a.z_idx[rs * Zc + cat] = pick; sm.zidx[r][cat] = pick; }
        }
      }
      lap(10);
      cbar();
      lap(1);
    }
  }
  if (prof) {
    unsigned long long* o = a.prof + 16 * blockIdx.x;
    for (int k = 0; k < 12; ++k) o[k] = (unsigned long long)pt[k];
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// bf16 [rows, K] K-major (row stride ld elements), box 64 k x 256 rows, SWIZZLE_128B
bool make_map(CUtensorMap* map, const void* ptr, long long rows, int K, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 256};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool dream_rr_supported(int N, int G, int D, int Z, int Zc, int Zk, int A, int L) {
  return N % kRowsR == 0 && N >= kRowsR && G == kDR && D == kDR && Z == kZR && Zc == kZcR && Zk == 32 && A >= 1 && A <= MA && L == 1;
}

cudaError_t launch_dream_rr(cudaStream_t s, const DreamTcArgs& a, const DreamRrExtra& x) {
  RParams P{};
  P.a = a;
  P.Wn_pre = reinterpret_cast<const __nv_bfloat16*>(x.Wn_pre); P.ld_n_pre = x.ld_n_pre;
  P.Wn_act[0] = reinterpret_cast<const __nv_bfloat16*>(x.Wn_act[0]); P.ld_n_act[0] = x.ld_n_act[0];
  P.Wn_act[1] = reinterpret_cast<const __nv_bfloat16*>(x.Wn_act[1]); P.ld_n_act[1] = x.ld_n_act[1];
  P.z0_idx = x.z0_idx;
  P.nets = a.discrete ? 1 : 2;
  const int G = a.G, D = a.D, Z = a.Z, W = a.G + a.Z;
  RMaps maps;
  bool ok = true;
  ok &= make_map(&maps.w[M_ACT], a.actor.wt[0], D, W, a.actor.ld_wt[0]);
  ok &= a.discrete ? make_map(&maps.w[M_STD], a.actor.wt[0], D, W, a.actor.ld_wt[0]) : make_map(&maps.w[M_STD], a.actor_std.wt[0], D, W, a.actor_std.ld_wt[0]);
  ok &= make_map(&maps.w[M_U], a.Wt_U, 3 * G, G, a.ld_U);
  ok &= make_map(&maps.w[M_K], a.Wt_K, 3 * G, D, a.ld_K);
  ok &= make_map(&maps.w[M_DYN], a.Wt_dyn, D, G, a.ld_dyn);
  ok &= make_map(&maps.w[M_ZP], a.Wt_zp, Z, D, a.ld_zp);
  if (!ok) return cudaErrorInvalidValue;
  const size_t smem = sizeof(RSmem) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dream_rr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dream_rr_kernel<<<dim3(a.N / kRowsR), dim3(kThreadsR), smem, s>>>(maps, P);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
