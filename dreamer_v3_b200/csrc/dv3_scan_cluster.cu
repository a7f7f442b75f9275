// RSSM observe scan (world_model.py:244-273) on ONE 16-CTA thread-block cluster, for model sizes whose recurrent
// weights fit the cluster's shared memory (XS: GRU 256, dense 256, 32x32 latents).
//
// The grid-wide version (dv3_scan.cu) spends most of a step waiting: every stage boundary is a grid barrier plus a
// chain of L2 round trips (publish, fence, atomic, poll, reload). Here the 16 CTAs of a cluster own 1/16 of every layer's
// output columns, keep exactly those weight columns resident in shared memory for all T steps, and hand activations to
// each other through distributed shared memory: a producer writes its [16 columns x 16 rows] block straight into the
// staging buffer of every peer and a hardware cluster barrier separates the stages. Nothing on the recurrence touches
// L2 except the one-hot row gather of the pre-GRU layer; everything the backward pass needs is written to global
// memory on the side (fire and forget).
//
// Per step t, 4 stages (cluster barrier after each):
//   A  x_pre[:, my 16 cols] = actW + sum_c W_pre[c*32 + idx_in[c], my cols]              -> all peers' buf0
//   B  u = SiLU(LN(x_pre)) in place (every CTA, one warp per row); gx = u . K[:, my 16 units x 3 gates];
//      GRU cell with gh (own units, kept locally from stage C of the previous step)        -> h[:, my units] to all peers' buf1
//   C  p_pre[:, my 16 cols] += h . W_post[E:][:, my cols] (weights in registers)          -> all peers' buf0
//      gh(t+1)[:, my units] = h . U[:, my gate cols] + b (kept locally; rows that restart at t+1 use h0 . U + b)
//   D  p_act = SiLU(LN(p_pre)) in place; logits[:, my 64 cols] = p_act . W_zq + b; softmax, 1% unimix, fp64 inverse-CDF
//      sample of my 2 categoricals                                                          -> idx to all peers
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "dv3_scan_common.cuh"

namespace cg = cooperative_groups;

namespace dv3 {
namespace {

using namespace scan;

constexpr int kC = 16;        // CTAs per cluster
constexpr int kU = 16;        // GRU units / dense columns owned per CTA (G = D = 16 * 16)
constexpr int kZC = 64;       // posterior logit columns owned per CTA (2 categoricals of 32 classes)
constexpr int kG = kC * kU;   // 256
constexpr int kLd = kInLd;    // 20: row stride of the [k][row] staging buffers

constexpr int kWLd = kG + 8; // k stride of a resident bf16 weight column (tensor-core variant): conflict-free fragment loads

struct CCommon {
  float buf0[kG][kLd];        // x_pre -> u, then p_pre -> p_act    20 KB   [k][row]
  float buf1[kG][kLd];        // h; cross-warp scratch in stage D   20 KB
  float red[4 * kRows * (3 * kU + 1)];   // k-slice partial sums of a 48-column pass ([4][16][49]); also the 16-column p_pre pass
  float gh[kRows][3 * kU];    // gh(t) of my units (z, r, c gate blocks of 16)
  float gh0[3 * kU];          // h0 . U + b for my units
  float hown[kRows][kU];      // h_{t-1} of my units
  float xch[3][kU][kLd];      // this CTA's [16 columns][16 rows (+ pad)] blocks of exchanges A, B, C, staged in the layout of its slice
                              // of buf0 / buf1 (one staging block per exchange: a block is rewritten one whole step later, when every
                              // peer is known to have received it)
  int idx[kC][kRows][2];      // z_{t-1} class indices: [owner CTA][row][its 2 categoricals] (categorical c -> idx[c / 2][row][c % 2])
  int xidx[kRows][2];         // this CTA's picks, staged
  unsigned long long xbar[4]; // mbarriers: arrival of all 16 blocks of exchange A (x_pre), B (h), C (p_pre), D (indices)
  int idx0[32];               // mode of the prior at the initial state (rows that restart)
  float lnp[4][kG];           // LayerNorm gain / offset of the pre-GRU and the posterior layer
};
struct CSmem : CCommon {      // fp32 variant: weights [k][column]
  float wK[kG][3 * kU];       // K[:, gate*G + my units]            48 KB
  float wU[kG][3 * kU];       // U[:, gate*G + my units]            48 KB
  float wZ[kG][kZC];          // W_zq[:, my 64 columns]             64 KB
};
struct CSmemTc : CCommon {    // tensor-core variant: bf16 weights [column][k] (the col-major B operand of mma.m16n8k16)
  __nv_bfloat16 wK[3 * kU][kWLd];   // 25 KB
  __nv_bfloat16 wU[3 * kU][kWLd];   // 25 KB
  __nv_bfloat16 wZ[kZC][kWLd];      // 33 KB
  __nv_bfloat16 wP[kU][kWLd];       // W_post[E:][:, my 16 columns]    8 KB
  __nv_bfloat16 wX[kC * kZC][kU];   // W_pre[:Z][:, my 16 columns]    32 KB: the one-hot row gather of stage A stays on chip
};
static_assert(sizeof(CSmem) <= 227 * 1024 && sizeof(CSmemTc) <= 227 * 1024, "cluster scan shared memory");

// Register-tiled contraction of the 16 staged rows with NC = 16 * CPL resident weight columns, K = 256:
// warp -> (k-slice of 64, column half), lane -> (4 rows, CPL columns): one LDS.128 of activations and CPL weights feed
// 4 * CPL FMAs, so the pass is bound by the FMA pipe, not by shared-memory wavefronts. Leaves the four k-slice partial sums
// in red[ks][row][col] (row stride NC + 1); the consumer adds them.
template <int CPL>
__device__ __forceinline__ void tile_pass(const float* in_s, const float* W, float* red) {
  constexpr int NC = 16 * CPL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = warp & 3, half = warp >> 2, rg = lane & 3, cg = lane >> 2;
  const int c0 = half * (NC / 2) + cg * CPL;
  float acc[4][CPL];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[r][c] = 0.f;
  const float* ap = in_s + (size_t)(64 * ks) * kLd + 4 * rg;
  const float* wp = W + (size_t)(64 * ks) * NC + c0;
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(ap + k * kLd);
    float w[CPL];
    if (CPL == 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(wp + k * NC);
      w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[CPL - 1] = w4.w;
    } else {
#pragma unroll
      for (int c = 0; c < CPL; ++c) w[c] = wp[k * NC + c];
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      acc[0][c] = fmaf(a.x, w[c], acc[0][c]); acc[1][c] = fmaf(a.y, w[c], acc[1][c]);
      acc[2][c] = fmaf(a.z, w[c], acc[2][c]); acc[3][c] = fmaf(a.w, w[c], acc[3][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CPL; ++c) red[(ks * kRows + 4 * rg + r) * (NC + 1) + c0 + c] = acc[r][c];
  __syncthreads();
}
template <int NC>
__device__ __forceinline__ float red_sum(const float* red, int b, int col) {
  return red[(0 * kRows + b) * (NC + 1) + col] + red[(1 * kRows + b) * (NC + 1) + col] + red[(2 * kRows + b) * (NC + 1) + col] +
         red[(3 * kRows + b) * (NC + 1) + col];
}

// Tensor-core variant of the contractions. The batch is 16 rows, so the shape is mma.sync m16n8k16 (tcgen05 starts at M = 64
// and takes a TMEM round trip per result; here the whole pass is a dozen warp-level MMAs per warp). A fragments come straight
// from the fp32 [k][row] staging buffer (rounded to bf16 here), B fragments from the resident bf16 columns W16[n][k].
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// A fragment of rows 0..15, k in [kb, kb + 16) (lane: g = lane / 4 -> rows g, g + 8; tq = lane % 4 -> k pairs 2 tq, 2 tq + 8)
__device__ __forceinline__ void load_a_frag(const float* in_s, int kb, int g, int tq, uint32_t (&a)[4]) {
  const float* ap = in_s + (size_t)(kb + 2 * tq) * kLd + g;
  a[0] = pack2_bf16(ap[0], ap[kLd]);
  a[1] = pack2_bf16(ap[8], ap[kLd + 8]);
  a[2] = pack2_bf16(ap[8 * kLd], ap[9 * kLd]);
  a[3] = pack2_bf16(ap[8 * kLd + 8], ap[9 * kLd + 8]);
}
// NC = 48 / 64 columns: warp -> (k-slice of 64, column half); leaves the four k-slice partial sums in red[ks][row][col] like tile_pass
template <int NC>
__device__ __forceinline__ void mma_pass(const float* in_s, const __nv_bfloat16* W16, float* red) {
  constexpr int NT = NC / 16;   // n-tiles (8 columns) per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = warp & 3, half = warp >> 2, g = lane >> 2, tq = lane & 3;
  float c[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int kb = 64 * ks + 16 * st;
    uint32_t a[4];
    load_a_frag(in_s, kb, g, tq, a);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint32_t* bp = reinterpret_cast<const uint32_t*>(W16 + (size_t)((half * NT + nt) * 8 + g) * kWLd + kb + 2 * tq);
      mma_bf16_16816(c[nt], a[0], a[1], a[2], a[3], bp[0], bp[4]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = (half * NT + nt) * 8 + 2 * tq;
    float* r0 = red + (size_t)(ks * kRows + g) * (NC + 1) + col;
    float* r1 = red + (size_t)(ks * kRows + g + 8) * (NC + 1) + col;
    r0[0] = c[nt][0]; r0[1] = c[nt][1]; r1[0] = c[nt][2]; r1[1] = c[nt][3];
  }
  __syncthreads();
}
// the 16-column p_pre pass: warp -> k-slice of 32, both n-tiles; partial sums in red[warp][row][17]
__device__ __forceinline__ void mma_pass16(const float* in_s, const __nv_bfloat16* W16, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  float c[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    const int kb = 32 * warp + 16 * st;
    uint32_t a[4];
    load_a_frag(in_s, kb, g, tq, a);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const uint32_t* bp = reinterpret_cast<const uint32_t*>(W16 + (size_t)(nt * 8 + g) * kWLd + kb + 2 * tq);
      mma_bf16_16816(c[nt], a[0], a[1], a[2], a[3], bp[0], bp[4]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = nt * 8 + 2 * tq;
    float* r0 = red + (size_t)(warp * kRows + g) * 17 + col;
    float* r1 = red + (size_t)(warp * kRows + g + 8) * 17 + col;
    r0[0] = c[nt][0]; r0[1] = c[nt][1]; r1[0] = c[nt][2]; r1[1] = c[nt][3];
  }
  __syncthreads();
}

// ---- exchange between the CTAs of the cluster -------------------------------------------------------------------------------
// A producer stages its block in its own shared memory in exactly the layout of its slice of the peers' receive buffer and
// hands it to the bulk-copy engine once per peer (cp.async.bulk shared::cta -> shared::cluster, one elected thread per peer):
// the 16 copies run at DSMEM bandwidth without occupying the LSU, and each signals the DESTINATION's mbarrier with its byte
// count. A consumer waits on its own mbarrier for the bytes of all 16 producers - the arrival of the data is the signal, so a
// step has no cluster-wide barrier at all. Write-after-read safety follows from the data flow: a CTA can only produce the block
// of exchange n + 1 after it has received every block of exchange n, i.e. after every peer has finished the stage that read the
// buffers of exchange n - 1 ... (each buffer is rewritten two exchanges later at the earliest; see the stage order).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void xbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
// one thread: this phase completes after `bytes` have landed in my receive buffer
__device__ __forceinline__ void xbar_arm(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) { printf("dv3 cluster scan: exchange wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// all threads call it after writing `staging` (bytes, multiple of 16): block -> offset rank * bytes of every peer's `recv`
__device__ __forceinline__ void xchg_send(void* recv, const void* staging, uint32_t bytes, unsigned long long* bar, unsigned rank) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy writes of the staging block -> async proxy
  __syncthreads();
  if (threadIdx.x < kC) {
    const uint32_t dst = map_to_rank(smem_addr(recv) + rank * bytes, threadIdx.x);
    const uint32_t mb = map_to_rank(smem_addr(bar), threadIdx.x);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(smem_addr(staging)), "r"(bytes), "r"(mb) : "memory");
  }
}
constexpr uint32_t kBlkBytes = kU * kLd * 4;   // a CTA's [16 columns][20] fp32 slice of buf0 / buf1
// this thread's value of (row b = tid / 16, column j = tid % 16) -> my slice of every peer's buf[k][row]
__device__ __forceinline__ void send_block(float (*xch)[kLd], float* buf, unsigned long long* bar, unsigned rank, float v) {
  xch[threadIdx.x & 15][threadIdx.x >> 4] = v;
  xchg_send(buf, &xch[0][0], kBlkBytes, bar, rank);
}

// In-place LayerNorm + SiLU of the 16 rows staged as buf[k][row]; one warp per row (2 rows per warp, both in flight).
// Row b's result is also written to global memory by CTA `rank == b` (spreads the side traffic over the cluster).
__device__ __forceinline__ void ln_silu_inplace(float* buf, const float* gamma, const float* beta, int B,
                                                unsigned rank, float* act_out, float* mean_out, float* rstd_out, size_t row_stride,
                                                size_t stat_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int EPL = kG / 32;
  float v[2][EPL], g[EPL], be[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = gamma[c]; be[i] = beta[c];
    v[0][i] = buf[c * kLd + warp];
    v[1][i] = buf[c * kLd + warp + kWarps];
  }
  float mean[2], rstd[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += v[r][i];
    mean[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mean[0] += __shfl_xor_sync(0xffffffffu, mean[0], o);
    mean[1] += __shfl_xor_sync(0xffffffffu, mean[1], o);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mean[r] /= (float)kG;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) { const float d = v[r][i] - mean[r]; q += d * d; }
    rstd[r] = q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rstd[0] += __shfl_xor_sync(0xffffffffu, rstd[0], o);
    rstd[1] += __shfl_xor_sync(0xffffffffu, rstd[1], o);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int b = warp + r * kWarps;
    rstd[r] = rsqrtf(rstd[r] / (float)kG + DV3_LN_EPS);
    const bool wr = (unsigned)b == rank && b < B;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = lane + 32 * i;
      const float n = (v[r][i] - mean[r]) * rstd[r] * g[i] + be[i];
      const float o = b < B ? n * sigmoid_fast(n) : 0.f;
      buf[c * kLd + b] = o;
      if (wr) act_out[(size_t)b * row_stride + c] = o;
    }
    if (wr && lane == 0) { mean_out[(size_t)b * stat_stride] = mean[r]; rstd_out[(size_t)b * stat_stride] = rstd[r]; }
  }
  __syncthreads();
}

template <bool kTc>
__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1) observe_scan_cluster_kernel(ScanFwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SmemT = typename std::conditional<kTc, CSmemTc, CSmem>::type;
  SmemT& sm = *reinterpret_cast<SmemT*>(smem_raw);
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const int B = p.B, T = p.T, G = kG, D = kG, Z = p.Z, Zc = p.Zc, Zk = p.Zk, E = p.E;
  const int W = G + Z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = rank * kU;        // my GRU units / dense columns
  const int z0c = rank * kZC;      // my posterior logit columns
  const float* W_post_h = p.W_post + (size_t)E * D;

  // ---- resident weights ----
  for (int e = tid; e < kG * 3 * kU; e += kThreads) {
    const int k = e / (3 * kU), j = e % (3 * kU), col = (j / kU) * G + u0 + (j % kU);
    if constexpr (kTc) {
      sm.wK[j][k] = __float2bfloat16_rn(p.gru_K[(size_t)k * 3 * G + col]);
      sm.wU[j][k] = __float2bfloat16_rn(p.gru_U[(size_t)k * 3 * G + col]);
    } else {
      sm.wK[k][j] = p.gru_K[(size_t)k * 3 * G + col];
      sm.wU[k][j] = p.gru_U[(size_t)k * 3 * G + col];
    }
  }
  for (int e = tid; e < kG * kZC; e += kThreads) {
    const int k = e / kZC, j = e % kZC;
    if constexpr (kTc) sm.wZ[j][k] = __float2bfloat16_rn(p.W_zq[(size_t)k * Z + z0c + j]);
    else sm.wZ[k][j] = p.W_zq[(size_t)k * Z + z0c + j];
  }
  // W_post[E:][:, my 16 columns]: fp32 variant in registers (thread (warp, lane) owns column lane % 16 and the k's of its slice),
  // tensor-core variant in shared memory next to W_pre[:Z][:, my 16 columns]
  float wP[kTc ? 1 : 16];
  if constexpr (kTc) {
    wP[0] = 0.f;
    for (int e = tid; e < kG * kU; e += kThreads) { const int k = e / kU, j = e % kU; sm.wP[j][k] = __float2bfloat16_rn(W_post_h[(size_t)k * D + u0 + j]); }
    for (int e = tid; e < Z * kU; e += kThreads) { const int k = e / kU, j = e % kU; sm.wX[k][j] = __float2bfloat16_rn(p.W_pre[(size_t)k * D + u0 + j]); }
  } else {
    const int k0 = warp * (kG / kWarps) + lane / 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) wP[i] = W_post_h[(size_t)(k0 + 2 * i) * D + u0 + (lane & 15)];
  }
  for (int e = tid; e < kC * kRows * 2; e += kThreads) (&sm.idx[0][0][0])[e] = p.z0_idx[(e / (2 * kRows)) * 2 + (e & 1)];
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) xbar_init(&sm.xbar[i]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) sm.idx0[tid] = p.z0_idx[tid];
  for (int e = tid; e < kG; e += kThreads) {
    sm.lnp[0][e] = p.pre_g[e]; sm.lnp[1][e] = p.pre_b[e]; sm.lnp[2][e] = p.post_g[e]; sm.lnp[3][e] = p.post_b[e];
  }
  for (int e = tid; e < kRows * kU; e += kThreads) sm.hown[e / kU][e % kU] = p.h0[u0 + (e % kU)];
  // every row starts from the learned initial state: buf1 <- h0 for all rows, gh(0) = gh0 = h0 . U + b1
  for (int e = tid; e < kG * kRows; e += kThreads) sm.buf1[e / kRows][e % kRows] = p.h0[e / kRows];
  __syncthreads();
  if constexpr (kTc) mma_pass<48>(&sm.buf1[0][0], &sm.wU[0][0], sm.red);
  else tile_pass<3>(&sm.buf1[0][0], &sm.wU[0][0], sm.red);
  if (tid < 3 * kU) sm.gh0[tid] = red_sum<48>(sm.red, 0, tid) + p.gru_b[3 * G + (tid / kU) * G + u0 + (tid % kU)];
  __syncthreads();
  for (int e = tid; e < kRows * 3 * kU; e += kThreads) sm.gh[e / (3 * kU)][e % (3 * kU)] = sm.gh0[e % (3 * kU)];
  cl.sync();

  const int b = tid >> 4, j = tid & 15;   // (row, owned column) of this thread in the 16 x 16 epilogues
  const bool row_ok = b < B;
  const float h0_own = p.h0[u0 + j];      // learned initial state of this thread's unit (rows that restart)
  // per-thread constants and the first step's streamed inputs (later steps are prefetched one step ahead, so that the
  // only global load a stage ever waits for is the row gather of stage A)
  float gb[6];
#pragma unroll
  for (int g2 = 0; g2 < 3; ++g2) { gb[g2] = p.gru_b[g2 * G + u0 + j]; gb[3 + g2] = p.gru_b[3 * G + g2 * G + u0 + j]; }
  const float bz0 = p.b_zq[z0c + lane], bz1 = p.b_zq[z0c + 32 + lane];
  auto stamp = [&](int i) {
    if (p.stamps != nullptr && rank == 0 && tid == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  stamp(0);
  bool first = row_ok;
  float actw = row_ok ? p.actW[((size_t)b * T) * D + u0 + j] : 0.f;

  for (int t = 0; t < T; ++t) {
    const size_t bt = (size_t)b * T + t;
    const uint32_t par = (uint32_t)t & 1u;   // phase parity of this step's four exchanges
    if (tid == 0) {   // (each barrier's previous phase was waited for in the previous step)
      xbar_arm(&sm.xbar[0], kC * kBlkBytes); xbar_arm(&sm.xbar[1], kC * kBlkBytes); xbar_arm(&sm.xbar[2], kC * kBlkBytes);
      xbar_arm(&sm.xbar[3], kC * (uint32_t)sizeof(sm.xidx));
    }
    // streamed inputs of this step, consumed in stages C / D
    const float pp_in = row_ok ? p.p_pre[bt * D + u0 + j] : 0.f;
    float up[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int pr = warp + q * kWarps, r = pr >> 1, cat = z0c / 32 + (pr & 1);
      up[q] = r < B ? p.u_post[((size_t)t * B + r) * Zc + cat] : 0.f;
    }
    bool first_n = false;
    float actw_n = 0.f;
    if (t + 1 < T && row_ok) { first_n = p.is_first[bt + 1] != 0.f; actw_n = p.actW[(bt + 1) * D + u0 + j]; }

    // ===== stage A: pre-GRU layer, my 16 columns: one-hot z -> 32 row gathers per output =====
    {
      float x = 0.f;
      if (row_ok) {
        x = actw;
        float w32[32];   // all 32 row gathers in flight at once: one L2 round trip instead of four (tensor-core variant: on chip)
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int cls = first ? sm.idx0[c] : sm.idx[c >> 1][b][c & 1];
          if constexpr (kTc) w32[c] = __bfloat162float(sm.wX[c * Zk + cls][j]);
          else w32[c] = p.W_pre[(size_t)(c * Zk + cls) * D + u0 + j];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) x += w32[c];
        p.x_pre[bt * D + u0 + j] = x;
        // masked recurrent inputs kept for the backward pass (world_model.py:246-250): my units / my 2 categoricals
        p.h_in[bt * G + u0 + j] = first ? h0_own : sm.hown[b][j];
      }
      send_block(sm.xch[0], &sm.buf0[0][0], &sm.xbar[0], rank, x);
      if (row_ok) {
        for (int q = j; q < kZC; q += 16) {
          const int cat = (z0c + q) / Zk, cls = (z0c + q) % Zk;
          const int pick = first ? sm.idx0[cat] : sm.idx[cat >> 1][b][cat & 1];
          p.z_in[bt * Z + z0c + q] = (pick == cls) ? 1.f : 0.f;
        }
      }
    }
    stamp(t * 8 + 1);
    xbar_wait(&sm.xbar[0], par);
    stamp(t * 8 + 2);

    // ===== stage B: u = SiLU(LN(x_pre)); gx for my units; GRU cell =====
    ln_silu_inplace(&sm.buf0[0][0], sm.lnp[0], sm.lnp[1], B, rank, p.u_act + (size_t)t * D, p.x_mean + t, p.x_rstd + t, (size_t)T * D, (size_t)T);
    if constexpr (kTc) mma_pass<48>(&sm.buf0[0][0], &sm.wK[0][0], sm.red);
    else tile_pass<3>(&sm.buf0[0][0], &sm.wK[0][0], sm.red);
    {
      const int u = u0 + j;
      float hn = 0.f;
      if (row_ok) {
        const float hp = first ? h0_own : sm.hown[b][j];
        const float ghz = sm.gh[b][j], ghr = sm.gh[b][kU + j], ghc = sm.gh[b][2 * kU + j];
        const float z = sigmoid_fast(red_sum<48>(sm.red, b, j) + gb[0] + ghz);
        const float r = sigmoid_fast(red_sum<48>(sm.red, b, kU + j) + gb[1] + ghr);
        const float c = tanh_fast(red_sum<48>(sm.red, b, 2 * kU + j) + gb[2] + r * ghc);
        hn = z * hp + (1.f - z) * c;
        p.hz[bt * W + u] = hn;
        float* gt = p.gates + bt * 3 * G;
        gt[u] = z; gt[G + u] = r; gt[2 * G + u] = c;
        float* gho = p.gh + bt * 3 * G;
        gho[u] = ghz; gho[G + u] = ghr; gho[2 * G + u] = ghc;
      }
      __syncthreads();   // every thread has read hown / gh of this step
      sm.hown[b][j] = hn;
      send_block(sm.xch[1], &sm.buf1[0][0], &sm.xbar[1], rank, hn);
    }
    stamp(t * 8 + 3);
    xbar_wait(&sm.xbar[1], par);
    stamp(t * 8 + 4);

    // ===== stage C: p_pre (my 16 columns, weights in registers); gh(t+1) for my units =====
    {
      if constexpr (kTc) {
        mma_pass16(&sm.buf1[0][0], &sm.wP[0][0], sm.red);
      } else {
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
        const float* in_s = &sm.buf1[0][0];
        const int k0 = warp * (kG / kWarps) + lane / 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4* iv = reinterpret_cast<const float4*>(in_s + (size_t)(k0 + 2 * i) * kLd);
          const float4 a0 = iv[0], a1 = iv[1], a2 = iv[2], a3 = iv[3];
          const float w_ = wP[i];
          acc[0] = fmaf(a0.x, w_, acc[0]); acc[1] = fmaf(a0.y, w_, acc[1]); acc[2] = fmaf(a0.z, w_, acc[2]); acc[3] = fmaf(a0.w, w_, acc[3]);
          acc[4] = fmaf(a1.x, w_, acc[4]); acc[5] = fmaf(a1.y, w_, acc[5]); acc[6] = fmaf(a1.z, w_, acc[6]); acc[7] = fmaf(a1.w, w_, acc[7]);
          acc[8] = fmaf(a2.x, w_, acc[8]); acc[9] = fmaf(a2.y, w_, acc[9]); acc[10] = fmaf(a2.z, w_, acc[10]); acc[11] = fmaf(a2.w, w_, acc[11]);
          acc[12] = fmaf(a3.x, w_, acc[12]); acc[13] = fmaf(a3.y, w_, acc[13]); acc[14] = fmaf(a3.z, w_, acc[14]); acc[15] = fmaf(a3.w, w_, acc[15]);
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 16);
        if (lane < 16) {
#pragma unroll
          for (int r = 0; r < kRows; ++r) sm.red[(warp * kRows + r) * 17 + lane] = acc[r];
        }
        __syncthreads();
      }
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += sm.red[(w * kRows + b) * 17 + j];
      float pv = 0.f;
      if (row_ok) {
        pv = pp_in + s;     // p_pre was pre-loaded with enc . W_post[:E]
        p.p_pre[bt * D + u0 + j] = pv;
      }
      send_block(sm.xch[2], &sm.buf0[0][0], &sm.xbar[2], rank, pv);
    }
    // everything the peers wait for (my p_pre columns) is on its way: gh(t+1), which only this CTA reads, is computed while the
    // blocks travel
    if (t + 1 < T) {
      if constexpr (kTc) mma_pass<48>(&sm.buf1[0][0], &sm.wU[0][0], sm.red);
      else tile_pass<3>(&sm.buf1[0][0], &sm.wU[0][0], sm.red);
#pragma unroll
      for (int gate = 0; gate < 3; ++gate) {
        const int q = gate * kU + j;
        sm.gh[b][q] = first_n ? sm.gh0[q] : red_sum<48>(sm.red, b, q) + gb[3 + gate];
      }
    }
    stamp(t * 8 + 5);
    xbar_wait(&sm.xbar[2], par);
    stamp(t * 8 + 6);

    // ===== stage D: p_act = SiLU(LN(p_pre)); logits for my 64 columns; sample my 2 categoricals =====
    ln_silu_inplace(&sm.buf0[0][0], sm.lnp[2], sm.lnp[3], B, rank, p.p_act + (size_t)t * D, p.p_mean + t, p.p_rstd + t, (size_t)T * D, (size_t)T);
    {
      float* red64 = &sm.buf1[0][0];   // h is dead until stage B of the next step: [4][16][65] k-slice partial sums
      if constexpr (kTc) mma_pass<64>(&sm.buf0[0][0], &sm.wZ[0][0], red64);
      else tile_pass<4>(&sm.buf0[0][0], &sm.wZ[0][0], red64);
      float lg[4];   // logits of this thread: rows (tid >> 5) + 8 i' ... laid out so that a warp owns one (row, categorical) pair
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int pr = warp + q4 * kWarps, r = pr >> 1, cl_ = pr & 1;
        const float v = red_sum<64>(red64, r, cl_ * 32 + lane) + (cl_ == 0 ? bz0 : bz1);
        lg[q4] = v;
        if (r < B) p.post_logits[((size_t)r * T + t) * Z + z0c + cl_ * 32 + lane] = v;
      }
      // 2 categoricals x 16 rows = 32 (row, categorical) pairs, one warp each (4 per warp). The four pairs of a warp go
      // through the shuffle reductions together (their chains interleave instead of running back to back).
      float mx[4], ex[4], sden[4], pu[4];
      double cdf[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) mx[q4] = lg[q4];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) mx[q4] = fmaxf(mx[q4], __shfl_xor_sync(0xffffffffu, mx[q4], o));
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) { ex[q4] = expf(lg[q4] - mx[q4]); sden[q4] = ex[q4]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) sden[q4] += __shfl_xor_sync(0xffffffffu, sden[q4], o);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) { pu[q4] = 0.99f * (ex[q4] / sden[q4]) + 0.01f * (1.0f / 32.f); cdf[q4] = (double)pu[q4]; }
      // fp64 inclusive scan of the fp32 probabilities (exact, as categorical_pick in dv3_device.cuh), four at a time
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const double tq = __shfl_up_sync(0xffffffffu, cdf[q4], o);
          if (lane >= o) cdf[q4] += tq;
        }
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int pr = warp + q4 * kWarps;
        const int r = pr >> 1, cl_ = pr & 1;
        const int cat = z0c / Zk + cl_;
        const double tot = __shfl_sync(0xffffffffu, cdf[q4], 31);
        const unsigned below = __ballot_sync(0xffffffffu, cdf[q4] <= (double)up[q4] * tot);
        const int pick = min(__popc(below), 31);
        if (r >= B) continue;
        const size_t rt = (size_t)r * T + t;
        p.post_probs[rt * Z + cat * 32 + lane] = pu[q4];
        if (lane == 0) p.z_idx[rt * Zc + cat] = pick;
        p.hz[rt * W + G + cat * 32 + lane] = (lane == pick) ? 1.f : 0.f;
        if (lane == 0) sm.xidx[r][cl_] = pick;
      }
      for (int pr = warp; pr < 2 * kRows; pr += kWarps) if ((pr >> 1) >= B && lane == 0) sm.xidx[pr >> 1][pr & 1] = 0;
      xchg_send(&sm.idx[0][0][0], &sm.xidx[0][0], (uint32_t)sizeof(sm.xidx), &sm.xbar[3], rank);
    }
    first = first_n;
    actw = actw_n;
    stamp(t * 8 + 7);
    xbar_wait(&sm.xbar[3], par);
    stamp(t * 8 + 8);
  }
  cl.sync();   // no CTA leaves while a peer may still be reading its staging blocks
}


// ------------------------------------------------------------------------------------------------------------------------
// Reverse-time BPTT through the scan on the same 16-CTA cluster (tensor-core mode; replaces tape.gradient over
// world_model.py:244-273, train_one_step.py:80). CTA r owns dense columns / GRU units [16 r, 16 r + 16) and posterior logit
// columns [64 r, 64 r + 64). Every contraction keeps its WEIGHTS stationary: the rows of the (untransposed) fp32 weight
// matrices that produce this CTA's output columns are converted to bf16 once and stay in shared memory for all T steps
// (125 KB per CTA), the 16-row operands travel between the CTAs through distributed shared memory in bf16 ([row][k], the A
// fragment layout of mma.m16n8k16) or fp32 ([k][row], where a LayerNorm backward over whole rows follows). Per step t
// (T-1 .. 0), 5 stages and 4 cluster barriers:
//   S1  dlogits[t] (my 2 categoricals) = softmax-Jacobian^T 0.99 (d_hz[t].z + carried dz_in + dKL/dprobs)   -> all peers (bf16)
//   S2  dp_act[:, my cols] = dlogits . W_zq^T                                                                 -> all peers (fp32)
//   S3  dp_pre = LN'SiLU'(dp_act) (every CTA, whole rows); dh = d_hz[t].h + carried dh_in + dp_pre . W_post[E:]^T for my units;
//       GRU gate gradients dz, dr, dc, dc r of my units                                                       -> all peers (bf16)
//   S4  dh_in[:, my units] = dh z + [dz dr dc r] . U^T;  du[:, my cols] = [dz dr dc] . K^T                    -> du to all peers (fp32)
//   S5  dx_pre = LN'SiLU'(du); dz_in[:, my 64 cols] = dx_pre . W_pre[:Z]^T: stays in this CTA (it is the z-gradient of the
//       categoricals S1 of step t-1 handles here); nothing of the recurrence goes through global memory.
// Everything a stage reads from the forward pass is loaded at the top of the step, before S1 (no dependent L2 round trip inside).
constexpr int kLdL = kC * kZC + 8;   // bf16 k stride of the resident W_zq rows   [column][1024 + 8]
constexpr int kLdG = 3 * kG + 8;     // bf16 k stride of the resident U / K rows  [unit][768 + 8]
constexpr int kLdXl = kZC + 8;       // row stride of a CTA's dlogits block       [row][64 + 8]   (fragment loads: 32 distinct banks)
constexpr int kLdXg = kU + 8;        // row stride of a CTA's gate-gradient block [row][16 + 8]   (same)

struct BSmem {
  __nv_bfloat16 wA[kU][kLdL];      // W_zq[my 16 rows][Z]             33 KB
  __nv_bfloat16 wB[kU][kWLd];      // W_post[E + my 16 units][D]       8 KB
  __nv_bfloat16 wC[kU][kLdG];      // U[my 16 units][3G]              24 KB
  __nv_bfloat16 wD[kU][kLdG];      // K[my 16 columns][3G]            24 KB
  __nv_bfloat16 wE[kZC][kWLd];     // W_pre[my 64 rows][D]            33 KB
  union {                          // receive buffers, one block per producing CTA (never live at the same time)
    __nv_bfloat16 dl[kC][kRows][kLdXl];      // dlogits[t]: [owner CTA][row][its 64 columns]                    36 KB
    __nv_bfloat16 dg[kC][4][kRows][kLdXg];   // dz, dr, dc, dc r: [owner CTA][gate block][row][its 16 units]     48 KB
  } x;
  float buf[kG][kLd];              // dp_act -> dp_pre, then du -> dx_pre: fp32 [k][row]   20 KB
  float red[2 * kWarps * kRows * 17];   // per-warp partial sums (two products in S4); [4][16][65] in S5
  float dzc[kRows][kZC + 1];       // carried z-gradient of my 64 columns: (1 - is_first[t+1]) dz_in[t+1]
  float xch[2][kU][kLd];           // my [16 columns][16 rows (+ pad)] fp32 blocks of dp_act / du, staged in the layout of my slice of buf
  __nv_bfloat16 xl[kRows][kLdXl];  // my dlogits block, staged
  __nv_bfloat16 xg[4][kRows][kLdXg];   // my gate-gradient blocks, staged
  float lnp[4][kG];                // post_g, post_b, pre_g, pre_b
  unsigned long long xbar[4];      // mbarriers: all 16 blocks of dlogits / dp_act / gate gradients / du have landed
};
static_assert(sizeof(BSmem) <= 227 * 1024, "cluster scan backward shared memory");
static_assert(offsetof(BSmem, x) % 16 == 0 && offsetof(BSmem, buf) % 16 == 0 && offsetof(BSmem, xch) % 16 == 0 && offsetof(BSmem, xl) % 16 == 0 &&
              offsetof(BSmem, xg) % 16 == 0 && offsetof(BSmem, xbar) % 8 == 0, "bulk copies need 16-byte aligned blocks");
static_assert(offsetof(CCommon, buf1) % 16 == 0 && offsetof(CCommon, xch) % 16 == 0 && offsetof(CCommon, idx) % 16 == 0 &&
              offsetof(CCommon, xidx) % 16 == 0 && offsetof(CCommon, xbar) % 8 == 0, "bulk copies need 16-byte aligned blocks");
static_assert(2 * kWarps * kRows * 17 >= 4 * kRows * (kZC + 1), "red holds the S5 partial sums");

__device__ __forceinline__ float silu_grad_fast(float n) {
  const float sg = sigmoid_fast(n);
  return sg * (1.f + n * (1.f - sg));
}
// A fragment (rows g, g + 8; k pairs 2 tq, 2 tq + 8 of [kb, kb + 16)) from a bf16 [row][k] operand with row stride ld
__device__ __forceinline__ void load_a_frag16(const __nv_bfloat16* a16, int ld, int kb, int g, int tq, uint32_t (&a)[4]) {
  const uint32_t* r0 = reinterpret_cast<const uint32_t*>(a16 + (size_t)g * ld + kb + 2 * tq);
  const uint32_t* r1 = reinterpret_cast<const uint32_t*>(a16 + (size_t)(g + 8) * ld + kb + 2 * tq);
  a[0] = r0[0]; a[1] = r1[0]; a[2] = r0[4]; a[3] = r1[4];
}
__device__ __forceinline__ void store_c_frag17(float* red, int slot, int nt, int g, int tq, const float (&c)[4]) {
  const int col = nt * 8 + 2 * tq;
  float* r0 = red + (size_t)(slot * kRows + g) * 17 + col;
  float* r1 = red + (size_t)(slot * kRows + g + 8) * 17 + col;
  r0[0] = c[0]; r0[1] = c[1]; r1[0] = c[2]; r1[1] = c[3];
}
// In-place LayerNorm+SiLU backward of the 16 rows staged as buf[k][row]: warp w handles rows w and w + 8, lane the columns
// lane + 32 i. xr / mu / rs: the forward pre-activations and statistics of those two rows (loaded by the caller at the top
// of the step). Row b of the result also goes to out (global, [B, T, D] at step t) from CTA `rank == b`.
__device__ __forceinline__ void ln_bwd_inplace(float* buf, const float (&xr)[2][kG / 32], const float (&mu)[2], const float (&rs)[2],
                                               const float* gamma, const float* beta, int B, unsigned rank, float* out, size_t row_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int EPL = kG / 32;
  float xh[2][EPL], d[2][EPL], s1[2], s2[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int b = warp + r * kWarps;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = lane + 32 * i;
      const float gm = gamma[c];
      const float xhat = (xr[r][i] - mu[r]) * rs[r];
      const float dv = buf[c * kLd + b] * silu_grad_fast(xhat * gm + beta[c]) * gm;
      xh[r][i] = xhat; d[r][i] = dv;
      a1 += dv; a2 += dv * xhat;
    }
    s1[r] = a1; s2[r] = a2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
      s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int b = warp + r * kWarps;
    const float m1 = s1[r] / (float)kG, m2 = s2[r] / (float)kG;
    const bool row_ok = b < B, wr = row_ok && (unsigned)b == rank;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = lane + 32 * i;
      const float o = row_ok ? rs[r] * (d[r][i] - m1 - xh[r][i] * m2) : 0.f;
      buf[c * kLd + b] = o;
      if (wr) out[(size_t)b * row_stride + c] = o;
    }
  }
  __syncthreads();
}
__global__ void __cluster_dims__(kC, 1, 1) __launch_bounds__(kThreads, 1) observe_scan_bwd_cluster_kernel(ScanBwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BSmem& sm = *reinterpret_cast<BSmem*>(smem_raw);
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const int B = p.B, T = p.T, G = kG, D = kG, Z = p.Z, W = kG + p.Z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, tq = lane & 3;
  const int u0 = rank * kU, z0c = rank * kZC;
  constexpr int EPL = kG / 32;

  // ---- resident weights: rows of the untransposed matrices, fp32 -> bf16 ----
  for (int e = tid; e < kU * Z; e += kThreads) { const int n = e / Z, k = e - n * Z; sm.wA[n][k] = __float2bfloat16_rn(p.W_zq[(size_t)(u0 + n) * Z + k]); }
  for (int e = tid; e < kU * D; e += kThreads) { const int n = e / D, k = e - n * D; sm.wB[n][k] = __float2bfloat16_rn(p.W_post_h[(size_t)(u0 + n) * D + k]); }
  for (int e = tid; e < kU * 3 * G; e += kThreads) {
    const int n = e / (3 * G), k = e - n * 3 * G;
    sm.wC[n][k] = __float2bfloat16_rn(p.gru_U[(size_t)(u0 + n) * 3 * G + k]);
    sm.wD[n][k] = __float2bfloat16_rn(p.gru_K[(size_t)(u0 + n) * 3 * G + k]);
  }
  for (int e = tid; e < kZC * D; e += kThreads) { const int n = e / D, k = e - n * D; sm.wE[n][k] = __float2bfloat16_rn(p.W_pre[(size_t)(z0c + n) * D + k]); }
  for (int e = tid; e < kG; e += kThreads) { sm.lnp[0][e] = p.post_g[e]; sm.lnp[1][e] = p.post_b[e]; sm.lnp[2][e] = p.pre_g[e]; sm.lnp[3][e] = p.pre_b[e]; }
  for (int e = tid; e < kRows * (kZC + 1); e += kThreads) (&sm.dzc[0][0])[e] = 0.f;
  for (int e = tid; e < kRows * kLdXl; e += kThreads) (&sm.xl[0][0])[e] = __float2bfloat16_rn(0.f);
  for (int e = tid; e < 4 * kRows * kLdXg; e += kThreads) (&sm.xg[0][0][0])[e] = __float2bfloat16_rn(0.f);
  for (int e = tid; e < 2 * kU * kLd; e += kThreads) (&sm.xch[0][0][0])[e] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) xbar_init(&sm.xbar[i]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cl.sync();

  const int b = tid >> 4, j = tid & 15;   // (row, owned column / unit) of this thread in the 16 x 16 epilogues
  const bool row_ok = b < B;
  float dh_carry = 0.f;                    // (1 - is_first[t+1]) dh_in[t+1] of (row b, unit u0 + j)
  auto stamp = [&](int i) {
    if (p.stamps != nullptr && rank == 0 && tid == 0) {
      unsigned long long v;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      p.stamps[i] = v;
    }
  };
  stamp(0);
  // S1 is the first thing a step does with data from global memory: its operands are fetched one step ahead (before S5 of
  // the step before), so that no L2 round trip sits between the LayerNorm backward of S5 and the categorical backward of S1
  float q_l[4], q_dp[4], q_dz[4];
  auto load_s1 = [&](int tt) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int pr = warp + q4 * kWarps, r = pr >> 1, cl_ = pr & 1;
      const bool ok = r < B;
      const size_t rt = (size_t)(ok ? r : 0) * T + tt;
      const int col = z0c + cl_ * 32 + lane;
      q_l[q4] = ok ? p.post_logits[rt * Z + col] : 0.f;
      q_dp[q4] = ok ? p.dpost[rt * Z + col] : 0.f;
      q_dz[q4] = ok ? p.d_hz[rt * W + G + col] : 0.f;
    }
  };
  load_s1(T - 1);

  for (int t = T - 1; t >= 0; --t) {
    const int si = (T - 1 - t) * 8;
    const size_t bt = (size_t)(row_ok ? b : 0) * T + t;
    const uint32_t par = (uint32_t)(T - 1 - t) & 1u;   // phase parity of this step's four exchanges
    if (tid == 0) {   // (each barrier's previous phase was waited for in the previous step)
      xbar_arm(&sm.xbar[0], kC * (uint32_t)sizeof(sm.xl)); xbar_arm(&sm.xbar[1], kC * kBlkBytes);
      xbar_arm(&sm.xbar[2], kC * (uint32_t)sizeof(sm.xg)); xbar_arm(&sm.xbar[3], kC * kBlkBytes);
    }
    // ===== S1: categorical backward (representation_layer.py:73-113) of my 2 categoricals x 16 rows, one warp per pair; the four
    // pairs of a warp go through the shuffle reductions together =====
    {
      float mx[4], ex[4], sden[4], prb[4], gz[4], dot[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) mx[q4] = q_l[q4];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) mx[q4] = fmaxf(mx[q4], __shfl_xor_sync(0xffffffffu, mx[q4], o));
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) { ex[q4] = __expf(q_l[q4] - mx[q4]); sden[q4] = ex[q4]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) sden[q4] += __shfl_xor_sync(0xffffffffu, sden[q4], o);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int pr = warp + q4 * kWarps, r = pr >> 1, cl_ = pr & 1;
        prb[q4] = __fdividef(ex[q4], sden[q4]);
        gz[q4] = 0.99f * (q_dz[q4] + sm.dzc[r][cl_ * 32 + lane] + q_dp[q4]);
        dot[q4] = gz[q4] * prb[q4];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) dot[q4] += __shfl_xor_sync(0xffffffffu, dot[q4], o);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int pr = warp + q4 * kWarps, r = pr >> 1, cl_ = pr & 1;
        const bool ok = r < B;
        const float dl = ok ? prb[q4] * (gz[q4] - dot[q4]) : 0.f;
        if (ok) p.dlogits[((size_t)r * T + t) * Z + z0c + cl_ * 32 + lane] = dl;
        sm.xl[r][cl_ * 32 + lane] = __float2bfloat16_rn(dl);
      }
    }
    stamp(T * 8 + 16 + (T - 1 - t) * 4 + 2);
    xchg_send(&sm.x.dl[0][0][0], &sm.xl[0][0], (uint32_t)sizeof(sm.xl), &sm.xbar[0], rank);   // my block -> dl[rank] of every peer
    stamp(T * 8 + 16 + (T - 1 - t) * 4 + 3);
    // ---- everything else this step reads from the forward pass / the head gradients: issued here, behind the broadcast the peers
    // wait for, and in flight across the barrier (first use in S3) ----
    // (the S1 operands - logits, dKL/dprobs, d_hz.z of my (row, categorical) pairs - were loaded one step ahead: q_l, q_dp, q_dz)
    float xp[2][EPL], mp[2], rp[2];   // S3: LayerNorm inputs of rows warp, warp + 8 (those of S5 are loaded behind the S3 send)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int rb = warp + r * kWarps;
      const bool ok = rb < B;
      const size_t rt = (size_t)(ok ? rb : 0) * T + t;
      mp[r] = ok ? p.p_mean[rt] : 0.f; rp[r] = ok ? p.p_rstd[rt] : 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) xp[r][i] = ok ? p.p_pre[rt * D + lane + 32 * i] : 0.f;
    }
    const int u = u0 + j;                      // S3 / S4: GRU operands of (row b, unit u)
    const float e_d = row_ok ? p.d_hz[bt * W + u] : 0.f;
    const float e_z = row_ok ? p.gates[bt * 3 * G + u] : 0.f;
    const float e_r = row_ok ? p.gates[bt * 3 * G + G + u] : 0.f;
    const float e_c = row_ok ? p.gates[bt * 3 * G + 2 * G + u] : 0.f;
    const float e_h = row_ok ? p.h_in[bt * G + u] : 0.f;
    const float e_g = row_ok ? p.gh[bt * 3 * G + 2 * G + u] : 0.f;
    const float keep = row_ok ? 1.f - p.is_first[bt] : 0.f;
    float keep4[4];                            // S5: rows of the 4 dz_in columns this thread finishes (same row b)
#pragma unroll
    for (int m = 0; m < 4; ++m) keep4[m] = keep;

    stamp(si + 1);
    xbar_wait(&sm.xbar[0], par);
    stamp(si + 2);

    // ===== S2: dp_act[:, my 16 columns] = dlogits . W_zq^T (K = 1024: 8 k-steps per warp) =====
    {
      float c[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
#pragma unroll
      for (int st = 0; st < 8; ++st) {
        const int kb = 128 * warp + 16 * st;
        uint32_t a[4];
        load_a_frag16(&sm.x.dl[kb >> 6][0][0], kLdXl, kb & 63, g8, tq, a);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t* bp = reinterpret_cast<const uint32_t*>(&sm.wA[nt * 8 + g8][kb + 2 * tq]);
          mma_bf16_16816(c[nt], a[0], a[1], a[2], a[3], bp[0], bp[4]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) store_c_frag17(sm.red, warp, nt, g8, tq, c[nt]);
      __syncthreads();
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += sm.red[(w * kRows + b) * 17 + j];
      if (row_ok) p.dp_act[bt * D + u0 + j] = s; else s = 0.f;
      send_block(sm.xch[0], &sm.buf[0][0], &sm.xbar[1], rank, s);
    }
    stamp(si + 3);
    xbar_wait(&sm.xbar[1], par);
    stamp(si + 4);

    // ===== S3: dp_pre (whole rows, every CTA); dh of my units; GRU gate gradients =====
    ln_bwd_inplace(&sm.buf[0][0], xp, mp, rp, sm.lnp[0], sm.lnp[1], B, rank, p.dp_pre + (size_t)t * D, (size_t)T * D);
    float dhz = 0.f;   // dh z of (row b, unit u): the direct path into dh_in
    {
      float c[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const int kb = 32 * warp + 16 * st;
        uint32_t a[4];
        load_a_frag(&sm.buf[0][0], kb, g8, tq, a);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t* bp = reinterpret_cast<const uint32_t*>(&sm.wB[nt * 8 + g8][kb + 2 * tq]);
          mma_bf16_16816(c[nt], a[0], a[1], a[2], a[3], bp[0], bp[4]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) store_c_frag17(sm.red, warp, nt, g8, tq, c[nt]);
      __syncthreads();
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += sm.red[(w * kRows + b) * 17 + j];
      float dz = 0.f, dr = 0.f, dc = 0.f, dcr = 0.f;
      if (row_ok) {
        const float d = s + e_d + dh_carry;
        dz = d * (e_h - e_c) * e_z * (1.f - e_z);
        dc = d * (1.f - e_z) * (1.f - e_c * e_c);
        dr = dc * e_g * e_r * (1.f - e_r);
        dcr = dc * e_r;
        dhz = d * e_z;
        float* gxo = p.dgx + bt * 3 * G;
        float* gho = p.dgh + bt * 3 * G;
        gxo[u] = dz; gxo[G + u] = dr; gxo[2 * G + u] = dc;
        gho[u] = dz; gho[G + u] = dr; gho[2 * G + u] = dcr;
      }
      sm.xg[0][b][j] = __float2bfloat16_rn(dz); sm.xg[1][b][j] = __float2bfloat16_rn(dr);
      sm.xg[2][b][j] = __float2bfloat16_rn(dc); sm.xg[3][b][j] = __float2bfloat16_rn(dcr);
      xchg_send(&sm.x.dg[0][0][0][0], &sm.xg[0][0][0], (uint32_t)sizeof(sm.xg), &sm.xbar[2], rank);   // my 4 blocks -> dg[rank] of every peer
    }
    float xx[2][EPL], mx_[2], rx[2];   // S5: LayerNorm inputs of rows warp, warp + 8, in flight while the gate gradients travel
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int rb = warp + r * kWarps;
      const bool ok = rb < B;
      const size_t rt = (size_t)(ok ? rb : 0) * T + t;
      mx_[r] = ok ? p.x_mean[rt] : 0.f; rx[r] = ok ? p.x_rstd[rt] : 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) xx[r][i] = ok ? p.x_pre[rt * D + lane + 32 * i] : 0.f;
    }
    stamp(si + 5);
    xbar_wait(&sm.xbar[2], par);
    stamp(si + 6);

    // ===== S4: dh_in[:, my units] = dh z + [dz dr dc r] . U^T;  du[:, my columns] = [dz dr dc] . K^T (K = 768: 6 k-steps per warp) =====
    {
      float ch[2][4], cu[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        ch[nt][0] = 0.f; ch[nt][1] = 0.f; ch[nt][2] = 0.f; ch[nt][3] = 0.f;
        cu[nt][0] = 0.f; cu[nt][1] = 0.f; cu[nt][2] = 0.f; cu[nt][3] = 0.f;
      }
#pragma unroll
      for (int st = 0; st < 6; ++st) {
        const int ks = 6 * warp + st, gate = ks >> 4, kk = (ks & 15) * 16, kw = gate * kG + kk;
        uint32_t ax[4], ah[4];
        load_a_frag16(&sm.x.dg[kk >> 4][gate][0][0], kLdXg, 0, g8, tq, ax);
        if (gate == 2) load_a_frag16(&sm.x.dg[kk >> 4][3][0][0], kLdXg, 0, g8, tq, ah);
        else { ah[0] = ax[0]; ah[1] = ax[1]; ah[2] = ax[2]; ah[3] = ax[3]; }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t* bu = reinterpret_cast<const uint32_t*>(&sm.wC[nt * 8 + g8][kw + 2 * tq]);
          const uint32_t* bk = reinterpret_cast<const uint32_t*>(&sm.wD[nt * 8 + g8][kw + 2 * tq]);
          mma_bf16_16816(ch[nt], ah[0], ah[1], ah[2], ah[3], bu[0], bu[4]);
          mma_bf16_16816(cu[nt], ax[0], ax[1], ax[2], ax[3], bk[0], bk[4]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { store_c_frag17(sm.red, warp, nt, g8, tq, ch[nt]); store_c_frag17(sm.red, kWarps + warp, nt, g8, tq, cu[nt]); }
      __syncthreads();
      float sh = 0.f, su = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { sh += sm.red[(w * kRows + b) * 17 + j]; su += sm.red[((kWarps + w) * kRows + b) * 17 + j]; }
      if (row_ok) {
        const float v = sh + dhz;
        p.dh_in[bt * G + u] = v;
        dh_carry = keep * v;
        p.du[bt * D + u0 + j] = su;
      } else {
        su = 0.f;
      }
      send_block(sm.xch[1], &sm.buf[0][0], &sm.xbar[3], rank, su);
    }
    stamp(si + 7);
    xbar_wait(&sm.xbar[3], par);
    stamp(si + 8);

    // ===== S5: dx_pre (whole rows, every CTA); dz_in of my 64 columns stays here for S1 of step t-1 =====
    if (t > 0) load_s1(t - 1);
    ln_bwd_inplace(&sm.buf[0][0], xx, mx_, rx, sm.lnp[2], sm.lnp[3], B, rank, p.dx_pre + (size_t)t * D, (size_t)T * D);
    stamp(T * 8 + 16 + (T - 1 - t) * 4 + 0);
    if (t > 0) {
      mma_pass<64>(&sm.buf[0][0], &sm.wE[0][0], sm.red);
#pragma unroll
      for (int m = 0; m < 4; ++m) sm.dzc[b][j + 16 * m] = keep4[m] * red_sum<64>(sm.red, b, j + 16 * m);
      __syncthreads();
    }
    stamp(T * 8 + 16 + (T - 1 - t) * 4 + 1);
  }
  cl.sync();   // no CTA leaves while a peer may still be reading its staging blocks
}

}  // namespace

// The cluster version needs the XS layer sizes (everything resident in 16 SMs' shared memory).
bool scan_cluster_supported(const ScanFwdParams& p) {
  return p.G == kG && p.D == kG && p.Z == kC * kZC && p.Zk == 32 && p.Zc == 32 && p.B <= kRows && std::getenv("DV3_NO_CLUSTER_SCAN") == nullptr;
}

template <bool kTc>
static cudaError_t launch_cluster_variant(cudaStream_t s, const ScanFwdParams& p) {
  using SmemT = typename std::conditional<kTc, CSmemTc, CSmem>::type;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(observe_scan_cluster_kernel<kTc>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(observe_scan_cluster_kernel<kTc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  observe_scan_cluster_kernel<kTc><<<dim3(kC), dim3(kThreads), sizeof(SmemT), s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  return e;
}

// bf16: the tensor-core variant (compute mode 1): bf16 weights resident on chip (the one-hot gather of W_pre included), mma.sync
// contractions with fp32 accumulation; everything else (LayerNorm, gates, softmax, the fp64 inverse-CDF draw) as in the fp32 variant
cudaError_t launch_observe_scan_cluster(cudaStream_t s, const ScanFwdParams& p, bool bf16) {
  const bool no_tc = std::getenv("DV3_NO_CLUSTER_TC") != nullptr;
  return (bf16 && !no_tc) ? launch_cluster_variant<true>(s, p) : launch_cluster_variant<false>(s, p);
}

// The backward cluster kernel: same size limits as the forward one; tensor-core mode only (its contractions are bf16)
bool scan_bwd_cluster_supported(const ScanBwdParams& p) {
  return p.G == kG && p.D == kG && p.Z == kC * kZC && p.Zk == 32 && p.Zc == 32 && p.B <= kRows && p.W_zq != nullptr &&
         std::getenv("DV3_NO_CLUSTER_SCAN") == nullptr && std::getenv("DV3_NO_CLUSTER_BWD") == nullptr;
}

cudaError_t launch_observe_scan_bwd_cluster(cudaStream_t s, const ScanBwdParams& p) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(observe_scan_bwd_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(observe_scan_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BSmem));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  observe_scan_bwd_cluster_kernel<<<dim3(kC), dim3(kThreads), sizeof(BSmem), s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) count_launch();
  return e;
}


}  // namespace dv3
