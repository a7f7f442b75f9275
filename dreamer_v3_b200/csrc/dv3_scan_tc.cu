// Weight-streaming tensor-core observe scan for the layer sizes whose step weights do not fit a thread-block cluster
// (S ... XL; world_model.py:244-273), tensor-core mode only (compute_mode 1). ONE persistent cooperative launch walks all T
// timesteps. The B <= 16 replay rows are the N operand of swap-AB tcgen05 MMAs:
//
//     out^T[128 weight rows, 16 batch rows] += Wt[128, 64 k] (bf16, TMA, SWIZZLE_128B) x act[16, 64 k] (bf16, staged by the CTA)
//
// i.e. the bf16 K-major weight shadows ([N_out, K], the same ones the N-row head contractions use) are the M = 128 operand and
// are streamed exactly once per step by TMA through a ring of 16 KB k-blocks; the fp32 accumulator tile lives in TMEM
// (16 columns). A stage's work = all (tile, k-block) units of its contractions in (k-chunk, tile, k-block) order, cut into 148
// equal contiguous ranges (stream-K): every SM streams the same number of weight bytes, a range touches at most two k-chunks of
// activations, and the partial [16 x 128] products of a range are added to the stage output with fp32 red.global (L2).
//
// Warp roles: warp 0 = TMA producer. The weights do not depend on the recurrence, so the producer never waits for a grid
//             barrier: while the CTA synchronises, the ring already holds the next stage's first k-blocks.
//             warp 1 = MMA issuer (one lane; owns the TMEM allocation).
//             warps 2-5 = activation stagers + epilogue + sampling. Every elementwise function of the step is applied by the
//             CONSUMER while it stages its k-chunk as the bf16 K-major B operand: LayerNorm + SiLU of the pre-GRU layer, the
//             GRU cell, LayerNorm + SiLU of the posterior layer. The CTA that owns tile 0 of a k-block also writes the fp32
//             values the backward pass needs (u, h_in, gates, h, p_act, row statistics).
// Per step, 4 stages separated by a one-atomic grid barrier:
//   S1  gx += u . K          gh += h_in . U             (u = SiLU(LN(x_pre)),  h_in = is_first ? h0 : h_{t-1})
//   S2  p_pre += h_t . W_post[E:]                        (h_t = GRU(gx, gh, h_in))
//   S3  logits += SiLU(LN(p_pre)) . W_zq
//   S4  one warp per (row, categorical): softmax, 1 % unimix, fp64 inverse-CDF sample -> z_t; then the pre-GRU layer of
//       step t+1 receives its one-hot row gather x_pre[t+1] += W_pre[c*Zk + z] (fp32 rows, vector red.global)
// Accumulators are pre-filled with their biases (gx: b0, gh: b1, logits: b_zq, x_pre: the hoisted action part) by one
// elementwise kernel before the launch.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "dv3_device.cuh"
#include "dv3_scan.h"

namespace dv3 {
namespace {

constexpr int kThreadsTc = 192;
constexpr int kWorkers = 128;        // warps 2..5
constexpr int kRing = 7;             // weight k-blocks (128 rows x 64 k bf16 = 16 KB) in flight per CTA
constexpr int kMaxKc = 16;           // k-blocks per activation chunk (<= 1024 k)
constexpr int kSlots = 2;            // activation chunks resident per stage
constexpr int kRows = 16;
constexpr int kKbBytes = 2048;       // one staged activation k-block: 16 rows x 128 B
constexpr int kWBytes = 128 * 128;   // one weight k-block

struct TcProblem { int n_out, K, tiles, kblocks, kc, span, unit_begin, units; };

struct TcMaps { CUtensorMap m[4]; };

struct ScanTcParams {
  ScanFwdParams p;
  float* gx_acc;        // [B, T, 3G], pre-filled with b0; p.gh is pre-filled with b1, p.post_logits with b_zq, p.x_pre with actW
  TcProblem prob[4];    // 0: gx (K = D), 1: gh (K = G), 2: posterior h-part (K = G), 3: posterior logits (K = D)
  int stage_units[3];
};

struct TcSmem {
  unsigned char ring[kRing][kWBytes];
  unsigned char bbuf[kSlots][kMaxKc][kKbBytes];
  float stats[kRows][2];
  uint64_t full[kRing], empty[kRing], acc_full[2], acc_empty[2], b_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, bf16 x bf16 -> fp32, both K-major, M = 128, N = 16
__device__ __forceinline__ uint32_t make_idesc16() { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) {  // never hang the GPU on a lost arrival
      printf("dv3 scan_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
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
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// grid barrier entered by the 128 worker threads of every CTA (the TMA / MMA warps never take part: the weights do not depend
// on the recurrence and the MMA warp is gated by b_ready). One monotonically increasing counter, zeroed by the launcher.
__device__ __forceinline__ void grid_bar_workers(unsigned int* ctr, unsigned int& target, int wt) {
  worker_sync();
  if (wt == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (++spins > (1u << 26)) { printf("dv3 scan_tc: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    } while (v < target);
  }
  worker_sync();
}

struct Seg { int prob, tile, chunk, kb0, kb1; };
// the CTA's contiguous unit range of a stage, cut into segments (one output tile, consecutive k-blocks of one k-chunk)
struct SegIter {
  int u, u_end, stage;
  __device__ SegIter(const ScanTcParams& P, int st) : stage(st) {
    const long long U = P.stage_units[st];
    u = (int)(U * blockIdx.x / gridDim.x);
    u_end = (int)(U * (blockIdx.x + 1) / gridDim.x);
  }
  __device__ bool next(const ScanTcParams& P, Seg& s) {
    if (u >= u_end) return false;
    int pr = stage == 0 ? 0 : stage + 1;
    if (stage == 0 && u >= P.prob[1].unit_begin) pr = 1;
    const TcProblem& q = P.prob[pr];
    const int local = u - q.unit_begin;
    const int chunk = local / q.span, r = local - chunk * q.span;
    const int tile = r / q.kc, kbi = r - tile * q.kc;
    const int n = min(q.kc - kbi, min(u_end, q.unit_begin + q.units) - u);
    s.prob = pr; s.tile = tile; s.chunk = chunk; s.kb0 = kbi; s.kb1 = kbi + n;
    u += n;
    return true;
  }
};
// activation-chunk slot of a segment: chunks are numbered in order of first use within the stage (<= kSlots per stage, checked
// by the launcher); identical in the MMA and the worker warps
struct SlotMap {
  int key[kSlots], n;
  __device__ SlotMap() : n(0) { key[0] = key[1] = -1; }
  __device__ int get(const Seg& s, bool& fresh) {
    const int k = s.prob * 4096 + s.chunk;
    fresh = false;
    for (int i = 0; i < n; ++i) if (key[i] == k) return i;
    fresh = true;
    key[n & (kSlots - 1)] = k;
    return (n++) & (kSlots - 1);
  }
};

// LayerNorm statistics of the 16 rows x[b, t, 0:D] (fp32, read through L2) into sm.stats; worker warp w owns rows 4w .. 4w+3.
template <int EPL>   // float4 per lane per row = D / 128
__device__ __forceinline__ void row_stats_t(TcSmem& sm, const float* x, int T, int t, int B, int D, int ww, int lane) {
  float4 v[4][EPL];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = ww * 4 + r;
#pragma unroll
    for (int i = 0; i < EPL; ++i)
      v[r][i] = b < B ? ldcg4(x + ((size_t)b * T + t) * D + (size_t)(lane + 32 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s[4], q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    s[r] /= (float)D;
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const float a = v[r][i].x - s[r], b_ = v[r][i].y - s[r], c = v[r][i].z - s[r], d = v[r][i].w - s[r];
      q[r] += (a * a + b_ * b_) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) { sm.stats[ww * 4 + r][0] = s[r]; sm.stats[ww * 4 + r][1] = rsqrtf(q[r] / (float)D + DV3_LN_EPS); }
  }
}
__device__ __forceinline__ void row_stats(TcSmem& sm, const float* x, int T, int t, int B, int D, int ww, int lane) {
  switch (D / 128) {
    case 4: row_stats_t<4>(sm, x, T, t, B, D, ww, lane); break;
    case 5: row_stats_t<5>(sm, x, T, t, B, D, ww, lane); break;
    case 6: row_stats_t<6>(sm, x, T, t, B, D, ww, lane); break;
    default: row_stats_t<8>(sm, x, T, t, B, D, ww, lane); break;
  }
}

__device__ __forceinline__ void silu8(float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] * sigmoid_fast(v[i]);
}
__device__ __forceinline__ void load8(float (&v)[8], const float* p) {
  const float4 a = ldcg4(p), b = ldcg4(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_ro(float (&v)[8], const float* __restrict__ p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// Stages activation k-block `kb` (absolute, 64 k) of problem `prob` at step t as the bf16 K-major SWIZZLE_128B B operand.
// Worker thread wt: row = wt / 8, 16-byte chunk (8 k) = wt % 8. `writer`: this CTA owns tile 0 of the k-block and stores the
// fp32 values for the backward pass.
__device__ __forceinline__ void stage_kb(const ScanTcParams& P, TcSmem& sm, unsigned char* dst, int prob, int kb, int t, bool writer, int wt) {
  const ScanFwdParams& p = P.p;
  const int b = wt >> 3, c8 = wt & 7, k = kb * 64 + c8 * 8;
  const int T = p.T, G = p.G, D = p.D;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (b < p.B) {
    const size_t bt = (size_t)b * T + t;
    if (prob == 0 || prob == 3) {   // SiLU(LayerNorm(x)) of the pre-GRU (0) / posterior (3) layer
      const float* x = (prob == 0 ? p.x_pre : p.p_pre) + bt * D + k;
      float g[8], be[8];
      load8(v, x);
      load8_ro(g, (prob == 0 ? p.pre_g : p.post_g) + k);
      load8_ro(be, (prob == 0 ? p.pre_b : p.post_b) + k);
      const float mean = sm.stats[b][0], rstd = sm.stats[b][1];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + be[i];
      silu8(v);
      if (writer) {
        store8((prob == 0 ? p.u_act : p.p_act) + bt * D + k, v);
        if (kb == 0 && c8 == 0) {
          (prob == 0 ? p.x_mean : p.p_mean)[bt] = mean;
          (prob == 0 ? p.x_rstd : p.p_rstd)[bt] = rstd;
        }
      }
    } else if (prob == 1) {          // masked recurrent input h_in (world_model.py:246-247)
      const bool first = (t == 0) || (p.is_first[bt] != 0.f);
      if (first) load8_ro(v, p.h0 + k);
      else load8(v, p.hz + (bt - 1) * (G + p.Z) + k);
      if (writer) store8(p.h_in + bt * G + k, v);
    } else {                         // h_t = GRU cell (Keras reset_after; gx / gh already hold their biases)
      const float* gx = P.gx_acc + bt * 3 * G + k;
      const float* gh = p.gh + bt * 3 * G + k;
      float xz[8], xr[8], xc[8], hz_[8], hr[8], hc[8], hin[8];
      load8(xz, gx); load8(xr, gx + G); load8(xc, gx + 2 * G);
      load8(hz_, gh); load8(hr, gh + G); load8(hc, gh + 2 * G);
      const bool first = (t == 0) || (p.is_first[bt] != 0.f);
      if (first) load8_ro(hin, p.h0 + k);
      else load8(hin, p.hz + (bt - 1) * (G + p.Z) + k);
      float z[8], r[8], c[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        z[i] = sigmoid_fast(xz[i] + hz_[i]);
        r[i] = sigmoid_fast(xr[i] + hr[i]);
        c[i] = tanh_fast(xc[i] + r[i] * hc[i]);
        v[i] = z[i] * hin[i] + (1.f - z[i]) * c[i];
      }
      if (writer) {
        float* gt = p.gates + bt * 3 * G + k;
        store8(gt, z); store8(gt + G, r); store8(gt + 2 * G, c);
        store8(p.hz + bt * (G + p.Z) + k, v);
      }
    }
  }
  const uint32_t a = smem_u32(dst) + (uint32_t)((b >> 3) * 1024 + (b & 7) * 128 + ((c8 ^ (b & 7)) << 4));
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
               "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
}

// one (row, categorical) of stage S4 per warp; also used for the prologue (t = -1: only the gather of step 0)
__device__ __forceinline__ void sample_and_gather(const ScanTcParams& P, int b, int cat, int t, int lane) {
  const ScanFwdParams& p = P.p;
  const int T = p.T, G = p.G, D = p.D, Z = p.Z, Zc = p.Zc, Zk = p.Zk, W = G + Z;
  int pick = 0;
  if (t >= 0) {
    const size_t bt = (size_t)b * T + t;
    const bool act = lane < Zk;
    const float l = act ? __ldcg(p.post_logits + bt * Z + cat * Zk + lane) : -INFINITY;
    const float mx = warp_max(l);
    const float ex = act ? expf(l - mx) : 0.f;
    const float s = warp_sum(ex);
    const float pu = 0.99f * (ex / s) + 0.01f * (1.0f / (float)Zk);
    if (act) p.post_probs[bt * Z + cat * Zk + lane] = pu;
    pick = categorical_pick(pu, act, Zk, p.u_post[((size_t)t * p.B + b) * Zc + cat], lane);
    if (lane == 0) p.z_idx[bt * Zc + cat] = pick;
    if (act) p.hz[bt * W + G + cat * Zk + lane] = (lane == pick) ? 1.f : 0.f;
  }
  if (t + 1 < T) {   // masked z input of the next step (world_model.py:249-250) and its row of the pre-GRU layer
    const size_t bn = (size_t)b * T + t + 1;
    const bool first = (t + 1 == 0) || (p.is_first[bn] != 0.f);
    const int idx = first ? p.z0_idx[cat] : pick;
    if (lane < Zk) p.z_in[bn * Z + cat * Zk + lane] = (lane == idx) ? 1.f : 0.f;
    const float* wrow = p.W_pre + (size_t)(cat * Zk + idx) * D;
    float* xrow = p.x_pre + bn * D;
    for (int j = lane * 4; j < D; j += 128) red_add_v4(xrow + j, __ldg(reinterpret_cast<const float4*>(wrow + j)));
  }
}

__global__ void __launch_bounds__(kThreadsTc, 1) observe_scan_tc_kernel(const __grid_constant__ TcMaps maps, const ScanTcParams P) {
  extern __shared__ unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ScanFwdParams& p = P.p;
  const int T = p.T;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
    for (int s = 0; s < kRing; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], kWorkers); }
    mbar_init(&sm.b_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: the whole weight stream of all T steps, throttled only by the ring ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = 0; t < T; ++t)
        for (int st = 0; st < 3; ++st) {
          SegIter si(P, st);
          Seg s;
          while (si.next(P, s)) {
            const TcProblem& q = P.prob[s.prob];
            for (int kb = s.kb0; kb < s.kb1; ++kb, ++it) {
              const int slot = it % kRing;
              mbar_wait(&sm.empty[slot], ((it / kRing) & 1) ^ 1);
              mbar_expect_tx(&sm.full[slot], kWBytes);
              tma_load_2d(smem_u32(sm.ring[slot]), &maps.m[s.prob], (s.chunk * q.kc + kb) * 64, s.tile * 128, &sm.full[slot]);
            }
          }
        }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc16();
      uint32_t it = 0, acc_it = 0, bphase = 0;
      for (int t = 0; t < T; ++t)
        for (int st = 0; st < 3; ++st) {
          SegIter si(P, st);
          SlotMap slots;
          Seg s;
          bool any = false;
          while (si.next(P, s)) {
            if (!any) {   // the stage's activation chunks have been staged by the worker warps
              mbar_wait(&sm.b_ready, bphase & 1);
              ++bphase;
              any = true;
            }
            bool fresh;
            const int sl = slots.get(s, fresh);
            const uint32_t as = acc_it & 1;
            mbar_wait(&sm.acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + as * 16;
            for (int kb = s.kb0; kb < s.kb1; ++kb, ++it) {
              const int slot = it % kRing;
              mbar_wait(&sm.full[slot], (it / kRing) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_addr = smem_u32(sm.ring[slot]), b_addr = smem_u32(sm.bbuf[sl][kb]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t da = make_desc(a_addr + j * 32), db = make_desc(b_addr + j * 32);
                const uint32_t accum = (kb == s.kb0 && j == 0) ? 0u : 1u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
                    : "memory");
              }
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[slot])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.acc_full[as])) : "memory");
            ++acc_it;
          }
        }
    }
  } else {
    // ---------------- workers: stage activations, drain accumulators, sample ----------------
    const int wt = tid - 64, ww = warp - 2, q4 = warp & 3;
    unsigned int bar_target = 0;
    uint32_t acc_it = 0;
    const int gwarp = blockIdx.x * 4 + ww, nwarps = gridDim.x * 4;
    const int ntask = p.B * p.Zc;
    // prologue: every row starts from the learned initial state: x_pre[0] += gather of z0, z_in[0] = z0
    for (int task = gwarp; task < ntask; task += nwarps) sample_and_gather(P, task / p.Zc, task % p.Zc, -1, lane);
    grid_bar_workers(p.bar, bar_target, wt);

    for (int t = 0; t < T; ++t) {
      for (int st = 0; st < 3; ++st) {
        // ---- stage the B operand of every segment of this CTA ----
        {
          SegIter si(P, st);
          SlotMap slots;
          Seg s;
          unsigned int staged[kSlots] = {0u, 0u};
          bool any = false, have_stats = false;
          while (si.next(P, s)) {
            any = true;
            if ((s.prob == 0 || s.prob == 3) && !have_stats) {
              row_stats(sm, s.prob == 0 ? p.x_pre : p.p_pre, T, t, p.B, p.D, ww, lane);
              worker_sync();
              have_stats = true;
            }
            bool fresh;
            const int sl = slots.get(s, fresh);
            if (fresh) staged[sl] = 0u;
            const TcProblem& q = P.prob[s.prob];
            for (int kb = s.kb0; kb < s.kb1; ++kb) {
              if (staged[sl] & (1u << kb)) continue;
              staged[sl] |= 1u << kb;
              stage_kb(P, sm, sm.bbuf[sl][kb], s.prob, s.chunk * q.kc + kb, t, s.tile == 0, wt);
            }
          }
          if (any) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            worker_sync();
            if (wt == 0) mbar_arrive(&sm.b_ready);
          }
        }
        // ---- epilogue of every segment: TMEM [128 weight rows x 16 batch rows] -> red.global into out[b, t, n] ----
        {
          SegIter si(P, st);
          Seg s;
          while (si.next(P, s)) {
            const uint32_t as = acc_it & 1;
            mbar_wait(&sm.acc_full[as], (acc_it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&sm.acc_empty[as]);
            ++acc_it;
            const TcProblem& q = P.prob[s.prob];
            float* out = s.prob == 0 ? P.gx_acc : s.prob == 1 ? p.gh : s.prob == 2 ? p.p_pre : p.post_logits;
            const int n = s.tile * 128 + q4 * 32 + lane;
            if (n < q.n_out) {
              float* o = out + (size_t)t * q.n_out + n;
              const size_t ldo = (size_t)T * q.n_out;
#pragma unroll
              for (int b = 0; b < kRows; ++b)
                if (b < p.B) red_add(o + (size_t)b * ldo, __uint_as_float(r[b]));
            }
          }
        }
        grid_bar_workers(p.bar, bar_target, wt);
      }
      // ---- S4: sample z_t, feed the pre-GRU layer of step t+1 ----
      for (int task = gwarp; task < ntask; task += nwarps) sample_and_gather(P, task / p.Zc, task % p.Zc, t, lane);
      grid_bar_workers(p.bar, bar_target, wt);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

// accumulators pre-filled with their biases / hoisted parts: gx <- b0, gh <- b1, logits <- b_zq, x_pre <- actW
__global__ void scan_tc_init_kernel(ScanFwdParams p, float* gx_acc) {
  const long long N = (long long)p.B * p.T;
  const int G3 = 3 * p.G, D = p.D, Z = p.Z;
  const long long per_row = (long long)(2 * G3 + Z + D) / 4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < N * per_row; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / per_row;
    int c = (int)(e - row * per_row) * 4;
    if (c < G3) { *reinterpret_cast<float4*>(gx_acc + row * G3 + c) = __ldg(reinterpret_cast<const float4*>(p.gru_b + c)); continue; }
    c -= G3;
    if (c < G3) { *reinterpret_cast<float4*>(p.gh + row * G3 + c) = __ldg(reinterpret_cast<const float4*>(p.gru_b + G3 + c)); continue; }
    c -= G3;
    if (c < Z) { *reinterpret_cast<float4*>(p.post_logits + row * Z + c) = __ldg(reinterpret_cast<const float4*>(p.b_zq + c)); continue; }
    c -= Z;
    *reinterpret_cast<float4*>(p.x_pre + row * D + c) = __ldg(reinterpret_cast<const float4*>(p.actW + row * D + c));
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
// bf16 [rows, K] K-major (row stride ld elements), box 64 k x 128 rows, SWIZZLE_128B, zero fill out of bounds
bool make_wmap(CUtensorMap* map, const void* ptr, int rows, int K, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_kc(int kblocks) {
  for (int kc = kMaxKc; kc >= 1; --kc) if (kblocks % kc == 0) return kc;
  return 1;
}

}  // namespace

bool scan_tc_supported(const ScanFwdParams& p) {
  return p.D % 128 == 0 && p.G % 128 == 0 && p.Z % 128 == 0 && p.Zk == 32 && p.D >= 512 && p.D <= 1024 && p.G >= 512 && p.B <= kRows &&
         std::getenv("DV3_NO_SCAN_TC") == nullptr;
}

int scan_tc_max_ctas(int device) {
  int sms = 0, per_sm = 0;
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaFuncSetAttribute(observe_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, observe_scan_tc_kernel, kThreadsTc, smem);
  return per_sm > 0 ? sms : 0;
}

cudaError_t launch_observe_scan_tc(cudaStream_t s, const ScanFwdParams& p, const ScanTcWeights& w, float* gx_acc, int ctas) {
  static_assert(sizeof(TcSmem) + 1024 <= 227 * 1024, "TcSmem exceeds the dynamic shared memory limit");
  ScanTcParams P{};
  P.p = p;
  P.gx_acc = gx_acc;
  const int G = p.G, D = p.D, Z = p.Z;
  const int n_out[4] = {3 * G, 3 * G, D, Z}, K[4] = {D, G, G, D};
  const void* ptr[4] = {w.Kt, w.Ut, w.Wpost_h_t, w.Wzq_t};
  const int ld[4] = {w.ld_Kt, w.ld_Ut, w.ld_Wpost_h_t, w.ld_Wzq_t};
  TcMaps maps;
  for (int i = 0; i < 4; ++i) {
    TcProblem& q = P.prob[i];
    q.n_out = n_out[i]; q.K = K[i]; q.tiles = n_out[i] / 128; q.kblocks = K[i] / 64; q.kc = pick_kc(q.kblocks);
    q.span = q.tiles * q.kc; q.units = q.tiles * q.kblocks; q.unit_begin = 0;
    if (!make_wmap(&maps.m[i], ptr[i], n_out[i], K[i], ld[i])) return cudaErrorInvalidValue;
  }
  P.prob[1].unit_begin = P.prob[0].units;
  P.stage_units[0] = P.prob[0].units + P.prob[1].units;
  P.stage_units[1] = P.prob[2].units;
  P.stage_units[2] = P.prob[3].units;
  // a CTA's range of a stage may touch at most kSlots activation chunks
  for (int st = 0; st < 3; ++st) {
    const int per = (P.stage_units[st] + ctas - 1) / ctas;
    const int span = st == 0 ? (P.prob[0].span < P.prob[1].span ? P.prob[0].span : P.prob[1].span) : P.prob[st + 1].span;
    if (per > span) return cudaErrorInvalidConfiguration;
  }
  {
    const long long n4 = (long long)p.B * p.T * (6 * G + Z + D) / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    scan_tc_init_kernel<<<blocks, 256, 0, s>>>(p, gx_acc);
    count_launch();
  }
  cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), s);
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaFuncSetAttribute(observe_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void* args[] = {&maps, &P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)observe_scan_tc_kernel, dim3(ctas), dim3(kThreadsTc), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
