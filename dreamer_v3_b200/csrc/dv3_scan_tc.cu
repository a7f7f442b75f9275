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
#include <vector>

#include "dv3_device.cuh"
#include "dv3_scan.h"

namespace dv3 {
namespace {

constexpr int kWorkers = 256;        // warps 2..9
constexpr int kWorkerWarps = kWorkers / 32;
constexpr int kThreadsTc = 64 + kWorkers;
constexpr int kRowsPerWarp = 16 / kWorkerWarps;      // rows of the row-resident staging owned by one worker warp
constexpr int kRing = 7;             // weight k-blocks (128 rows x 64 k bf16 = 16 KB) in flight per CTA
constexpr int kMaxKc = 16;           // k-blocks per activation chunk (<= 1024 k)
constexpr int kSlots = 2;            // activation chunks resident per stage
constexpr int kRows = 16;
constexpr int kKbBytes = 2048;       // one staged activation k-block: 16 rows x 128 B
constexpr int kWBytes = 128 * 128;   // one weight k-block

struct TcProblem { int n_out, K, tiles, kblocks, kc, span, unit_begin, units; };

constexpr int kMaxProb = 5, kMaxStages = 4;
struct TcMaps { CUtensorMap m[kMaxProb]; };

// forward: problems 0 gx (K = D), 1 gh (K = G) | 2 posterior h-part (K = G) | 3 posterior logits (K = D)   + sampling stage
// backward: 0 dp_act (K = Z) | 1 dh (K = D) | 2 dh_in (K = 3G), 3 du (K = 3G) | 4 dz_in (K = D)
struct ScanTcParams {
  ScanFwdParams p;      // forward tensors (unused by the backward kernel)
  ScanBwdParams q;      // backward tensors (unused by the forward kernel)
  float* gx_acc;        // forward: [B, T, 3G], pre-filled with b0; p.gh is pre-filled with b1, p.post_logits with b_zq, p.x_pre with actW
  TcProblem prob[kMaxProb];
  int stage_units[kMaxStages], stage_p0[kMaxStages], stage_np[kMaxStages];
  int B, T;
  unsigned int* bar;
  unsigned long long* stamps;   // optional globaltimer stamps of CTA 0 after every grid barrier (profiling), may be null
  int evict_first;              // weights exceed the L2: stream them with the evict_first policy
};

constexpr int kMaxSeg = 12;              // segments of one CTA in one stage (checked by the launcher)
constexpr int kMaxItems = 2 * kMaxKc + 2;  // staging events of one CTA in one stage
constexpr int kMaxFlags = 4096;          // B * T is_first flags kept in shared memory

// The work of a CTA is the same at every timestep: its schedule is built once at kernel start (one thread) and read from
// shared memory by the three roles afterwards.
struct SegRec {                // one output tile x consecutive k-blocks of one k-chunk
  unsigned char prob, slot, kb0, kb1;
  unsigned short tile, kb_abs0;          // first absolute k-block (chunk * kc + kb0)
  unsigned char need[kMaxKc];            // need[kb - kb0]: staging events of this stage that must have completed before k-block kb
};
struct ItemRec {               // one staging event, in the order the MMA warp first touches the staged k-blocks
  unsigned char rows, prob, slot, kb;    // rows = 1: the whole K = D activation (kc k-blocks) from full rows; else one k-block
  unsigned short kb_abs, writer;         // writer: this CTA stores the fp32 values (rows: bit mask over the k-blocks)
};
struct StageSched { int nseg, nitem, count; SegRec seg[kMaxSeg]; ItemRec item[kMaxItems]; };

struct TcSmem {
  unsigned char ring[kRing][kWBytes];
  unsigned char bbuf[kSlots][kMaxKc][kKbBytes];
  float ln[4][1024];                 // pre_g, pre_b, post_g, post_b (the grid barrier invalidates L1: keep them on chip)
  unsigned char first[kMaxFlags];    // is_first[b * T + t] != 0
  StageSched sched[kMaxStages];
  uint64_t full[kRing], empty[kRing], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t staged_seq;      // running count of staging events completed by the worker warps (read by the MMA warp)
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
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); }
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
    atomicAdd(ctr, 1u);   // (measured: a single red.release.gpu arrive is 1 us / step slower)
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
  __host__ __device__ SegIter(const ScanTcParams& P, int st, int blk, int nblk) : stage(st) {
    const long long U = P.stage_units[st];
    u = (int)(U * blk / nblk);
    u_end = (int)(U * (blk + 1) / nblk);
  }
  __host__ __device__ bool next(const ScanTcParams& P, Seg& s) {
    if (u >= u_end) return false;
    int pr = P.stage_p0[stage];
    if (P.stage_np[stage] > 1 && u >= P.prob[pr + 1].unit_begin) ++pr;
    const TcProblem& q = P.prob[pr];
    const int local = u - q.unit_begin;
    const int chunk = local / q.span, r = local - chunk * q.span;
    const int tile = r / q.kc, kbi = r - tile * q.kc;
    int n = q.kc - kbi;
    const int lim = (u_end < q.unit_begin + q.units ? u_end : q.unit_begin + q.units) - u;
    if (lim < n) n = lim;
    s.prob = pr; s.tile = tile; s.chunk = chunk; s.kb0 = kbi; s.kb1 = kbi + n;
    u += n;
    return true;
  }
};
// activation-chunk slot of a segment: chunks are numbered in order of first use within the stage (<= kSlots per stage, checked
// by the launcher); identical in the MMA and the worker warps
struct SlotMap {
  int key[kSlots], n;
  __host__ __device__ SlotMap() : n(0) { key[0] = key[1] = -1; }
  __host__ __device__ int get(const Seg& s, bool& fresh) {
    const int k = s.prob * 4096 + s.chunk;
    fresh = false;
    for (int i = 0; i < n; ++i) if (key[i] == k) return i;
    fresh = true;
    key[n & (kSlots - 1)] = k;
    return (n++) & (kSlots - 1);
  }
};

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
// writes 8 staged values (row b, 16-byte chunk c8 of a 64-k block) as bf16 into the K-major SWIZZLE_128B B-operand block
__device__ __forceinline__ void put_b(unsigned char* dst, int b, int c8, const float (&v)[8]) {
  const uint32_t a = smem_u32(dst) + (uint32_t)((b >> 3) * 1024 + (b & 7) * 128 + ((c8 ^ (b & 7)) << 4));
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
               "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])));
}
// row-resident staging: 4 values of row b at columns c .. c+3 (c % 4 == 0) of a K = D activation whose k-blocks lie back to
// back from `chunk` (bbuf[slot][0])
__device__ __forceinline__ void put_b4(unsigned char* chunk, int b, int c, float v0, float v1, float v2, float v3) {
  const int kb = c >> 6, c8 = (c & 63) >> 3, half = (c >> 2) & 1;
  const uint32_t a = smem_u32(chunk) + (uint32_t)(kb * kKbBytes + (b >> 3) * 1024 + (b & 7) * 128 + ((c8 ^ (b & 7)) << 4) + half * 8);
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(pack_bf16(v0, v1)), "r"(pack_bf16(v2, v3)));
}


// =====================================================================================================================
// Forward policy (world_model.py:244-273)
// =====================================================================================================================
struct FwdPolicy {
  static constexpr int kGemmStages = 3;
  static constexpr bool kExtraStage = true;
  __device__ static int time_of(const ScanTcParams& P, int i) { return i; }
  __device__ static const float* ln_vec(const ScanTcParams& P, int i) { return i == 0 ? P.p.pre_g : i == 1 ? P.p.pre_b : i == 2 ? P.p.post_g : P.p.post_b; }
  __device__ static const float* first_flags(const ScanTcParams& P) { return P.p.is_first; }
  __device__ static int width_D(const ScanTcParams& P) { return P.p.D; }
  // K = D problems: the CTA reads the 16 full rows once (LayerNorm statistics need them anyway) and stages every k-block of the
  // activation straight from those registers
  __host__ __device__ static bool row_resident(int prob) { return prob == 0 || prob == 3; }
  __device__ static int batch_of(int prob) { return prob == 1 ? 8 : 4; }   // k-blocks per staging event (two thread groups)
  __device__ static bool has_pre_stage(int) { return false; }
  __device__ static void pre_stage(const ScanTcParams&, TcSmem&, int, int, int) {}

  // SiLU(LayerNorm(x[b, t, :])) of the pre-GRU (prob 0) / posterior (prob 3) layer for all 16 rows: worker warp ww owns rows
  // 4ww .. 4ww+3, a lane holds D / 128 float4 per row. writer_mask: k-blocks whose fp32 values this CTA stores for the backward
  // pass.
  template <int EPL>
  __device__ static void stage_rows_t(const ScanTcParams& P, TcSmem& sm, unsigned char* chunk, int prob, int t, unsigned int writer_mask, int ww, int lane) {
    const ScanFwdParams& p = P.p;
    const int T = p.T, B = p.B, D = p.D;
    const float* x = prob == 0 ? p.x_pre : p.p_pre;
    const float* gamma = sm.ln[prob == 0 ? 0 : 2];   // shared memory copies
    const float* beta = sm.ln[prob == 0 ? 1 : 3];
    float4 v[kRowsPerWarp][EPL];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int b = ww * kRowsPerWarp + r;
#pragma unroll
      for (int i = 0; i < EPL; ++i)
        v[r][i] = b < B ? ldcg4(x + ((size_t)b * T + t) * D + (size_t)(lane + 32 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s[kRowsPerWarp], q[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      s[r] = 0.f;
#pragma unroll
      for (int i = 0; i < EPL; ++i) s[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
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
      for (int r = 0; r < kRowsPerWarp; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) q[r] = rsqrtf(q[r] / (float)D + DV3_LN_EPS);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(gamma + c), b4 = *reinterpret_cast<const float4*>(beta + c);
      const bool wr = (writer_mask >> (c >> 6)) & 1u;
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int b = ww * kRowsPerWarp + r;
        float4 o;
        o.x = (v[r][i].x - s[r]) * q[r] * g4.x + b4.x; o.y = (v[r][i].y - s[r]) * q[r] * g4.y + b4.y;
        o.z = (v[r][i].z - s[r]) * q[r] * g4.z + b4.z; o.w = (v[r][i].w - s[r]) * q[r] * g4.w + b4.w;
        o.x *= sigmoid_fast(o.x); o.y *= sigmoid_fast(o.y); o.z *= sigmoid_fast(o.z); o.w *= sigmoid_fast(o.w);
        if (b >= B) o = make_float4(0.f, 0.f, 0.f, 0.f);
        put_b4(chunk, b, c, o.x, o.y, o.z, o.w);
        if (wr && b < B) *reinterpret_cast<float4*>((prob == 0 ? p.u_act : p.p_act) + ((size_t)b * T + t) * D + c) = o;
      }
    }
    if ((writer_mask & 1u) && lane == 0) {
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int b = ww * kRowsPerWarp + r;
        if (b < B) {
          (prob == 0 ? p.x_mean : p.p_mean)[(size_t)b * T + t] = s[r];
          (prob == 0 ? p.x_rstd : p.p_rstd)[(size_t)b * T + t] = q[r];
        }
      }
    }
  }
  __device__ static void stage_rows(const ScanTcParams& P, TcSmem& sm, unsigned char* chunk, int prob, int t, unsigned int writer_mask, int ww, int lane) {
    switch (P.p.D / 128) {
      case 4: stage_rows_t<4>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      case 5: stage_rows_t<5>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      case 6: stage_rows_t<6>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      default: stage_rows_t<8>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
    }
  }

  // Stages NB k-blocks (64 k each) of problem 1 (masked recurrent input h_in, world_model.py:246-247) or 2 (h_t = GRU cell,
  // Keras reset_after; gx / gh already hold their biases) as the bf16 K-major SWIZZLE_128B B operand. Worker thread wt: row =
  // wt / 8, 16-byte chunk (8 k) = wt % 8. Every load of the batch is issued before the first use (one L2 round trip per batch).
  // `writer`: this CTA owns tile 0 of the k-block and stores the fp32 values for the backward pass.
  template <int NB>
  __device__ static void stage_batch(const ScanTcParams& P, TcSmem& sm, const ItemRec* it, int n, int t, int wt) {
    const ScanFwdParams& p = P.p;
    const int grp = wt >> 7, b = (wt & 127) >> 3, c8 = wt & 7;   // two groups of 128 threads take alternate k-blocks of a batch
    const int T = p.T, G = p.G, W = G + p.Z;
    const bool row_ok = b < p.B;
    const size_t bt = (size_t)(row_ok ? b : 0) * T + t;
    const bool first = (t == 0) || (sm.first[bt] != 0);
    const int prob = it[0].prob;
    it += grp; n = (n - grp + 1) / 2;   // this group's items: grp, grp + 2, ...
    if (prob == 1) {
      float v[NB][8];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          if (first) load8_ro(v[j], p.h0 + k);
          else load8(v[j], p.hz + (bt - 1) * W + k);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          if (!row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
          } else if (it[2 * j].writer) store8(p.h_in + bt * G + k, v[j]);
          put_b(sm.bbuf[it[2 * j].slot][it[2 * j].kb], b, c8, v[j]);
        }
      }
    } else {
      float xz[NB][8], xr[NB][8], xc[NB][8], hz_[NB][8], hr[NB][8], hc[NB][8], hin[NB][8];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          const float* gx = P.gx_acc + bt * 3 * G + k;
          const float* gh = p.gh + bt * 3 * G + k;
          load8(xz[j], gx); load8(xr[j], gx + G); load8(xc[j], gx + 2 * G);
          load8(hz_[j], gh); load8(hr[j], gh + G); load8(hc[j], gh + 2 * G);
          if (first) load8_ro(hin[j], p.h0 + k);
          else load8(hin[j], p.hz + (bt - 1) * W + k);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          float z[8], r[8], c[8], v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            z[i] = sigmoid_fast(xz[j][i] + hz_[j][i]);
            r[i] = sigmoid_fast(xr[j][i] + hr[j][i]);
            c[i] = tanh_fast(xc[j][i] + r[i] * hc[j][i]);
            v[i] = row_ok ? z[i] * hin[j][i] + (1.f - z[i]) * c[i] : 0.f;
          }
          if (it[2 * j].writer && row_ok) {
            float* gt = p.gates + bt * 3 * G + k;
            store8(gt, z); store8(gt + G, r); store8(gt + 2 * G, c);
            store8(p.hz + bt * W + k, v);
          }
          put_b(sm.bbuf[it[2 * j].slot][it[2 * j].kb], b, c8, v);
        }
      }
    }
  }
  __device__ static void stage_items(const ScanTcParams& P, TcSmem& sm, const ItemRec* it, int n, int t, int wt) {
    if (it[0].prob == 1) stage_batch<4>(P, sm, it, n, t, wt);
    else stage_batch<2>(P, sm, it, n, t, wt);
  }

  // accumulator column n (= weight row) of problem `prob`: r[j] for batch rows b0 + j -> out[b, t, n]
  __device__ static void epilogue(const ScanTcParams& P, TcSmem&, int prob, int t, int n, int b0, const uint32_t (&r)[8]) {
    const ScanFwdParams& p = P.p;
    const TcProblem& q = P.prob[prob];
    float* out = prob == 0 ? P.gx_acc : prob == 1 ? p.gh : prob == 2 ? p.p_pre : p.post_logits;
    float* o = out + (size_t)t * q.n_out + n;
    const size_t ldo = (size_t)p.T * q.n_out;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (b0 + j < p.B) red_add(o + (size_t)(b0 + j) * ldo, __uint_as_float(r[j]));
  }

  // one (row, categorical) of stage S4 per warp; t = -1: the prologue (only the gather of step 0)
  __device__ static void sample_and_gather(const ScanTcParams& P, TcSmem& sm, int b, int cat, int t, int lane) {
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
      const bool first = (t + 1 == 0) || (sm.first[bn] != 0);
      const int idx = first ? p.z0_idx[cat] : pick;
      if (lane < Zk) p.z_in[bn * Z + cat * Zk + lane] = (lane == idx) ? 1.f : 0.f;
      const float* wrow = p.W_pre + (size_t)(cat * Zk + idx) * D;
      float* xrow = p.x_pre + bn * D;
      for (int j = lane * 4; j < D; j += 128) red_add_v4(xrow + j, __ldg(reinterpret_cast<const float4*>(wrow + j)));
    }
  }
  __device__ static void extra(const ScanTcParams& P, TcSmem& sm, int t, int gwarp, int nwarps, int lane) {
    const int ntask = P.B * P.p.Zc;
    for (int task = gwarp; task < ntask; task += nwarps) sample_and_gather(P, sm, task / P.p.Zc, task % P.p.Zc, t, lane);
  }
  __device__ static void prologue(const ScanTcParams& P, TcSmem& sm, int gwarp, int nwarps, int lane) { extra(P, sm, -1, gwarp, nwarps, lane); }
};

// =====================================================================================================================
// Backward policy: reverse-time BPTT of the same scan (input gradients only; every weight gradient is one N-row contraction
// after the kernel). Per step t = T-1 ... 0, 4 stages:
//   B1  dp_act += dlogits . W_zq^T        dlogits = softmax-backward(0.99 (dz_t + dKL/dpost)), formed by the consumer
//   B2  dh_t   += dp_pre . W_post[E:]^T   dp_pre = LayerNorm+SiLU backward of dp_act (row sums recomputed per CTA)
//   B3a (dgx, dgh) = GRU-cell backward of dh_t, each element once over the grid (elementwise stage); dh_in = dh_t * z
//   B3  dh_in  += dgh . U^T,  du += dgx . K^T
//   B4  dz_in  = dx_pre . W_pre[:Z]^T     dx_pre = LayerNorm+SiLU backward of du
// dz_in is passed to step t-1 in the epilogue of B4 (d_hz[t-1, G:] += (1 - is_first[t]) dz_in), dh_in is picked up by B3a of
// step t-1.
// Weight operands: the plain bf16 shadows ([out, in] row-major = K-major for these transposed contractions).
// =====================================================================================================================
struct BwdPolicy {
  static constexpr int kGemmStages = 4;
  static constexpr bool kExtraStage = false;
  __device__ static int time_of(const ScanTcParams& P, int i) { return P.T - 1 - i; }
  __device__ static const float* ln_vec(const ScanTcParams& P, int i) { return i == 0 ? P.q.pre_g : i == 1 ? P.q.pre_b : i == 2 ? P.q.post_g : P.q.post_b; }
  __device__ static const float* first_flags(const ScanTcParams& P) { return P.q.is_first; }
  __device__ static int width_D(const ScanTcParams& P) { return P.q.D; }
  __host__ __device__ static bool row_resident(int prob) { return prob == 1 || prob == 4; }

  // LayerNorm + SiLU backward of the posterior (prob 1: dp_act -> dp_pre) / pre-GRU (prob 4: du -> dx_pre) layer for all 16
  // rows, two rows per warp and round: dn = dy silu'(n) g, dx = rstd (dn - mean(dn) - xhat mean(dn xhat))
  template <int EPL>
  __device__ static void stage_rows_t(const ScanTcParams& P, TcSmem& sm, unsigned char* chunk, int prob, int t, unsigned int writer_mask, int ww, int lane) {
    const ScanBwdParams& q = P.q;
    const int T = P.T, B = P.B, D = q.D;
    const float* dy = prob == 1 ? q.dp_act : q.du;
    const float* x = prob == 1 ? q.p_pre : q.x_pre;
    const float* mean_in = prob == 1 ? q.p_mean : q.x_mean;
    const float* rstd_in = prob == 1 ? q.p_rstd : q.x_rstd;
    const float* gamma = sm.ln[prob == 1 ? 2 : 0];   // shared memory copies
    const float* beta = sm.ln[prob == 1 ? 3 : 1];
    float* out = prob == 1 ? q.dp_pre : q.dx_pre;
    {
      float4 xv[kRowsPerWarp][EPL], dv[kRowsPerWarp][EPL];
      float mean[kRowsPerWarp], rstd[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int b = ww * kRowsPerWarp + r;
        const size_t bt = (size_t)(b < B ? b : 0) * T + t;
        mean[r] = mean_in[bt]; rstd[r] = rstd_in[bt];
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
          const size_t off = bt * D + (size_t)(lane + 32 * i) * 4;
          xv[r][i] = ldcg4(x + off);
          dv[r][i] = b < B ? ldcg4(dy + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float s1[kRowsPerWarp], s2[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) { s1[r] = 0.f; s2[r] = 0.f; }
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const float4 g4 = *(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        const float4 b4 = *(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
          float xs[4] = {xv[r][i].x, xv[r][i].y, xv[r][i].z, xv[r][i].w};
          float ds[4] = {dv[r][i].x, dv[r][i].y, dv[r][i].z, dv[r][i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xhat = (xs[j] - mean[r]) * rstd[r];
            const float n = xhat * g[j] + be[j];
            const float sg = sigmoid_fast(n);
            const float dng = ds[j] * (sg * (1.f + n * (1.f - sg))) * g[j];
            s1[r] += dng; s2[r] += dng * xhat;
            xs[j] = xhat; ds[j] = dng;   // kept for the second pass
          }
          xv[r][i] = make_float4(xs[0], xs[1], xs[2], xs[3]);
          dv[r][i] = make_float4(ds[0], ds[1], ds[2], ds[3]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) { s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o); s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o); }
      }
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        const int c = (lane + 32 * i) * 4;
        const bool wr = (writer_mask >> (c >> 6)) & 1u;
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
          const int b = ww * kRowsPerWarp + r;
          const float m1 = s1[r] / (float)D, m2 = s2[r] / (float)D;
          float4 o;
          o.x = rstd[r] * (dv[r][i].x - m1 - xv[r][i].x * m2); o.y = rstd[r] * (dv[r][i].y - m1 - xv[r][i].y * m2);
          o.z = rstd[r] * (dv[r][i].z - m1 - xv[r][i].z * m2); o.w = rstd[r] * (dv[r][i].w - m1 - xv[r][i].w * m2);
          if (b >= B) o = make_float4(0.f, 0.f, 0.f, 0.f);
          put_b4(chunk, b, c, o.x, o.y, o.z, o.w);
          if (wr && b < B) *reinterpret_cast<float4*>(out + ((size_t)b * T + t) * D + c) = o;
        }
      }
    }
  }
  __device__ static void stage_rows(const ScanTcParams& P, TcSmem& sm, unsigned char* chunk, int prob, int t, unsigned int writer_mask, int ww, int lane) {
    switch (P.q.D / 128) {
      case 4: stage_rows_t<4>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      case 5: stage_rows_t<5>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      case 6: stage_rows_t<6>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
      default: stage_rows_t<8>(P, sm, chunk, prob, t, writer_mask, ww, lane); break;
    }
  }

  template <int NB>
  __device__ static void stage_batch(const ScanTcParams& P, TcSmem& sm, const ItemRec* it, int n, int t, int wt) {
    const ScanBwdParams& q = P.q;
    const int grp = wt >> 7, b = (wt & 127) >> 3, c8 = wt & 7;   // two groups of 128 threads take alternate k-blocks of a batch
    const int T = P.T, G = q.G, Z = q.Z, W = G + Z;
    const bool row_ok = b < P.B;
    const size_t bt = (size_t)(row_ok ? b : 0) * T + t;
    const int prob = it[0].prob;
    it += grp; n = (n - grp + 1) / 2;   // this group's items: grp, grp + 2, ...
    if (prob == 0) {
      // dlogits of the posterior categorical (representation_layer.py:73-113 backward): 4 neighbouring threads hold the 32
      // classes of one categorical (Zk = 32); straight-through dz + KL dprobs, through the 1 % unimix and the softmax
      float l[NB][8], dz[NB][8], dk[NB][8];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          load8(l[j], q.post_logits + bt * Z + k);
          load8(dz[j], q.d_hz + bt * W + G + k);
          load8(dk[j], q.dpost + bt * Z + k);
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {   // (n is uniform over the CTA: every lane takes part in the shuffles)
          const int k = it[2 * j].kb_abs * 64 + c8 * 8;
          float g[8], e[8], v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = 0.99f * (dz[j][i] + dk[j][i]);
          float mx = l[j][0];
#pragma unroll
          for (int i = 1; i < 8; ++i) mx = fmaxf(mx, l[j][i]);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { e[i] = expf(l[j][i] - mx); s += e[i]; }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          const float inv = 1.f / s;
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { e[i] *= inv; dot += g[i] * e[i]; }
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = row_ok ? e[i] * (g[i] - dot) : 0.f;
          if (it[2 * j].writer && row_ok) store8(q.dlogits + bt * Z + k, v);
          put_b(sm.bbuf[it[2 * j].slot][it[2 * j].kb], b, c8, v);
        }
      }
    } else {
      // dgh (prob 2) / dgx (prob 3), formed once per step by the elementwise stage B3a: plain fp32 -> bf16 staging
      const float* src = (prob == 2 ? q.dgh : q.dgx) + bt * 3 * G + c8 * 8;
      float v[NB][8];
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j < n) load8(v[j], src + (size_t)it[2 * j].kb_abs * 64);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < n) {
          if (!row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
          }
          put_b(sm.bbuf[it[2 * j].slot][it[2 * j].kb], b, c8, v[j]);
        }
      }
    }
  }
  // B3a: GRU-cell backward (Keras reset_after) of step t for all rows, each (row, unit) exactly once over the grid:
  //   dh = d_hz[t] (heads + posterior) + (1 - is_first[t+1]) dh_in[t+1];   dgx = (dz, dr, dc), dgh = (dz, dr, dc r), dh_in[t] = dh z
  // (the contraction dgh . U^T of stage B3 is then added to dh_in[t] with red.global)
  __device__ static bool has_pre_stage(int st) { return st == 2; }
  __device__ static void pre_stage(const ScanTcParams& P, TcSmem& sm, int t, int gthread, int nthreads) {
    const ScanBwdParams& q = P.q;
    const int T = P.T, G = q.G, W = G + q.Z, G4 = G / 4;
    for (int e = gthread; e < P.B * G4; e += nthreads) {
      const int b = e / G4, u = (e - b * G4) * 4;
      const size_t bt = (size_t)b * T + t;
      float4 d = ldcg4(q.d_hz + bt * W + u);
      const bool carry = t + 1 < T && sm.first[bt + 1] == 0;
      const float* gt = q.gates + bt * 3 * G + u;
      const float4 z = ldcg4(gt), r = ldcg4(gt + G), c = ldcg4(gt + 2 * G);
      const float4 hh = ldcg4(q.gh + bt * 3 * G + 2 * G + u), hp = ldcg4(q.h_in + bt * G + u);
      if (carry) {
        const float4 nx = ldcg4(q.dh_in + (bt + 1) * G + u);
        d.x += nx.x; d.y += nx.y; d.z += nx.z; d.w += nx.w;
      }
      float4 dz, dr, dc, dcr, dhz;
#define DV3_GRU_BWD(f) \
      dz.f = d.f * (hp.f - c.f) * z.f * (1.f - z.f); dc.f = d.f * (1.f - z.f) * (1.f - c.f * c.f); \
      dr.f = dc.f * hh.f * r.f * (1.f - r.f); dcr.f = dc.f * r.f; dhz.f = d.f * z.f;
      DV3_GRU_BWD(x) DV3_GRU_BWD(y) DV3_GRU_BWD(z) DV3_GRU_BWD(w)
#undef DV3_GRU_BWD
      float* gx = q.dgx + bt * 3 * G + u;
      float* gh = q.dgh + bt * 3 * G + u;
      *reinterpret_cast<float4*>(gx) = dz; *reinterpret_cast<float4*>(gx + G) = dr; *reinterpret_cast<float4*>(gx + 2 * G) = dc;
      *reinterpret_cast<float4*>(gh) = dz; *reinterpret_cast<float4*>(gh + G) = dr; *reinterpret_cast<float4*>(gh + 2 * G) = dcr;
      *reinterpret_cast<float4*>(q.dh_in + bt * G + u) = dhz;
      *reinterpret_cast<float4*>(q.d_hz + bt * W + u) = d;   // total gradient w.r.t. h_t
    }
  }
  __device__ static int batch_of(int prob) { return prob == 0 ? 4 : 16; }   // k-blocks per staging event (two thread groups)
  __device__ static void stage_items(const ScanTcParams& P, TcSmem& sm, const ItemRec* it, int n, int t, int wt) {
    if (it[0].prob == 0) stage_batch<2>(P, sm, it, n, t, wt);
    else stage_batch<8>(P, sm, it, n, t, wt);
  }

  __device__ static void epilogue(const ScanTcParams& P, TcSmem& sm, int prob, int t, int n, int b0, const uint32_t (&r)[8]) {
    const ScanBwdParams& q = P.q;
    const int T = P.T, G = q.G, D = q.D, Z = q.Z, W = G + Z;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j;
      if (b >= P.B) continue;
      const size_t bt = (size_t)b * T + t;
      const float val = __uint_as_float(r[j]);
      if (prob == 0) red_add(q.dp_act + bt * D + n, val);
      else if (prob == 1) red_add(q.d_hz + bt * W + n, val);
      else if (prob == 3) red_add(q.du + bt * D + n, val);
      else if (prob == 2) red_add(q.dh_in + bt * G + n, val);
      else if (t > 0 && sm.first[bt] == 0) red_add(q.d_hz + (bt - 1) * W + G + n, val);
    }
  }
  __device__ static void extra(const ScanTcParams&, TcSmem&, int, int, int, int) {}
  __device__ static void prologue(const ScanTcParams&, TcSmem&, int, int, int) {}
};


// Builds the CTA's schedule of one stage (run once, by one thread): segments in unit order, the activation-chunk slot of each,
// the staging events in the order the MMA warp first touches what they stage, and per k-block the number of events it needs.
// Returns 0, or -1 / -2 when the segment / staging table would overflow (the launcher's bounds make that impossible; the host-side
// check dv3_scan_tc_schedule_check walks every CTA of every size through this very function).
template <class Pol>
__host__ __device__ int build_stage_sched(const ScanTcParams& P, int st, StageSched& sc, int blk, int nblk) {
  SegIter si(P, st, blk, nblk);
  SlotMap slots;
  Seg s;
  unsigned int staged[kSlots] = {0u, 0u};
  int rows_item[kSlots] = {-1, -1};
  sc.nseg = 0; sc.nitem = 0; sc.count = 0;
  while (si.next(P, s)) {
    bool fresh;
    const int sl = slots.get(s, fresh);
    if (fresh) { staged[sl] = 0u; rows_item[sl] = -1; }
    const TcProblem& q = P.prob[s.prob];
    if (sc.nseg >= kMaxSeg) return -1;
    SegRec& r = sc.seg[sc.nseg++];
    r.prob = (unsigned char)s.prob; r.slot = (unsigned char)sl; r.kb0 = (unsigned char)s.kb0; r.kb1 = (unsigned char)s.kb1;
    r.tile = (unsigned short)s.tile; r.kb_abs0 = (unsigned short)(s.chunk * q.kc + s.kb0);
    for (int kb = s.kb0; kb < s.kb1; ++kb) {
      if (!((staged[sl] >> kb) & 1u)) {
        if (sc.nitem >= kMaxItems) return -2;
        ItemRec& it = sc.item[sc.nitem];
        if (Pol::row_resident(s.prob)) {
          it.rows = 1; it.prob = (unsigned char)s.prob; it.slot = (unsigned char)sl; it.kb = 0; it.kb_abs = 0; it.writer = 0;
          rows_item[sl] = sc.nitem;
          staged[sl] = 0xFFFFFFFFu;
        } else {
          it.rows = 0; it.prob = (unsigned char)s.prob; it.slot = (unsigned char)sl; it.kb = (unsigned char)kb;
          it.kb_abs = (unsigned short)(s.chunk * q.kc + kb); it.writer = s.tile == 0 ? 1 : 0;
          staged[sl] |= 1u << kb;
        }
        ++sc.nitem;
        sc.count = sc.nitem;
      }
      r.need[kb - s.kb0] = (unsigned char)sc.nitem;   // every event up to the one that staged this k-block
    }
    if (Pol::row_resident(s.prob) && s.tile == 0 && rows_item[sl] >= 0)
      sc.item[rows_item[sl]].writer |= (unsigned short)(((1u << (s.kb1 - s.kb0)) - 1u) << s.kb0);
  }
  return slots.n > kSlots ? -3 : 0;   // (-3: a CTA's range of this stage touched more activation chunks than there are slots)
}

template <class Pol>
__global__ void __launch_bounds__(kThreadsTc, 1) scan_tc_kernel(const __grid_constant__ TcMaps maps, const ScanTcParams P) {
  extern __shared__ unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = P.T;

  if (tid == 0) {
    for (int i = 0; i < kMaxProb; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
    for (int s = 0; s < kRing; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], kWorkers); }
    sm.staged_seq = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid >= 64 && tid < 64 + Pol::kGemmStages && build_stage_sched<Pol>(P, tid - 64, sm.sched[tid - 64], blockIdx.x, gridDim.x) != 0) {
    printf("dv3 scan_tc: schedule table overflow (block %d stage %d)\n", blockIdx.x, tid - 64);
    __trap();
  }
  {   // LayerNorm gains / offsets and the is_first flags: read thousands of times, and the grid barrier invalidates L1
    const int D = Pol::width_D(P);
    for (int i = tid; i < 4 * D; i += kThreadsTc) sm.ln[i / D][i % D] = Pol::ln_vec(P, i / D)[i % D];
    const float* f = Pol::first_flags(P);
    for (int i = tid; i < P.B * T; i += kThreadsTc) sm.first[i] = f[i] != 0.f ? 1 : 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: the whole weight stream of all T steps, throttled only by the ring ----------------
    if (lane == 0) {
      uint64_t policy = 0;
      if (P.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      uint32_t it = 0;
      for (int i = 0; i < T; ++i)
        for (int st = 0; st < Pol::kGemmStages; ++st) {
          const StageSched& sc = sm.sched[st];
          for (int g = 0; g < sc.nseg; ++g) {
            const SegRec& r = sc.seg[g];
            const int n = r.kb1 - r.kb0;
            for (int j = 0; j < n; ++j, ++it) {
              const int slot = it % kRing;
              mbar_wait(&sm.empty[slot], ((it / kRing) & 1) ^ 1);
              mbar_expect_tx(&sm.full[slot], kWBytes);
              // weights larger than the L2 are streamed with evict_first so that they do not push the step's activations
              // (a few hundred KB, re-read by every CTA) out of the L2
              if (P.evict_first) tma_load_2d_hint(smem_u32(sm.ring[slot]), &maps.m[r.prob], (r.kb_abs0 + j) * 64, r.tile * 128, &sm.full[slot], policy);
              else tma_load_2d(smem_u32(sm.ring[slot]), &maps.m[r.prob], (r.kb_abs0 + j) * 64, r.tile * 128, &sm.full[slot]);
            }
          }
        }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc16();
      uint32_t it = 0, acc_it = 0, base_seq = 0, have = 0;
      for (int i = 0; i < T; ++i)
        for (int st = 0; st < Pol::kGemmStages; ++st) {
          const StageSched& sc = sm.sched[st];
          for (int g = 0; g < sc.nseg; ++g) {
            const SegRec& r = sc.seg[g];
            const uint32_t as = acc_it & 1;
            mbar_wait(&sm.acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + as * 16;
            for (int kb = r.kb0; kb < r.kb1; ++kb, ++it) {
              // the activation k-block must have been staged: the worker warps publish a running count of staging events
              const uint32_t need = base_seq + r.need[kb - r.kb0];
              if ((int32_t)(have - need) < 0) {
                uint32_t spins = 0;
                do {
                  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(&sm.staged_seq)) : "memory");
                  if (++spins > (1u << 26)) { printf("dv3 scan_tc: staging wait timed out (block %d)\n", blockIdx.x); __trap(); }
                } while ((int32_t)(have - need) < 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              }
              const int slot = it % kRing;
              mbar_wait(&sm.full[slot], (it / kRing) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_addr = smem_u32(sm.ring[slot]), b_addr = smem_u32(sm.bbuf[r.slot][kb]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t da = make_desc(a_addr + j * 32), db = make_desc(b_addr + j * 32);
                const uint32_t accum = (kb == r.kb0 && j == 0) ? 0u : 1u;
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
          base_seq += (uint32_t)sc.count;
        }
    }
  } else {
    // ---------------- workers: stage activations, drain accumulators, sample ----------------
    const int wt = tid - 64, ww = warp - 2, q4 = warp & 3;
    unsigned int bar_target = 0;
    uint32_t acc_it = 0, seq = 0;
    const int gwarp = blockIdx.x * kWorkerWarps + ww, nwarps = gridDim.x * kWorkerWarps;
    int n_stamp = 0;
    auto stamp = [&]() {
      if (P.stamps != nullptr && blockIdx.x == 0 && wt == 0) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
        P.stamps[n_stamp] = v;
      }
      ++n_stamp;
    };
    // optional per-CTA cycle totals (P.stamps + 2048 + 8 * blockIdx.x): 1 row-resident staging, 2 batched staging, 3 publish,
    // 4 accumulator wait, 5 epilogue reds, 6 grid barrier, 7 extra stage
    long long prof1 = 0, prof2 = 0, prof3 = 0, prof4 = 0, prof5 = 0, prof6 = 0, prof7 = 0;
    long long tc0 = clock64();
    const bool profiling = P.stamps != nullptr;
    auto lap = [&](long long& acc) { if (profiling) { const long long c = clock64(); acc += c - tc0; tc0 = c; } };
    // staged k-blocks become visible to the tensor core (generic -> async proxy), then their running count to the MMA warp
    auto publish = [&](uint32_t count) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      worker_sync();
      seq += count;
      if (wt == 0) asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&sm.staged_seq)), "r"(seq) : "memory");
    };
    Pol::prologue(P, sm, gwarp, nwarps, lane);
    grid_bar_workers(P.bar, bar_target, wt);
    stamp();
    tc0 = clock64();

    for (int i = 0; i < T; ++i) {
      const int t = Pol::time_of(P, i);
      for (int st = 0; st < Pol::kGemmStages; ++st) {
        const StageSched& sc = sm.sched[st];
        if (Pol::has_pre_stage(st)) {   // an elementwise stage of its own (every element once over the grid), then a barrier
          Pol::pre_stage(P, sm, t, blockIdx.x * kWorkers + wt, gridDim.x * kWorkers);
          lap(prof7);
          stamp();
          grid_bar_workers(P.bar, bar_target, wt);
          lap(prof6);
          stamp();
        }
        // ---- stage the B operand of every segment of this CTA, in the order the MMA warp consumes it ----
        for (int done = 0; done < sc.nitem;) {
          const ItemRec& it0 = sc.item[done];
          if (it0.rows) {
            Pol::stage_rows(P, sm, sm.bbuf[it0.slot][0], it0.prob, t, it0.writer, ww, lane);
            lap(prof1);
            publish(1u);
            lap(prof3);
            ++done;
          } else {
            const int nb = Pol::batch_of(it0.prob);
            int n = 1;
            while (n < nb && done + n < sc.nitem && !sc.item[done + n].rows && sc.item[done + n].prob == it0.prob) ++n;
            Pol::stage_items(P, sm, &sc.item[done], n, t, wt);
            lap(prof2);
            publish((uint32_t)n);
            lap(prof3);
            done += n;
          }
        }
        stamp();   // operands staged
        // ---- epilogue of every segment: TMEM [128 weight rows x 16 batch rows] -> red.global ----
        for (int g = 0; g < sc.nseg; ++g) {
          const SegRec& r_ = sc.seg[g];
          const uint32_t as = acc_it & 1;
          mbar_wait(&sm.acc_full[as], (acc_it >> 1) & 1);
          lap(prof4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t r[8];
          const int b0 = (ww >> 2) * 8;   // two warps per TMEM lane quarter: accumulator columns (batch rows) b0 .. b0 + 7
          const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 16 + b0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&sm.acc_empty[as]);
          ++acc_it;
          const int n = (int)r_.tile * 128 + q4 * 32 + lane;
          if (n < P.prob[r_.prob].n_out) Pol::epilogue(P, sm, r_.prob, t, n, b0, r);
          lap(prof5);
        }
        stamp();   // accumulators drained
        grid_bar_workers(P.bar, bar_target, wt);
        lap(prof6);
        stamp();   // stage complete on every CTA
      }
      if (Pol::kExtraStage) {
        Pol::extra(P, sm, t, gwarp, nwarps, lane);
        lap(prof7);
        stamp();
        grid_bar_workers(P.bar, bar_target, wt);
        lap(prof6);
        stamp();
      }
    }
    if (profiling && wt == 0) {
      unsigned long long* o = P.stamps + 2048 + 8 * blockIdx.x;
      o[0] = 0; o[1] = prof1; o[2] = prof2; o[3] = prof3; o[4] = prof4; o[5] = prof5; o[6] = prof6; o[7] = prof7;
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

// problem table + stage table + tensor maps; false when a CTA's range of some stage could touch more than kSlots chunks
bool build_tables(ScanTcParams& P, int nprob, const int* n_out, const int* K, int nstage, const int* stage_p0, const int* stage_np, int ctas) {
  for (int i = 0; i < nprob; ++i) {
    TcProblem& q = P.prob[i];
    if (n_out[i] % 128 != 0 || K[i] % 64 != 0) return false;
    q.n_out = n_out[i]; q.K = K[i]; q.tiles = n_out[i] / 128; q.kblocks = K[i] / 64; q.kc = pick_kc(q.kblocks);
    q.span = q.tiles * q.kc; q.units = q.tiles * q.kblocks; q.unit_begin = 0;
  }
  for (int st = 0; st < nstage; ++st) {
    P.stage_p0[st] = stage_p0[st]; P.stage_np[st] = stage_np[st];
    int units = 0, span = 1 << 30;
    for (int j = 0; j < stage_np[st]; ++j) {
      TcProblem& q = P.prob[stage_p0[st] + j];
      q.unit_begin = units;
      units += q.units;
      if (q.span < span) span = q.span;
    }
    P.stage_units[st] = units;
    const int per = (units + ctas - 1) / ctas;
    if (per > span) return false;
    int kc_min = kMaxKc;
    for (int j = 0; j < stage_np[st]; ++j) if (P.prob[stage_p0[st] + j].kc < kc_min) kc_min = P.prob[stage_p0[st] + j].kc;
    if (per / kc_min + 4 > kMaxSeg) return false;   // segment table of a CTA (kMaxSeg per stage)
  }
  long long wbytes = 0;
  for (int i = 0; i < nprob; ++i) wbytes += 2LL * n_out[i] * K[i];
  P.evict_first = wbytes > (90LL << 20) ? 1 : 0;   // the bf16 weights of one step do not fit the 126 MB L2 next to everything else
  return P.B * P.T <= kMaxFlags;
}

// problem table + stage table + tensor maps; false when a CTA's range of some stage could touch more than kSlots chunks
bool build_schedule(ScanTcParams& P, TcMaps& maps, int nprob, const int* n_out, const int* K, const void* const* ptr, const int* ld,
                    int nstage, const int* stage_p0, const int* stage_np, int ctas) {
  if (!build_tables(P, nprob, n_out, K, nstage, stage_p0, stage_np, ctas)) return false;
  for (int i = 0; i < nprob; ++i)
    if (!make_wmap(&maps.m[i], ptr[i], n_out[i], K[i], ld[i])) return false;
  for (int i = nprob; i < kMaxProb; ++i) maps.m[i] = maps.m[0];
  return true;
}

void fwd_shapes(int G, int D, int Z, int* n_out, int* K) {
  const int n[4] = {3 * G, 3 * G, D, Z}, k[4] = {D, G, G, D};
  for (int i = 0; i < 4; ++i) { n_out[i] = n[i]; K[i] = k[i]; }
}
void bwd_shapes(int G, int D, int Z, int* n_out, int* K) {
  const int n[5] = {D, G, G, D, Z}, k[5] = {Z, D, 3 * G, 3 * G, D};
  for (int i = 0; i < 5; ++i) { n_out[i] = n[i]; K[i] = k[i]; }
}

// Host-side walk of the device schedule: every CTA's tables are built with the code the kernel runs and every (contraction,
// tile, k-block) unit of every stage must be covered exactly once, within the table bounds, with at most kSlots activation chunks
// per CTA and stage, and with staging counts that never decrease along a CTA's consumption order.
template <class Pol>
int check_schedule(ScanTcParams& P, int nprob, int nstage, int ctas) {
  for (int st = 0; st < nstage; ++st) {
    std::vector<int> seen;
    std::vector<int> base(nprob + 1, 0);
    for (int i = 0; i < nprob; ++i) base[i + 1] = base[i] + P.prob[i].tiles * P.prob[i].kblocks;
    seen.assign(base[nprob], 0);
    for (int blk = 0; blk < ctas; ++blk) {
      StageSched sc;
      const int rc = build_stage_sched<Pol>(P, st, sc, blk, ctas);
      if (rc != 0) return -10 + rc;
      int last_need = 0;
      for (int g = 0; g < sc.nseg; ++g) {
        const SegRec& r = sc.seg[g];
        const TcProblem& q = P.prob[r.prob];
        if (r.prob < P.stage_p0[st] || r.prob >= P.stage_p0[st] + P.stage_np[st] || r.tile >= q.tiles || r.kb1 > q.kc || r.slot >= kSlots) return -20;
        for (int kb = r.kb0; kb < r.kb1; ++kb) {
          const int kb_abs = r.kb_abs0 + (kb - r.kb0);
          if (kb_abs >= q.kblocks) return -21;
          ++seen[base[r.prob] + r.tile * q.kblocks + kb_abs];
          const int need = r.need[kb - r.kb0];
          if (need < 1 || need > sc.count) return -22;
          if (need > last_need) last_need = need;
        }
      }
      if (last_need != sc.count) return -23;
    }
    for (int i = P.stage_p0[st]; i < P.stage_p0[st] + P.stage_np[st]; ++i)
      for (int u = base[i]; u < base[i + 1]; ++u)
        if (seen[u] != 1) return -30;
    for (int i = 0; i < nprob; ++i)
      if (i < P.stage_p0[st] || i >= P.stage_p0[st] + P.stage_np[st])
        for (int u = base[i]; u < base[i + 1]; ++u)
          if (seen[u] != 0) return -31;
  }
  return 0;
}

template <class Pol>
int max_ctas(int device) {
  int sms = 0, per_sm = 0;
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaFuncSetAttribute(scan_tc_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_tc_kernel<Pol>, kThreadsTc, smem);
  return per_sm > 0 ? sms : 0;
}

template <class Pol>
cudaError_t launch(cudaStream_t s, TcMaps& maps, ScanTcParams& P, int ctas) {
  static_assert(sizeof(TcSmem) + 1024 <= 227 * 1024, "TcSmem exceeds the dynamic shared memory limit");
  cudaMemsetAsync(P.bar, 0, sizeof(unsigned int), s);
  const size_t smem = sizeof(TcSmem) + 1024;
  cudaFuncSetAttribute(scan_tc_kernel<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void* args[] = {&maps, &P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)scan_tc_kernel<Pol>, dim3(ctas), dim3(kThreadsTc), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace

bool scan_tc_supported(int B, int G, int D, int Z, int Zk) {
  return D % 128 == 0 && G % 128 == 0 && Z % 128 == 0 && Zk == 32 && D >= 512 && D <= 1024 && G >= 512 && B <= kRows &&
         std::getenv("DV3_NO_SCAN_TC") == nullptr;
}

// 0 = the stream-K schedules of both scan kernels are valid for these sizes on `ctas` CTAs; 1 = sizes not handled by the kernels
// (the engine then uses the other scan paths); < 0 = a defect of the schedule (see check_schedule)
int scan_tc_schedule_check(int B, int T, int G, int D, int Z, int ctas) {
  if (!(D % 128 == 0 && G % 128 == 0 && Z % 128 == 0 && D >= 512 && D <= 1024 && G >= 512 && B <= kRows) || ctas < 1) return 1;
  const int sp0f[3] = {0, 2, 3}, snpf[3] = {2, 1, 1}, sp0b[4] = {0, 1, 2, 4}, snpb[4] = {1, 1, 2, 1};
  int n_out[5], K[5];
  {
    ScanTcParams P{};
    P.B = B; P.T = T;
    fwd_shapes(G, D, Z, n_out, K);
    if (!build_tables(P, 4, n_out, K, 3, sp0f, snpf, ctas)) return 1;
    const int rc = check_schedule<FwdPolicy>(P, 4, 3, ctas);
    if (rc != 0) return rc;
  }
  {
    ScanTcParams P{};
    P.B = B; P.T = T;
    bwd_shapes(G, D, Z, n_out, K);
    if (!build_tables(P, 5, n_out, K, 4, sp0b, snpb, ctas)) return 1;
    const int rc = check_schedule<BwdPolicy>(P, 5, 4, ctas);
    if (rc != 0) return rc - 100;
  }
  return 0;
}

int scan_tc_max_ctas(int device) {
  const int a = max_ctas<FwdPolicy>(device), b = max_ctas<BwdPolicy>(device);
  return a < b ? a : b;
}

cudaError_t launch_observe_scan_tc(cudaStream_t s, const ScanFwdParams& p, const ScanTcWeights& w, float* gx_acc, int ctas) {
  ScanTcParams P{};
  P.p = p; P.gx_acc = gx_acc; P.B = p.B; P.T = p.T; P.bar = p.bar; P.stamps = p.stamps;
  const int G = p.G, D = p.D, Z = p.Z;
  int n_out[4], K[4];
  fwd_shapes(G, D, Z, n_out, K);
  const void* ptr[4] = {w.Kt, w.Ut, w.Wpost_h_t, w.Wzq_t};
  const int ld[4] = {w.ld_Kt, w.ld_Ut, w.ld_Wpost_h_t, w.ld_Wzq_t};
  const int sp0[3] = {0, 2, 3}, snp[3] = {2, 1, 1};
  TcMaps maps;
  if (!build_schedule(P, maps, 4, n_out, K, ptr, ld, 3, sp0, snp, ctas)) return cudaErrorInvalidConfiguration;
  {
    const long long n4 = (long long)p.B * p.T * (6 * G + Z + D) / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    scan_tc_init_kernel<<<blocks, 256, 0, s>>>(p, gx_acc);
    count_launch();
  }
  return launch<FwdPolicy>(s, maps, P, ctas);
}

// q.dp_act and q.du must be zero on entry (accumulators); q.d_hz holds the head / prior gradients
cudaError_t launch_observe_scan_tc_bwd(cudaStream_t s, const ScanBwdParams& q, const ScanTcWeightsBwd& w, int ctas) {
  ScanTcParams P{};
  P.q = q; P.B = q.B; P.T = q.T; P.bar = q.bar; P.stamps = q.stamps;
  const int G = q.G, D = q.D, Z = q.Z;
  int n_out[5], K[5];
  bwd_shapes(G, D, Z, n_out, K);
  const void* ptr[5] = {w.Wzq, w.Wpost_h, w.U, w.K, w.Wpre_z};
  const int ld[5] = {w.ld_Wzq, w.ld_Wpost_h, w.ld_U, w.ld_K, w.ld_Wpre_z};
  const int sp0[4] = {0, 1, 2, 4}, snp[4] = {1, 1, 2, 1};
  TcMaps maps;
  if (!build_schedule(P, maps, 5, n_out, K, ptr, ld, 4, sp0, snp, ctas)) return cudaErrorInvalidConfiguration;
  return launch<BwdPolicy>(s, maps, P, ctas);
}

}  // namespace dv3
