// Persistent imagination rollout (DreamerModel.dream_trajectory, dreamer_model.py:189-208) for the small layer sizes (XS, S) in
// the tensor-core mode: ONE cooperative launch walks all H dreamed steps. A step is a chain of grid-wide stages separated by a
// one-atomic grid barrier (~1.7 us) instead of ~12 dependent kernel launches (~7 us each in the replayed graph). Contraction
// stages (G) alternate with elementwise stages (E) that touch every element exactly once over the whole grid (one warp per
// dream row) and write the bf16 operand of the next contraction:
//
//   G0  [h, z]_{i-1}  x  { actor layer 0 | actor-std layer 0 (Box) | z-part of the pre-GRU layer | U }
//   (E  SiLU(LN(.)) of the actor's hidden layer;  G  actor layer 1 ...   for L = 2)
//   E1  actor: LN+SiLU, the D x A output layer, unimix + inverse-CDF sample or the reparameterised Normal; the pre-GRU
//       pre-activation gets a . W_pre[Z:]; LN+SiLU -> u
//   G1  gx = u . K        E2  GRU cell -> h_i        G2  q_pre = h_i . W_dyn        E3  LN+SiLU -> q
//   G3  prior logits = q . W_z  in standard orientation (dream rows = TMEM lanes): the epilogue thread owns whole categoricals
//       of its row: softmax, 1 % unimix, fp64 inverse-CDF sample without a shuffle -> z_i (one-hot, fp32 + bf16)
//
// Work unit of a contraction stage = (contraction, 128-row weight tile, 128 dream rows) over the full K, dealt round-robin to
// the CTAs; both operands arrive by TMA (weights from the bf16 K-major shadows, activations from the bf16 twins written by
// the elementwise stages), the MMA is tcgen05 128 x 128 x 16 with a double-buffered fp32 accumulator in TMEM. Every tensor the
// per-kernel path produces is produced here in the same buffers, so the heads, the losses and the BPTT of the Box actor run
// unchanged afterwards.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "dv3_device.cuh"
#include "dv3_dream_tc.h"

namespace dv3 {
namespace {

constexpr int kThreadsD = 320;       // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 workers
constexpr int kWorkersD = 256;
constexpr int kWRing = 6;            // (weight, activation) k-block pairs in flight: 2 x [128 rows x 64 k] bf16 = 32 KB each
constexpr int kBlk = 128 * 128;      // bytes of one [128 rows x 64 k] bf16 k-block
constexpr int kMaxProbD = 9, kMaxStagesD = 5, kMaxSeqD = 12;

// b_map: activation source (0 [h, z], 1 u, 2 q, 3 actor hidden, 4 actor-std hidden); b_col0: first column; b_curr: rows of
// dream step i (1) or i - 1 (0)
struct DProblem { int n_out, K, tiles, kblocks, b_map, b_col0, b_curr, unit_begin, units; };

struct DMaps { CUtensorMap w[kMaxProbD]; CUtensorMap b[5]; };

enum { SQ_GEMM = 0, SQ_ELEM_MLP, SQ_ELEM_ACTOR, SQ_ELEM_GRU, SQ_ELEM_Q };
struct DParams {
  DreamTcArgs a;
  DProblem prob[kMaxProbD];
  int nstage, stage_p0[kMaxStagesD], stage_np[kMaxStagesD], stage_units[kMaxStagesD];   // contraction stages
  int nseq, seq_kind[kMaxSeqD], seq_arg[kMaxSeqD];   // the step: SQ_GEMM (arg = contraction stage) or an elementwise stage
  int groups;   // N / 128
};

struct DSmem {
  unsigned char wring[kWRing][kBlk];
  unsigned char bbuf[kWRing][kBlk];
  float vec[12][512];           // LayerNorm gain / offset pairs: actor l0, l1, actor-std l0, l1, pre-GRU layer, prior net
  uint64_t full[kWRing], empty[kWRing], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t gemm_seq;            // contraction stages whose inputs are visible (the TMA producer waits on it)
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
// kind::f16, bf16 x bf16 -> fp32, both K-major, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc128() { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) {
      printf("dv3 dream_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
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
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkersD) : "memory"); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t wait_seq(const uint32_t* flag, uint32_t need, uint32_t have) {
  uint32_t spins = 0;
  while ((int32_t)(have - need) < 0) {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(flag)) : "memory");
    if (++spins > (1u << 26)) { printf("dv3 dream_tc: flag wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
  return have;
}
__device__ __forceinline__ void grid_bar_workers(unsigned int* ctr, unsigned int& target, int wt) {
  worker_sync();
  if (wt == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (++spins > (1u << 26)) { printf("dv3 dream_tc: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    } while (v < target);
  }
  worker_sync();
}

struct Unit { int prob, tile, group; };
__device__ __forceinline__ bool unit_of(const DParams& P, int st, int j, Unit& u) {
  if (j >= P.stage_units[st]) return false;
  int pr = P.stage_p0[st];
  for (int i = 1; i < P.stage_np[st]; ++i)
    if (j >= P.prob[P.stage_p0[st] + i].unit_begin) pr = P.stage_p0[st] + i;
  const int r = j - P.prob[pr].unit_begin;
  u.prob = pr; u.tile = r / P.groups; u.group = r - u.tile * P.groups;
  return true;
}

__device__ __forceinline__ void st_bf16x4(void* base16, size_t idx, float v0, float v1, float v2, float v3) {
  uint2 o; o.x = pack_bf16(v0, v1); o.y = pack_bf16(v2, v3);
  *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base16) + idx) = o;
}

// vec slots in shared memory (gain, offset pairs)
enum { V_ACT_G0 = 0, V_ACT_B0, V_ACT_G1, V_ACT_B1, V_STD_G0, V_STD_B0, V_STD_G1, V_STD_B1, V_PRE_G, V_PRE_B, V_DYN_G, V_DYN_B };

// LayerNorm (eps 1e-3, biased variance, two-pass) + SiLU of one row held as NPL float4 per lane (lane owns columns
// (lane + 32 i) * 4 .. +3); returns mean / rstd, overwrites v with the activations
template <int NPL>
__device__ __forceinline__ void ln_silu_row(float4 (&v)[NPL], int D, const float* g, const float* be, float& mean, float& rstd, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  rstd = rsqrtf(q / (float)D + DV3_LN_EPS);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(g + c), b4 = *reinterpret_cast<const float4*>(be + c);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g4.x + b4.x; o.y = (v[i].y - mean) * rstd * g4.y + b4.y;
    o.z = (v[i].z - mean) * rstd * g4.z + b4.z; o.w = (v[i].w - mean) * rstd * g4.w + b4.w;
    o.x *= sigmoidf_(o.x); o.y *= sigmoidf_(o.y); o.z *= sigmoidf_(o.z); o.w *= sigmoidf_(o.w);
    v[i] = o;
  }
}

// ---- elementwise stages: every element once over the whole grid (one warp per dream row); all loads of a row are issued
// before the first use (the grid barrier invalidated L1: each dependent load is a full L2 round trip) --------------------------

// actor hidden layer l:  act = SiLU(LN(pre[l])) (+ bf16 twin = the B operand of layer l + 1)
template <int NPL>
__device__ void elem_mlp_layer(const DParams& P, DSmem& sm, int l, size_t row, int lane) {
  const DreamTcArgs& a = P.a;
  const int D = a.D;
  const bool box = !a.discrete;
  float4 v[2][NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const size_t c = (size_t)(lane + 32 * i) * 4;
    v[0][i] = ldcg4(a.actor.pre[l] + row * D + c);
    v[1][i] = box ? ldcg4(a.actor_std.pre[l] + row * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    if (net == 1 && !box) break;
    const DreamTcMlp& m = net == 0 ? a.actor : a.actor_std;
    const float* g = sm.vec[(net == 0 ? V_ACT_G0 : V_STD_G0) + 2 * l];
    float mean, rstd;
    ln_silu_row<NPL>(v[net], D, g, g + 512, mean, rstd, lane);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = (lane + 32 * i) * 4;
      *reinterpret_cast<float4*>(m.act[l] + row * D + c) = v[net][i];
      st_bf16x4(m.act16[l], row * D + c, v[net][i].x, v[net][i].y, v[net][i].z, v[net][i].w);
    }
    if (lane == 0) { m.mean[l][row] = mean; m.rstd[l][row] = rstd; }
  }
}

// finish the actor of state i-1 (dream row ra), sample the action, complete the pre-GRU pre-activation of step i (row rs of the
// per-step buffers, row rl of the per-step scratch) and write u = SiLU(LN(x_pre)) (+ bf16 twin = the B operand of gx)
template <int NPL>
__device__ void elem_actor_and_pre(const DParams& P, DSmem& sm, size_t ra, size_t rs, size_t rl, int lane) {
  const DreamTcArgs& a = P.a;
  const int D = a.D, A = a.A, L = a.L;
  const bool box = !a.discrete;
  float4 pm[NPL], ps[NPL], x[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const size_t c = (size_t)(lane + 32 * i) * 4;
    pm[i] = ldcg4(a.actor.pre[L - 1] + ra * D + c);
    ps[i] = box ? ldcg4(a.actor_std.pre[L - 1] + ra * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    x[i] = ldcg4(a.zpart + rl * D + c);
  }
  // ---- actor output layer(s) on SiLU(LN(pre[L-1])) ----
  float out_m[kDreamTcMaxA], out_s[kDreamTcMaxA];
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    if (net == 1 && !box) break;
    const DreamTcMlp& m = net == 0 ? a.actor : a.actor_std;
    float4 (&v)[NPL] = net == 0 ? pm : ps;
    const float* g = sm.vec[(net == 0 ? V_ACT_G0 : V_STD_G0) + 2 * (L - 1)];
    float mean, rstd;
    ln_silu_row<NPL>(v, D, g, g + 512, mean, rstd, lane);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = (lane + 32 * i) * 4;
      *reinterpret_cast<float4*>(m.act[L - 1] + ra * D + c) = v[i];
      if (m.act16[L - 1] != nullptr) st_bf16x4(m.act16[L - 1], ra * D + c, v[i].x, v[i].y, v[i].z, v[i].w);
    }
    if (lane == 0) { m.mean[L - 1][ra] = mean; m.rstd[L - 1][ra] = rstd; }
    const float* wo = net == 0 ? a.actor_out_w : a.std_out_w;   // [D, A]
    const float* bo = net == 0 ? a.actor_out_b : a.std_out_b;
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) {
      float s = 0.f;
      if (j < A) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const int c = (lane + 32 * i) * 4;
          s += v[i].x * __ldg(wo + (size_t)c * A + j) + v[i].y * __ldg(wo + (size_t)(c + 1) * A + j) +
               v[i].z * __ldg(wo + (size_t)(c + 2) * A + j) + v[i].w * __ldg(wo + (size_t)(c + 3) * A + j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += __ldg(bo + j);
      }
      if (net == 0) out_m[j] = s; else out_s[j] = s;
    }
  }
  // ---- the action (actor_network.py:63-118) ----
  float act[kDreamTcMaxA];
  if (!box) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) if (j < A) mx = fmaxf(mx, out_m[j]);
    float e[kDreamTcMaxA], s = 0.f;
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) { e[j] = j < A ? expf(out_m[j] - mx) : 0.f; s += e[j]; }
    double cdf = 0.0, tot = 0.0;
    float pu[kDreamTcMaxA];
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) { pu[j] = j < A ? 0.99f * (e[j] / s) + 0.01f * (1.0f / (float)A) : 0.f; tot += (double)pu[j]; }
    const double thr = (double)__ldg(a.act_noise + ra) * tot;
    int pick = 0;
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) { if (j < A) { cdf += (double)pu[j]; if (cdf <= thr) ++pick; } }
    pick = min(pick, A - 1);
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) act[j] = j == pick ? 1.f : 0.f;
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < kDreamTcMaxA; ++j)
        if (j < A) { a.actor_logits[ra * A + j] = out_m[j]; a.actor_probs[ra * A + j] = pu[j]; a.dr_act[ra * A + j] = act[j]; }
      a.a_idx[ra] = pick;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) {
      act[j] = 0.f;
      if (j < A) {
        const float sd = 0.9f * sigmoidf_(out_s[j] + 2.0f) + 0.1f;
        act[j] = tanhf(out_m[j]) + sd * __ldg(a.act_noise + ra * A + j);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < kDreamTcMaxA; ++j)
        if (j < A) { a.actor_logits[ra * A + j] = out_m[j]; a.actor_s[ra * A + j] = out_s[j]; a.dr_act[ra * A + j] = act[j]; }
    }
  }
  // ---- x_pre = z . W_pre[:Z] (contraction stage 0) + a . W_pre[Z:];  u = SiLU(LN(x_pre)) ----
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = (lane + 32 * i) * 4;
#pragma unroll
    for (int j = 0; j < kDreamTcMaxA; ++j) {
      if (j < A) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.W_pre_a + (size_t)j * D + c));
        x[i].x += act[j] * w.x; x[i].y += act[j] * w.y; x[i].z += act[j] * w.z; x[i].w += act[j] * w.w;
      }
    }
    *reinterpret_cast<float4*>(a.x_pre + rs * D + c) = x[i];
  }
  float mean, rstd;
  ln_silu_row<NPL>(x, D, sm.vec[V_PRE_G], sm.vec[V_PRE_B], mean, rstd, lane);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = (lane + 32 * i) * 4;
    *reinterpret_cast<float4*>(a.u + rs * D + c) = x[i];
    st_bf16x4(a.u16, rs * D + c, x[i].x, x[i].y, x[i].z, x[i].w);
  }
  if (lane == 0) { a.x_mean[rs] = mean; a.x_rstd[rs] = rstd; }
}

// GRU cell (Keras reset_after; gx / gh hold their biases): 4 units of one row; h_i (+ bf16 twin = the B operand of the prior net)
__device__ __forceinline__ void elem_gru(const DParams& P, size_t rs, size_t rp, size_t rc, size_t rl, int c) {
  const DreamTcArgs& a = P.a;
  const int G = a.G, W = a.G + a.Z;
  const float* gx = a.gx_raw + rl * 3 * G + c;
  const float* gh = a.gh + rs * 3 * G + c;
  const float4 xz = ldcg4(gx), xr = ldcg4(gx + G), xc = ldcg4(gx + 2 * G);
  const float4 hz_ = ldcg4(gh), hr = ldcg4(gh + G), hc = ldcg4(gh + 2 * G);
  const float4 hp = ldcg4(a.hz + rp * W + c);
  float4 z, rg, cc, h;
#define DV3_GRU(f) \
  z.f = sigmoidf_(xz.f + hz_.f); rg.f = sigmoidf_(xr.f + hr.f); cc.f = tanhf(xc.f + rg.f * hc.f); \
  h.f = z.f * hp.f + (1.f - z.f) * cc.f;
  DV3_GRU(x) DV3_GRU(y) DV3_GRU(z) DV3_GRU(w)
#undef DV3_GRU
  float* gt = a.gx + rs * 3 * G + c;   // the gates (what gru_fwd leaves in gx for the backward pass)
  *reinterpret_cast<float4*>(gt) = z; *reinterpret_cast<float4*>(gt + G) = rg; *reinterpret_cast<float4*>(gt + 2 * G) = cc;
  *reinterpret_cast<float4*>(a.hz + rc * W + c) = h;
  st_bf16x4(a.hz16, rc * W + c, h.x, h.y, h.z, h.w);
}

// prior net: q = SiLU(LN(q_pre)) (+ bf16 twin = the B operand of the prior logits)
template <int NPL>
__device__ void elem_q(const DParams& P, DSmem& sm, size_t rs, int lane) {
  const DreamTcArgs& a = P.a;
  const int D = a.D;
  float4 v[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) v[i] = ldcg4(a.q_pre + rs * D + (size_t)(lane + 32 * i) * 4);
  float mean, rstd;
  ln_silu_row<NPL>(v, D, sm.vec[V_DYN_G], sm.vec[V_DYN_B], mean, rstd, lane);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = (lane + 32 * i) * 4;
    *reinterpret_cast<float4*>(a.q_act + rs * D + c) = v[i];
    st_bf16x4(a.q_act16, rs * D + c, v[i].x, v[i].y, v[i].z, v[i].w);
  }
  if (lane == 0) { a.q_mean[rs] = mean; a.q_rstd[rs] = rstd; }
}

// problem ids
enum { PB_ACT0 = 0, PB_STD0, PB_ZPART, PB_GH, PB_ACT1, PB_STD1, PB_GX, PB_Q, PB_LOGITS };

#define DV3_TMEM_LD64(r, taddr)                                                                                          \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                           \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "                                  \
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "                                  \
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"                          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),       \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),            \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),           \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),           \
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),           \
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),           \
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),           \
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])                         \
      : "r"(taddr))

// epilogue of the prior-logits unit (standard orientation): this thread owns dream row `row` and the 64 logits of two whole
// categoricals; RepresentationLayer forward (representation_layer.py:73-113) without a shuffle
__device__ __forceinline__ void sample_epilogue(const DreamTcArgs& a, const uint32_t (&r)[64], size_t rs, size_t rc, int n0) {
  const int Z = a.Z, W = a.G + a.Z, Zc = a.Zc;
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int cat = (n0 >> 5) + h2;
    float l[32];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k) { l[k] = __uint_as_float(r[h2 * 32 + k]) + __ldg(a.zp_bias + n0 + h2 * 32 + k); mx = fmaxf(mx, l[k]); }
    float* lo = a.prior_logits + rs * Z + n0 + h2 * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(lo + k) = make_float4(l[k], l[k + 1], l[k + 2], l[k + 3]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) { l[k] = expf(l[k] - mx); s += l[k]; }
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) { l[k] = 0.99f * (l[k] / s) + 0.01f * (1.0f / 32.0f); tot += (double)l[k]; }
    float* po = a.prior_probs + rs * Z + n0 + h2 * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(po + k) = make_float4(l[k], l[k + 1], l[k + 2], l[k + 3]);
    const double thr = (double)__ldg(a.u_dyn + rs * Zc + cat) * tot;
    double cdf = 0.0;
    int pick = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) { cdf += (double)l[k]; if (cdf <= thr) ++pick; }
    pick = min(pick, 31);
    a.z_idx[rs * Zc + cat] = pick;
    float* zo = a.hz + rc * W + a.G + n0 + h2 * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 4)
      *reinterpret_cast<float4*>(zo + k) = make_float4(k == pick ? 1.f : 0.f, k + 1 == pick ? 1.f : 0.f, k + 2 == pick ? 1.f : 0.f, k + 3 == pick ? 1.f : 0.f);
    if (a.hz16 != nullptr) {
      unsigned short* z16 = reinterpret_cast<unsigned short*>(a.hz16) + rc * W + a.G + n0 + h2 * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        uint4 o;
        o.x = (k == pick ? 0x3F80u : 0u) | (k + 1 == pick ? 0x3F800000u : 0u);
        o.y = (k + 2 == pick ? 0x3F80u : 0u) | (k + 3 == pick ? 0x3F800000u : 0u);
        o.z = (k + 4 == pick ? 0x3F80u : 0u) | (k + 5 == pick ? 0x3F800000u : 0u);
        o.w = (k + 6 == pick ? 0x3F80u : 0u) | (k + 7 == pick ? 0x3F800000u : 0u);
        *reinterpret_cast<uint4*>(z16 + k) = o;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreadsD, 1) dream_tc_kernel(const __grid_constant__ DMaps maps, const DParams P) {
  extern __shared__ unsigned char smem_raw[];
  DSmem& sm = *reinterpret_cast<DSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const DreamTcArgs& a = P.a;
  const int H = a.H, N = a.N, G = a.G, D = a.D;

  if (tid == 0) {
    for (int i = 0; i < kMaxProbD; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[i]) : "memory");
    for (int i = 0; i < 5; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[i]) : "memory");
    for (int s = 0; s < kWRing; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], kWorkersD); }
    sm.gemm_seq = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {   // LayerNorm vectors into shared memory (the grid barrier's acquire invalidates L1)
    const float* src[12] = {a.actor.g[0], a.actor.b[0], a.actor.g[1], a.actor.b[1], a.actor_std.g[0], a.actor_std.b[0],
                            a.actor_std.g[1], a.actor_std.b[1], a.pre_g, a.pre_b, a.dyn_g, a.dyn_b};
    for (int i = tid; i < 12 * D; i += kThreadsD) { const int v = i / D, c = i - v * D; if (src[v] != nullptr) sm.vec[v][c] = src[v][c]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: every contraction stage reads activations written (with ordinary stores, by other CTAs)
    // in the stage before it: wait for the workers' word that the barrier has been passed, then order the async-proxy reads
    if (lane == 0) {
      uint32_t it = 0, have = 0, round = 0;
      for (int i = 1; i <= H; ++i) {
        const int prev_row = (i - 1) * N, curr_row = i * N;
        for (int sq = 0; sq < P.nseq; ++sq) {
          if (P.seq_kind[sq] != SQ_GEMM) continue;
          const int st = P.seq_arg[sq];
          ++round;
          have = wait_seq(&sm.gemm_seq, round, have);
          asm volatile("fence.proxy.async;" ::: "memory");
          Unit u;
          for (int j = blockIdx.x; unit_of(P, st, j, u); j += gridDim.x) {
            const DProblem& q = P.prob[u.prob];
            const int brow = (q.b_curr ? curr_row : prev_row) + u.group * 128;
            for (int kb = 0; kb < q.kblocks; ++kb, ++it) {
              const int slot = it % kWRing;
              mbar_wait(&sm.empty[slot], ((it / kWRing) & 1) ^ 1);
              mbar_expect_tx(&sm.full[slot], 2 * kBlk);
              tma_load_2d(smem_u32(sm.wring[slot]), &maps.w[u.prob], kb * 64, u.tile * 128, &sm.full[slot]);
              tma_load_2d(smem_u32(sm.bbuf[slot]), &maps.b[q.b_map], q.b_col0 + kb * 64, brow, &sm.full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc128();
      uint32_t it = 0, acc_it = 0;
      for (int i = 1; i <= H; ++i)
        for (int sq = 0; sq < P.nseq; ++sq) {
          if (P.seq_kind[sq] != SQ_GEMM) continue;
          const int st = P.seq_arg[sq];
          Unit u;
          for (int j = blockIdx.x; unit_of(P, st, j, u); j += gridDim.x) {
            const DProblem& q = P.prob[u.prob];
            const uint32_t as = acc_it & 1;
            mbar_wait(&sm.acc_empty[as], ((acc_it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + as * 128;
            const bool standard = u.prob == PB_LOGITS;   // dream rows as the M operand (TMEM lanes)
            for (int kb = 0; kb < q.kblocks; ++kb, ++it) {
              const int slot = it % kWRing;
              mbar_wait(&sm.full[slot], (it / kWRing) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t w_addr = smem_u32(sm.wring[slot]), x_addr = smem_u32(sm.bbuf[slot]);
              const uint32_t m_addr = standard ? x_addr : w_addr, n_addr = standard ? w_addr : x_addr;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const uint64_t da = make_desc(m_addr + jj * 32), db = make_desc(n_addr + jj * 32);
                const uint32_t accum = (kb == 0 && jj == 0) ? 0u : 1u;
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
    // ---------------- workers: epilogues of the contraction stages, elementwise stages ----------------
    const int wt = tid - 64, ww = warp - 2, q4 = warp & 3, half = ww >> 2;
    unsigned int bar_target = 0;
    uint32_t acc_it = 0, gseq = 0;
    const int NPL = D / 128;
    const int gwarp = blockIdx.x * (kWorkersD / 32) + ww, nwarps = gridDim.x * (kWorkersD / 32);
    const int gthread = blockIdx.x * kWorkersD + wt, nthreads = gridDim.x * kWorkersD;
    // optional cycle totals of worker thread 0: 1 elementwise, 3 accumulator wait, 4 epilogue, 5 grid barrier, 6 stage fence
    long long pf1 = 0, pf3 = 0, pf4 = 0, pf5 = 0, pf6 = 0, tc0 = clock64();
    long long work[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per stage of the step: cycles from the barrier before it to the barrier arrive
    const bool profiling = a.prof != nullptr;
    auto lap = [&](long long& acc) { if (profiling) { const long long c = clock64(); acc += c - tc0; tc0 = c; } };
    for (int i = 1; i <= H; ++i) {
      const size_t prev = (size_t)(i - 1) * N, curr = (size_t)i * N, stp = (size_t)(i - 1) * N;
      for (int sq = 0; sq < P.nseq; ++sq) {
        const int kind = P.seq_kind[sq];
        const long long tw0 = profiling ? clock64() : 0;
        if (kind == SQ_GEMM) {
          const int st = P.seq_arg[sq];
          // everything this stage reads through TMA is visible now (kernel start / the barrier that closed the stage before)
          asm volatile("fence.proxy.async;" ::: "memory");
          worker_sync();
          ++gseq;
          if (wt == 0) asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&sm.gemm_seq)), "r"(gseq) : "memory");
          lap(pf6);
          Unit u;
          for (int j = blockIdx.x; unit_of(P, st, j, u); j += gridDim.x) {
            const size_t g0 = (size_t)u.group * 128;
            const uint32_t as = acc_it & 1;
            // output pointer / bias of this unit before the accumulator is waited for (the bias load is an L2 round trip)
            const int n = u.tile * 128 + q4 * 32 + lane;
            float* out = nullptr; size_t rowbase = 0; int ld = D; float bias = 0.f;
            switch (u.prob) {
              case PB_ACT0: out = a.actor.pre[0]; rowbase = prev; break;
              case PB_STD0: out = a.actor_std.pre[0]; rowbase = prev; break;
              case PB_ACT1: out = a.actor.pre[1]; rowbase = prev; break;
              case PB_STD1: out = a.actor_std.pre[1]; rowbase = prev; break;
              case PB_ZPART: out = a.zpart; rowbase = 0; break;
              case PB_GH: out = a.gh; rowbase = stp; ld = 3 * G; bias = __ldg(a.gru_b + 3 * G + n); break;
              case PB_GX: out = a.gx_raw; rowbase = 0; ld = 3 * G; bias = __ldg(a.gru_b + n); break;
              case PB_Q: out = a.q_pre; rowbase = stp; break;
              default: break;
            }
            mbar_wait(&sm.acc_full[as], (acc_it >> 1) & 1);
            lap(pf3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (u.prob == PB_LOGITS) {
              uint32_t r[64];
              const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 128 + half * 64);
              DV3_TMEM_LD64(r, taddr);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(&sm.acc_empty[as]);
              const size_t row = g0 + q4 * 32 + lane;
              sample_epilogue(a, r, stp + row, curr + row, u.tile * 128 + half * 64);
            } else {
              // swap-AB: lane = output column n, accumulator columns = dream rows g0 + half * 64 + j; two rounds of 32 columns
              float* o = out + (rowbase + g0 + (size_t)half * 64) * ld + n;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 128 + half * 64 + hh * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (hh == 1) {
                  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                  mbar_arrive(&sm.acc_empty[as]);
                }
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) o[(size_t)(hh * 32 + jj) * ld] = __uint_as_float(r[jj]) + bias;
              }
            }
            ++acc_it;
            lap(pf4);
          }
        } else if (kind == SQ_ELEM_MLP) {
          for (int row = gwarp; row < N; row += nwarps) {
            if (NPL == 2) elem_mlp_layer<2>(P, sm, P.seq_arg[sq], prev + row, lane);
            else elem_mlp_layer<4>(P, sm, P.seq_arg[sq], prev + row, lane);
          }
          lap(pf1);
        } else if (kind == SQ_ELEM_ACTOR) {
          for (int row = gwarp; row < N; row += nwarps) {
            if (NPL == 2) elem_actor_and_pre<2>(P, sm, prev + row, stp + row, row, lane);
            else elem_actor_and_pre<4>(P, sm, prev + row, stp + row, row, lane);
          }
          lap(pf1);
        } else if (kind == SQ_ELEM_GRU) {
          const int per_row = G / 4;
          for (int e = gthread; e < N * per_row; e += nthreads) {
            const int row = e / per_row, c = (e - row * per_row) * 4;
            elem_gru(P, stp + row, prev + row, curr + row, row, c);
          }
          lap(pf1);
        } else {
          for (int row = gwarp; row < N; row += nwarps) {
            if (NPL == 2) elem_q<2>(P, sm, stp + row, lane);
            else elem_q<4>(P, sm, stp + row, lane);
          }
          lap(pf1);
        }
        // writers of the bf16 twins: make them visible to the async proxy of every SM before the barrier releases
        asm volatile("fence.proxy.async;" ::: "memory");
        if (profiling && sq < 8) work[sq] += clock64() - tw0;
        grid_bar_workers(a.bar, bar_target, wt);
        lap(pf5);
      }
    }
    if (profiling && wt == 0) {
      unsigned long long* o = a.prof + 16 * blockIdx.x;
      o[0] = 0; o[1] = pf1; o[2] = 0; o[3] = pf3; o[4] = pf4; o[5] = pf5; o[6] = pf6; o[7] = 0;
      for (int k = 0; k < 8; ++k) o[8 + k] = work[k];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
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
bool make_map(CUtensorMap* map, const void* ptr, long long rows, int K, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool dream_tc_supported(int N, int G, int D, int Z, int Zk, int A, int L) {
  return N % 128 == 0 && N >= 128 && (D == 256 || D == 512) && G % 64 == 0 && G >= 128 && G <= 512 && (G + Z) % 64 == 0 && Z % 128 == 0 &&
         Zk == 32 && A >= 1 && A <= kDreamTcMaxA && L >= 1 && L <= kDreamTcMaxL;
}

int dream_tc_max_ctas(int device) {
  int sms = 0, per_sm = 0;
  const size_t smem = sizeof(DSmem) + 1024;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaFuncSetAttribute(dream_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dream_tc_kernel, kThreadsD, smem);
  return per_sm > 0 ? sms : 0;
}

cudaError_t launch_dream_tc(cudaStream_t s, const DreamTcArgs& a, int ctas) {
  static_assert(sizeof(DSmem) + 1024 <= 227 * 1024, "DSmem exceeds the dynamic shared memory limit");
  DParams P{};
  P.a = a;
  P.groups = a.N / 128;
  const int G = a.G, D = a.D, Z = a.Z, W = a.G + a.Z;
  const bool box = !a.discrete;
  const long long R = (long long)(a.H + 1) * a.N, HN = (long long)a.H * a.N;
  DMaps maps;
  auto set = [&](int id, const void* w, int ld, int n_out, int K, int bmap, int col0, int curr) -> bool {
    DProblem& q = P.prob[id];
    q.n_out = n_out; q.K = K; q.tiles = (n_out + 127) / 128; q.kblocks = (K + 63) / 64; q.b_map = bmap; q.b_col0 = col0; q.b_curr = curr;
    q.units = q.tiles * P.groups; q.unit_begin = 0;
    return make_map(&maps.w[id], w, n_out, K, ld);
  };
  bool ok = true;
  ok &= set(PB_ACT0, a.actor.wt[0], a.actor.ld_wt[0], D, W, 0, 0, 0);
  ok &= box ? set(PB_STD0, a.actor_std.wt[0], a.actor_std.ld_wt[0], D, W, 0, 0, 0) : set(PB_STD0, a.actor.wt[0], a.actor.ld_wt[0], D, W, 0, 0, 0);
  ok &= set(PB_ZPART, a.Wt_pre_z, a.ld_pre_z, D, Z, 0, G, 0);
  ok &= set(PB_GH, a.Wt_U, a.ld_U, 3 * G, G, 0, 0, 0);
  if (a.L == 2) {
    ok &= set(PB_ACT1, a.actor.wt[1], a.actor.ld_wt[1], D, D, 3, 0, 0);
    ok &= box ? set(PB_STD1, a.actor_std.wt[1], a.actor_std.ld_wt[1], D, D, 4, 0, 0) : set(PB_STD1, a.actor.wt[1], a.actor.ld_wt[1], D, D, 3, 0, 0);
  } else {
    maps.w[PB_ACT1] = maps.w[PB_ACT0]; maps.w[PB_STD1] = maps.w[PB_ACT0];
  }
  ok &= set(PB_GX, a.Wt_K, a.ld_K, 3 * G, D, 1, 0, 0);
  ok &= set(PB_Q, a.Wt_dyn, a.ld_dyn, D, G, 0, 0, 1);
  ok &= set(PB_LOGITS, a.Wt_zp, a.ld_zp, Z, D, 2, 0, 0);
  ok &= make_map(&maps.b[0], a.hz16, R, W, W);
  ok &= make_map(&maps.b[1], a.u16, HN, D, D);
  ok &= make_map(&maps.b[2], a.q_act16, HN, D, D);
  if (a.L == 2) {
    ok &= make_map(&maps.b[3], a.actor.act16[0], R, D, D);
    ok &= box ? make_map(&maps.b[4], a.actor_std.act16[0], R, D, D) : make_map(&maps.b[4], a.actor.act16[0], R, D, D);
  } else {
    maps.b[3] = maps.b[1]; maps.b[4] = maps.b[1];
  }
  if (!ok) return cudaErrorInvalidValue;
  int nst = 0, nsq = 0;
  auto add_gemm = [&](std::initializer_list<int> ids) {   // (the ids of a stage are consecutive by construction of the enum)
    int units = 0, first = -1, cnt = 0;
    for (int id : ids) {
      if (first < 0) first = id;
      P.prob[id].unit_begin = units;
      units += P.prob[id].units;
      ++cnt;
    }
    P.stage_p0[nst] = first; P.stage_np[nst] = cnt; P.stage_units[nst] = units;
    P.seq_kind[nsq] = SQ_GEMM; P.seq_arg[nsq] = nst;
    ++nst; ++nsq;
  };
  auto add_elem = [&](int kind, int arg) { P.seq_kind[nsq] = kind; P.seq_arg[nsq] = arg; ++nsq; };
  if (!box) { P.prob[PB_STD0].units = 0; P.prob[PB_STD1].units = 0; }   // the Discrete actor has no std net
  add_gemm({PB_ACT0, PB_STD0, PB_ZPART, PB_GH});
  if (a.L == 2) { add_elem(SQ_ELEM_MLP, 0); add_gemm({PB_ACT1, PB_STD1}); }
  add_elem(SQ_ELEM_ACTOR, 0);
  add_gemm({PB_GX});
  add_elem(SQ_ELEM_GRU, 0);
  add_gemm({PB_Q});
  add_elem(SQ_ELEM_Q, 0);
  add_gemm({PB_LOGITS});
  P.nstage = nst; P.nseq = nsq;
  cudaMemsetAsync(a.bar, 0, sizeof(unsigned int), s);
  const size_t smem = sizeof(DSmem) + 1024;
  cudaFuncSetAttribute(dream_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  void* args[] = {&maps, &P};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)dream_tc_kernel, dim3(ctas), dim3(kThreadsD), args, smem, s);
  if (e == cudaSuccess) count_launch();
  return e;
}

}  // namespace dv3
