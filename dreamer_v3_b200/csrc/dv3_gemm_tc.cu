// bf16 tensor-core GEMM for sm_100a on fp32 operands: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the accumulator in
// TMEM. This is the fallback of the tensor-core mode for contractions whose operands have no bf16 copy in global memory
// (255-wide logits and their gradients, a few small wgrads); everything else runs on the TMA-fed kernel in dv3_gemm_tma.cu.
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N])      same contract as gemm_f32 (fp32 in / fp32 out)
//
// The engine keeps fp32 master tensors, so the operands are converted on the way into shared memory: every
// CTA (256 threads) loads fp32 tiles with coalesced 128-bit loads, rounds to bf16 and writes them in the
// canonical K-major SWIZZLE_128B UMMA layout (8-row x 128-byte swizzle atoms); operands whose contraction
// index is not contiguous in memory (B of a forward layer, both operands of a weight gradient) are transposed
// by that fill. One elected thread issues the MMAs (4 x K=16 per 64-wide k-block) and commits them to an
// mbarrier; the fill of the next k-block overlaps the tensor-core work of the current one (2-stage ring).
// Epilogue: tcgen05.ld 32 lanes x 32 columns per warp -> bias / accumulate -> fp32 global stores.
#include <cuda_bf16.h>

#include <cstdlib>

#include "dv3_common.h"

namespace dv3 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kStages = 2;
constexpr int kTcThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, k) inside a K-major SWIZZLE_128B tile (k < 64): 1024-byte atoms of 8 rows
__device__ __forceinline__ uint32_t sw128_chunk_off(int r, int kchunk) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((kchunk ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- fills: fp32 global -> registers (all loads of a k-block in flight at once) -> bf16 K-major swizzled smem ----
// A tile of ROWS rows x 64 k is cut into ROWS*8 tasks of 8 consecutive k; thread t owns tasks t, t+256, ...
template <int ROWS, bool TRANSPOSE>
struct Fill {
  static constexpr int kTasks = ROWS * 8 / kTcThreads;
  float v[kTasks][8];

  // TRANSPOSE = false: source is k-contiguous, src(r, k) = src[r*ld + k]
  // TRANSPOSE = true : source is row-contiguous, src(r, k) = src[k*ld + r] (transposed by the fill)
  __device__ __forceinline__ void load(const float* __restrict__ src, int ld, int r0, int rmax, int k0, int kmax, bool vec) {
#pragma unroll
    for (int i = 0; i < kTasks; ++i) {
      const int task = threadIdx.x + i * kTcThreads;
      if (!TRANSPOSE) {
        const int r = task >> 3, kc = task & 7;
        const int gr = r0 + r, gk = k0 + kc * 8;
        const float* p = src + (size_t)gr * ld + gk;
        if (gr < rmax && vec && gk + 7 < kmax) {
          const float4 a = *reinterpret_cast<const float4*>(p);
          const float4 b = *reinterpret_cast<const float4*>(p + 4);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = (gr < rmax && gk + j < kmax) ? p[j] : 0.f;
        }
      } else if (vec && (kTasks % 4 == 0)) {
        // quad task q = (4 consecutive rows, 8 k): 8 x LDG.128 along the contiguous row index, transposed in registers.
        // kTasks is a multiple of 4 here; sub-task i of quad (i / 4) is row (i % 4) of that quad.
        if ((i & 3) == 0) {
          const int quad = threadIdx.x + (i >> 2) * kTcThreads;
          const int r4 = (quad % (ROWS / 4)) * 4, kc = quad / (ROWS / 4);
          const int gr = r0 + r4, gk = k0 + kc * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk + j < kmax) {
              const float* p = src + (size_t)(gk + j) * ld + gr;
              if (gr + 3 < rmax) t = *reinterpret_cast<const float4*>(p);
              else {
                if (gr + 0 < rmax) t.x = p[0];
                if (gr + 1 < rmax) t.y = p[1];
                if (gr + 2 < rmax) t.z = p[2];
              }
            }
            v[i + 0][j] = t.x; v[i + 1][j] = t.y; v[i + 2][j] = t.z; v[i + 3][j] = t.w;
          }
        }
      } else {
        const int r = task % ROWS, kc = task / ROWS;  // consecutive threads -> consecutive r: coalesced reads
        const int gr = r0 + r, gk = k0 + kc * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = (gr < rmax && gk + j < kmax) ? src[(size_t)(gk + j) * ld + gr] : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(uint32_t tile, bool vec) const {  // tile = shared-space address
#pragma unroll
    for (int i = 0; i < kTasks; ++i) {
      const int task = threadIdx.x + i * kTcThreads;
      int r, kc;
      if (!TRANSPOSE) { r = task >> 3; kc = task & 7; }
      else if (vec && (kTasks % 4 == 0)) {
        const int quad = threadIdx.x + (i >> 2) * kTcThreads;
        r = (quad % (ROWS / 4)) * 4 + (i & 3); kc = quad / (ROWS / 4);
      } else { r = task % ROWS; kc = task / ROWS; }
      uint4 o;
      o.x = pack_bf16(v[i][0], v[i][1]); o.y = pack_bf16(v[i][2], v[i][3]);
      o.z = pack_bf16(v[i][4], v[i][5]); o.w = pack_bf16(v[i][6], v[i][7]);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + sw128_chunk_off(r, kc)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
  }
};

// ---- descriptors ---------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(=1)<<16 | SBO(=1024>>4)<<32 | version 1 | layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = f32 (1<<4), A = B = bf16 (1<<7, 1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 22)) {  // never hang the GPU on a lost arrival: fail loudly instead
      printf("dv3 gemm_tc: mbarrier wait timed out (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
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

template <int BN, int STAGES>
struct TcSmem {
  unsigned char a[STAGES][BM * 128];   // 16 KB per stage
  unsigned char b[STAGES][BN * 128];
  uint64_t bar[STAGES];
  uint32_t tmem_base;
};

template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(GemmArgs g, int k_per_split, int vecA, int vecB, int use_atomic) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  constexpr int STAGES = kStages;
  using SmemT = TcSmem<BN, STAGES>;
  SmemT& sm = *reinterpret_cast<SmemT*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&sm.bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates BN fp32 accumulator columns of tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch: prologue above overlaps
  asm volatile("griddepcontrol.wait;" ::: "memory");                 // the previous kernel's tail; its results are visible from here
  const uint32_t idesc = make_idesc(BN);

  // issue the MMAs of one staged k-block (one elected thread) and commit them to the stage's mbarrier
  auto issue_mma = [&](int kb, int s) {
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(sm.a[s]), b_addr = smem_u32(sm.b[s]);
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        const uint64_t da = make_smem_desc(a_addr + j * 32);  // +16 bf16 along K inside the 128-byte swizzle row
        const uint64_t db = make_smem_desc(b_addr + j * 32);
        const uint32_t accum = (kb > 0 || j > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
      }
      // arrives on the stage's mbarrier once every MMA issued so far has completed
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.bar[s])) : "memory");
    }
  };
  {
    Fill<BM, TA> fa;
    Fill<BN, !TB> fb;
    fa.load(g.A, g.lda, m0, g.M, kbeg, kend, vecA);
    fb.load(g.B, g.ldb, n0, g.N, kbeg, kend, vecB);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      if (kb >= kStages) mbar_wait(&sm.bar[s], (uint32_t)(((kb / kStages) - 1) & 1));  // MMAs that read this stage are done
      fa.store(smem_u32(sm.a[s]), vecA);
      fb.store(smem_u32(sm.b[s]), vecB);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
      __syncthreads();
      issue_mma(kb, s);
      if (kb + 1 < nkb) {  // next k-block's global loads fly while the tensor core works on this one
        const int k0 = kbeg + (kb + 1) * BK;
        fa.load(g.A, g.lda, m0, g.M, k0, kend, vecA);
        fb.load(g.B, g.ldb, n0, g.N, k0, kend, vecB);
      }
    }
  }
  // wait for the last commit (it covers all earlier MMAs)
  {
    const int last = nkb - 1;
    mbar_wait(&sm.bar[last % STAGES], (uint32_t)((last / STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: warp w reads TMEM lanes (= tile rows) 32(w%4) .. +31, column half w/4 ----
  const int q = warp & 3;                 // TMEM lane quarter this warp may read
  const int cbeg = (warp >> 2) * (BN / 2), cend = cbeg + BN / 2;
  // all MMAs have completed, so the operand stages are free: 4 KB per warp to transpose 32x32 blocks so that
  // the global stores are coalesced along N (XOR swizzle keeps both phases bank-conflict free)
  const uint32_t epi = smem_u32(sm.a[0]) + warp * 4096;
#pragma unroll 1
  for (int c0 = cbeg; c0 < cend; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
#pragma unroll
    for (int j = 0; j < 32; ++j)                                                        // thread = row
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(epi + (uint32_t)((lane * 32 + (j ^ lane)) * 4)), "r"(r[j]) : "memory");
    __syncwarp();
    const int gn = n0 + c0 + lane;                                                      // thread = column
    const float bv = (g.bias != nullptr && blockIdx.z == 0 && gn < g.N) ? g.bias[gn] : 0.f;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const int gm = m0 + q * 32 + i;
      if (gm < g.M && gn < g.N) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(epi + (uint32_t)((i * 32 + (lane ^ i)) * 4)) : "memory");
        v += bv;
        float* dst = g.C + (size_t)gm * g.ldc + gn;
        if (use_atomic) atomicAdd(dst, v);
        else if (g.beta != 0.f) *dst += v;
        else *dst = v;
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
}


// =====================================================================================================================
// Warp-specialised variant for a bf16-shadow B operand (weights): two producer groups of 4 warps alternate k-blocks
// (global loads -> convert -> swizzled st.shared -> proxy fence -> mbarrier arrive), a ninth warp only issues
// tcgen05.mma and commits completions back to the producers ("stage empty"). There is no CTA-wide barrier in the
// main loop, so a group's fence only waits for its own loads while the other group's loads stay in flight.
// =====================================================================================================================
constexpr int kWsStages = 4;
constexpr int kWsGroup = 128;                 // threads per producer group
constexpr int kWsThreads = 2 * kWsGroup + 32;  // + the MMA warp

template <int BN>
struct WsSmem {
  unsigned char a[kWsStages][BM * 128];
  unsigned char b[kWsStages][BN * 128];
  uint64_t full[kWsStages], empty[kWsStages], done;
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, bool TA>
__global__ void __launch_bounds__(kWsThreads, 1) gemm_tc_ws_kernel(GemmArgs g, int k_per_split, int vecA, int use_atomic) {
  extern __shared__ unsigned char smem_raw[];
  using SmemT = WsSmem<BN>;
  SmemT& sm = *reinterpret_cast<SmemT*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < kWsStages; ++s) { mbar_init(&sm.full[s], kWsGroup); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch: prologue above overlaps
  asm volatile("griddepcontrol.wait;" ::: "memory");                 // the previous kernel's tail; its results are visible from here

  if (warp == 8) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kWsStages;
        mbar_wait(&sm.full[s], (uint32_t)((kb / kWsStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sm.a[s]), b_addr = smem_u32(sm.b[s]);
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t da = make_smem_desc(a_addr + j * 32);
          const uint64_t db = make_smem_desc(b_addr + j * 32);
          const uint32_t accum = (kb > 0 || j > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.done)) : "memory");
    }
  } else {
    // ---------------- producers: group 0 takes even k-blocks, group 1 odd ones ----------------
    const int group = warp >> 2, t = tid & (kWsGroup - 1);
    const __nv_bfloat16* Bh = reinterpret_cast<const __nv_bfloat16*>(g.Bh);
    constexpr int kTasksA = BM * 8 / kWsGroup;   // 8
    constexpr int kTasksB = BN * 8 / kWsGroup;   // 4 / 8 / 16
    for (int kb = group; kb < nkb; kb += 2) {
      const int s = kb % kWsStages, use = kb / kWsStages;
      const int k0 = kbeg + kb * BK;
      // A: fp32 -> registers
      float va[kTasksA][8];
#pragma unroll
      for (int i = 0; i < kTasksA; ++i) {
        const int task = t + i * kWsGroup;
        if (!TA) {
          const int r = task >> 3, kc = task & 7;
          const int gr = m0 + r, gk = k0 + kc * 8;
          const float* p = g.A + (size_t)gr * g.lda + gk;
          if (gr < g.M && vecA && gk + 7 < kend) {
            const float4 x = *reinterpret_cast<const float4*>(p);
            const float4 y = *reinterpret_cast<const float4*>(p + 4);
            va[i][0] = x.x; va[i][1] = x.y; va[i][2] = x.z; va[i][3] = x.w; va[i][4] = y.x; va[i][5] = y.y; va[i][6] = y.z; va[i][7] = y.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) va[i][j] = (gr < g.M && gk + j < kend) ? p[j] : 0.f;
          }
        } else {
          const int r = task % BM, kc = task / BM;
          const int gr = m0 + r, gk = k0 + kc * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) va[i][j] = (gr < g.M && gk + j < kend) ? g.A[(size_t)(gk + j) * g.lda + gr] : 0.f;
        }
      }
      // B: bf16 shadow -> registers
      uint4 vb[kTasksB];
#pragma unroll
      for (int i = 0; i < kTasksB; ++i) {
        const int task = t + i * kWsGroup;
        const int r = task >> 3, kc = task & 7;
        const int gr = n0 + r, gk = k0 + kc * 8;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (gr < g.N && gk < kend) {
          const __nv_bfloat16* p = Bh + (size_t)gr * g.ldbh + gk;
          if (gk + 7 < kend) x = *reinterpret_cast<const uint4*>(p);
          else {
            unsigned short h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = (gk + j < kend) ? reinterpret_cast<const unsigned short*>(p)[j] : (unsigned short)0;
            x.x = h[0] | ((uint32_t)h[1] << 16); x.y = h[2] | ((uint32_t)h[3] << 16);
            x.z = h[4] | ((uint32_t)h[5] << 16); x.w = h[6] | ((uint32_t)h[7] << 16);
          }
        }
        vb[i] = x;
      }
      if (use > 0) mbar_wait(&sm.empty[s], (uint32_t)((use - 1) & 1));   // the MMAs that read this stage have completed
      const uint32_t ta_ = smem_u32(sm.a[s]), tb_ = smem_u32(sm.b[s]);
#pragma unroll
      for (int i = 0; i < kTasksA; ++i) {
        const int task = t + i * kWsGroup;
        const int r = TA ? task % BM : task >> 3, kc = TA ? task / BM : task & 7;
        const uint32_t x = pack_bf16(va[i][0], va[i][1]), y = pack_bf16(va[i][2], va[i][3]);
        const uint32_t z = pack_bf16(va[i][4], va[i][5]), w = pack_bf16(va[i][6], va[i][7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ta_ + sw128_chunk_off(r, kc)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
      }
#pragma unroll
      for (int i = 0; i < kTasksB; ++i) {
        const int task = t + i * kWsGroup;
        const int r = task >> 3, kc = task & 7;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb_ + sw128_chunk_off(r, kc)), "r"(vb[i].x), "r"(vb[i].y), "r"(vb[i].z), "r"(vb[i].w) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&sm.full[s]);
    }
    // ---------------- epilogue (the 8 producer warps) ----------------
    mbar_wait(&sm.done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int cbeg = (warp >> 2) * (BN / 2), cend = cbeg + BN / 2;
    const uint32_t epi = smem_u32(sm.a[0]) + warp * 4096;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
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
#pragma unroll
      for (int j = 0; j < 32; ++j)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(epi + (uint32_t)((lane * 32 + (j ^ lane)) * 4)), "r"(r[j]) : "memory");
      __syncwarp();
      const int gn = n0 + c0 + lane;
      const float bv = (g.bias != nullptr && blockIdx.z == 0 && gn < g.N) ? g.bias[gn] : 0.f;
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int gm = m0 + q * 32 + i;
        if (gm < g.M && gn < g.N) {
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(epi + (uint32_t)((i * 32 + (lane ^ i)) * 4)) : "memory");
          v += bv;
          float* dst = g.C + (size_t)gm * g.ldc + gn;
          if (use_atomic) atomicAdd(dst, v);
          else if (g.beta != 0.f) *dst += v;
          else *dst = v;
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
}

template <int BN>
void launch_tc(cudaStream_t s, const GemmArgs& g) {
  const int gx = cdiv(g.N, BN), gy = cdiv(g.M, BM);
  const long long tiles = (long long)gx * gy;
  int splits = 1;
  if (tiles <= 48 && g.K >= 16 * BK) {
    splits = (int)min((long long)(148 / tiles), (long long)(g.K / (8 * BK)));
    if (splits < 1) splits = 1;
    if (splits > 148) splits = 148;   // (conv weight gradients: a 48 x 32 output over 10^6 pixel rows is one tile - every SM takes a K slice)
  }
  const int kps = cdiv(cdiv(g.K, splits), BK) * BK;
  splits = cdiv(g.K, kps);
  const int use_atomic = splits > 1;
  if (use_atomic && g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
  const int vA = ((uintptr_t)g.A % 16 == 0) && (g.lda % 4 == 0);
  const int vB = ((uintptr_t)g.B % 16 == 0) && (g.ldb % 4 == 0);
  const bool bh = g.Bh != nullptr && (g.ldbh % 8 == 0) && ((uintptr_t)g.Bh % 16 == 0);
  const size_t smem = sizeof(TcSmem<BN, kStages>) + 1024;
  dim3 grid(gx, gy, splits), block(kTcThreads);
#define DV3_TC_LAUNCH(TA_, TB_)                                                                                   \
  do {                                                                                                            \
    static bool configured = false;                                                                               \
    if (!configured) {                                                                                            \
      cudaFuncSetAttribute(gemm_tc_kernel<BN, TA_, TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      configured = true;                                                                                          \
    }                                                                                                             \
    launch_pdl(gemm_tc_kernel<BN, TA_, TB_>, grid, block, smem, s, g, kps, vA, vB, use_atomic);                    \
  } while (0)
  if (!bh) {
    if (!g.ta && !g.tb) DV3_TC_LAUNCH(false, false);
    else if (!g.ta && g.tb) DV3_TC_LAUNCH(false, true);
    else if (g.ta && !g.tb) DV3_TC_LAUNCH(true, false);
    else DV3_TC_LAUNCH(true, true);
  }
  if (bh) {
    const size_t smem_ws = sizeof(WsSmem<BN>) + 1024;
    if (!g.ta) {
      static bool cfgd = false;
      if (!cfgd) { cudaFuncSetAttribute(gemm_tc_ws_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws); cfgd = true; }
      launch_pdl(gemm_tc_ws_kernel<BN, false>, grid, dim3(kWsThreads), smem_ws, s, g, kps, vA, use_atomic);
    } else {
      static bool cfgd = false;
      if (!cfgd) { cudaFuncSetAttribute(gemm_tc_ws_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws); cfgd = true; }
      launch_pdl(gemm_tc_ws_kernel<BN, true>, grid, dim3(kWsThreads), smem_ws, s, g, kps, vA, use_atomic);
    }
  }
#undef DV3_TC_LAUNCH
  count_launch();
}

}  // namespace

namespace {
__global__ void weight_shadow_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out_t,
                                     __nv_bfloat16* __restrict__ out_n, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    const float v = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (out_n != nullptr && r < rows && c < cols) out_n[(size_t)r * cols + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (out_t != nullptr && r2 < rows && c2 < cols) out_t[(size_t)c2 * rows + r2] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
}  // namespace

namespace {
__global__ void weight_shadow_batched_kernel(const ShadowDesc* __restrict__ descs, int n) {
  __shared__ float tile[32][33];
  // find the matrix this block belongs to (descs sorted by tile_begin)
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ShadowDesc d = descs[lo];
  const int local = blockIdx.x - d.tile_begin;
  const int tiles_x = (d.cols + 31) / 32;
  const int bx = local % tiles_x, by = local / tiles_x;
  __nv_bfloat16* out_t = reinterpret_cast<__nv_bfloat16*>(d.out_t);
  __nv_bfloat16* out_n = reinterpret_cast<__nv_bfloat16*>(d.out_n);
  const int c = bx * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = by * 32 + i;
    const float v = (r < d.rows && c < d.cols) ? d.in[(size_t)r * d.cols + c] : 0.f;
    tile[i][threadIdx.x] = v;
    if (r < d.rows && c < d.cols) out_n[(size_t)r * d.cols + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int r2 = by * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = bx * 32 + i;
    if (r2 < d.rows && c2 < d.cols) out_t[(size_t)c2 * d.ld_t + r2] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
}  // namespace

void weight_shadow_batched(cudaStream_t s, const ShadowDesc* descs_dev, int n, int total_tiles) {
  if (n <= 0 || total_tiles <= 0) return;
  weight_shadow_batched_kernel<<<total_tiles, dim3(32, 8), 0, s>>>(descs_dev, n);
  count_launch();
}

void weight_shadow_bf16(cudaStream_t s, const float* in, void* out_t, void* out_n, int rows, int cols) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
  weight_shadow_kernel<<<grid, block, 0, s>>>(in, reinterpret_cast<__nv_bfloat16*>(out_t), reinterpret_cast<__nv_bfloat16*>(out_n), rows, cols);
  count_launch();
}

// Large contractions go to the tensor cores; small-M scan steps and tiny shapes stay on the exact fp32 path.
bool gemm_tc_eligible(const GemmArgs& g) {
  if (g.N < 16 || g.K < 32 || (long long)g.M * g.N * g.K < (1LL << 22)) return false;
  return g.M >= 128 || (g.M >= 16 && g.K >= 4096);  // skinny outputs only when the contraction is long (weight gradients)
}

void gemm_tc(cudaStream_t s, const GemmArgs& g) {
  if (const char* f = std::getenv("DV3_TC_BN")) {  // probing aid
    const int bn = std::atoi(f);
    if (bn == 256) return launch_tc<256>(s, g);
    if (bn == 128) return launch_tc<128>(s, g);
    if (bn == 64) return launch_tc<64>(s, g);
  }
  // widest N tile that still yields enough CTAs to occupy the GPU (split-K costs an atomic epilogue, so the
  // tile shape is the first lever for parallelism)
  const long long mt = cdiv(g.M, BM);
  if (g.N > 128 && mt * cdiv(g.N, 256) >= 96) launch_tc<256>(s, g);
  else if (g.N > 64 && mt * cdiv(g.N, 128) >= 64) launch_tc<128>(s, g);
  else if (g.N > 128 && mt * cdiv(g.N, 64) < 16) launch_tc<256>(s, g);  // tiny grid either way: fewer, fatter CTAs + split-K
  else launch_tc<64>(s, g);
}

}  // namespace dv3
