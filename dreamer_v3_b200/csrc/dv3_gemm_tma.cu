// TMA-fed bf16 tcgen05 GEMM for sm_100a: both operands already exist in bf16, K-major, in global memory
// (activations written in bf16 by their producer kernels, weights as the engine's bf16 shadows), so operand
// tiles move global -> SWIZZLE_128B shared memory with cp.async.bulk.tensor and no thread touches them.
//
//   C[M,N] (+)= A16[M,K] * B16[N,K]^T (+ bias[N])          fp32 accumulate in TMEM, fp32 output
//
// Persistent, warp-specialised: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane, owns the TMEM
// allocation), warps 2..5 = epilogue. The accumulator is double buffered in TMEM (2 x BN columns), so the epilogue
// of work item i overlaps the main loop of item i+1. Work items are (tile, k-split) pairs, strided over the grid;
// consecutive items share the A row block, so concurrent CTAs hit the same A tile in L2.
// Out-of-bounds rows / k are zero-filled by the TMA unit, which is what makes ragged M, N and K correct. The
// epilogue goes back through the TMA unit too (swizzled 32x32 staging blocks -> cp.async.bulk.tensor store, or
// cp.reduce...add for accumulate / split-K), which clips at M / N and keeps the store loop to ~50 instructions.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "dv3_common.h"

namespace dv3 {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 per k-block = one 128-byte swizzle row
constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
__host__ __device__ constexpr int epi_bufs(int bn) { return bn >= 256 ? 1 : 2; }
// operand ring depth = what is left of the 227 KB after the epilogue staging blocks, bias rows, barriers and alignment slack
__host__ __device__ constexpr int tma_stages(int bn) {
  return (227 * 1024 - 1024 - 4 * epi_bufs(bn) * 4096 - 4 * bn * 4 - 512) / (BM * 128 + bn * 128);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, SWIZZLE_128B: 64 MN elements (128 B) x 8 k per swizzle atom; k-groups 1024 B apart (SBO), the next 64 MN
// elements 8192 B apart (LBO) -- exactly how two [64 k x 64 mn] TMA boxes land back to back
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n, bool mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 24)) {  // never hang the GPU on a lost arrival
      printf("dv3 gemm_tma: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
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

template <int BN>
struct TmaSmem {
  static constexpr int S = tma_stages(BN);
  unsigned char a[S][BM * 128];
  unsigned char b[S][BN * 128];
  static constexpr int EB = epi_bufs(BN);
  unsigned char epi[kEpiWarps][EB][4096];   // per warp: 32x32 fp32 staging blocks (SWIZZLE_128B rows) for the TMA stores
  float bias_s[kEpiWarps][BN];
  uint64_t full[S], empty[S], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct TmaGemmArgs {
  float* C; int ldc;
  int M, N, K;
  const float* bias;
  float beta;
  int tiles_n, tiles, splits, k_per_split, use_atomic;
};

template <int BN, bool MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapC, TmaGemmArgs g) {
  extern __shared__ unsigned char smem_raw[];
  using SmemT = TmaSmem<BN>;
  constexpr int S = SmemT::S;
  SmemT& sm = *reinterpret_cast<SmemT*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwork = g.tiles * g.splits;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
    for (int s = 0; s < S; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // 2 accumulators of BN fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
  // tail of the previous kernel; nothing below may run before its results are visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      constexpr uint32_t kBytes = BM * 128 + BN * 128;
      uint32_t it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int tile = w / g.splits, split = w - tile * g.splits;
        const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN;
        const int kbeg = split * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
        for (int k0 = kbeg; k0 < kend; k0 += BK, ++it) {
          const int s = it % S;
          mbar_wait(&sm.empty[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&sm.full[s], kBytes);
          if (!MN) {
            tma_load_2d(smem_u32(sm.a[s]), &mapA, k0, m0, &sm.full[s]);
            tma_load_2d(smem_u32(sm.b[s]), &mapB, k0, n0, &sm.full[s]);
          } else {   // operands stored [k][mn]: boxes of 64 mn x 64 k, 8 KB each
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(smem_u32(sm.a[s]) + i * 8192, &mapA, m0 + 64 * i, k0, &sm.full[s]);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(smem_u32(sm.b[s]) + i * 8192, &mapB, n0 + 64 * i, k0, &sm.full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN, MN);
      uint32_t it = 0, acc_it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++acc_it) {
        const int split = w % g.splits;
        const int kbeg = split * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
        const uint32_t as = acc_it & 1;
        mbar_wait(&sm.acc_empty[as], ((acc_it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem + as * BN;
        bool first = true;
        for (int k0 = kbeg; k0 < kend; k0 += BK, ++it) {
          const int s = it % S;
          mbar_wait(&sm.full[s], (it / S) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(sm.a[s]), b_addr = smem_u32(sm.b[s]);
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t da = MN ? make_smem_desc_mn(a_addr + j * 2048) : make_smem_desc(a_addr + j * 32);
            const uint64_t db = MN ? make_smem_desc_mn(b_addr + j * 2048) : make_smem_desc(b_addr + j * 32);
            const uint32_t accum = (first && j == 0) ? 0u : 1u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
                : "memory");
          }
          first = false;
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.empty[s])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sm.acc_full[as])) : "memory");
      }
    }
  } else {
    // ---------------- epilogue: warp -> TMEM lane quarter (warp % 4), all BN columns ----------------
    // thread = accumulator row: tcgen05.ld 32 columns -> (+bias) -> 8 x st.shared.v4 into a swizzled 32x32 block ->
    // one lane hands the block to the TMA unit (plain store, or reduce-add for accumulate / split-K), which also
    // clips rows / columns beyond M / N. Two staging blocks per warp keep one store in flight.
    const int q = warp & 3, ew = warp - 2;
    const uint32_t epi0 = smem_u32(sm.epi[ew][0]);
    const bool reduce = g.use_atomic || g.beta != 0.f;
    uint32_t acc_it = 0, blk = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++acc_it) {
      const int tile = w / g.splits, split = w - tile * g.splits;
      const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN;
      const bool has_bias = g.bias != nullptr && split == 0;
      if (has_bias) {
        __syncwarp();
        for (int j = lane; j < BN; j += 32) sm.bias_s[ew][j] = (n0 + j < g.N) ? g.bias[n0 + j] : 0.f;
        __syncwarp();
      }
      const uint32_t as = acc_it & 1;
      mbar_wait(&sm.acc_full[as], (acc_it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // Last work item of this CTA: every MMA has completed and the producer has nothing left to load, so the B ring
      // is free -- each 32-column block gets its own staging block there and no store ever waits for an earlier one.
      const bool last_item = w + (int)gridDim.x >= nwork;
      const uint32_t ring = smem_u32(sm.b[0]) + (uint32_t)ew * (BN / 32) * 4096;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32, ++blk) {
        if (n0 + c0 >= g.N) break;
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        // the staging block about to be reused must have been read by the TMA unit before it is overwritten
        const uint32_t blk_base = last_item ? ring + (uint32_t)(c0 / 32) * 4096 : epi0 + (blk & (SmemT::EB - 1)) * 4096;
        if (!last_item) {
          if (lane == 0) {
            if (SmemT::EB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const uint32_t buf = blk_base + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          if (has_bias) {
            const float4 b = *reinterpret_cast<const float4*>(&sm.bias_s[ew][c0 + 4 * j]);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)((j ^ (lane & 7)) << 4)), "f"(v.x), "f"(v.y),
                       "f"(v.z), "f"(v.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const uint32_t src = blk_base;
          if (reduce)
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&mapC), "r"(n0 + c0), "r"(m0 + q * 32), "r"(src) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&mapC), "r"(n0 + c0), "r"(m0 + q * 32), "r"(src) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&sm.acc_empty[as]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may go away once it has been read; kernel end covers the writes
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN));
}

// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint), so that libdv3.so has no link-time
// dependency on libcuda.so.1 and still loads on a machine without a driver (CPU-side tests of the C ABI).
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

// 2-D bf16 tensor map: dims {K (contiguous), rows}, box {64, box_rows}, SWIZZLE_128B, zero fill out of bounds
bool make_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp32 output map: dims {N, M}, box 32 x 32 (128-byte rows), SWIZZLE_128B staging blocks
bool make_map_c(CUtensorMap* map, float* ptr, int M, int N, int ldc) {
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ldc * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, bool MN>
bool launch_tma(cudaStream_t s, const GemmArgs& g) {
  CUtensorMap mapA, mapB, mapC;
  if (!MN) {
    if (!make_map(&mapA, g.Ah, g.M, g.K, g.ldah, BM)) return false;
    if (!make_map(&mapB, g.Bh, g.N, g.K, g.ldbh, BN)) return false;
  } else {   // [K rows, M (or N) contiguous]: boxes of 64 k x 64 mn
    if (!make_map(&mapA, g.Ah, g.K, g.M, g.ldah, 64)) return false;
    if (!make_map(&mapB, g.Bmn, g.K, g.N, g.ldbmn, 64)) return false;
  }
  if (!make_map_c(&mapC, g.C, g.M, g.N, g.ldc)) return false;
  TmaGemmArgs a;
  a.C = g.C; a.ldc = g.ldc; a.M = g.M; a.N = g.N; a.K = g.K; a.bias = g.bias; a.beta = g.beta;
  a.tiles_n = cdiv(g.N, BN);
  a.tiles = a.tiles_n * cdiv(g.M, BM);
  int splits = 1;
  if (a.tiles <= 48 && g.K >= 12 * BK) {   // (K = 768, the GRU backward products of the XS dream: 7.3 us unsplit as a graph node, ~5 split)
    splits = min(148 / a.tiles, g.K / (4 * BK));
    splits = max(1, min(splits, 32));
  }
  a.k_per_split = cdiv(cdiv(g.K, splits), BK) * BK;
  a.splits = cdiv(g.K, a.k_per_split);
  a.use_atomic = a.splits > 1;
  if (a.use_atomic && g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
  const size_t smem = sizeof(TmaSmem<BN>) + 1024;
  static_assert(sizeof(TmaSmem<BN>) + 1024 <= 227 * 1024, "TmaSmem exceeds the 227 KB dynamic shared memory limit");
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(gemm_tma_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  const int grid = min(a.tiles * a.splits, 148);
  launch_pdl(gemm_tma_kernel<BN, MN>, dim3(grid), dim3(kThreads), smem, s, mapA, mapB, mapC, a);
  count_launch();
  return true;
}

}  // namespace

bool gemm_tma_eligible(const GemmArgs& g) {
  if ((g.ldc % 4 != 0) || ((uintptr_t)g.C % 16 != 0) || g.N < 16 || g.K < 32) return false;
  if (g.ta && !g.tb)   // weight gradient X^T dY: both operands bf16 copies as stored ([K, M] and [K, N] row-major)
    return g.Ah != nullptr && g.Bmn != nullptr && (g.ldah % 8 == 0) && (g.ldbmn % 8 == 0) && ((uintptr_t)g.Ah % 16 == 0) &&
           ((uintptr_t)g.Bmn % 16 == 0) && g.M >= 64 && g.K >= 512;
  return g.Ah != nullptr && g.Bh != nullptr && !g.ta && (g.ldah % 8 == 0) && (g.ldbh % 8 == 0) &&
         ((uintptr_t)g.Ah % 16 == 0) && ((uintptr_t)g.Bh % 16 == 0) && g.M >= 128;
}

template <bool MN>
static bool gemm_tma_sel(cudaStream_t s, const GemmArgs& g) {
  if (const char* f = std::getenv("DV3_TC_BN")) {  // probing aid
    const int bn = std::atoi(f);
    if (bn == 256) return launch_tma<256, MN>(s, g);
    if (bn == 128) return launch_tma<128, MN>(s, g);
    if (bn == 64) return launch_tma<64, MN>(s, g);
    if constexpr (!MN) { if (bn == 32) return launch_tma<32, false>(s, g); }
  }
  const long long mt = cdiv(g.M, BM);
  // (measured: 32-column tiles without split-K are slower than 64-column tiles with split-K for the [1024 x 256 x 1280] steps of
  //  the dream - 9.7 vs 6.8 us per node in the replayed graph - so they are only reachable through DV3_TC_BN=32)
  if (g.N > 128 && mt * cdiv(g.N, 256) >= 96) return launch_tma<256, MN>(s, g);
  if (g.N > 64 && mt * cdiv(g.N, 128) >= 64) return launch_tma<128, MN>(s, g);
  if (g.N > 128 && mt * cdiv(g.N, 64) < 16) return launch_tma<256, MN>(s, g);
  return launch_tma<64, MN>(s, g);
}

bool gemm_tma(cudaStream_t s, const GemmArgs& g) {
  return (g.ta && !g.tb) ? gemm_tma_sel<true>(s, g) : gemm_tma_sel<false>(s, g);
}

}  // namespace dv3
