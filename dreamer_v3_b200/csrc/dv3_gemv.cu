// Weight-streaming contractions for the 16 replay rows of the observe scan at the large model sizes (XL: GRU 4096,
// recurrent kernel 4096 x 12288 = 201 MB fp32), where the scan runs per op and every step re-reads its weights from HBM:
//
//     C[M <= 16, N] (+)= A[M, K] * op(B)[K, N] (+ bias)          fp32, exact in every compute mode
//
// At 16 rows a weight element is used 16 times: 8 flop per byte, so the bound is the HBM stream of B and the job of the
// kernel is to keep ~40 KB of 128-bit loads in flight per SM and touch each weight once. Two layouts:
//   * gemv16_n_kernel (tb = 0, B [K, N], N contiguous - forward layers): a CTA owns 128 columns x a K-chunk; its 8 warps
//     split the chunk, a lane owns 4 columns (one LDG.128 per weight row, 4 rows in flight), the 16 activations of a k are
//     four broadcast LDS.128 from the staged A chunk; 64 accumulators per lane; the warps' partial tiles are reduced
//     through shared memory and stored (or added to C with atomics when K is split over CTAs).
//   * gemv16_t_kernel (tb = 1, B [N, K], K contiguous - input gradients, B = W): a warp owns 4 output columns (4 weight
//     rows), lanes stride K with LDG.128, activations come from the staged A chunk as LDS.128 shared by the 4 columns;
//     64 accumulators per lane, butterfly reduction at the end.
// The tiled SIMT kernel ran [16 x 4096] x [4096 x 12288] in 300 us (0.67 TB/s), the tcgen05 kernel on bf16 shadows in
// 123 us; the stream bound is 31 us.
#include "dv3_common.h"

namespace dv3 {
namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_ROWS = 16;

// ---- B [K, N] ------------------------------------------------------------------------------------------------------
constexpr int GN_COLS = 128;     // columns per CTA
constexpr int GN_KMAX = 1024;    // K-chunk staged per CTA (64 KB of A^T)

constexpr int GN_LD = 20;        // row stride of the staged A^T chunk (16 rows + pad: conflict-free float4 stores)
constexpr int GN_RED = (GV_THREADS / 32) * GV_ROWS * (GN_COLS + 4);   // floats of the cross-warp reduction buffer (aliases the chunk)

__global__ void __launch_bounds__(GV_THREADS) gemv16_n_kernel(GemmArgs g, int kchunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* at = reinterpret_cast<float*>(smem_raw);   // A chunk transposed: at[k * GN_LD + row]
  float* red = at;                                   // [warp][row][GN_COLS + 4], reused after the main loop
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * GN_COLS, k0 = blockIdx.y * kchunk, kn = min(kchunk, g.K - k0);
  for (int k = threadIdx.x; k < kn; k += GV_THREADS) {   // a thread gathers the 16 rows of its k (coalesced along k)
    float v[GV_ROWS];
#pragma unroll
    for (int r = 0; r < GV_ROWS; ++r) v[r] = r < g.M ? g.A[(size_t)r * g.lda + k0 + k] : 0.f;
    float4* dst = reinterpret_cast<float4*>(at + (size_t)k * GN_LD);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]); dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    dst[2] = make_float4(v[8], v[9], v[10], v[11]); dst[3] = make_float4(v[12], v[13], v[14], v[15]);
  }
  __syncthreads();
  float acc[GV_ROWS][4];
#pragma unroll
  for (int r = 0; r < GV_ROWS; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
  const int c = n0 + 4 * lane;
  const bool vec = (c + 3 < g.N) && (g.ldb % 4 == 0) && (((uintptr_t)g.B & 15) == 0);
  const float* bp = g.B + (size_t)k0 * g.ldb + c;
  const int per = (kn + 7) / 8, kb = warp * per, ke = min(kn, kb + per);
  auto fma_row = [&](const float4& w, int k) {
    const float4* a4 = reinterpret_cast<const float4*>(at + (size_t)k * GN_LD);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = a4[q];
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[4 * q + i][0] = fmaf(av[i], w.x, acc[4 * q + i][0]);
        acc[4 * q + i][1] = fmaf(av[i], w.y, acc[4 * q + i][1]);
        acc[4 * q + i][2] = fmaf(av[i], w.z, acc[4 * q + i][2]);
        acc[4 * q + i][3] = fmaf(av[i], w.w, acc[4 * q + i][3]);
      }
    }
  };
  int k = kb;
  if (vec) {
    for (; k + 4 <= ke; k += 4) {   // four weight rows in flight per lane
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = __ldcs(reinterpret_cast<const float4*>(bp + (size_t)(k + j) * g.ldb));
#pragma unroll
      for (int j = 0; j < 4; ++j) fma_row(w[j], k + j);
    }
  }
  for (; k < ke; ++k) {
    float4 w;
    const float* row = bp + (size_t)k * g.ldb;
    w.x = c + 0 < g.N ? row[0] : 0.f; w.y = c + 1 < g.N ? row[1] : 0.f; w.z = c + 2 < g.N ? row[2] : 0.f; w.w = c + 3 < g.N ? row[3] : 0.f;
    fma_row(w, k);
  }
  __syncthreads();   // every warp is done with the staged chunk: the buffer becomes the reduction area
#pragma unroll
  for (int r = 0; r < GV_ROWS; ++r)
    *reinterpret_cast<float4*>(red + ((size_t)warp * GV_ROWS + r) * (GN_COLS + 4) + 4 * lane) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  __syncthreads();
  const bool split = gridDim.y > 1;
  for (int e = threadIdx.x; e < g.M * GN_COLS; e += GV_THREADS) {
    const int r = e / GN_COLS, j = e % GN_COLS;
    if (n0 + j >= g.N) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < GV_THREADS / 32; ++w) s += red[((size_t)w * GV_ROWS + r) * (GN_COLS + 4) + j];
    if (g.bias != nullptr && blockIdx.y == 0) s += g.bias[n0 + j];
    float* dst = g.C + (size_t)r * g.ldc + n0 + j;
    if (split) atomicAdd(dst, s);                       // C was zeroed (beta = 0) or holds the value to accumulate into
    else *dst = g.beta != 0.f ? *dst + s : s;
  }
}

// ---- B [N, K] ------------------------------------------------------------------------------------------------------
constexpr int GT_KMAX = 1024;    // K-chunk staged per CTA: A [16][GT_KMAX] (64 KB)
constexpr int GT_COLS = 4;       // output columns (weight rows) per warp

// gpw: column groups (of GT_COLS columns) per warp - the staged A chunk is reused for 8 * gpw * GT_COLS weight rows
__global__ void __launch_bounds__(GV_THREADS) gemv16_t_kernel(GemmArgs g, int kchunk, int gpw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* as = reinterpret_cast<float*>(smem_raw);      // as[row][k], row stride kchunk (multiple of 4)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.y * kchunk, kn = min(kchunk, g.K - k0);
  for (int e = threadIdx.x; e < kchunk * GV_ROWS; e += GV_THREADS) {
    const int k = e % kchunk, r = e / kchunk;
    as[(size_t)r * kchunk + k] = (r < g.M && k < kn) ? g.A[(size_t)r * g.lda + k0 + k] : 0.f;
  }
  __syncthreads();
  const bool split = gridDim.y > 1;
  for (int grp = 0; grp < gpw; ++grp) {
  const int n0 = ((blockIdx.x * gpw + grp) * (GV_THREADS / 32) + warp) * GT_COLS;
  float acc[GV_ROWS][GT_COLS];
#pragma unroll
  for (int r = 0; r < GV_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < GT_COLS; ++j) acc[r][j] = 0.f;
  const bool vec = (g.ldb % 4 == 0) && (((uintptr_t)g.B & 15) == 0) && (k0 % 4 == 0);
  const float* brow[GT_COLS];
#pragma unroll
  for (int j = 0; j < GT_COLS; ++j) brow[j] = g.B + (size_t)min(n0 + j, g.N - 1) * g.ldb + k0;
  if (n0 < g.N) {
    const int kv = vec ? (kn & ~3) : 0;
    for (int k = 4 * lane; k < kv; k += 128) {
      float4 w[GT_COLS];
#pragma unroll
      for (int j = 0; j < GT_COLS; ++j) w[j] = __ldcs(reinterpret_cast<const float4*>(brow[j] + k));
#pragma unroll
      for (int r = 0; r < GV_ROWS; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(as + (size_t)r * kchunk + k);
#pragma unroll
        for (int j = 0; j < GT_COLS; ++j)
          acc[r][j] = fmaf(a.x, w[j].x, fmaf(a.y, w[j].y, fmaf(a.z, w[j].z, fmaf(a.w, w[j].w, acc[r][j]))));
      }
    }
    for (int k = kv + lane; k < kn; k += 32) {
      float w[GT_COLS];
#pragma unroll
      for (int j = 0; j < GT_COLS; ++j) w[j] = brow[j][k];
#pragma unroll
      for (int r = 0; r < GV_ROWS; ++r) {
        const float a = as[(size_t)r * kchunk + k];
#pragma unroll
        for (int j = 0; j < GT_COLS; ++j) acc[r][j] = fmaf(a, w[j], acc[r][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < GV_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < GT_COLS; ++j) {
      float v = acc[r][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[r][j] = v;
    }
  // 64 outputs per warp, two per lane
#pragma unroll
  for (int r = 0; r < GV_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < GT_COLS; ++j) {
      if (lane == ((r * GT_COLS + j) & 31) && r < g.M && n0 + j < g.N) {
        float s = acc[r][j];
        if (g.bias != nullptr && blockIdx.y == 0) s += g.bias[n0 + j];
        float* dst = g.C + (size_t)r * g.ldc + n0 + j;
        if (split) atomicAdd(dst, s);
        else *dst = g.beta != 0.f ? *dst + s : s;
      }
    }
  }
}

}  // namespace

bool gemv16_eligible(const GemmArgs& g) {
  // few rows, a weight matrix large enough that streaming it is the cost (>= 4 MB), plain A
  return !g.ta && g.M >= 1 && g.M <= GV_ROWS && (long long)g.K * g.N >= (1 << 20) && g.K >= 256 && g.N >= 64;
}

// Number of K-splits in [lo, hi] that fills whole waves of `cap` co-resident CTAs best (a 1.6-wave grid leaves 40 % of the
// chip idle for the second half of the kernel); ties go to fewer splits (fewer atomics).
static int pick_splits(int units, int lo, int hi, int cap) {
  int best = lo;
  double best_fill = 0.0;
  for (int sp = lo; sp <= hi; ++sp) {
    const long long ctas = (long long)units * sp;
    const double fill = (double)ctas / (double)(cdiv((int)ctas, cap) * (long long)cap);
    if (fill > best_fill + 0.02) { best_fill = fill; best = sp; }
  }
  return best;
}

void gemv16(cudaStream_t s, const GemmArgs& g) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int cap = 2 * sms;   // both kernels: 2 CTAs per SM (registers)
  if (!g.tb) {
    const int slabs = cdiv(g.N, GN_COLS);
    int splits = pick_splits(slabs, cdiv(g.K, GN_KMAX), max(cdiv(g.K, GN_KMAX), min(cdiv(2 * cap, slabs), cdiv(g.K, 128))), cap);
    const int kchunk = cdiv(cdiv(g.K, splits), 4) * 4;
    splits = cdiv(g.K, kchunk);
    if (splits > 1 && g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(gemv16_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_KMAX * GN_LD * 4); configured = true; }
    const size_t smem = (size_t)max(kchunk * GN_LD, GN_RED) * 4;
    gemv16_n_kernel<<<dim3(slabs, splits), GV_THREADS, smem, s>>>(g, kchunk);
  } else {
    // each CTA reuses its staged A chunk for gpw column groups per warp: as many as still leave a full wave of CTAs
    const int per_cta = (GV_THREADS / 32) * GT_COLS;
    const int kmin = cdiv(g.K, GT_KMAX);
    int gpw = 1;
    while (gpw < 8 && (long long)cdiv(g.N, per_cta * gpw * 2) * kmin >= cap) gpw *= 2;
    const int groups = cdiv(g.N, per_cta * gpw);
    int splits = pick_splits(groups, kmin, max(kmin, min(cdiv(2 * cap, groups), cdiv(g.K, 256))), cap);
    const int kchunk = cdiv(cdiv(g.K, splits), 4) * 4;
    splits = cdiv(g.K, kchunk);
    if (splits > 1 && g.beta == 0.f) zero_2d(s, g.C, g.M, g.N, g.ldc);
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(gemv16_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_KMAX * GV_ROWS * 4); configured = true; }
    gemv16_t_kernel<<<dim3(groups, splits), GV_THREADS, (size_t)kchunk * GV_ROWS * 4, s>>>(g, kchunk, gpw);
  }
  count_launch();
}

}  // namespace dv3
