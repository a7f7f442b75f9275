// Elementwise, LayerNorm+SiLU, GRU cell, categorical representation kernels (fp32).
#include <cuda_bf16.h>

#include <cstdlib>

#include "dv3_device.cuh"

namespace dv3 {
namespace {

constexpr int kThreads = 256;
inline int grid_for(long long n, int per_block = kThreads, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// SiLU's sigmoid in the LayerNorm kernels: exact (expf + IEEE division) in the fp32 parity mode, ex2.approx based (rel. error
// ~1e-6) in the tensor-core mode, where it halves the time of these 8 - 16 byte / element passes. The engine sets the switch at
// every entry point from its compute mode (set_ln_fast); the isolated ops run the fast kernels.
static bool g_ln_fast = true;
__device__ __forceinline__ float silu_sigmoid(float n, int fast) { return fast ? sigmoid_fast(n) : sigmoidf_(n); }

__global__ void symlog_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = symlog_dev(x[i]);
}
__global__ void inv_symlog_kernel(const float* __restrict__ y, float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = inv_symlog_dev(y[i]);
  }
}
__global__ void fill_kernel(float* p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void tanh_kernel(const float* x, float* y, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = tanhf(x[i]);
}
// mode 0: dst = 0; 1: dst = src; 2: dst += src
template <int MODE>
__global__ void op_2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    float* d = dst + (size_t)r * ldd + c;
    if (MODE == 0) *d = 0.f;
    else if (MODE == 1) *d = src[(size_t)r * lds + c];
    else *d += src[(size_t)r * lds + c];
  }
}

// fp32 -> bf16 copy of a 2-D block (cols a multiple of 4, 8-byte aligned destination rows)
__global__ void to_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd, long long rows, int cols4) {
  pdl_trigger();
  pdl_wait();
  const long long n = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4; const int c = (int)(i - r * cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + r * ldd + c) = o;
  }
}
__global__ void to_bf16_scalar_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd, long long rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols; const int c = (int)(i - r * cols);
    dst[r * ldd + c] = __float2bfloat16_rn(src[r * lds + c]);
  }
}

// ---- LayerNorm + SiLU: one warp per row ----------------------------------------------------------------
// (SiLU's sigmoid is the ex2.approx one of dv3_device.cuh in every LayerNorm kernel of this file, rel. error ~1e-6: the exact
//  expf + IEEE division cost as much as the memory traffic - 10^6 x 32 rows forward 165 -> 64 us with it)
// templated on the per-lane element count so that narrow rows (conv channels, D = 24..96) do not pay for 32 predicated
// iterations: EPL = ceil(D / 32) rounded up to a power of two
// RU rows per warp iteration: on narrow rows (conv channels over ~10^6 pixel rows) one 128-byte load in flight per warp
// leaves the kernel latency-bound (1.6 TB/s); RU = 4 rows go through the loads and the shuffle reductions together.
template <int kLnMaxPerLane, int RU>
__global__ void ln_silu_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int sstride, long long rows, int D,
                                   __nv_bfloat16* __restrict__ y16, int fast) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float gm[kLnMaxPerLane], bt[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    gm[i] = (c < D) ? gamma[c] : 0.f;
    bt[i] = (c < D) ? beta[c] : 0.f;
  }
  for (long long r0 = warp0; r0 < rows; r0 += nwarps * RU) {
    float v[RU][kLnMaxPerLane], s[RU], q[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = r0 + u * nwarps;
      const float* xr = x + (size_t)(r < rows ? r : r0) * ldx;
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        v[u][i] = (c < D) ? xr[c] : 0.f;
        s[u] += v[u][i];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      s[u] /= (float)D;   // mean
      q[u] = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        const float d = (c < D) ? (v[u][i] - s[u]) : 0.f;
        q[u] += d * d;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = r0 + u * nwarps;
      if (r >= rows) continue;
      const float mean = s[u];
      const float rstd = rsqrtf(q[u] / (float)D + DV3_LN_EPS);
      float* yr = y + (size_t)r * ldy;
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
          const float n = (v[u][i] - mean) * rstd * gm[i] + bt[i];
          const float o = n * silu_sigmoid(n, fast);
          yr[c] = o;
          if (y16 != nullptr) y16[(size_t)r * ldy + c] = __float2bfloat16_rn(o);  // bf16 twin for the TMA-fed GEMM
        }
      }
      if (lane == 0) {
        if (mean_out) mean_out[r * sstride] = mean;
        if (rstd_out) rstd_out[r * sstride] = rstd;
      }
    }
  }
}

template <int kLnMaxPerLane, int RU>
__global__ void ln_silu_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ dx, int lddx, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, int sstride, long long rows, int D,
                                   __nv_bfloat16* __restrict__ dx16, int fast) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float ag[kLnMaxPerLane], ab[kLnMaxPerLane], gm[kLnMaxPerLane], bt[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    ag[i] = ab[i] = 0.f;
    gm[i] = (c < D) ? gamma[c] : 0.f;
    bt[i] = (c < D) ? beta[c] : 0.f;
  }
  for (long long r0 = warp0; r0 < rows; r0 += nwarps * RU) {
    float xh[RU][kLnMaxPerLane], dxh[RU][kLnMaxPerLane], s1[RU], s2[RU], rs[RU];
    // every load of the RU rows first
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = r0 + u * nwarps;
      const bool rok = r < rows;
      const long long rr = rok ? r : r0;
      const float* xr = x + (size_t)rr * ldx;
      const float* dyr = dy + (size_t)rr * lddy;
      s1[u] = mean_in[rr * sstride];   // (mean, until the sums below)
      rs[u] = rstd_in[rr * sstride];
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        xh[u][i] = (c < D) ? xr[c] : 0.f;
        dxh[u][i] = (c < D && rok) ? dyr[c] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const float mean = s1[u];
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
          const float xhat = (xh[u][i] - mean) * rs[u];
          const float n = xhat * gm[i] + bt[i];
          const float sg = silu_sigmoid(n, fast);
          const float dn = dxh[u][i] * (sg * (1.f + n * (1.f - sg)));   // 0 for rows beyond the end (dy = 0)
          ag[i] += dn * xhat;
          ab[i] += dn;
          xh[u][i] = xhat;
          dxh[u][i] = dn * gm[i];
          a1 += dxh[u][i];
          a2 += dxh[u][i] * xhat;
        } else {
          xh[u][i] = 0.f; dxh[u][i] = 0.f;
        }
      }
      s1[u] = a1; s2[u] = a2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
        s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
      }
    }
    if (dx != nullptr) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const long long r = r0 + u * nwarps;
        if (r >= rows) continue;
        const float m1 = s1[u] / (float)D, m2 = s2[u] / (float)D;
        float* dxr = dx + (size_t)r * lddx;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) {
          const int c = lane + 32 * i;
          if (c < D) {
            const float o = rs[u] * (dxh[u][i] - m1 - xh[u][i] * m2);
            dxr[c] = o;
            if (dx16 != nullptr) dx16[(size_t)r * lddx + c] = __float2bfloat16_rn(o);
          }
        }
      }
    }
  }
  if (dgamma != nullptr) {
    // block-level reduction in shared memory first: one global atomic per column per block instead of per warp
    __shared__ float sdg[32 * kLnMaxPerLane], sdb[32 * kLnMaxPerLane];
    for (int c = threadIdx.x; c < 32 * kLnMaxPerLane; c += blockDim.x) { sdg[c] = 0.f; sdb[c] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        atomicAdd(&sdg[c], ag[i]);
        atomicAdd(&sdb[c], ab[i]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      atomicAdd(&dgamma[c], sdg[c]);
      atomicAdd(&dbeta[c], sdb[c]);
    }
  }
}

// ---- narrow rows (conv channels over ~10^6 pixel rows: D = 24 ... 192, D % 4 == 0, 16-byte aligned rows) -------------------
// A warp-per-row layout moves 128 bytes per warp-load at D = 32 and spends 10 shuffles per row: latency-bound at ~1.6 - 2 TB/s.
// Here LPR = 8 / 16 / 32 lanes share a row (32 / LPR rows per warp-load, each lane one float4 per chunk, VPL chunks per lane), so
// every load instruction of a warp moves up to 512 contiguous bytes and a row reduction is log2(LPR) shuffles; RU such steps are
// in flight together. The rows of one warp iteration are consecutive in memory. The sigmoid is the ex2.approx one of the scan
// kernels (rel. error ~1e-6): at 8 bytes per element the exact expf + division was half of the kernel's time.
template <int LPR, int VPL, int RU>
__global__ void __launch_bounds__(256) ln_silu_fwd_narrow_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                                                  float* __restrict__ mean_out, float* __restrict__ rstd_out, int sstride,
                                                                  long long rows, int D, __nv_bfloat16* __restrict__ y16, int fast) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invD = 1.f / (float)D;
  float4 gm[VPL], bt[VPL];
  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (cl + LPR * i);
    act[i] = c < D;
    gm[i] = act[i] ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = act[i] ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long base = warp0 * (RPW * RU); base < rows; base += nwarps * (RPW * RU)) {
    float4 v[RU][VPL];
    float s[RU], q[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = base + u * RPW + sub;
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        v[u][i] = (r < rows && act[i]) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx + 4 * (cl + LPR * i)))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        s[u] += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      s[u] *= invD;   // mean
      q[u] = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (act[i]) {
          const float dx = v[u][i].x - s[u], dy = v[u][i].y - s[u], dz = v[u][i].z - s[u], dw = v[u][i].w - s[u];
          q[u] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = base + u * RPW + sub;
      if (r >= rows) continue;
      const float mean = s[u];
      const float rstd = rsqrtf(q[u] * invD + DV3_LN_EPS);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (!act[i]) continue;
        const int c = 4 * (cl + LPR * i);
        float4 o;
        { const float n = (v[u][i].x - mean) * rstd * gm[i].x + bt[i].x; o.x = n * silu_sigmoid(n, fast); }
        { const float n = (v[u][i].y - mean) * rstd * gm[i].y + bt[i].y; o.y = n * silu_sigmoid(n, fast); }
        { const float n = (v[u][i].z - mean) * rstd * gm[i].z + bt[i].z; o.z = n * silu_sigmoid(n, fast); }
        { const float n = (v[u][i].w - mean) * rstd * gm[i].w + bt[i].w; o.w = n * silu_sigmoid(n, fast); }
        *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) = o;
        if (y16 != nullptr) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(y16 + (size_t)r * ldy + c) = pk;
        }
      }
      if (cl == 0) {
        if (mean_out) mean_out[r * sstride] = mean;
        if (rstd_out) rstd_out[r * sstride] = rstd;
      }
    }
  }
}

template <int LPR, int VPL, int RU>
__global__ void __launch_bounds__(256) ln_silu_bwd_narrow_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* __restrict__ dx, int lddx, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, int sstride, long long rows, int D,
                                                                  __nv_bfloat16* __restrict__ dx16, int fast) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invD = 1.f / (float)D;
  float4 gm[VPL], bt[VPL], ag[VPL], ab[VPL];
  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = 4 * (cl + LPR * i);
    act[i] = c < D;
    gm[i] = act[i] ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = act[i] ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long base = warp0 * (RPW * RU); base < rows; base += nwarps * (RPW * RU)) {
    float4 xh[RU][VPL], dh[RU][VPL];
    float s1[RU], s2[RU], rs[RU], mu[RU];
    // every load of the RU x RPW rows first
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long r = base + u * RPW + sub;
      const bool rok = r < rows;
      mu[u] = rok ? mean_in[r * sstride] : 0.f;
      rs[u] = rok ? rstd_in[r * sstride] : 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const bool ok = rok && act[i];
        const int c = 4 * (cl + LPR * i);
        xh[u][i] = ok ? __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        dh[u][i] = ok ? __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * lddy + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (!act[i]) continue;
        // per component: xhat, n = xhat g + b, dn = dy silu'(n); accumulates dgamma / dbeta, leaves (xhat, dn g) in place
#define DV3_LN_COMP(F_)                                                                  \
        {                                                                                \
          const float xhat = (xh[u][i].F_ - mu[u]) * rs[u];                              \
          const float n = xhat * gm[i].F_ + bt[i].F_;                                    \
          const float sg = silu_sigmoid(n, fast);                                                 \
          const float dn = dh[u][i].F_ * (sg * (1.f + n * (1.f - sg)));                  \
          ag[i].F_ += dn * xhat; ab[i].F_ += dn;                                         \
          xh[u][i].F_ = xhat; dh[u][i].F_ = dn * gm[i].F_;                               \
          a1 += dh[u][i].F_; a2 += dh[u][i].F_ * xhat;                                   \
        }
        DV3_LN_COMP(x) DV3_LN_COMP(y) DV3_LN_COMP(z) DV3_LN_COMP(w)
#undef DV3_LN_COMP
      }
      s1[u] = a1; s2[u] = a2;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
        s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
      }
    }
    if (dx != nullptr) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const long long r = base + u * RPW + sub;
        if (r >= rows) continue;
        const float m1 = s1[u] * invD, m2 = s2[u] * invD;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          if (!act[i]) continue;
          const int c = 4 * (cl + LPR * i);
          float4 o;
          o.x = rs[u] * (dh[u][i].x - m1 - xh[u][i].x * m2);
          o.y = rs[u] * (dh[u][i].y - m1 - xh[u][i].y * m2);
          o.z = rs[u] * (dh[u][i].z - m1 - xh[u][i].z * m2);
          o.w = rs[u] * (dh[u][i].w - m1 - xh[u][i].w * m2);
          *reinterpret_cast<float4*>(dx + (size_t)r * lddx + c) = o;
          if (dx16 != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dx16 + (size_t)r * lddx + c) = pk;
          }
        }
      }
    }
  }
  if (dgamma != nullptr) {
    // the row groups of a warp hold partial sums of the same columns: fold them with shuffles, then one shared-memory atomic
    // per column and warp, one global atomic per column and block
    __shared__ float sdg[4 * LPR * VPL], sdb[4 * LPR * VPL];
    for (int c = threadIdx.x; c < 4 * LPR * VPL; c += blockDim.x) { sdg[c] = 0.f; sdb[c] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        ag[i].x += __shfl_xor_sync(0xffffffffu, ag[i].x, o); ag[i].y += __shfl_xor_sync(0xffffffffu, ag[i].y, o);
        ag[i].z += __shfl_xor_sync(0xffffffffu, ag[i].z, o); ag[i].w += __shfl_xor_sync(0xffffffffu, ag[i].w, o);
        ab[i].x += __shfl_xor_sync(0xffffffffu, ab[i].x, o); ab[i].y += __shfl_xor_sync(0xffffffffu, ab[i].y, o);
        ab[i].z += __shfl_xor_sync(0xffffffffu, ab[i].z, o); ab[i].w += __shfl_xor_sync(0xffffffffu, ab[i].w, o);
      }
      if (sub == 0 && act[i]) {
        const int c = 4 * (cl + LPR * i);
        atomicAdd(&sdg[c], ag[i].x); atomicAdd(&sdg[c + 1], ag[i].y); atomicAdd(&sdg[c + 2], ag[i].z); atomicAdd(&sdg[c + 3], ag[i].w);
        atomicAdd(&sdb[c], ab[i].x); atomicAdd(&sdb[c + 1], ab[i].y); atomicAdd(&sdb[c + 2], ab[i].z); atomicAdd(&sdb[c + 3], ab[i].w);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      atomicAdd(&dgamma[c], sdg[c]);
      atomicAdd(&dbeta[c], sdb[c]);
    }
  }
}

// ---- vectorised LayerNorm + SiLU (D a multiple of 128, 16-byte aligned rows): lane owns V chunks of 4 consecutive columns,
// chunk i at column 4 * (lane + 32 i); all loads of a row are 128-bit and issued together ----
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 o; o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = o;
}

template <int V>
__global__ void __launch_bounds__(256) ln_silu_fwd_v4_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y, int ldy,
                                                             float* __restrict__ mean_out, float* __restrict__ rstd_out, int sstride,
                                                             long long rows, __nv_bfloat16* __restrict__ y16, int fast) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 g[V], be[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * i));
    be[i] = *reinterpret_cast<const float4*>(beta + 4 * (lane + 32 * i));
  }
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* xr = x + (size_t)r * ldx;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = *reinterpret_cast<const float4*>(xr + 4 * (lane + 32 * i));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + DV3_LN_EPS);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c0 = 4 * (lane + 32 * i);
      float4 o;
      float n;
      n = (v[i].x - mean) * rstd * g[i].x + be[i].x; o.x = n * silu_sigmoid(n, fast);
      n = (v[i].y - mean) * rstd * g[i].y + be[i].y; o.y = n * silu_sigmoid(n, fast);
      n = (v[i].z - mean) * rstd * g[i].z + be[i].z; o.z = n * silu_sigmoid(n, fast);
      n = (v[i].w - mean) * rstd * g[i].w + be[i].w; o.w = n * silu_sigmoid(n, fast);
      *reinterpret_cast<float4*>(y + (size_t)r * ldy + c0) = o;
      if (y16 != nullptr) store_bf16x4(y16 + (size_t)r * ldy + c0, o.x, o.y, o.z, o.w);
    }
    if (lane == 0) {
      if (mean_out) mean_out[r * sstride] = mean;
      if (rstd_out) rstd_out[r * sstride] = rstd;
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256) ln_silu_bwd_v4_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                             const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ dx, int lddx, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int sstride, long long rows,
                                                             __nv_bfloat16* __restrict__ dx16, int fast) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = V * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 g[V], be[V], ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * i));
    be[i] = *reinterpret_cast<const float4*>(beta + 4 * (lane + 32 * i));
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = warp0; r < rows; r += nwarps) {
    const float mean = mean_in[r * sstride], rstd = rstd_in[r * sstride];
    float4 xh[V], dh[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      xh[i] = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + 4 * (lane + 32 * i));
      dh[i] = *reinterpret_cast<const float4*>(dy + (size_t)r * lddy + 4 * (lane + 32 * i));
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
#define DV3_LNB(c_)                                                              \
      {                                                                          \
        const float xhat = (xh[i].c_ - mean) * rstd;                             \
        const float n = xhat * g[i].c_ + be[i].c_;                               \
        const float sg = silu_sigmoid(n, fast);                                           \
        const float dn = dh[i].c_ * (sg * (1.f + n * (1.f - sg)));               \
        ag[i].c_ += dn * xhat; ab[i].c_ += dn;                                   \
        xh[i].c_ = xhat; dh[i].c_ = dn * g[i].c_;                                \
        s1 += dh[i].c_; s2 += dh[i].c_ * xhat;                                   \
      }
      DV3_LNB(x) DV3_LNB(y) DV3_LNB(z) DV3_LNB(w)
#undef DV3_LNB
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    if (dx != nullptr) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c0 = 4 * (lane + 32 * i);
        float4 o;
        o.x = rstd * (dh[i].x - s1 - xh[i].x * s2); o.y = rstd * (dh[i].y - s1 - xh[i].y * s2);
        o.z = rstd * (dh[i].z - s1 - xh[i].z * s2); o.w = rstd * (dh[i].w - s1 - xh[i].w * s2);
        *reinterpret_cast<float4*>(dx + (size_t)r * lddx + c0) = o;
        if (dx16 != nullptr) store_bf16x4(dx16 + (size_t)r * lddx + c0, o.x, o.y, o.z, o.w);
      }
    }
  }
  if (dgamma != nullptr) {
    __shared__ float sdg[D], sdb[D];
    for (int c = threadIdx.x; c < D; c += blockDim.x) { sdg[c] = 0.f; sdb[c] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c0 = 4 * (lane + 32 * i);
      atomicAdd(&sdg[c0], ag[i].x); atomicAdd(&sdg[c0 + 1], ag[i].y); atomicAdd(&sdg[c0 + 2], ag[i].z); atomicAdd(&sdg[c0 + 3], ag[i].w);
      atomicAdd(&sdb[c0], ab[i].x); atomicAdd(&sdb[c0 + 1], ab[i].y); atomicAdd(&sdb[c0 + 2], ab[i].z); atomicAdd(&sdb[c0 + 3], ab[i].w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      atomicAdd(&dgamma[c], sdg[c]);
      atomicAdd(&dbeta[c], sdb[c]);
    }
  }
}

// ---- GRU cell ---------------------------------------------------------------------------------------
__global__ void gru_fwd_kernel(float* __restrict__ gx, int ldgx, const float* __restrict__ gh, int ldgh,
                               const float* __restrict__ h_prev, int ldhp, float* __restrict__ h_out, int ldho,
                               int M, int G, __nv_bfloat16* __restrict__ h16) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)M * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / G), j = (int)(i % G);
    float* gxr = gx + (size_t)r * ldgx;
    const float* ghr = gh + (size_t)r * ldgh;
    const float z = sigmoidf_(gxr[j] + ghr[j]);
    const float rg = sigmoidf_(gxr[G + j] + ghr[G + j]);
    const float c = tanhf(gxr[2 * G + j] + rg * ghr[2 * G + j]);
    const float hp = h_prev[(size_t)r * ldhp + j];
    const float hn = z * hp + (1.f - z) * c;
    h_out[(size_t)r * ldho + j] = hn;
    if (h16 != nullptr) h16[(size_t)r * ldho + j] = __float2bfloat16_rn(hn);
    gxr[j] = z; gxr[G + j] = rg; gxr[2 * G + j] = c;
  }
}

__global__ void gru_bwd_kernel(const float* __restrict__ dh, int lddh, const float* __restrict__ gates, int ldg,
                               const float* __restrict__ gh, int ldgh, const float* __restrict__ h_prev, int ldhp,
                               float* __restrict__ dgx, int lddgx, float* __restrict__ dgh, int lddgh,
                               float* __restrict__ dh_prev, int lddhp, int accumulate, int M, int G,
                               __nv_bfloat16* __restrict__ dgx16, __nv_bfloat16* __restrict__ dgh16) {
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)M * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / G), j = (int)(i % G);
    const float* gr = gates + (size_t)r * ldg;
    const float z = gr[j], rg = gr[G + j], c = gr[2 * G + j];
    const float hh = gh[(size_t)r * ldgh + 2 * G + j];
    const float hp = h_prev[(size_t)r * ldhp + j];
    const float d = dh[(size_t)r * lddh + j];
    const float dz = d * (hp - c) * z * (1.f - z);
    const float dc = d * (1.f - z) * (1.f - c * c);
    const float dr = dc * hh * rg * (1.f - rg);
    float* a = dgx + (size_t)r * lddgx;
    float* b = dgh + (size_t)r * lddgh;
    a[j] = dz; a[G + j] = dr; a[2 * G + j] = dc;
    b[j] = dz; b[G + j] = dr; b[2 * G + j] = dc * rg;
    if (dgx16 != nullptr) {   // bf16 twins for the two input-gradient contractions that follow
      __nv_bfloat16* a16 = dgx16 + (size_t)r * lddgx;
      __nv_bfloat16* b16 = dgh16 + (size_t)r * lddgh;
      const __nv_bfloat16 hz_ = __float2bfloat16_rn(dz), hr = __float2bfloat16_rn(dr);
      a16[j] = hz_; a16[G + j] = hr; a16[2 * G + j] = __float2bfloat16_rn(dc);
      b16[j] = hz_; b16[G + j] = hr; b16[2 * G + j] = __float2bfloat16_rn(dc * rg);
    }
    float* o = dh_prev + (size_t)r * lddhp + j;
    if (accumulate) *o += d * z; else *o = d * z;
  }
}

// ---- categorical representation: one warp per (row, categorical); lane = class -----------------------------
__global__ void repr_fwd_kernel(const float* __restrict__ logits, int ldl, float* __restrict__ probs, int ldp,
                                const float* __restrict__ u, int ldu, int32_t* __restrict__ idx, int ldi,
                                float* __restrict__ z, int ldz, int M, int Zc, int Zk, int mode,
                                __nv_bfloat16* __restrict__ z16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)M * Zc;
  for (long long w = warp0; w < total; w += nwarps) {
    const int r = (int)(w / Zc), c = (int)(w % Zc);
    const bool act = lane < Zk;
    const float l = act ? logits[(size_t)r * ldl + c * Zk + lane] : -INFINITY;
    const float mx = warp_max(l);
    const float e = act ? expf(l - mx) : 0.f;
    const float s = warp_sum(e);
    const float p = 0.99f * (e / s) + 0.01f * (1.0f / (float)Zk);
    if (act && probs != nullptr) probs[(size_t)r * ldp + c * Zk + lane] = p;
    if (mode == 0) continue;
    int pick;
    if (mode == 1) {
      pick = categorical_pick(p, act, Zk, u[(size_t)r * ldu + c], lane);
    } else {
      // argmax, first index on ties
      const float pm = warp_max(act ? p : -1.f);
      const unsigned is = __ballot_sync(0xffffffffu, act && (p == pm));
      pick = __ffs(is) - 1;
    }
    if (idx != nullptr && lane == 0) idx[(size_t)r * ldi + c] = pick;
    if (z != nullptr && act) z[(size_t)r * ldz + c * Zk + lane] = (lane == pick) ? 1.f : 0.f;
    if (z16 != nullptr && act) z16[(size_t)r * ldz + c * Zk + lane] = __float2bfloat16_rn((lane == pick) ? 1.f : 0.f);
  }
}

// Same, 32-class categoricals, four per warp at a time: their loads are issued together and their shuffle chains (max, sum,
// fp64 scan) interleave instead of running back to back - the kernel is a latency chain per warp, not a bandwidth problem
__global__ void __launch_bounds__(256) repr_fwd4_kernel(const float* __restrict__ logits, int ldl, float* __restrict__ probs, int ldp,
                                                        const float* __restrict__ u, int ldu, int32_t* __restrict__ idx, int ldi,
                                                        float* __restrict__ z, int ldz, int M, int Zc, int mode,
                                                        __nv_bfloat16* __restrict__ z16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int groups = Zc / 4;
  const long long total = (long long)M * groups;
  for (long long w = warp0; w < total; w += nwarps) {
    const int r = (int)(w / groups), c0 = (int)(w % groups) * 4;
    float l[4], uu[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      l[q] = logits[(size_t)r * ldl + (c0 + q) * 32 + lane];
      uu[q] = mode == 1 ? u[(size_t)r * ldu + c0 + q] : 0.f;
    }
    float mx[4], e[4], sden[4], p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mx[q] = l[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { e[q] = expf(l[q] - mx[q]); sden[q] = e[q]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sden[q] += __shfl_xor_sync(0xffffffffu, sden[q], o);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      p[q] = 0.99f * (e[q] / sden[q]) + 0.01f * (1.0f / 32.0f);
      if (probs != nullptr) probs[(size_t)r * ldp + (c0 + q) * 32 + lane] = p[q];
    }
    if (mode == 0) continue;
    int pick[4];
    if (mode == 1) {
      double cdf[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) cdf[q] = (double)p[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double t = __shfl_up_sync(0xffffffffu, cdf[q], o);
          if (lane >= o) cdf[q] += t;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // as categorical_pick (dv3_device.cuh): index = #(cdf <= u * total), clamped
        const double tot = __shfl_sync(0xffffffffu, cdf[q], 31);
        const unsigned below = __ballot_sync(0xffffffffu, cdf[q] <= (double)uu[q] * tot);
        pick[q] = min(__popc(below), 31);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // argmax, first index on ties
        const float pm = warp_max(p[q]);
        pick[q] = __ffs(__ballot_sync(0xffffffffu, p[q] == pm)) - 1;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (idx != nullptr && lane == 0) idx[(size_t)r * ldi + c0 + q] = pick[q];
      const float hot = (lane == pick[q]) ? 1.f : 0.f;
      if (z != nullptr) z[(size_t)r * ldz + (c0 + q) * 32 + lane] = hot;
      if (z16 != nullptr) z16[(size_t)r * ldz + (c0 + q) * 32 + lane] = __float2bfloat16_rn(hot);
    }
  }
}

__global__ void repr_bwd_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ dz, int lddz,
                                const float* __restrict__ dprobs, int lddp, float* __restrict__ dlogits, int lddl,
                                int M, int Zc, int Zk, __nv_bfloat16* __restrict__ dl16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)M * Zc;
  for (long long w = warp0; w < total; w += nwarps) {
    const int r = (int)(w / Zc), c = (int)(w % Zc);
    const bool act = lane < Zk;
    const int col = c * Zk + lane;
    const float l = act ? logits[(size_t)r * ldl + col] : -INFINITY;
    const float mx = warp_max(l);
    const float e = act ? expf(l - mx) : 0.f;
    const float s = warp_sum(e);
    const float p = e / s;
    float g = 0.f;
    if (act) {
      if (dz != nullptr) g += dz[(size_t)r * lddz + col];
      if (dprobs != nullptr) g += dprobs[(size_t)r * lddp + col];
      g *= 0.99f;
    }
    const float dot = warp_sum(g * p);
    if (act) {
      const float o = p * (g - dot);
      dlogits[(size_t)r * lddl + col] = o;
      if (dl16 != nullptr) dl16[(size_t)r * lddl + col] = __float2bfloat16_rn(o);
    }
  }
}

// Same, 32-class categoricals, four per warp at a time (loads issued together, shuffle chains interleaved; see repr_fwd4_kernel)
__global__ void __launch_bounds__(256) repr_bwd4_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ dz, int lddz,
                                                        const float* __restrict__ dprobs, int lddp, float* __restrict__ dlogits, int lddl,
                                                        int M, int Zc, __nv_bfloat16* __restrict__ dl16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int groups = Zc / 4;
  const long long total = (long long)M * groups;
  for (long long w = warp0; w < total; w += nwarps) {
    const int r = (int)(w / groups), c0 = (int)(w % groups) * 4;
    float l[4], g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = (c0 + q) * 32 + lane;
      l[q] = logits[(size_t)r * ldl + col];
      float gg = 0.f;
      if (dz != nullptr) gg += dz[(size_t)r * lddz + col];
      if (dprobs != nullptr) gg += dprobs[(size_t)r * lddp + col];
      g[q] = gg * 0.99f;
    }
    float mx[4], e[4], sden[4], p[4], dot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mx[q] = l[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { e[q] = expf(l[q] - mx[q]); sden[q] = e[q]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) sden[q] += __shfl_xor_sync(0xffffffffu, sden[q], o);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { p[q] = e[q] / sden[q]; dot[q] = g[q] * p[q]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = (c0 + q) * 32 + lane;
      const float o = p[q] * (g[q] - dot[q]);
      dlogits[(size_t)r * lddl + col] = o;
      if (dl16 != nullptr) dl16[(size_t)r * lddl + col] = __float2bfloat16_rn(o);
    }
  }
}

// ---- observe-scan masking (world_model.py:246-253) -----------------------------------------------------------
__global__ void scan_mask_inputs_kernel(const float* __restrict__ is_first, int ldf, const float* __restrict__ h_prev,
                                        int ldhp, const float* __restrict__ z_prev, int ldzp,
                                        const float* __restrict__ h0, const float* __restrict__ z0,
                                        const float* __restrict__ actW, int ldaw, float* __restrict__ h_in, int ldhi,
                                        float* __restrict__ z_in, int ldzi, float* __restrict__ x_pre, int ldx, int B,
                                        int G, int Z, int D) {
  const int W = G + Z + D;
  const long long n = (long long)B * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / W);
    int j = (int)(i % W);
    const float f = is_first[(size_t)b * ldf];
    if (j < G) {
      const float hp = (h_prev != nullptr) ? h_prev[(size_t)b * ldhp + j] : h0[j];
      h_in[(size_t)b * ldhi + j] = hp * (1.f - f) + h0[j] * f;
    } else if (j < G + Z) {
      j -= G;
      const float zp = (z_prev != nullptr) ? z_prev[(size_t)b * ldzp + j] : z0[j];
      z_in[(size_t)b * ldzi + j] = zp * (1.f - f) + z0[j] * f;
    } else {
      j -= G + Z;
      x_pre[(size_t)b * ldx + j] = actW[(size_t)b * ldaw + j];
    }
  }
}

__global__ void scan_mask_grads_kernel(const float* __restrict__ is_first, int ldf, const float* __restrict__ d_in,
                                       int ldin, float* __restrict__ d_out, int ldout, int B, int W) {
  const long long n = (long long)B * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / W), j = (int)(i % W);
    d_out[(size_t)b * ldout + j] += d_in[(size_t)b * ldin + j] * (1.f - is_first[(size_t)b * ldf]);
  }
}


// out[c] += sum_r x[r*ld + c]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int ld, long long rows, int cols,
                                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (long long r = blockIdx.y * 8 + ty; r < rows; r += (long long)gridDim.y * 8) s += x[(size_t)r * ld + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(&out[c], t);
  }
}

// a_in[b,t,:] = (1 - f[b,t]) * actions[b,(t-1) mod T,:]; wf[b,t] = t==0 ? 1 : f[b,t]   (world_model.py:246-253)
__global__ void scan_prepare_kernel(const float* __restrict__ actions, const float* __restrict__ is_first,
                                    float* __restrict__ a_in, float* __restrict__ wf, int B, int T, int A) {
  const long long n = (long long)B * T * (A + 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % (A + 1));
    const long long bt = i / (A + 1);
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float f = is_first[bt];
    if (j < A) a_in[bt * A + j] = (1.f - f) * actions[((size_t)b * T + (t + T - 1) % T) * A + j];
    else wf[bt] = (t == 0) ? 1.f : f;
  }
}
// acting-step masking (world_model.py:160-172): h/z fall back to the learned initial state, actions to zero
__global__ void infer_mask_kernel(const float* __restrict__ f, const float* __restrict__ ph, const float* __restrict__ pz,
                                  const float* __restrict__ pa, const float* __restrict__ h0, const float* __restrict__ z0,
                                  float* __restrict__ h, float* __restrict__ z, float* __restrict__ a, int rows, int G, int Z, int A) {
  const int W = G + Z + A;
  const long long n = (long long)rows * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / W);
    int j = (int)(i % W);
    const float m = f[b];
    if (j < G) h[(size_t)b * G + j] = ph[(size_t)b * G + j] * (1.f - m) + h0[j] * m;
    else if (j < G + Z) { j -= G; z[(size_t)b * Z + j] = pz[(size_t)b * Z + j] * (1.f - m) + z0[j] * m; }
    else { j -= G + Z; a[(size_t)b * A + j] = pa[(size_t)b * A + j] * (1.f - m); }
  }
}
// dinit[j] += dh0[j] * (1 - h0[j]^2)
__global__ void tanh_bwd_acc_kernel(const float* __restrict__ dh0, const float* __restrict__ h0, float* __restrict__ dinit, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dinit[i] += dh0[i] * (1.f - h0[i] * h0[i]);
}

}  // namespace

void symlog_fwd(cudaStream_t s, const float* x, float* y, long long n) {
  if (n <= 0) return;
  symlog_kernel<<<grid_for(n), kThreads, 0, s>>>(x, y, n); count_launch();
}
void inverse_symlog_fwd(cudaStream_t s, const float* y, float* x, long long n) {
  if (n <= 0) return;
  inv_symlog_kernel<<<grid_for(n), kThreads, 0, s>>>(y, x, n); count_launch();
}
void fill_f32(cudaStream_t s, float* p, float v, long long n) {
  if (n <= 0) return;
  fill_kernel<<<grid_for(n), kThreads, 0, s>>>(p, v, n); count_launch();
}
void tanh_fwd(cudaStream_t s, const float* x, float* y, int n) {
  tanh_kernel<<<grid_for(n), kThreads, 0, s>>>(x, y, n); count_launch();
}
void zero_2d(cudaStream_t s, float* p, int rows, int cols, int ld) {
  if (rows <= 0 || cols <= 0) return;
  if (ld == cols) { cudaMemsetAsync(p, 0, (size_t)rows * cols * sizeof(float), s); return; }
  launch_pdl(op_2d_kernel<0>, dim3(grid_for((long long)rows * cols)), dim3(kThreads), 0, s, nullptr, 0, p, ld, rows, cols); count_launch();
}
void to_bf16(cudaStream_t s, const float* src, int lds, void* dst, int ldd, long long rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 8 == 0)
    launch_pdl(to_bf16_kernel, dim3(grid_for(rows * (cols / 4))), dim3(kThreads), 0, s, src, lds, d, ldd, rows, cols / 4);
  else
    launch_pdl(to_bf16_scalar_kernel, dim3(grid_for(rows * cols)), dim3(kThreads), 0, s, src, lds, d, ldd, rows, cols);
  count_launch();
}
void copy_2d(cudaStream_t s, const float* src, int lds, float* dst, int ldd, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  launch_pdl(op_2d_kernel<1>, dim3(grid_for((long long)rows * cols)), dim3(kThreads), 0, s, src, lds, dst, ldd, rows, cols); count_launch();
}
void add_2d(cudaStream_t s, const float* src, int lds, float* dst, int ldd, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  launch_pdl(op_2d_kernel<2>, dim3(grid_for((long long)rows * cols)), dim3(kThreads), 0, s, src, lds, dst, ldd, rows, cols); count_launch();
}

void set_ln_fast(bool on) { g_ln_fast = on; }

void ln_silu_fwd(cudaStream_t s, const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                 float* mean, float* rstd, int sstride, long long rows, int D, void* y16_) {
  if (rows <= 0) return;
  const int fast = g_ln_fast ? 1 : 0;
  const int grid = grid_for(rows, kThreads / 32, 148 * 8);
  __nv_bfloat16* y16 = reinterpret_cast<__nv_bfloat16*>(y16_);
  if (D % 128 == 0 && D <= 1024 && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 &&
      (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 && (uintptr_t)y16_ % 8 == 0) {
#define DV3_LN_FWD4(V_) launch_pdl(ln_silu_fwd_v4_kernel<V_>, dim3(grid), dim3(kThreads), 0, s, x, ldx, gamma, beta, y, ldy, mean, rstd, sstride, rows, y16, fast)
    switch (D / 128) {   // every width that is a multiple of 128 (dense 256 / 512 / 640 / 768 / 1024, wide conv channels)
      case 1: DV3_LN_FWD4(1); break; case 2: DV3_LN_FWD4(2); break; case 3: DV3_LN_FWD4(3); break; case 4: DV3_LN_FWD4(4); break;
      case 5: DV3_LN_FWD4(5); break; case 6: DV3_LN_FWD4(6); break; case 7: DV3_LN_FWD4(7); break; default: DV3_LN_FWD4(8); break;
    }
#undef DV3_LN_FWD4
    count_launch();
    return;
  }
  static const bool no_narrow = std::getenv("DV3_NO_LN_NARROW") != nullptr;   // (A/B switch: the warp-per-row kernels below)
  if (fast && !no_narrow && D % 4 == 0 && D <= 256 && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 &&
      (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 && (uintptr_t)y16_ % 8 == 0) {
    // conv channel counts 24 ... 192 (and the small test sizes): several rows per warp, 128-bit accesses
#define DV3_LN_FWDN(L_, V_, RU_)                                                                                                   \
    launch_pdl(ln_silu_fwd_narrow_kernel<L_, V_, RU_>, dim3(grid_for(cdiv(rows, (32 / L_) * RU_), kThreads / 32, 148 * 8)), dim3(kThreads), \
               0, s, x, ldx, gamma, beta, y, ldy, mean, rstd, sstride, rows, D, y16, fast)
    if (D <= 32) DV3_LN_FWDN(8, 1, 4);
    else if (D <= 64) DV3_LN_FWDN(16, 1, 4);
    else if (D <= 128) DV3_LN_FWDN(32, 1, 4);
    else DV3_LN_FWDN(32, 2, 2);
#undef DV3_LN_FWDN
    count_launch();
    return;
  }
#define DV3_LN_FWD(E_, RU_) launch_pdl(ln_silu_fwd_kernel<E_, RU_>, dim3(grid), dim3(kThreads), 0, s, x, ldx, gamma, beta, y, ldy, mean, rstd, sstride, rows, D, y16, fast)
  if (D <= 32) DV3_LN_FWD(1, 4);
  else if (D <= 64) DV3_LN_FWD(2, 4);
  else if (D <= 128) DV3_LN_FWD(4, 2);
  else if (D <= 256) DV3_LN_FWD(8, 1);
  else if (D <= 512) DV3_LN_FWD(16, 1);
  else DV3_LN_FWD(32, 1);
#undef DV3_LN_FWD
  count_launch();
}
void ln_silu_bwd(cudaStream_t s, const float* dy, int lddy, const float* x, int ldx, const float* mean,
                 const float* rstd, const float* gamma, const float* beta, float* dx, int lddx, float* dgamma,
                 float* dbeta, int sstride, long long rows, int D, void* dx16_) {
  if (rows <= 0) return;
  const int fast = g_ln_fast ? 1 : 0;
  const int grid = grid_for(rows, kThreads / 32, 148 * 4);
  __nv_bfloat16* dx16 = reinterpret_cast<__nv_bfloat16*>(dx16_);
  if (D % 128 == 0 && D <= 1024 && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)dy % 16 == 0 &&
      (uintptr_t)dx % 16 == 0 && (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 && (uintptr_t)dx16_ % 8 == 0) {
    // with gain / offset gradients every block ends with 2 D global atomics: fewer, longer blocks (measured at 16384 x 256:
    // 17.4 / 13.1 / 18.5 / 19.7 us for 1 / 2 / 4 / 8 blocks per SM)
    const int grid4 = grid_for(rows, kThreads / 32, dgamma != nullptr ? 148 * 2 : 148 * 8);
#define DV3_LN_BWD4(V_) launch_pdl(ln_silu_bwd_v4_kernel<V_>, dim3(grid4), dim3(kThreads), 0, s, dy, lddy, x, ldx, mean, rstd, gamma, beta, dx, lddx, dgamma, dbeta, sstride, rows, dx16, fast)
    switch (D / 128) {
      case 1: DV3_LN_BWD4(1); break; case 2: DV3_LN_BWD4(2); break; case 3: DV3_LN_BWD4(3); break; case 4: DV3_LN_BWD4(4); break;
      case 5: DV3_LN_BWD4(5); break; case 6: DV3_LN_BWD4(6); break; case 7: DV3_LN_BWD4(7); break; default: DV3_LN_BWD4(8); break;
    }
#undef DV3_LN_BWD4
    count_launch();
    return;
  }
  static const bool no_narrow = std::getenv("DV3_NO_LN_NARROW") != nullptr;
  if (fast && !no_narrow && D % 4 == 0 && D <= 256 && ldx % 4 == 0 && lddy % 4 == 0 && (dx == nullptr || lddx % 4 == 0) && (uintptr_t)x % 16 == 0 &&
      (uintptr_t)dy % 16 == 0 && (uintptr_t)dx % 16 == 0 && (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 && (uintptr_t)dx16_ % 8 == 0) {
#define DV3_LN_BWDN(L_, V_, RU_)                                                                                                   \
    launch_pdl(ln_silu_bwd_narrow_kernel<L_, V_, RU_>, dim3(grid_for(cdiv(rows, (32 / L_) * RU_), kThreads / 32, 148 * 4)), dim3(kThreads), \
               0, s, dy, lddy, x, ldx, mean, rstd, gamma, beta, dx, lddx, dgamma, dbeta, sstride, rows, D, dx16, fast)
    if (D <= 32) DV3_LN_BWDN(8, 1, 4);
    else if (D <= 64) DV3_LN_BWDN(16, 1, 4);
    else if (D <= 128) DV3_LN_BWDN(32, 1, 2);
    else DV3_LN_BWDN(32, 2, 1);
#undef DV3_LN_BWDN
    count_launch();
    return;
  }
#define DV3_LN_BWD(E_, RU_) launch_pdl(ln_silu_bwd_kernel<E_, RU_>, dim3(grid), dim3(kThreads), 0, s, dy, lddy, x, ldx, mean, rstd, gamma, beta, dx, lddx, dgamma, dbeta, sstride, rows, D, dx16, fast)
  if (D <= 32) DV3_LN_BWD(1, 4);
  else if (D <= 64) DV3_LN_BWD(2, 4);
  else if (D <= 128) DV3_LN_BWD(4, 2);
  else if (D <= 256) DV3_LN_BWD(8, 1);
  else if (D <= 512) DV3_LN_BWD(16, 1);
  else DV3_LN_BWD(32, 1);
#undef DV3_LN_BWD
  count_launch();
}

void gru_fwd(cudaStream_t s, float* gx, int ldgx, const float* gh, int ldgh, const float* h_prev, int ldhp,
             float* h_out, int ldho, int M, int G, void* h16) {
  launch_pdl(gru_fwd_kernel, dim3(grid_for((long long)M * G)), dim3(kThreads), 0, s, gx, ldgx, gh, ldgh, h_prev, ldhp, h_out, ldho, M, G,
                                                                 reinterpret_cast<__nv_bfloat16*>(h16));
  count_launch();
}
void gru_bwd(cudaStream_t s, const float* dh, int lddh, const float* gates, int ldg, const float* gh, int ldgh,
             const float* h_prev, int ldhp, float* dgx, int lddgx, float* dgh, int lddgh, float* dh_prev, int lddhp,
             int accumulate_dh_prev, int M, int G, void* dgx16, void* dgh16) {
  launch_pdl(gru_bwd_kernel, dim3(grid_for((long long)M * G)), dim3(kThreads), 0, s, dh, lddh, gates, ldg, gh, ldgh, h_prev, ldhp, dgx, lddgx, dgh, lddgh, dh_prev, lddhp, accumulate_dh_prev, M, G,
                                                                 reinterpret_cast<__nv_bfloat16*>(dgx16), reinterpret_cast<__nv_bfloat16*>(dgh16));
  count_launch();
}

void repr_fwd(cudaStream_t s, const float* logits, int ldl, float* probs, int ldp, const float* u, int ldu,
              int32_t* idx, int ldi, float* z, int ldz, int M, int Zc, int Zk, int mode, void* z16) {
  if (M <= 0) return;
  if (Zk == 32 && Zc % 4 == 0) {
    launch_pdl(repr_fwd4_kernel, dim3(grid_for((long long)M * (Zc / 4), kThreads / 32)), dim3(kThreads), 0, s, logits, ldl, probs, ldp, u, ldu, idx, ldi, z, ldz, M, Zc,
               mode, reinterpret_cast<__nv_bfloat16*>(z16));
    count_launch();
    return;
  }
  launch_pdl(repr_fwd_kernel, dim3(grid_for((long long)M * Zc, kThreads / 32)), dim3(kThreads), 0, s, logits, ldl, probs, ldp, u, ldu, idx, ldi, z, ldz, M, Zc, Zk, mode,
                                                                                  reinterpret_cast<__nv_bfloat16*>(z16));
  count_launch();
}
void repr_bwd(cudaStream_t s, const float* logits, int ldl, const float* dz, int lddz, const float* dprobs, int lddp,
              float* dlogits, int lddl, int M, int Zc, int Zk, void* dl16) {
  if (M <= 0) return;
  if (Zk == 32 && Zc % 4 == 0) {
    launch_pdl(repr_bwd4_kernel, dim3(grid_for((long long)M * (Zc / 4), kThreads / 32)), dim3(kThreads), 0, s, logits, ldl, dz, lddz, dprobs, lddp, dlogits, lddl, M, Zc,
               reinterpret_cast<__nv_bfloat16*>(dl16));
    count_launch();
    return;
  }
  launch_pdl(repr_bwd_kernel, dim3(grid_for((long long)M * Zc, kThreads / 32)), dim3(kThreads), 0, s, logits, ldl, dz, lddz, dprobs, lddp, dlogits, lddl, M, Zc, Zk, reinterpret_cast<__nv_bfloat16*>(dl16));
  count_launch();
}

void scan_mask_inputs(cudaStream_t s, const float* is_first, int ldf, const float* h_prev, int ldhp,
                      const float* z_prev, int ldzp, const float* h0, const float* z0, const float* actW, int ldaw,
                      float* h_in, int ldhi, float* z_in, int ldzi, float* x_pre, int ldx, int B, int G, int Z, int D) {
  scan_mask_inputs_kernel<<<grid_for((long long)B * (G + Z + D)), kThreads, 0, s>>>(is_first, ldf, h_prev, ldhp, z_prev, ldzp, h0, z0, actW, ldaw, h_in, ldhi, z_in, ldzi, x_pre, ldx, B, G, Z, D);
  count_launch();
}
void scan_mask_grads(cudaStream_t s, const float* is_first, int ldf, const float* d_in, int ldin, float* d_out,
                     int ldout, int B, int W) {
  scan_mask_grads_kernel<<<grid_for((long long)B * W), kThreads, 0, s>>>(is_first, ldf, d_in, ldin, d_out, ldout, B, W);
  count_launch();
}

}  // namespace dv3

namespace dv3 {
void colsum_acc(cudaStream_t s, const float* x, int ld, long long rows, int cols, float* out) {
  if (rows <= 0 || cols <= 0) return;
  long long gy = (rows + 63) / 64; if (gy > 148 * 2) gy = 148 * 2; if (gy < 1) gy = 1;
  dim3 grid(cdiv(cols, 32), (unsigned)gy);
  launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, s, x, ld, rows, cols, out); count_launch();
}
void scan_prepare(cudaStream_t s, const float* actions, const float* is_first, float* a_in, float* wf, int B, int T, int A) {
  long long n = (long long)B * T * (A + 1);
  scan_prepare_kernel<<<grid_for(n), kThreads, 0, s>>>(actions, is_first, a_in, wf, B, T, A); count_launch();
}
void infer_mask(cudaStream_t s, const float* f, const float* ph, const float* pz, const float* pa, const float* h0, const float* z0,
                float* h, float* z, float* a, int rows, int G, int Z, int A) {
  infer_mask_kernel<<<grid_for((long long)rows * (G + Z + A)), kThreads, 0, s>>>(f, ph, pz, pa, h0, z0, h, z, a, rows, G, Z, A); count_launch();
}
void tanh_bwd_acc(cudaStream_t s, const float* dh0, const float* h0, float* dinit, int n) {
  tanh_bwd_acc_kernel<<<grid_for(n), kThreads, 0, s>>>(dh0, h0, dinit, n); count_launch();
}
}  // namespace dv3
