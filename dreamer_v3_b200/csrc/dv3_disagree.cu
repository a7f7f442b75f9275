// Plan2Explore curiosity: the elementwise / reduction kernels around the disagree nets (models/disagree_networks.py:47-95,
// losses/disagree_loss.py:11-41). The nets themselves (MLP -> RepresentationLayer probs) run on the engine's contraction,
// LayerNorm and categorical kernels; what is here is HBM-bound row work: the disagreement statistic over the nets, the
// batch-mean subtraction, and the squared-error loss with its gradient w.r.t. the predicted probabilities.
#include "dv3_device.cuh"

namespace dv3 {

static inline int grid_warps(long long rows, int cap = 148 * 8) {
  long long b = (rows + 7) / 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// out[r] = mean_j std_n probs[n][r][j] (population stddev over the nets, tf.math.reduce_std; disagree_networks.py:65-69).
// One warp per row; every element of every net is read exactly once: a lane keeps the (<= 8) nets' values of its 4 consecutive
// probabilities in registers (128-bit loads, a warp covers 512 contiguous bytes of each net's row per iteration) and forms
// mean and variance from them. VEC = false: scalar loads (rows not 16-byte aligned or Z % 4 != 0); nets > 8: two passes.
template <bool VEC>
__global__ void __launch_bounds__(256) disagree_std_rows_kernel(const float* __restrict__ probs, size_t net_stride, int nets,
                                                                  long long rows, int Z, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_n = 1.f / (float)nets;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* p = probs + (size_t)r * Z;
    float acc = 0.f;
    if (VEC && nets <= 8) {
      for (int j = 4 * lane; j < Z; j += 128) {
        float4 v[8];
#pragma unroll
        for (int n = 0; n < 8; ++n)
          if (n < nets) v[n] = __ldg(reinterpret_cast<const float4*>(p + (size_t)n * net_stride + j));
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int n = 0; n < 8; ++n)
          if (n < nets) { m.x += v[n].x; m.y += v[n].y; m.z += v[n].z; m.w += v[n].w; }
        m.x *= inv_n; m.y *= inv_n; m.z *= inv_n; m.w *= inv_n;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int n = 0; n < 8; ++n)
          if (n < nets) {
            const float dx = v[n].x - m.x, dy = v[n].y - m.y, dz = v[n].z - m.z, dw = v[n].w - m.w;
            q.x += dx * dx; q.y += dy * dy; q.z += dz * dz; q.w += dw * dw;
          }
        acc += sqrtf(q.x * inv_n) + sqrtf(q.y * inv_n) + sqrtf(q.z * inv_n) + sqrtf(q.w * inv_n);
      }
    } else {
      for (int j = lane; j < Z; j += 32) {
        float m = 0.f;
        for (int n = 0; n < nets; ++n) m += p[(size_t)n * net_stride + j];
        m *= inv_n;
        float v = 0.f;
        for (int n = 0; n < nets; ++n) {
          const float d = p[(size_t)n * net_stride + j] - m;
          v += d * d;
        }
        acc += sqrtf(v * inv_n);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc / (float)Z;
  }
}

// ri[r] = (s[r] - mean_r s) * scale over all rows (the reference's `stddevs_B_mean -= reduce_mean(...)`, :71-73); scalars[1] =
// mean of ri over rows [first_row, rows) (DISAGREE_intrinsic_rewards: rows of t = 1..H). One block, fixed-order fp64 tree.
__global__ void __launch_bounds__(1024) disagree_intrinsic_kernel(const float* __restrict__ s, long long rows, long long first_row,
                                                                    float scale, float* __restrict__ ri, float* __restrict__ scalars) {
  __shared__ double red[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 1024) a += (double)s[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float mean = (float)(red[0] / (double)rows);
  __syncthreads();
  double b = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 1024) {
    const float v = (s[i] - mean) * scale;
    ri[i] = v;
    if (i >= first_row) b += (double)v;
  }
  red[threadIdx.x] = b;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scalars[1] = (float)(red[0] / (double)(rows - first_row));
}

// One net: mse[r] = sum_j (p[r][j] - z_next[r][j])^2, dprobs[r][j] = 2 (p - z_next) * inv_count for the rows of t < H
// (disagree_loss.py:24-40: prediction at t against the dreamed prior z of t + 1; mean over [H, N], no loss weights).
// 12 bytes per probability (two reads, one write), 128-bit accesses when the rows allow it.
template <bool VEC>
__global__ void __launch_bounds__(256) disagree_loss_rows_kernel(const float* __restrict__ probs, const float* __restrict__ z_next,
                                                                   int ldz, long long rows, int Z, float inv_count,
                                                                   float* __restrict__ dprobs, float* __restrict__ mse) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float k2 = 2.f * inv_count;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* p = probs + (size_t)r * Z;
    const float* z = z_next + (size_t)r * ldz;
    float* d = dprobs + (size_t)r * Z;
    float acc = 0.f;
    if (VEC) {
      for (int j = 4 * lane; j < Z; j += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + j));
        const float4 b = __ldg(reinterpret_cast<const float4*>(z + j));
        const float ex = a.x - b.x, ey = a.y - b.y, ez = a.z - b.z, ew = a.w - b.w;
        acc += ex * ex + ey * ey + ez * ez + ew * ew;
        *reinterpret_cast<float4*>(d + j) = make_float4(k2 * ex, k2 * ey, k2 * ez, k2 * ew);
      }
    } else {
      for (int j = lane; j < Z; j += 32) {
        const float e = p[j] - z[j];
        acc += e * e;
        d[j] = k2 * e;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) mse[r] = acc;
  }
}

// scalars[0] = sum over nets of mean over rows of mse (DISAGREE_L_total); one block, fixed-order fp64 tree
__global__ void __launch_bounds__(1024) disagree_loss_total_kernel(const float* __restrict__ mse, long long n, long long rows,
                                                                     float* __restrict__ scalars) {
  __shared__ double red[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) a += (double)mse[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) scalars[0] = (float)(red[0] / (double)rows);
}

void disagree_std_rows(cudaStream_t s, const float* probs, size_t net_stride, int nets, long long rows, int Z, float* out) {
  const bool vec = Z % 4 == 0 && net_stride % 4 == 0 && (uintptr_t)probs % 16 == 0;
  if (vec) disagree_std_rows_kernel<true><<<grid_warps(rows), 256, 0, s>>>(probs, net_stride, nets, rows, Z, out);
  else disagree_std_rows_kernel<false><<<grid_warps(rows), 256, 0, s>>>(probs, net_stride, nets, rows, Z, out);
  count_launch();
}
void disagree_intrinsic(cudaStream_t s, const float* std_rows, long long rows, long long first_row, float scale, float* ri,
                        float* scalars) {
  disagree_intrinsic_kernel<<<1, 1024, 0, s>>>(std_rows, rows, first_row, scale, ri, scalars); count_launch();
}
void disagree_loss_rows(cudaStream_t s, const float* probs, const float* z_next, int ldz, long long rows, int Z, float inv_count,
                        float* dprobs, float* mse) {
  const bool vec = Z % 4 == 0 && ldz % 4 == 0 && (uintptr_t)probs % 16 == 0 && (uintptr_t)z_next % 16 == 0 && (uintptr_t)dprobs % 16 == 0;
  if (vec) disagree_loss_rows_kernel<true><<<grid_warps(rows), 256, 0, s>>>(probs, z_next, ldz, rows, Z, inv_count, dprobs, mse);
  else disagree_loss_rows_kernel<false><<<grid_warps(rows), 256, 0, s>>>(probs, z_next, ldz, rows, Z, inv_count, dprobs, mse);
  count_launch();
}
void disagree_loss_total(cudaStream_t s, const float* mse, long long n, long long rows, float* scalars) {
  disagree_loss_total_kernel<<<1, 1024, 0, s>>>(mse, n, rows, scalars); count_launch();
}

}  // namespace dv3
