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
// One warp per row, lanes stride the Z probabilities (coalesced 128-byte reads of every net's row).
__global__ void __launch_bounds__(256) disagree_std_rows_kernel(const float* __restrict__ probs, size_t net_stride, int nets,
                                                                  long long rows, int Z, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_n = 1.f / (float)nets;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* p = probs + (size_t)r * Z;
    float acc = 0.f;
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
__global__ void __launch_bounds__(256) disagree_loss_rows_kernel(const float* __restrict__ probs, const float* __restrict__ z_next,
                                                                   int ldz, long long rows, int Z, float inv_count,
                                                                   float* __restrict__ dprobs, float* __restrict__ mse) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* p = probs + (size_t)r * Z;
    const float* z = z_next + (size_t)r * ldz;
    float* d = dprobs + (size_t)r * Z;
    float acc = 0.f;
    for (int j = lane; j < Z; j += 32) {
      const float e = p[j] - z[j];
      acc += e * e;
      d[j] = 2.f * e * inv_count;
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
  disagree_std_rows_kernel<<<grid_warps(rows), 256, 0, s>>>(probs, net_stride, nets, rows, Z, out); count_launch();
}
void disagree_intrinsic(cudaStream_t s, const float* std_rows, long long rows, long long first_row, float scale, float* ri,
                        float* scalars) {
  disagree_intrinsic_kernel<<<1, 1024, 0, s>>>(std_rows, rows, first_row, scale, ri, scalars); count_launch();
}
void disagree_loss_rows(cudaStream_t s, const float* probs, const float* z_next, int ldz, long long rows, int Z, float inv_count,
                        float* dprobs, float* mse) {
  disagree_loss_rows_kernel<<<grid_warps(rows), 256, 0, s>>>(probs, z_next, ldz, rows, Z, inv_count, dprobs, mse); count_launch();
}
void disagree_loss_total(cudaStream_t s, const float* mse, long long n, long long rows, float* scalars) {
  disagree_loss_total_kernel<<<1, 1024, 0, s>>>(mse, n, rows, scalars); count_launch();
}

}  // namespace dv3
