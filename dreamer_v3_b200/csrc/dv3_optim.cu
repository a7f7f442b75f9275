// Global-norm clip + Keras Adam (+ critic EMA) on a group's flat buffers, and the Philox noise generator.
#include "dv3_device.cuh"

namespace dv3 {
namespace {

__global__ void __launch_bounds__(256) grad_norm_kernel(const float* __restrict__ g, long long n, double* sumsq,
                                                        float* stats) {
  __shared__ double red[8];
  __shared__ float redm[8];
  double s = 0.0;
  float mx = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += (double)v * (double)v;
    mx = fmaxf(mx, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp] = s; redm[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; float m = 0.f;
    for (int i = 0; i < 8; ++i) { t += red[i]; m = fmaxf(m, redm[i]); }
    atomicAdd(sumsq, t);
    atomicMax(reinterpret_cast<int*>(stats + 1), __float_as_int(m));  // non-negative floats order as ints
  }
}

// tf.clip_by_global_norm: g * clip / max(||g||, clip); Keras Adam: eps outside the bias correction.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float clip, float lr, float eps, const double* sumsq,
                                                   float* stats, const int* step, float* __restrict__ ema_w,
                                                   float ema_decay) {
  const float gn = (float)sqrt(*sumsq);
  const float scale = clip / fmaxf(gn, clip);
  const int t = *step;
  const double b1t = pow(0.9, (double)t), b2t = pow(0.999, (double)t);
  const float lr_t = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[0] = gn;
    stats[2] = stats[1] * scale;
    stats[3] = scale;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    float mi = m[i], vi = v[i];
    mi += (gi - mi) * (1.0f - 0.9f);
    vi += (gi * gi - vi) * (1.0f - 0.999f);
    m[i] = mi; v[i] = vi;
    const float wi = w[i] - lr_t * mi / (sqrtf(vi) + eps);
    w[i] = wi;
    if (ema_w != nullptr) ema_w[i] = ema_decay * ema_w[i] + (1.0f - ema_decay) * wi;
  }
}

__global__ void adam_prologue_kernel(double* sumsq, float* stats, int* step) {
  *sumsq = 0.0;
  stats[0] = stats[1] = stats[2] = stats[3] = 0.f;
  *step += 1;
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ w, long long n, float decay) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ema[i] = decay * ema[i] + (1.0f - decay) * w[i];
}

// Philox4x32-10 counter-based generator.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__global__ void bump_kernel(unsigned long long* c) { *c += 1ull; }
__global__ void philox_kernel(float* __restrict__ out, long long n, uint64_t seed, uint64_t salt, const unsigned long long* dev_step, int normal) {
  const uint64_t stream_id = salt + (dev_step ? (uint64_t)(*dev_step) * 1000003ull : 0ull);
  const long long quads = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = (float)(c[j] >> 8) * (1.0f / 16777216.0f);  // [0,1), 24 bits
    if (normal) {  // Box-Muller on pairs
      const float r0 = sqrtf(-2.f * logf(1.f - f[0])), r1 = sqrtf(-2.f * logf(1.f - f[2]));
      const float a0 = 6.283185307179586f * f[1], a1 = 6.283185307179586f * f[3];
      f[0] = r0 * cosf(a0); f[1] = r0 * sinf(a0); f[2] = r1 * cosf(a1); f[3] = r1 * sinf(a1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = f[j];
  }
}

inline int grid_threads(long long n, int cap = 148 * 8) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

}  // namespace

void grad_norm(cudaStream_t s, const float* g, long long n, AdamState st) {
  adam_prologue_kernel<<<1, 1, 0, s>>>(st.sumsq, st.stats, st.step); count_launch();
  grad_norm_kernel<<<grid_threads(n, 148 * 4), 256, 0, s>>>(g, n, st.sumsq, st.stats); count_launch();
}
void adam_step(cudaStream_t s, float* w, const float* g, float* m, float* v, long long n, float clip, float lr,
               float eps, AdamState st, float* ema_w, float ema_decay) {
  adam_kernel<<<grid_threads(n), 256, 0, s>>>(w, g, m, v, n, clip, lr, eps, st.sumsq, st.stats, st.step, ema_w, ema_decay);
  count_launch();
}
void ema_update(cudaStream_t s, float* ema, const float* w, long long n, float decay) {
  ema_kernel<<<grid_threads(n), 256, 0, s>>>(ema, w, n, decay); count_launch();
}
void philox_fill(cudaStream_t s, float* out, long long n, uint64_t seed, uint64_t salt, const unsigned long long* dev_step, int normal) {
  if (n <= 0) return;
  philox_kernel<<<grid_threads((n + 3) / 4), 256, 0, s>>>(out, n, seed, salt, dev_step, normal); count_launch();
}
void bump_counter(cudaStream_t s, unsigned long long* ctr) { bump_kernel<<<1, 1, 0, s>>>(ctr); count_launch(); }

}  // namespace dv3
