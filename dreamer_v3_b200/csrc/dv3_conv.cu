// Patch gather/scatter for the 4x4 stride-2 "same" convolutions of CNNAtari / ConvTransposeAtari
// (components/cnn_atari.py:33-71, components/conv_transpose_atari.py:48-92). The contractions themselves
// run through the GEMM path:
//   conv fwd        : Y = im2col(X) * W[16*Cin, Cout]
//   conv bwd data   : dX = col2im(dY * W^T)            conv bwd weight: dW = im2col(X)^T * dY
//   convT fwd       : Y = col2im(X * Wt^T) (+bias)     with Wt = Keras kernel [4,4,Cout,Cin] viewed [16*Cout, Cin]
//   convT bwd data  : dX = im2col(dY) * Wt             convT bwd weight: dWt = im2col(dY)^T * X
#include <cuda_bf16.h>

#include "dv3_common.h"

namespace dv3 {
namespace {

__global__ void im2col_kernel(const float* __restrict__ img, float* __restrict__ cols, int n, int H, int W, int C,
                              __nv_bfloat16* __restrict__ cols16) {
  // thread = one (pixel row, kernel tap): 32-bit index arithmetic (the host guarantees n*Ho*Wo*16 < 2^31), C contiguous
  // channels copied per thread, so neighbouring threads write neighbouring runs of the patch row
  const int Ho = H / 2, Wo = W / 2;
  const unsigned total = (unsigned)n * Ho * Wo * 16;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned kx = i & 3, ky = (i >> 2) & 3;
    unsigned t = i >> 4;
    const unsigned ox = t % Wo; t /= Wo;
    const unsigned oy = t % Ho;
    const unsigned b = t / Ho;
    const int y = 2 * (int)oy - 1 + (int)ky, x = 2 * (int)ox - 1 + (int)kx;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    const float* src = img + (((size_t)b * H + (in ? y : 0)) * W + (in ? x : 0)) * C;
    float* dst = cols + (size_t)i * C;
    for (int c = 0; c < C; ++c) {
      const float v = in ? src[c] : 0.f;
      if (cols != nullptr) dst[c] = v;
      if (cols16 != nullptr) cols16[(size_t)i * C + c] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void col2im_kernel(const float* __restrict__ cols, float* __restrict__ img, int n, int H, int W, int C,
                              const float* __restrict__ bias, float add_const) {
  const int Ho = H / 2, Wo = W / 2;
  const unsigned total = (unsigned)n * H * W * C;   // 32-bit index arithmetic (12.6 M elements at B16 x T64 x 64 x 64 x 3)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % (unsigned)C);
    unsigned t = i / (unsigned)C;
    const int x = (int)(t % (unsigned)W); t /= (unsigned)W;
    const int y = (int)(t % (unsigned)H);
    const int b = (int)(t / (unsigned)H);
    float acc = add_const + (bias != nullptr ? bias[c] : 0.f);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        acc += cols[((((size_t)b * Ho + oy) * Wo + ox) * 16 + ky * 4 + kx) * C + c];
      }
    }
    img[i] = acc;
  }
}

// grid.y = one row of output pixels (b, oy); grid.x x 256 threads cover its (ox, tap, 4-channel group) elements: a single
// 32-bit division per thread (the flat 64-bit index decode of the first version cost more than the memory traffic)
__global__ void __launch_bounds__(256) im2col_v4_kernel(const float4* __restrict__ img, float4* __restrict__ cols, int n, int H, int W, int C4,
                                                        uint2* __restrict__ cols16) {
  const int Ho = H / 2, Wo = W / 2;
  const unsigned per_row = (unsigned)Wo * 16u * (unsigned)C4;
  const unsigned j = blockIdx.x * 256u + threadIdx.x;
  if (j >= per_row) return;
  const unsigned t = j / (unsigned)C4, c = j - t * (unsigned)C4;
  const int kx = (int)(t & 3u), ky = (int)((t >> 2) & 3u), ox = (int)(t >> 4);
  const int x = 2 * ox - 1 + kx;
  for (unsigned row = blockIdx.y; row < (unsigned)n * (unsigned)Ho; row += gridDim.y) {
    const unsigned b = row / (unsigned)Ho, oy = row - b * (unsigned)Ho;
    const int y = 2 * (int)oy - 1 + ky;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W) v = img[(((size_t)b * H + y) * W + x) * C4 + c];
    const size_t i = (size_t)row * per_row + j;
    if (cols != nullptr) cols[i] = v;   // (null: every consumer reads the bf16 patch matrix)
    if (cols16 != nullptr) {   // bf16 twin of the patch matrix for the TMA-fed contraction
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 o; o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
      cols16[i] = o;
    }
  }
}

// grid.y = one row of image pixels (b, y); grid.x x 256 threads cover its (x, 4-channel group) elements
__global__ void __launch_bounds__(256) col2im_v4_kernel(const float4* __restrict__ cols, float4* __restrict__ img, int n, int H, int W, int C4,
                                                        const float* __restrict__ bias, float add_const) {
  const int Ho = H / 2, Wo = W / 2;
  const unsigned per_row = (unsigned)W * (unsigned)C4;
  const unsigned j = blockIdx.x * 256u + threadIdx.x;
  if (j >= per_row) return;
  const int x = (int)(j / (unsigned)C4), c = (int)(j - (unsigned)x * (unsigned)C4);
  float4 acc0 = make_float4(add_const, add_const, add_const, add_const);
  if (bias != nullptr) { acc0.x += bias[4 * c]; acc0.y += bias[4 * c + 1]; acc0.z += bias[4 * c + 2]; acc0.w += bias[4 * c + 3]; }
  for (unsigned row = blockIdx.y; row < (unsigned)n * (unsigned)H; row += gridDim.y) {
    const int b = (int)(row / (unsigned)H), y = (int)(row - (unsigned)b * (unsigned)H);
    float4 acc = acc0;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        const float4 v = cols[((((size_t)b * Ho + oy) * Wo + ox) * 16 + ky * 4 + kx) * C4 + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    img[(size_t)row * per_row + j] = acc;
  }
}

inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

void im2col_k4s2(cudaStream_t s, const float* img, float* cols, int n, int H, int W, int C, void* cols16) {
  if (C % 4 == 0 && ((uintptr_t)img % 16 == 0) && ((uintptr_t)cols % 16 == 0) && ((uintptr_t)cols16 % 8 == 0))
    im2col_v4_kernel<<<dim3((unsigned)(((W / 2) * 16 * (C / 4) + 255) / 256), (unsigned)(n * (H / 2) < 65535 ? n * (H / 2) : 65535)), 256, 0, s>>>(
        reinterpret_cast<const float4*>(img), reinterpret_cast<float4*>(cols), n, H, W, C / 4, reinterpret_cast<uint2*>(cols16));
  else   // (n * Ho * Wo * 16 < 2^31 for every configuration: 16 M at B16 x T64 x 64 x 64 inputs)
    im2col_kernel<<<grid_for((long long)n * (H / 2) * (W / 2) * 16), 256, 0, s>>>(img, cols, n, H, W, C,
                                                                                   reinterpret_cast<__nv_bfloat16*>(cols16));
  count_launch();
}
void col2im_k4s2(cudaStream_t s, const float* cols, float* img, int n, int H, int W, int C, const float* bias,
                 float add_const) {
  if (C % 4 == 0 && ((uintptr_t)img % 16 == 0) && ((uintptr_t)cols % 16 == 0))
    col2im_v4_kernel<<<dim3((unsigned)((W * (C / 4) + 255) / 256), (unsigned)(n * H < 65535 ? n * H : 65535)), 256, 0, s>>>(
        reinterpret_cast<const float4*>(cols), reinterpret_cast<float4*>(img), n, H, W, C / 4, bias, add_const);
  else
    col2im_kernel<<<grid_for((long long)n * H * W * C), 256, 0, s>>>(cols, img, n, H, W, C, bias, add_const);
  count_launch();
}

}  // namespace dv3
