// Row packing of the replay sampler (reference utils/episode_replay_buffer.py:98-211) as ONE gather kernel over a
// paged, HBM-resident episode store.
//
// Store layout ("pages" of PAGE timesteps, owned by episodes, see dreamer_v3_b200/utils/episode_replay_buffer.py):
//   obs  [slots, obs_elems]  uint8 (image observations as the env emits them) or float32
//   act  [slots, A] float32, rew [slots] float32        slot = page * PAGE + timestep % PAGE
// Plan (built on the host from the reference's own index arithmetic, 16 B per output timestep):
//   int32 plan[n][4] = { obs_slot, act_slot, rew_slot (-1: reward 0.0, the "reset" reward of :134-153), flags }
//   flags bit0 is_first, bit1 is_last, bit2 is_terminated
// Output: the [B, T, ...] sample dict, float32, written straight into the engine's in.* tensors.
//
// Bound: HBM.  Per output timestep: read obs_elems (u8) or 4*obs_elems (f32) bytes, write 4*obs_elems bytes (+ 4A + 20).
// One CTA moves CHUNK = 256 threads x 16 elements of one row with 128-bit loads / stores; uint8 is normalised on the fly
// (x / 128 - 1, exact in float32: env_runner_v2.py:53-54) so image frames cost 1 B/pixel of HBM instead of 4.
#include "dv3_common.h"
#include <algorithm>

namespace dv3 {

constexpr int RP_THREADS = 256;
constexpr int RP_VEC = 16;  // elements per thread per chunk

template <bool U8, bool NORM>
__global__ void __launch_bounds__(RP_THREADS)
replay_gather_kernel(const void* __restrict__ obs_store, const float* __restrict__ act_store, const float* __restrict__ rew_store,
                     const int4* __restrict__ plan, long long obs_elems, int A,
                     float* __restrict__ obs_out, float* __restrict__ act_out, float* __restrict__ rew_out,
                     float* __restrict__ is_first, float* __restrict__ is_last, float* __restrict__ is_term) {
  const long long row = blockIdx.x;
  const int4 p = plan[row];
  const long long e0 = ((long long)blockIdx.y * RP_THREADS + threadIdx.x) * RP_VEC;
  float* dst = obs_out + row * obs_elems;
  if (U8) {
    const unsigned char* src = (const unsigned char*)obs_store + (long long)p.x * obs_elems;
    if (e0 + RP_VEC <= obs_elems && ((obs_elems & 15) == 0)) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + e0);
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 o;
        float f0 = (float)(w[q] & 0xff), f1 = (float)((w[q] >> 8) & 0xff), f2 = (float)((w[q] >> 16) & 0xff), f3 = (float)(w[q] >> 24);
        if (NORM) { f0 = f0 * 0.0078125f - 1.0f; f1 = f1 * 0.0078125f - 1.0f; f2 = f2 * 0.0078125f - 1.0f; f3 = f3 * 0.0078125f - 1.0f; }
        o.x = f0; o.y = f1; o.z = f2; o.w = f3;
        __stcs(reinterpret_cast<float4*>(dst + e0) + q, o);  // streaming store: the sample is consumed once
      }
    } else {
      for (long long e = e0; e < e0 + RP_VEC && e < obs_elems; ++e) {
        float f = (float)src[e];
        dst[e] = NORM ? f * 0.0078125f - 1.0f : f;
      }
    }
  } else {
    const float* src = (const float*)obs_store + (long long)p.x * obs_elems;
    if (e0 + RP_VEC <= obs_elems && ((obs_elems & 3) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<float4*>(dst + e0) + q, __ldg(reinterpret_cast<const float4*>(src + e0) + q));
    } else {
      for (long long e = e0; e < e0 + RP_VEC && e < obs_elems; ++e) dst[e] = src[e];
    }
  }
  if (blockIdx.y == 0) {
    for (int a = threadIdx.x; a < A; a += RP_THREADS) act_out[row * A + a] = act_store[(long long)p.y * A + a];
    if (threadIdx.x == 0) {
      rew_out[row] = p.z >= 0 ? rew_store[p.z] : 0.0f;
      is_first[row] = (p.w & 1) ? 1.0f : 0.0f;
      is_last[row] = (p.w & 2) ? 1.0f : 0.0f;
      is_term[row] = (p.w & 4) ? 1.0f : 0.0f;
    }
  }
}

// Wide-row path (image frames): a grid of a few CTAs per SM strides over the sample in units of 4 elements - one 32-bit
// load of 4 uint8 (or one float4) becomes one float4 store, so a warp reads 128 (512) contiguous bytes and writes 512
// contiguous bytes per instruction - and every thread keeps UNROLL independent loads in flight before it converts and
// stores, which is what a copy needs to approach the HBM roofline.
template <bool U8, bool NORM, int UNROLL>
__global__ void __launch_bounds__(RP_THREADS, 8)   // 8 CTAs per SM resident: <= 32 registers, 32-bit index arithmetic
replay_gather_wide_kernel(const void* __restrict__ obs_store, const float* __restrict__ act_store, const float* __restrict__ rew_store,
                          const int4* __restrict__ plan, long long n, long long obs_elems, int A,
                          float* __restrict__ obs_out, float* __restrict__ act_out, float* __restrict__ rew_out,
                          float* __restrict__ is_first, float* __restrict__ is_last, float* __restrict__ is_term) {
  const long long tid = (long long)blockIdx.x * RP_THREADS + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * RP_THREADS;
  const unsigned upr = (unsigned)(obs_elems >> 2);   // units per row; the host guarantees n * obs_elems < 2^32
  const unsigned total = (unsigned)n * upr;
  const unsigned stride = gridDim.x * RP_THREADS * UNROLL;
  for (unsigned base = blockIdx.x * RP_THREADS * UNROLL + threadIdx.x; base < total; base += stride) {
    uint4 v[U8 ? 1 : UNROLL];
    unsigned w[U8 ? UNROLL : 1];
    unsigned off[UNROLL];   // unit index in the output
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const unsigned u = base + j * RP_THREADS;
      off[j] = u;
      if (u < total) {
        const unsigned row = u / upr, col = u - row * upr;
        const long long slot = plan[row].x;
        if (U8) w[U8 ? j : 0] = __ldcs(reinterpret_cast<const unsigned*>((const unsigned char*)obs_store + slot * obs_elems) + col);
        else v[U8 ? 0 : j] = __ldcs(reinterpret_cast<const uint4*>((const float*)obs_store + slot * obs_elems) + col);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (off[j] >= total) continue;
      float4* dst = reinterpret_cast<float4*>(obs_out) + off[j];   // rows are contiguous in the output: unit u -> float4 u
      if (U8) {
        const unsigned x = w[U8 ? j : 0];
        float4 o;
        o.x = (float)(x & 0xff); o.y = (float)((x >> 8) & 0xff); o.z = (float)((x >> 16) & 0xff); o.w = (float)(x >> 24);
        if (NORM) { o.x = o.x * 0.0078125f - 1.0f; o.y = o.y * 0.0078125f - 1.0f; o.z = o.z * 0.0078125f - 1.0f; o.w = o.w * 0.0078125f - 1.0f; }
        __stcs(dst, o);
      } else {
        __stcs(reinterpret_cast<uint4*>(dst), v[U8 ? 0 : j]);
      }
    }
  }
  for (long long r = tid; r < n; r += nthreads) {  // the narrow columns of the sample dict
    const int4 p = plan[r];
    for (int a = 0; a < A; ++a) act_out[r * A + a] = act_store[(long long)p.y * A + a];
    rew_out[r] = p.z >= 0 ? rew_store[p.z] : 0.0f;
    is_first[r] = (p.w & 1) ? 1.0f : 0.0f;
    is_last[r] = (p.w & 2) ? 1.0f : 0.0f;
    is_term[r] = (p.w & 4) ? 1.0f : 0.0f;
  }
}

int replay_gather(cudaStream_t s, const void* obs_store, int obs_is_u8, int normalize_u8, long long obs_elems, const float* act_store, int A,
                  const float* rew_store, const int32_t* plan, long long n, float* obs_out, float* act_out, float* rew_out,
                  float* is_first, float* is_last, float* is_term) {
  if (n <= 0) return 0;
  if (obs_elems % 4 == 0 && obs_elems >= 1024 && n * obs_elems < (1LL << 32)) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const int unroll = obs_is_u8 ? 8 : 4;
    const long long total = n * (obs_elems / 4);
    const long long want = (total + (long long)RP_THREADS * unroll - 1) / ((long long)RP_THREADS * unroll);
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 8);  // 8 CTAs of 256 threads per SM, all resident
    const int4* p4w = reinterpret_cast<const int4*>(plan);
#define RP_WIDE(U8, NORM, UN) replay_gather_wide_kernel<U8, NORM, UN><<<grid, RP_THREADS, 0, s>>>(obs_store, act_store, rew_store, p4w, n, obs_elems, A, \
                                                                                            obs_out, act_out, rew_out, is_first, is_last, is_term)
    if (obs_is_u8 && normalize_u8) RP_WIDE(true, true, 8);
    else if (obs_is_u8) RP_WIDE(true, false, 8);
    else RP_WIDE(false, false, 4);
#undef RP_WIDE
    count_launch();
    return 0;
  }
  const long long per = (long long)RP_THREADS * RP_VEC;
  dim3 grid((unsigned)n, (unsigned)((obs_elems + per - 1) / per));
  const int4* p4 = reinterpret_cast<const int4*>(plan);
#define RP_LAUNCH(U8, NORM) replay_gather_kernel<U8, NORM><<<grid, RP_THREADS, 0, s>>>(obs_store, act_store, rew_store, p4, obs_elems, A, \
                                                                                 obs_out, act_out, rew_out, is_first, is_last, is_term)
  if (obs_is_u8 && normalize_u8) RP_LAUNCH(true, true);
  else if (obs_is_u8) RP_LAUNCH(true, false);
  else RP_LAUNCH(false, false);
#undef RP_LAUNCH
  count_launch();
  return 0;
}

// Host side of EpisodeReplayBuffer.sample (utils/episode_replay_buffer.py:98-189): the reference's sampling loop as index
// arithmetic over the episode tables. Consumes `draws` (the caller's np.random.randint(0, n_indices, size = B*10) blocks,
// concatenated) exactly as the reference consumes them: one per segment, in order.
// Returns the number of draws consumed, or -(needed) when `draws` ran out (the caller appends a block and calls again).
long long replay_plan(int B, int T, int page, const long long* chunk_cum, const int* chunk_ep, const int* chunk_ts0, long long n_chunks,
                      const int* ep_len, const unsigned char* ep_term, const long long* ep_page_off, const int* pages,
                      const long long* draws, long long n_draws, int32_t* plan) {
  long long cursor = 0;
  auto slot = [&](int ep, int t) -> int32_t { return (int32_t)((long long)pages[ep_page_off[ep] + t / page] * page + t % page); };
  for (int b = 0; b < B; ++b) {
    int fill = 0;
    while (fill < T) {
      if (cursor >= n_draws) return -(cursor + 1);
      const long long flat = draws[cursor++];
      // chunk that holds index `flat`: last c with chunk_cum[c] <= flat
      long long lo = 0, hi = n_chunks;
      while (hi - lo > 1) { const long long mid = (lo + hi) / 2; if (chunk_cum[mid] <= flat) lo = mid; else hi = mid; }
      const int ep = chunk_ep[lo];
      int ts = chunk_ts0[lo] + (int)(flat - chunk_cum[lo]);
      if (fill > 0) ts = 0;                              // continuation segments start at the reset observation (:139-141)
      const int len = ep_len[ep];
      const int seg = std::min(len - ts + 1, T - fill);  // observations ts .. len
      for (int i = 0; i < seg; ++i) {
        const int t = ts + i;
        int32_t* row = plan + ((size_t)b * T + fill + i) * 4;
        row[0] = slot(ep, t);
        row[1] = slot(ep, std::min(t, len - 1));        // last action repeated (:157-158)
        row[2] = t > 0 ? slot(ep, t - 1) : -1;           // reward that led to the observation; 0.0 at a reset
        row[3] = 0;
      }
      int32_t* first = plan + ((size_t)b * T + fill) * 4;
      if (fill > 0) first[2] = -1;
      first[3] |= 1;
      int32_t* last = plan + ((size_t)b * T + fill + seg - 1) * 4;
      last[3] |= 2;
      if (ep_term[ep] && ts + seg == len + 1) last[3] |= 4;
      fill += seg;
    }
  }
  return cursor;
}

}  // namespace dv3
