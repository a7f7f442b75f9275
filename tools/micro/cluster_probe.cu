// Micro-benchmark: cost of one "exchange + cluster barrier" round inside a single thread-block cluster
// (DSMEM broadcast of `bytes` per CTA to every peer, then barrier.cluster). Build: nvcc -arch=sm_100a -o cluster_probe cluster_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void probe(int rounds, int floats_per_cta, unsigned long long* out, float* sink) {
  extern __shared__ float buf[];   // [cluster][floats_per_cta]
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), n = cl.num_blocks();
  float acc = 0.f;
  cl.sync();
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (int r = 0; r < rounds; ++r) {
    // every thread writes its share of this CTA's block into every peer's buffer
    for (unsigned peer = 0; peer < n; ++peer) {
      float* dst = cl.map_shared_rank(buf, peer) + rank * floats_per_cta;
      for (int i = threadIdx.x; i < floats_per_cta; i += blockDim.x) dst[i] = (float)(r + i) + acc;
    }
    cl.sync();
    for (int i = threadIdx.x; i < (int)n * floats_per_cta; i += blockDim.x) acc += buf[i];
    cl.sync();   // buffer reuse
  }
  unsigned long long t1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
  if (rank == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  unsigned long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 16 * 256 * 4);
  for (int cs : {8, 16}) {
    for (int fl : {64, 256, 1024}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = (size_t)cs * fl * 4;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
      const int rounds = 2000;
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, rounds, fl, out, sink);
      cudaError_t e2 = cudaDeviceSynchronize();
      unsigned long long ns = 0;
      cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost);
      printf("cluster %2d, %5d B per CTA broadcast: %s / %s, %.3f us per round (2 barriers)\n", cs, fl * 4, cudaGetErrorString(e), cudaGetErrorString(e2), ns / 1e3 / rounds);
    }
  }
  return 0;
}
