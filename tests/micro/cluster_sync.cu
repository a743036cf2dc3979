// Measuring aid (not a test): cost of barrier.cluster for 512-thread CTAs in clusters of 1..8, and of a DSMEM row pull.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512) k_sync(long long* out, int iters, int mode) {
  __shared__ uint32_t buf[4096];
  cg::cluster_group cl = cg::this_cluster();
  const int CL = cl.num_blocks(), rank = cl.block_rank();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
  cl.sync();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    if (mode == 0) __syncthreads();
    if (mode == 1) cl.sync();
    if (mode == 2) {  // split arrive / wait
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (mode == 3) {  // sync + pull 330 words from the owner + __syncthreads
      cl.sync();
      const int owner = it % CL;
      if (rank != owner) {
        const uint32_t* r = cl.map_shared_rank(buf, owner);
        for (int i = 2 * threadIdx.x; i < 330; i += 2 * blockDim.x)
          *reinterpret_cast<uint2*>(buf + 1024 + i) = *reinterpret_cast<const uint2*>(r + i);
      }
      __syncthreads();
    }
    if (mode == 4) {  // relaxed arrive (no release), wait
      asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
    acc += buf[(it * 7 + threadIdx.x) & 4095];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
  if (acc == 12345) out[1] = acc;
  cl.sync();
}
int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int CL : {1, 2, 4, 8}) {
    printf("cluster of %d x 512 threads:", CL);
    for (int mode = 0; mode < 5; mode++) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CL), cfg.blockDim = dim3(512);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
      cfg.attrs = at, cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_sync, d, 1000, mode);
      cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const char* nm[] = {"__syncthreads", "cluster.sync", "arrive.release+wait.acquire", "sync+pull+syncthreads", "arrive.relaxed+wait"};
      printf("  %s %lld", nm[mode], h);
    }
    printf("  (cycles)\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
