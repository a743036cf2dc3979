// Micro-benchmark (measuring aid, not part of the library): issue rate of tcgen05.mma for kind::i8 / f8f6f4 / f16
// with smem operands already resident, for several N and accumulator patterns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo16, uint32_t layout_type) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)sbo16 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// pattern: 0 = all MMAs into one accumulator, 1 = round-robin over 512/N accumulators
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int pattern, int layout_bk, long long* out, int commit_every = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    // idesc: c_format [4,6): 1 = F32, 2 = S32; a/b format [7,10)/[10,13): i8: 1 = S8; f8f6f4: 0 = E4M3; f16: 0 = F16
    uint32_t cfmt = KIND == 0 ? 2u : 1u, abfmt = KIND == 0 ? 1u : 0u;
    uint32_t idesc = (cfmt << 4) | (abfmt << 7) | (abfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t sbo16 = (8u * layout_bk) >> 4;
    uint32_t lt = layout_bk == 128 ? 2u : (layout_bk == 64 ? 4u : 6u);
    uint64_t adesc = make_desc(smem_u32(smem), sbo16, lt), bdesc = make_desc(smem_u32(smem) + 32768, sbo16, lt);
    const int nslots = pattern ? 512 / N : 1;
    long long t0 = clock64();
    int slot = 0;
    for (int i = 0; i < iters; i++) {
      if (elect_one()) {
        mma<KIND>(tmem_base + slot * N, adesc, bdesc, idesc, 1u);
        if (commit_every && (i % commit_every) == commit_every - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
      if (++slot == nslots) slot = 0;
    }
    long long t1 = clock64();
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int KIND>
void run(const char* name, int kbytes) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int bk : {128, 64, 32})
    for (int N : {32, 64, 128, 256})
      for (int pattern : {0, 1}) {
        const int iters = 2000;
        bench<KIND><<<1, 128, 80 * 1024>>>(N, iters, pattern, bk, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        double cyc = (double)h[1] / iters;
        printf("%-7s swz%-3d N=%-3d %-11s issue %.1f cyc/mma, complete %.1f cyc/mma -> %.0f MAC/clk/SM (%s)\n", name, bk, N,
               pattern ? "round-robin" : "one-acc", (double)h[0] / iters, cyc, 128.0 * N * kbytes / cyc, cudaGetErrorString(e));
      }
  for (int ce : {1, 4, 16}) {
    bench<KIND><<<1, 128, 80 * 1024>>>(64, 2000, 1, 64, d, ce);
    cudaDeviceSynchronize();
    long long h2[2];
    cudaMemcpy(h2, d, 16, cudaMemcpyDeviceToHost);
    printf("%-7s N=64 round-robin, tcgen05.commit every %d MMAs: issue %.1f cyc/mma, complete %.1f cyc/mma\n", name, ce, (double)h2[0] / 2000, (double)h2[1] / 2000);
  }
  // all SMs busy: does the rate hold chip-wide?
  bench<KIND><<<148, 128, 80 * 1024>>>(256, 2000, 0, 128, d);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-7s 148 CTAs N=256: %.1f cyc/mma\n", name, (double)h[1] / 2000);
  cudaFree(d);
}

int main() {
  run<0>("i8", 32);

  return 0;
}
