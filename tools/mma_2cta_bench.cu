// Microbenchmark (round-2 queue, NOT yet run on hardware): tcgen05.mma.cta_group::2 issue rate for the conv kernel's MMA shapes.
//
// Single-CTA SS-mode MMAs cost max(N/2, 32 + N/4) cycles (profiles/r01_mma_microbench_smem_model.txt): the N_eff = 96 MMAs of
// rdb.conv1-4 are bound by the shared-memory operand fetch (56 cycles against 48 on the tensor pipe).  In a CTA pair each SM
// fetches its own 128 rows of A but only HALF of B, so the model predicts max(N/2, 32 + N/8): 48 cycles at N = 96, 40 at
// N = 64, 36 at N = 32 — 14 % fewer MMA cycles per 8-row tile of the N = 32 layers (440 vs 512) if the
// weight image is split per stack shape.  This measures the rate before anyone restructures the kernel around it.
//
// One cluster of 2 CTAs per SM pair; the leader issues M = 256 MMAs (both CTAs' TMEM receive 128 rows each), commits to an
// mbarrier in both CTAs (multicast) and times `iters` MMAs with clock64.  Operands are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc \
//        tools/mma_2cta_bench.cu -o build/mma_2cta_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_2cta_kernel(int n, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0 (fp16)
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers and operands are ready before the leader issues
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const bool leader_lane = ptx::elect_one();
    if (rank == 0 && leader_lane) {
      const uint64_t ad = ptx::smem_desc_sw128(base, 1024, 0), bd = ptx::smem_desc_sw128(base + 128 * 1024, 1024, 0);
      const uint32_t idesc = make_idesc_f16(256, n, true);
      for (int rep = 0; rep < 3; rep++) {
        t0 = clock64();
        for (int i = 0; i < iters; i += 12) {
          const uint32_t col = (uint32_t)(((i / 12) % 6) * 32);  // sliding accumulator block like the stacked conv
#pragma unroll
          for (int kx = 0; kx < 3; kx++)
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              const uint64_t a = ad + (uint64_t)(((i / 12) % 6) * 17408 >> 4) + (uint64_t)(kx * 8 + ks * 2);
              const uint64_t b = bd + (uint64_t)(((kx * 8192) >> 4) + ks * 2);
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + col),
                           "l"(a), "l"(b), "r"(idesc)
                           : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
        ptx::mbar_wait(ptx::smem_u32(&bar), rep & 1);
        t1 = clock64();
      }
      out[blockIdx.x >> 1] = (unsigned long long)(t1 - t0);
    } else if (rank == 1 && leader_lane) {
      for (int rep = 0; rep < 3; rep++) ptx::mbar_wait(ptx::smem_u32(&bar), rep & 1);  // the peer must outlive the pair's MMAs
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// usage: mma_2cta_bench [n]   (no argument: N = 32, 64, 96, 128, 192, one per launch; a fault ends the run)
int main(int argc, char** argv) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d;
  cudaMalloc(&d, sms * 8);
  const int smem = 160 * 1024 + 2048;
  cudaFuncSetAttribute(mma_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 12 * 2000;
  const int grid = sms / 2 * 2;
  printf("# M = 256 (CTA pair), K = 16; single-CTA model max(N/2, 32 + N/4), pair model max(N/2, 32 + N/8)\n");
  for (int n : {32, 64, 96, 128, 192}) {
    if (argc == 2 && atoi(argv[1]) != n) continue;
    cudaMemset(d, 0, sms * 8);
    mma_2cta_kernel<<<grid, 128, smem>>>(n, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=%d FAILED: %s\n", n, cudaGetErrorString(e));
      return 1;
    }
    std::vector<unsigned long long> h(grid / 2);
    cudaMemcpy(h.data(), d, grid / 2 * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += (double)v;
    avg /= h.size();
    const double one = n / 2.0 > 32 + n / 4.0 ? n / 2.0 : 32 + n / 4.0, two = n / 2.0 > 32 + n / 8.0 ? n / 2.0 : 32 + n / 8.0;
    printf("N=%-4d cyc/MMA=%-7.1f  tensor-bound=%-5.1f  1-CTA model=%-5.1f  pair model=%-5.1f\n", n, avg / iters, n / 2.0, one, two);
  }
  return 0;
}
