// Numerics probe for tcgen05.mma.cta_group::2 (CTA pair, M = 256): which rows of B does each CTA of the pair supply, and where
// do they land in D?  The rolling conv kernel (csrc/roll_kernel.cuh, PAIR mode) splits the stacked weight image [W(ky=2) |
// W(ky=1) | W(ky=0)] (N_eff = 3 * Cout rows) between the two CTAs of a pair and relies on:
//   * CTA r's shared memory provides B rows [r * N/2, (r + 1) * N/2) at the SAME descriptor offset in both CTAs,
//   * CTA r's TMEM lanes 0..127 receive D rows [128 r, 128 r + 128) (the pixels of ITS A tile), all N columns.
// Set-up: one cluster of 2 CTAs.  A (128 rows x K = 16, K-major, SWIZZLE_128B rows of 128 B) holds 1.0 in CTA 0 and 2.0 in
// CTA 1.  B row j of CTA r holds the value (64 r + j) in every K element.  One MMA (N = n, default 96) with accumulate = 0, then
// every CTA dumps lane 0 of its TMEM columns [0, n).  Expected: CTA 0 prints 16 * {0..n/2-1, 64..64+n/2-1}, CTA 1 twice that.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc \
//        tools/mma_2cta_probe.cu -o build/mma_2cta_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "ptx.cuh"

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_kernel(int n, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  uint8_t* sm = smem + (base - ptx::smem_u32(smem));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // A: 128 rows x 128 B at offset 0; B: up to 128 rows x 128 B at offset 16 KB.  Constant along K, so the swizzle is irrelevant.
  __half* a = reinterpret_cast<__half*>(sm);
  __half* b = reinterpret_cast<__half*>(sm + 16384);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) a[i] = __float2half(rank ? 2.0f : 1.0f);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) b[i] = __float2half((float)(64 * rank + (i >> 6)));
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (rank == 0 && ptx::elect_one()) {
      const uint64_t ad = ptx::smem_desc_sw128(base, 1024, 0), bd = ptx::smem_desc_sw128(base + 16384, 1024, 0);
      const uint32_t idesc = make_idesc_f16(256, n, true);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad),
                   "l"(bd), "r"(idesc)
                   : "memory");
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(&bar)),
                   "h"((uint16_t)3)
                   : "memory");
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);  // both CTAs: the multicast commit arrives on each CTA's own barrier
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < n; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tmem + c0, r);
      ptx::tmem_ld_wait();
      if (lane == 0)
        for (int i = 0; i < 16; i++) out[rank * 256 + c0 + i] = __uint_as_float(r[i]);
      if (lane == 5)
        for (int i = 0; i < 16; i++) out[512 + rank * 256 + c0 + i] = __uint_as_float(r[i]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 96;
  float* d;
  cudaMalloc(&d, 1024 * 4);
  cudaMemset(d, 0xFF, 1024 * 4);
  const int smem = 16384 * 2 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<2, 128, smem>>>(n, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAILED: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> h(1024);
  cudaMemcpy(h.data(), d, 1024 * 4, cudaMemcpyDeviceToHost);
  int ok = 1;
  for (int r = 0; r < 2; r++) {
    printf("CTA %d lane 0, D columns 0..%d (value / 16 / A):", r, n - 1);
    for (int c = 0; c < n; c++) {
      const float v = h[r * 256 + c] / 16.0f / (r ? 2.0f : 1.0f);
      printf(" %g", v);
      const float want = c < n / 2 ? (float)c : (float)(64 + c - n / 2);
      if (v != want || h[512 + r * 256 + c] != h[r * 256 + c]) ok = 0;
    }
    printf("\n");
  }
  printf(ok ? "MAPPING AS ASSUMED: CTA r supplies B rows [r*N/2, (r+1)*N/2); D columns in that order; each CTA holds its own A rows\n"
            : "MAPPING DIFFERS FROM THE ASSUMPTION\n");
  return 0;
}
