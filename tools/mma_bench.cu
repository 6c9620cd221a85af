// Microbenchmark: sustained tcgen05.mma (kind::f16, M=128, cta_group::1, SS) issue rate for the operand
// patterns the conv kernel uses.  One CTA per SM, operands are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc tools/mma_bench.cu -o build/mma_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

struct Cfg {
  int n;         // MMA N
  int shift;     // 0: A always 1024-aligned; 1: A start cycles +0/+128/+256 B (the kx taps)
  int acc_mode;  // 0: one accumulator block; 1: sliding window over 8 blocks like the stacked conv
  int kmode;     // 0: 4 K-steps walk +32 B inside the swizzled row (conv); 1: every MMA uses a fresh 1 KB-aligned A tile
  int iters;     // MMAs per measurement
  int commit;    // 1: tcgen05.commit to a (free-running) mbarrier after every 12 MMAs, like the conv stages
  int hs;        // 0: free-running; 1: full/empty mbarrier handshake with a producer warp per 12-MMA stage (wait, fence, MMAs, commit); 2: same without the tcgen05 fence; 3: wait for the next stage issued after the 8th MMA
  int nslot;     // pipeline slots of the handshake (<= 12)
  int copy;      // 1: a second warp streams 16.6 KB bulk copies global->smem (stand-in for the TMA row loads)
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(Cfg c, unsigned long long* out, const uint8_t* src) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3, cbar[8], fullb[12], emptyb[12];
  __shared__ volatile int done;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0 (fp16)
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::mbar_init(ptx::smem_u32(&bar2), 1);
    ptx::mbar_init(ptx::smem_u32(&bar3), 1);
    for (int k = 0; k < 8; k++) ptx::mbar_init(ptx::smem_u32(&cbar[k]), 1);
    for (int k = 0; k < 12; k++) { ptx::mbar_init(ptx::smem_u32(&fullb[k]), 1); ptx::mbar_init(ptx::smem_u32(&emptyb[k]), 1); }
    done = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const bool leader = ptx::elect_one();
    const uint32_t a_smem = base, b_smem = base + 128 * 1024;  // A region 128 KB, B region 64 KB
    const uint64_t ad = ptx::smem_desc_sw128(a_smem, 1024, 0), bd = ptx::smem_desc_sw128(b_smem, 1024, 0);
    const uint32_t idesc = make_idesc_f16(128, c.n, true);
    long long t0 = 0, t1 = 0;
    uint32_t fph = 0;  // phase of the full barriers (all stages advance together per wrap)
    int hstage = 0;
    for (int rep = 0; rep < 3; rep++) {
      __syncwarp();
      t0 = clock64();
      if (c.hs == 0) {
        if (leader) {
          int stage = 0, blk = 0;
          for (int i = 0; i < c.iters; i += 12) {
            const uint32_t col = c.acc_mode ? (uint32_t)((blk % 6) * 32) : 0u;
#pragma unroll
            for (int kx = 0; kx < 3; kx++)
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                uint64_t a = ad + (uint64_t)(stage * 17408 >> 4) + (uint64_t)((c.shift ? kx * 8 : 0) + ks * 2);
                const uint64_t b = bd + (uint64_t)(((kx * 8192) >> 4) + ks * 2);
                ptx::mma_f16_ss(tmem + col, a, b, idesc, 1);
              }
            if (c.commit) ptx::mma_commit(ptx::smem_u32(&bar2));
            stage = (stage + 1) % 6;
            blk++;
          }
          ptx::mma_commit(ptx::smem_u32(&bar));
        }
      } else {
        // whole warp runs the loop (like the conv kernel); producer warp 2 refills "full" when "empty" fires
        int blk = 0;
        if (c.hs == 3) ptx::mbar_wait(ptx::smem_u32(&fullb[hstage]), fph);
        for (int i = 0; i < c.iters; i += 12) {
          const uint32_t col = c.acc_mode ? (uint32_t)((blk % 6) * 32) : 0u;
          if (c.hs != 3) {
            ptx::mbar_wait(ptx::smem_u32(&fullb[hstage]), fph);
            if (c.hs == 1) ptx::tc_fence_after();
          }
          int ns = hstage + 1; uint32_t np = fph; if (ns == c.nslot) { ns = 0; np ^= 1; }
          if (leader) {
#pragma unroll
            for (int kx = 0; kx < 2; kx++)
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                uint64_t a = ad + (uint64_t)((hstage % 6) * 17408 >> 4) + (uint64_t)(kx * 8 + ks * 2);
                const uint64_t b = bd + (uint64_t)(((kx * 8192) >> 4) + ks * 2);
                ptx::mma_f16_ss(tmem + col, a, b, idesc, 1);
              }
          }
          if (c.hs == 3 && i + 12 < c.iters) { ptx::mbar_wait(ptx::smem_u32(&fullb[ns]), np); ptx::tc_fence_after(); }
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              uint64_t a = ad + (uint64_t)((hstage % 6) * 17408 >> 4) + (uint64_t)(2 * 8 + ks * 2);
              const uint64_t b = bd + (uint64_t)(((2 * 8192) >> 4) + ks * 2);
              ptx::mma_f16_ss(tmem + col, a, b, idesc, 1);
            }
            ptx::mma_commit(ptx::smem_u32(&emptyb[hstage]));
          }
          __syncwarp();
          hstage = ns; fph = np;
          blk++;
        }
        if (leader) ptx::mma_commit(ptx::smem_u32(&bar));
      }
      __syncwarp();
      ptx::mbar_wait(ptx::smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (leader) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    done = 1;
  } else if (warp == 2 && c.hs) {
    if (threadIdx.x == 64) {
      // initial fill of all 6 stages, then refill each stage when its MMAs complete
      const long long total = 3LL * (c.iters / 12);
      uint32_t eph = 0; int st = 0;
      for (long long k = 0; k < total; k++) {
        if (k >= c.nslot) { ptx::mbar_wait(ptx::smem_u32(&emptyb[st]), eph); }
        ptx::mbar_arrive(ptx::smem_u32(&fullb[st]));
        if (++st == c.nslot) { st = 0; if (k >= c.nslot) eph ^= 1; }
      }
    }
  } else if (warp == 1 && c.copy) {
    // keeps `c.copy` row-sized bulk copies in flight into a scratch ring (stand-in for the TMA row loads)
    if (threadIdx.x == 32) {
      const int K = c.copy;
      uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      unsigned long long n = 0;
      for (int k = 0; k < K; k++) {
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&cbar[k]), 16640);
        ptx::bulk_load(base + 128 * 1024 + 50 * 1024 + (k % 1) * 17408, src + ((size_t)blockIdx.x * 8 + k) * 16640, 16640, ptx::smem_u32(&cbar[k]));
      }
      while (!done) {
        for (int k = 0; k < K; k++) {
          ptx::mbar_wait(ptx::smem_u32(&cbar[k]), ph[k]);
          ph[k] ^= 1;
          n++;
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&cbar[k]), 16640);
          ptx::bulk_load(base + 128 * 1024 + 50 * 1024, src + ((size_t)blockIdx.x * 8 + k) * 16640, 16640, ptx::smem_u32(&cbar[k]));
        }
      }
      for (int k = 0; k < K; k++) ptx::mbar_wait(ptx::smem_u32(&cbar[k]), ph[k]);
      out[gridDim.x + blockIdx.x] = n;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d;
  cudaMalloc(&d, sms * 16);
  uint8_t* src;
  cudaMalloc(&src, (size_t)sms * 8 * 16640);
  cudaMemset(src, 0, (size_t)sms * 8 * 16640);
  const int smem = 200 * 1024 + 2048;
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 12 * 2000;
  printf("%-6s %-6s %-8s %-6s %-12s %-10s %-8s\n", "N", "shift", "acc", "kmode", "cyc/MMA", "ideal", "eff");
  for (int grid : {sms})
    for (int n : {32, 96, 192})
      for (int hs : {1})
        for (int nslot : {1, 2, 4, 6, 9, 12}) {
            const int commit = 1, copy = 0;
            const int shift = 1, acc = 1, kmode = 0;
            Cfg c{n, shift, acc, kmode, iters, commit, hs, nslot, copy};
            mma_bench_kernel<<<grid, 128, smem>>>(c, d, src);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("N=%d failed: %s\n", n, cudaGetErrorString(e));
              return 1;
            }
            std::vector<unsigned long long> h(2 * grid);
            cudaMemcpy(h.data(), d, 2 * grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0;
            double ncopy = 0;
            for (int i = 0; i < grid; i++) { avg += (double)h[i]; ncopy += copy ? (double)h[grid + i] : 0.0; }
            avg /= grid;
            ncopy /= grid;
            double per = avg / iters, ideal = n / 2.0;
            printf("N=%-4d hs=%d nslot=%d copies_in_flight=%d  cyc/MMA=%-8.1f ideal=%-6.1f eff=%.2f  copy B/cyc/SM=%.1f\n", n, hs, nslot, copy, per, ideal, ideal / per, ncopy * 16640.0 / (3.0 * avg));
          }
  return 0;
}
