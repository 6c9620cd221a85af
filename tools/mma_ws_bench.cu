// Microbenchmark (round-2 queue, NOT yet run on hardware): can the collector buffers of tcgen05.mma remove the
// B-operand (weight) fetch from the N = 32 conv layers?
//
// Background (profiles/r01_mma_microbench_smem_model.txt): a single-CTA SS-mode tcgen05.mma M=128 costs
// max(N/2, 32 + N/4) cycles — 32 cycles to fetch A (128 rows x 32 B) and N/4 to fetch B (N rows x 32 B) from shared
// memory at 128 B/clk.  The stacked-tap N_eff = 96 MMAs of rdb.conv1-4 therefore run at 56 instead of 48 cycles, and in
// the full kernel (TMA writes competing for the same shared-memory bandwidth) at ~75.  Within a tile the SAME weight
// block B(kx, kstep) multiplies every input row, so a kstep-outer loop order could hold B in a collector buffer
// (tcgen05.mma.ws ... collector::bN::fill / ::use / ::lastuse) and fetch it once per R + 2 rows.
//
// What this measures, per mode, on one CTA per SM (cycles per MMA and the first accumulator element as a semantics probe):
//   0  plain tcgen05.mma, N_eff = n                                     (baseline: expect max(n/2, 32 + n/4))
//   1  tcgen05.mma.ws, collector::b0::fill on every MMA                 (is .ws legal for this N, and what is its base rate?)
//   2  tcgen05.mma.ws, b0::fill once, then b0::use for `reuse` - 1 MMAs (B-stationary: expect ~max(n/2, 32))
//   3  plain tcgen05.mma, collector::a::fill then a::use                (A-stationary, for completeness)
// Semantics probe: operand memory holds 1.0 everywhere except the B block that the ::use MMAs point their descriptor at,
// which holds 2.0.  D[0][0] after the run tells whether ::use consumed the collector (1.0-weights) or re-read smem (2.0).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc \
//        tools/mma_ws_bench.cu -o build/mma_ws_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

struct Cfg {
  int n;      // MMA N (N_eff of the stacked taps: 32, 64, 96, 128, 192)
  int mode;   // see above
  int reuse;  // MMAs per B (or A) fill in modes 2 / 3
  int iters;  // MMAs per measurement
};

#define WS_MMA(SUFFIX)                                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma" SUFFIX " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), \
               "l"(a), "l"(b), "r"(idesc), "r"(acc)                                                                      \
               : "memory")

__device__ __forceinline__ void mma_plain(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".cta_group::1.kind::f16"); }
__device__ __forceinline__ void mma_ws_fill(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".ws.cta_group::1.kind::f16.collector::b0::fill"); }
__device__ __forceinline__ void mma_ws_use(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".ws.cta_group::1.kind::f16.collector::b0::use"); }
__device__ __forceinline__ void mma_ws_last(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".ws.cta_group::1.kind::f16.collector::b0::lastuse"); }
__device__ __forceinline__ void mma_a_fill(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".cta_group::1.kind::f16.collector::a::fill"); }
__device__ __forceinline__ void mma_a_use(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".cta_group::1.kind::f16.collector::a::use"); }
__device__ __forceinline__ void mma_a_last(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { WS_MMA(".cta_group::1.kind::f16.collector::a::lastuse"); }

__global__ void __launch_bounds__(128, 1) mma_ws_bench_kernel(Cfg c, unsigned long long* cycles, float* probe) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* words = reinterpret_cast<uint32_t*>(smem + (base - ptx::smem_u32(smem)));
  // A region [0, 64 KB): 1.0; B region 0 [64 KB, 96 KB): 1.0; B region 1 [96 KB, 128 KB): 2.0 (fp16)
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) words[i] = i < 96 * 1024 / 4 ? 0x3c003c00u : 0x40004000u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
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
    const uint64_t ad = ptx::smem_desc_sw128(base, 1024, 0);
    const uint64_t b_one = ptx::smem_desc_sw128(base + 64 * 1024, 1024, 0), b_two = ptx::smem_desc_sw128(base + 96 * 1024, 1024, 0);
    const uint32_t idesc = make_idesc_f16(128, c.n, true);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; rep++) {
      __syncwarp();
      t0 = clock64();
      if (leader) {
        for (int i = 0; i < c.iters; i += c.reuse) {
          // a different A tile per MMA (8 rows of 1 KB apart, like consecutive input rows of a stage), one accumulator block
          for (int r = 0; r < c.reuse; r++) {
            const uint64_t a = ad + (uint64_t)(((r & 7) * 4096) >> 4);
            const uint32_t acc = (rep == 2 && i == 0 && r == 0) ? 0u : 1u;  // the last repetition starts from zero for the probe
            switch (c.mode) {
              case 0: mma_plain(tmem, a, b_one, idesc, acc); break;
              case 1: mma_ws_fill(tmem, a, b_one, idesc, acc); break;
              case 2:
                if (r == 0) mma_ws_fill(tmem, a, b_one, idesc, acc);
                else if (r == c.reuse - 1) mma_ws_last(tmem, a, b_two, idesc, acc);
                else mma_ws_use(tmem, a, b_two, idesc, acc);
                break;
              default:  // A-stationary: same A, B alternates between the two regions
                if (r == 0) mma_a_fill(tmem, ad, b_one, idesc, acc);
                else if (r == c.reuse - 1) mma_a_last(tmem, ad + 256, (r & 1) ? b_two : b_one, idesc, acc);
                else mma_a_use(tmem, ad + 256, (r & 1) ? b_two : b_one, idesc, acc);
                break;
            }
          }
        }
        ptx::mma_commit(ptx::smem_u32(&bar));
      }
      __syncwarp();
      ptx::mbar_wait(ptx::smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    ptx::tc_fence_after();
    uint32_t rr[16];
    ptx::tmem_ld16(tmem, rr);
    ptx::tmem_ld_wait();
    if (lane == 0) {
      cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
      probe[blockIdx.x] = __uint_as_float(rr[0]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

// usage: mma_ws_bench [mode n]   (no arguments: all modes; a faulting mode poisons the context, so the round-2 queue script
// runs one (mode, n) per process)
int main(int argc, char** argv) {
  const int only_mode = argc == 3 ? atoi(argv[1]) : -1, only_n = argc == 3 ? atoi(argv[2]) : -1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d;
  float* p;
  cudaMalloc(&d, sms * 8);
  cudaMalloc(&p, sms * 4);
  const int smem = 128 * 1024 + 2048;
  cudaFuncSetAttribute(mma_ws_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reuse = 10;  // R + 2 input rows of an R = 8 tile share one weight block
  const int iters = reuse * 2000;
  printf("# K = 16 per MMA, M = 128; probe = D[0][0] of the last repetition: all-1.0 weights give 16 * iters = %d, the 2.0 block gives more\n", 16 * iters);
  for (int mode = 0; mode < 4; mode++)
    for (int n : {32, 64, 96, 128, 192}) {
      if (only_mode >= 0 && (mode != only_mode || n != only_n)) continue;
      Cfg c{n, mode, reuse, iters};
      mma_ws_bench_kernel<<<sms, 128, smem>>>(c, d, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        // an illegal-instruction fault poisons the context: report and stop (the remaining modes need a fresh process)
        printf("mode=%d N=%d FAILED: %s\n", mode, n, cudaGetErrorString(e));
        return 1;
      }
      std::vector<unsigned long long> h(sms);
      std::vector<float> hp(sms);
      cudaMemcpy(h.data(), d, sms * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(hp.data(), p, sms * 4, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < sms; i++) avg += (double)h[i];
      avg /= sms;
      printf("mode=%d N=%-4d cyc/MMA=%-7.1f  tensor-bound=%-5.1f  ss-model=%-5.1f  probe=%.0f\n", mode, n, avg / iters, n / 2.0,
             n / 2.0 > 32 + n / 4.0 ? n / 2.0 : 32 + n / 4.0, hp[0]);
    }
  return 0;
}
