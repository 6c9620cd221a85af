// Probe (round-2 queue, NOT yet run on hardware): can a TMA tensor map replicate pixels, i.e. fold the nearest-x2 upsample of
// cnn_super_resolution.py:150-153 into the consumer conv's address generation instead of the producer's 4x replicated store?
//
// The source is an NHWC image [H][W][64] of 16-bit values.  The map describes the x-upsampled image as the 5-D tensor
// (C = 64, rep = 2 with a ZERO byte stride, W, H, N) and loads the box (64, 2, 66, 1, 1): if the encoder accepts a zero stride,
// shared memory receives 132 pixels x 128 B in pixel order 2 * xs + rep — exactly the row stage the conv kernel's UMMA
// descriptors walk (SWIZZLE_128B, checked here through the same address-bit swizzle the kernel relies on).  The y replication
// needs no map support: the producer issues one load per stage row with row >> 1.
// Prints: the CUresult of cuTensorMapEncodeTiled, whether the loaded tile equals the replicated source (incl. the zero fill
// left of x = 0 and right of x = W - 1), the same for the second-row destination the kernel uses (+17408, 1024-byte aligned), and —
// for information — for destinations that are only 128-byte aligned (+16896, +128).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sentinel2-super-resolution-poc_b200/csrc \
//        tools/tma_stride0_probe.cu -o build/tma_stride0_probe
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

constexpr int W = 200, H = 3, C = 64, BOX_W = 66;

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int xs0, int row, uint32_t dst_off, uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t base = ((ptx::smem_u32(smem) + 1023u) & ~1023u) + dst_off;
  constexpr uint32_t BYTES = BOX_W * 2 * C * 2;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&bar), BYTES);
    tma_load_5d(base, &tm, ptx::smem_u32(&bar), 0, 0, xs0, row, 0);
  }
  const bool ok = ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  __syncthreads();
  // un-swizzle (16-byte unit index ^ bits [7,10) of the absolute shared address) into a plain [pixel][channel] array
  const uint8_t* s = smem + (base - ptx::smem_u32(smem));
  for (int i = threadIdx.x; i < BOX_W * 2 * C; i += blockDim.x) {
    const int px = i / C, ch = i % C;
    const uint32_t lin = (uint32_t)px * 128u + (uint32_t)ch * 2u;
    const uint32_t abs_addr = base + lin;
    const uint32_t sw = abs_addr ^ (((abs_addr >> 7) & 7u) << 4);
    out[i] = ok ? *reinterpret_cast<const uint16_t*>(s + (sw - base)) : (uint16_t)0xDEAD;
  }
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      printf("no cuTensorMapEncodeTiled entry point\n");
      return 1;
    }
    enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  std::vector<uint16_t> src((size_t)H * W * C);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < C; c++) src[((size_t)y * W + x) * C + c] = (uint16_t)(1 + ((y * 7 + x) * 64 + c) % 60000);
  uint16_t *d_src, *d_out;
  cudaMalloc(&d_src, src.size() * 2);
  cudaMalloc(&d_out, (size_t)BOX_W * 2 * C * 2);
  cudaMemcpy(d_src, src.data(), src.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[5] = {C, 2, W, H, 1};
  cuuint64_t strides[4] = {0, (cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};  // rep dimension: zero stride
  cuuint32_t box[5] = {C, 2, BOX_W, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d_src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("cuTensorMapEncodeTiled with a zero stride on the replication dimension -> CUresult %d%s\n", (int)r,
         r == CUDA_SUCCESS ? " (accepted)" : " (rejected: the folded upsample needs another route)");
  if (r != CUDA_SUCCESS) return 0;
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<uint16_t> out((size_t)BOX_W * 2 * C);
  const int cases[][3] = {{-1, 1, 0}, {W - 60, 2, 0}, {40, 0, 17408}, {40, 0, 16896}, {40, 0, 128}};  // {first source x, row, destination offset}
  for (auto& cs : cases) {
    const int xs0 = cs[0], row = cs[1];
    probe_kernel<<<1, 128, smem>>>(tm, xs0, row, (uint32_t)cs[2], d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("xs0=%d row=%d dst+%d: launch failed: %s\n", xs0, row, cs[2], cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(out.data(), d_out, out.size() * 2, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int p = 0; p < BOX_W * 2; p++) {
      const int xs = xs0 + p / 2;
      for (int c = 0; c < C; c++) {
        const uint16_t want = (xs < 0 || xs >= W) ? 0 : src[((size_t)row * W + xs) * C + c];
        bad += out[(size_t)p * C + c] != want;
      }
    }
    printf("xs0=%4d row=%d dst+%-5d: %zu of %d values differ from the x-replicated source%s\n", xs0, row, cs[2], bad, BOX_W * 2 * C,
           bad ? "" : "  -> replication + zero fill + address-bit swizzle as needed");
  }
  return 0;
}
