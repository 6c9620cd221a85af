// Product path of the two upsample convs (option tail_fold_upsample, default 1).  The zero-stride tensor map it needs is
// accepted by the driver and replicates as required (tools/tma_stride0_probe.cu, profiles/r02_queue_tma_stride0_probe.txt);
// output bit-identical to the producer-side replicated store, HR tail 7.5 -> 5.8 ms on 16 windows of 276x276
// (profiles/r02_queue_fold_upsample.txt).
//
// conv3x3_tc_ups_kernel — conv_up1 / conv_up2 of RRDBNet.forward (cnn_super_resolution.py:150-153:
// `lrelu(conv(F.interpolate(x, scale_factor=2, mode="nearest")))`) with the nearest-x2 upsample folded into the CONSUMER's
// TMA address generation: the input stays at source resolution [Nw][h][w][64] and is never written out replicated.
//   * x replication: a 5-D tensor map (C = 64, rep = 2 with a zero byte stride, W, H, N); the box (64, 2, 66, 1, 1) lands in
//     shared memory as 132 pixels x 128 B in upsampled pixel order 2 * xs + rep — a row stage like the product kernel's, one
//     pixel longer at the front, hence the +128 B on the A descriptors (mma_issuer<.., UPS = true>);
//   * y replication: the producer issues one load per stage row with source row (Y >> 1); row -1 and row h are out of bounds
//     and zero filled, which is the conv's zero padding of the UPSAMPLED image because (-1) >> 1 = -1 and (2h) >> 1 = h;
//   * everything else (stacked-tap MMA issue, TMEM double buffering, the plain epilogue with coalesced quad stores, vertical
//     tiles of the remainder strip through the transposed map) is the product kernel's.
// Against the product path (producer stores every pixel 2 x 2 times, 2 026 B per LR pixel written by conv_up1 alone,
// profiles/r01_launch_report_cfg2s.txt) the HR tail moves 29 % fewer HBM bytes and conv_body / conv_up1 store through the
// coalesced quad path.  Only the shape the two upsample convs have is built: Cin = 64 (one resident weight chunk), N = 64.
#pragma once
#include "conv_kernels.cuh"

__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_ups_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_v, const ConvParams P) {
  constexpr int N = 64, R = 4;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vert = (int)blockIdx.x >= P.grid_h;
  const CUtensorMap& tmap = vert ? tmap_v : tmap_h;
  const uint8_t* wpack = vert ? P.wpack_v : P.wpack;
  const int tile0 = vert ? (int)blockIdx.x - P.grid_h : (int)blockIdx.x;
  const int tile_step = vert ? (int)gridDim.x - P.grid_h : P.grid_h;
  const int tile_end = vert ? P.n_tiles_v : P.n_tiles;
  const uint32_t a_smem = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t w_smem = a_smem + P.n_stage * P.astage;
  const uint32_t ctl_addr = w_smem + P.w_chunk_bytes;  // one resident 64-channel weight chunk
  TcSmemCtl* ctl = reinterpret_cast<TcSmemCtl*>(smem + (ctl_addr - ptx::smem_u32(smem)));
  constexpr uint32_t tmem_cols = 512u;  // 2 x R x N

  if (threadIdx.x == TC_WARP_TMA * 32) {
    ptx::prefetch_tmap(&tmap);
    for (int i = 0; i < P.n_stage; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->a_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->a_empty[i]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&ctl->w_full[0]), 1);
    ptx::mbar_init(ptx::smem_u32(&ctl->w_empty[0]), 1);
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->t_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->t_empty[i]), TC_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == TC_WARP_MMA) {
    ptx::tmem_alloc(ptx::smem_u32(&ctl->tmem_base), tmem_cols);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < 64) ctl->bias[threadIdx.x] = P.bias[threadIdx.x];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, ctl->tmem_base, 0);

  uint32_t wd = 1u << 18;
  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer =====================
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t aphase = 0;
    for (int tile = tile0, it = 0; tile < tile_end; tile += tile_step, it++) {
      const TileCoord tc = decode_tile(P, vert, tile);
      if (it == 0 && leader) {  // the weights stay resident across this CTA's tiles
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&ctl->w_full[0]), P.w_chunk_bytes);
        ptx::bulk_load(w_smem, wpack, P.w_chunk_bytes, ptx::smem_u32(&ctl->w_full[0]));
      }
      for (int sp = 0; sp < (R + 2) / 2; sp++) {
        if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->a_empty[stage]), aphase ^ 1, wd)) tc_fail(P, 12);
        if (leader) {
          const uint32_t full = ptx::smem_u32(&ctl->a_full[stage]);
          ptx::mbar_arrive_expect_tx(full, 2 * TC_UPS_BOXB);
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const int row_up = tc.v0 - 1 + 2 * sp + half;  // row of the upsampled image (or column, for vertical tiles)
            ptx::tma_load_5d(a_smem + stage * P.astage + half * TC_UPS_ROWB, &tmap, full, 0, 0, (tc.u0 >> 1) - 1, row_up >> 1, tc.n);
          }
        }
        if (++stage == P.n_stage) { stage = 0; aphase ^= 1; }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ===================== MMA issuer =====================
    const bool elected = ptx::elect_one();
    const int n_my = tile0 < tile_end ? (tile_end - tile0 + tile_step - 1) / tile_step : 0;
    mma_issuer<N, R, false, true>(P, ctl, elected, elected, a_smem, w_smem, w_smem, tmem_base, n_my);
  } else {
    // ===================== epilogue warps: bias, LeakyReLU, 16-bit store =====================
    const int q = warp & 3, r_first = warp >> 2;
    const int u_lim = vert ? P.h : P.w, v_lim = vert ? P.w : P.h;
    const EpiConst E = make_epi_const(P);
    const long long run_step = (vert ? (long long)P.w : 1LL) * E.out_stride;
    for (int tile = tile0, it = 0; tile < tile_end; tile += tile_step, it++) {
      const TileCoord tc = decode_tile(P, vert, tile);
      const int n = tc.n, u = tc.u0 + q * 32 + lane;
      const int accbuf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->t_full[accbuf]), acc_phase, wd)) tc_fail(P, 31);
      ptx::tc_fence_after();
      for (int r = r_first; r < R; r += TC_EPI_WARPS / 4) {
        const int v = tc.v0 + r;
        if (v >= v_lim) break;
        const int y = vert ? u : v, x = vert ? v : u;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + accbuf * R * N + r * N;
        for (int c32 = 0; c32 < N / 32; c32++) {
          uint32_t rr[32];
          ptx::tmem_ld32(taddr + c32 * 32, rr);
          ptx::tmem_ld_wait();
          float vv[32];
#pragma unroll
          for (int i = 0; i < 32; i++) vv[i] = __uint_as_float(rr[i]);
          uint16_t* px = E.out_t + (((long long)n * P.h + y) * P.w + x) * E.out_stride + c32 * 32;
          epi_plain32(E, vv, ctl->bias + c32 * 32, px, run_step, u, u_lim);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->t_empty[accbuf]));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) ptx::tmem_dealloc(tmem_base, tmem_cols);
}
