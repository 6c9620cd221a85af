// EXPERIMENTAL — round-2 work in progress.  NOT on the product path (option trunk_dataflow=1, set before loading the
// network, selects it).  First hardware runs (profiles/r01_dataflow_trunk_trial.txt): output equal to the layer-by-layer
// path within the operand-rounding noise floor (uint8 within 1 LSB on 100 % of pixels; 1 / 9 / 25 windows), no speed-up
// yet: 25 windows of 276 x 276, 23 blocks: 72 ms layer-by-layer, 90 ms with groups of 2, 74-78 ms with groups of 4 —
// window-level dependencies leave bubbles at 2 windows per group, and 4 windows no longer fit the L2.
//
// rdb_trunk_kernel — the whole residual trunk (num_block x 3 RDBs x 5 convs, cnn_super_resolution.py:85-107) of a small
// GROUP of windows in ONE persistent launch, so that the RDB dense buffer of the group (2 windows of 276 x 276: 58 MB)
// stays in the 126 MB L2 across the five convs that re-read it and across RDBs (the layer-by-layer
// path moves 2 044 B of HBM per RDB pixel, 1 280 of them re-reads of that buffer).
//
// Work is a list of tasks (rdb, layer k, window of the group, tile), enumerated layer-major; CTA b runs tasks b, b + grid,
// ... in order.  There is no grid-wide barrier: a task's TMA producer waits (acquire) until every tile of layer k-1 of
// ITS window has been published by the epilogues that wrote it (one counter per (rdb, k, window)); with two windows per
// group the other window's tasks fill that wait.  Everything inside a task is the per-tile body of conv3x3_tc_kernel in
// its 32-channel-chunk form (conv_kernels.cuh): stacked-tap MMA issue, identity K-step, the two specialised epilogues.
// All layers stream their weights per task (18 / 36 KB chunks, double buffered), which makes the tensor maps layer
// independent: every conv reads 32-channel boxes of the same two 192-channel dense buffers.
#pragma once
#include "conv_kernels.cuh"

struct TrunkLayerW {  // one trunk conv: 32-channel-chunk weight images (horizontal / transposed taps) and bias
  const uint8_t* w32;
  const uint8_t* w32v;
  const float* bias;
};

struct TrunkKind {  // tile geometry of one layer kind over ONE window: [0] N = 32 / R = 8 (conv1-4), [1] N = 64 / R = 4 (conv5)
  int tiles_x, tiles_y, n_h;  // horizontal tiles over x in [0, strip_x0): runs per row, row blocks, total
  int v_runs, v_rows, n_v;    // vertical tiles of the remainder strip
  int n;                      // tasks per window and layer = n_h + n_v
};

struct TrunkParams {
  int G, win0;         // windows in this group, index of its first window in the batch buffers
  int h, w, strip_x0;  // window size, first strip column (== w: no strip)
  int n_rdb;           // 3 * num_block
  int n_stage;         // activation stage slots of TC_ASTAGE32 bytes
  int fp16, last_fp16; // operand type of the trunk / of the hi output of the very last conv5 (the tail's operand type)
  uint32_t idesc_base;
  TrunkKind kind[2];
  F32Layout f32;
  uint16_t* dense[2];  // NHWC [nb][h][w][192]; RDB j reads dense[j & 1] and writes its conv5 hi output to dense[(j + 1) & 1]
  uint16_t* lo;
  float* rrdb;
  const TrunkLayerW* layers;  // [n_rdb * 5]
  unsigned int* counters;     // [n_rdb * 5 * G] published epilogue warps per (rdb, k, window); zeroed before the launch
  int* err_flag;
};

constexpr int TRUNK_WBUF_BYTES = 3 * 3 * 64 * 64;  // largest 32-channel weight chunk (N = 64)

struct TrunkTask {
  int rdb, k, wgi, vert, u0, v0;
};

__device__ __forceinline__ int trunk_tasks_per_rdb(const TrunkParams& T) { return T.G * (4 * T.kind[0].n + T.kind[1].n); }

__device__ __forceinline__ TrunkTask trunk_decode(const TrunkParams& T, int task) {
  TrunkTask t;
  const int per_rdb = trunk_tasks_per_rdb(T);
  t.rdb = task / per_rdb;
  int rem = task - t.rdb * per_rdb;
  const int n0 = T.kind[0].n, n1 = T.kind[1].n;
  int tile;
  if (rem < 4 * T.G * n0) {
    t.k = rem / (T.G * n0);
    rem -= t.k * T.G * n0;
    t.wgi = rem / n0;
    tile = rem - t.wgi * n0;
  } else {
    t.k = 4;
    rem -= 4 * T.G * n0;
    t.wgi = rem / n1;
    tile = rem - t.wgi * n1;
  }
  const TrunkKind& K = T.kind[t.k == 4];
  const int R = t.k == 4 ? 4 : 8;
  if (tile < K.n_h) {
    t.vert = 0;
    const int vb = tile / K.tiles_x, ur = tile - vb * K.tiles_x;
    t.u0 = ur * TC_RUN;
    t.v0 = vb * R;
  } else {
    t.vert = 1;
    const int tv = tile - K.n_h;
    const int vb = tv / K.v_runs, ur = tv - vb * K.v_runs;
    t.u0 = ur * TC_RUN;
    t.v0 = T.strip_x0 + vb * R;
  }
  return t;
}

// Bounded acquire-wait on a dependency counter (a protocol bug becomes an error code, not a hung GPU).
__device__ __forceinline__ bool trunk_dep_wait(const unsigned int* ctr, unsigned int target, uint32_t& budget) {
#pragma unroll 1
  for (uint32_t i = 0; i < budget; i++) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) return true;
    __nanosleep(100);
  }
  budget = 4;
  return false;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
rdb_trunk_kernel(const __grid_constant__ CUtensorMap tm_h0, const __grid_constant__ CUtensorMap tm_v0,
                 const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_v1, const TrunkParams T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_smem = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t w_smem = a_smem + T.n_stage * TC_ASTAGE32;
  const uint32_t id_smem = w_smem + 2 * TRUNK_WBUF_BYTES;  // two 64 x 64 B identity tiles (input channels 0..31 / 32..63)
  const uint32_t ctl_addr = id_smem + 8192u;
  TcSmemCtl* ctl = reinterpret_cast<TcSmemCtl*>(smem + (ctl_addr - ptx::smem_u32(smem)));
  const int n_tasks = T.n_rdb * trunk_tasks_per_rdb(T);
  const int task0 = (int)blockIdx.x, task_step = (int)gridDim.x;
  const int n_my = task0 < n_tasks ? (n_tasks - task0 + task_step - 1) / task_step : 0;

  if (threadIdx.x == TC_WARP_TMA * 32) {
    ptx::prefetch_tmap(&tm_h0);
    ptx::prefetch_tmap(&tm_h1);
    for (int i = 0; i < T.n_stage; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->a_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->a_empty[i]), 1);
    }
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->w_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->w_empty[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->t_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->t_empty[i]), TC_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == TC_WARP_MMA) {
    ptx::tmem_alloc(ptx::smem_u32(&ctl->tmem_base), 512u);
    ptx::tmem_relinquish();
  }
  {  // B = 5 * I as two K-major SWIZZLE_64B tiles (see conv3x3_tc_kernel)
    const uint32_t five = T.fp16 ? 0x4500u : 0x40A0u;
    uint32_t* idw = reinterpret_cast<uint32_t*>(smem + (id_smem - ptx::smem_u32(smem)));
    for (int i = threadIdx.x; i < 2048; i += TC_THREADS) {
      const int t = i >> 10, wd_i = i & 1023;
      const int row = wd_i >> 4, b = (wd_i & 15) * 4;
      const int c0 = 32 * t + (((b >> 4) ^ ((row >> 1) & 3)) << 3) + ((b & 15) >> 1);
      idw[i] = (c0 == row ? five : 0u) | (c0 + 1 == row ? five << 16 : 0u);
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, ctl->tmem_base, 0);

  ConvParams Pm;  // the fields the shared issue path reads
  Pm.n_stage = T.n_stage;
  Pm.astage = TC_ASTAGE32;
  Pm.err_flag = T.err_flag;

  uint32_t wd = 1u << 18;
  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer =====================
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t aphase = 0, wcount = 0, dep_budget = 1u << 21;
    for (int i = 0; i < n_my; i++) {
      const TrunkTask t = trunk_decode(T, task0 + i * task_step);
      const int N = t.k == 4 ? 64 : 32, R = t.k == 4 ? 4 : 8;
      const int n_chunks = (64 + 32 * t.k) / 32;
      // every tile of the previous layer of this window must have been published
      if (t.k > 0 || t.rdb > 0) {
        const int pk = t.k > 0 ? t.k - 1 : 4, prdb = t.k > 0 ? t.rdb : t.rdb - 1;
        const unsigned int target = (unsigned int)T.kind[pk == 4].n * TC_EPI_WARPS;
        if (!trunk_dep_wait(T.counters + ((size_t)(prdb * 5 + pk) * T.G + t.wgi), target, dep_budget)) tc_fail(Pm, 41);
        ptx::fence_proxy_async_all();  // generic-proxy writes of the producers before this warp's async-proxy reads
      }
      const TrunkLayerW L = T.layers[t.rdb * 5 + t.k];
      const uint8_t* wsrc = t.vert ? L.w32v : L.w32;
      const uint32_t wbytes = 3u * 3u * (uint32_t)N * 64u;
      const CUtensorMap* tm = (t.rdb & 1) ? (t.vert ? &tm_v1 : &tm_h1) : (t.vert ? &tm_v0 : &tm_h0);
      for (int c = 0; c < n_chunks; c++) {
        const uint32_t b = wcount & 1, use = wcount >> 1;
        if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->w_empty[b]), (use & 1) ^ 1, wd)) tc_fail(Pm, 11);
        if (leader) {
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&ctl->w_full[b]), wbytes);
          ptx::bulk_load(w_smem + b * TRUNK_WBUF_BYTES, wsrc + (size_t)c * wbytes, wbytes, ptx::smem_u32(&ctl->w_full[b]));
        }
        wcount++;
        for (int sp = 0; sp < (R + 2) / 2; sp++) {
          if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->a_empty[stage]), aphase ^ 1, wd)) tc_fail(Pm, 12);
          if (leader) {
            ptx::mbar_arrive_expect_tx(ptx::smem_u32(&ctl->a_full[stage]), TC_ABYTES);
            ptx::tma_load_4d(a_smem + stage * TC_ASTAGE32, tm, ptx::smem_u32(&ctl->a_full[stage]), c * 32, t.u0 - 1,
                             t.v0 - 1 + 2 * sp, T.win0 + t.wgi);
          }
          if (++stage == T.n_stage) { stage = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    // ===================== MMA issuer =====================
    const bool elected = ptx::elect_one();
    const uint64_t adesc64 = ptx::smem_desc_sw64(a_smem, 512), bdesc64 = ptx::smem_desc_sw64(w_smem, 512);
    const uint64_t id_desc64 = ptx::smem_desc_sw64(id_smem, 512);
    const uint32_t full0 = ptx::smem_u32(&ctl->a_full[0]), empty0 = ptx::smem_u32(&ctl->a_empty[0]);
    IssueState S{0, 0u, 1u << 18};
    uint32_t wcount = 0;
    for (int i = 0; i < n_my; i++) {
      const TrunkTask t = trunk_decode(T, task0 + i * task_step);
      const int n_chunks = (64 + 32 * t.k) / 32;
      const int accbuf = i & 1;
      const uint32_t acc_phase = (i >> 1) & 1;
      if (!ptx::mbar_wait_hot(ptx::smem_u32(&ctl->t_empty[accbuf]), acc_phase ^ 1, S.wd)) tc_fail(Pm, 21);
      // The first stage of a task is waited for HERE, not prefetch-waited inside the previous task's last stage: the
      // next task's loads are gated by its dependencies, which may include this CTA's previous task — whose completion
      // must therefore never wait on them.
      if (!ptx::mbar_wait_hot(full0 + 8 * S.stage, S.aphase, S.wd)) tc_fail(Pm, 23);
      ptx::tc_fence_after();
      const uint32_t acc_base = tmem_base + accbuf * 256;  // R * N = 256 columns for both layer kinds
      const bool ident_layer = t.k == 4 && (t.rdb % 3) != 0;  // rdb2 / rdb3 of an RRDB take the trunk's hi half through the MMA
      for (int c = 0; c < n_chunks; c++) {
        const uint32_t wb = wcount & 1;
        if (!ptx::mbar_wait_hot(ptx::smem_u32(&ctl->w_full[wb]), (wcount >> 1) & 1, S.wd)) tc_fail(Pm, 22);
        wcount++;
        ptx::tc_fence_after();
        const uint64_t bd = bdesc64 + (uint64_t)((wb * TRUNK_WBUF_BYTES) >> 4);
        const bool last_chunk = c == n_chunks - 1;  // no prefetch-wait across a task boundary (see above)
        const uint64_t idd = id_desc64 + (uint64_t)((c & 1) * (4096 >> 4));
        if (t.k < 4) {
          if (c == 0) issue_chunk<32, 8, true, true, false, false>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base);
          else issue_chunk<32, 8, false, true, false, false>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base);
        } else if (ident_layer && c < 2) {
          if (c == 0) issue_chunk<64, 4, true, true, false, true>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base, idd);
          else issue_chunk<64, 4, false, true, false, true>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base, idd);
        } else {
          if (c == 0) issue_chunk<64, 4, true, true, false, false>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base);
          else issue_chunk<64, 4, false, true, false, false>(Pm, S, elected, elected, last_chunk, full0, empty0, adesc64, bd, acc_base, T.idesc_base);
        }
        if (elected) ptx::mma_commit(ptx::smem_u32(&ctl->w_empty[wb]));
      }
      if (elected) ptx::mma_commit(ptx::smem_u32(&ctl->t_full[accbuf]));
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, r_first = warp >> 2;
    for (int i = 0; i < n_my; i++) {
      const TrunkTask t = trunk_decode(T, task0 + i * task_step);
      const int N = t.k == 4 ? 64 : 32, R = t.k == 4 ? 4 : 8;
      const bool vert = t.vert != 0;
      const int u_lim = vert ? T.h : T.w, v_lim = vert ? T.w : T.h;
      const int n = T.win0 + t.wgi, u = t.u0 + q * 32 + lane;
      const int accbuf = i & 1;
      const uint32_t acc_phase = (i >> 1) & 1;
      const TrunkLayerW L = T.layers[t.rdb * 5 + t.k];
      EpiConst E;
      const int r3 = t.rdb % 3;
      E.do_act = t.k < 4;
      E.slope = 0.2f;
      E.out_fp16 = (t.k == 4 && t.rdb == T.n_rdb - 1) ? T.last_fp16 != 0 : T.fp16 != 0;
      E.scale1 = 0.2f;
      E.scale2 = 0.2f;
      E.has_res2 = t.k == 4 && r3 == 2;
      E.res1 = (t.k == 4 && r3 == 0) ? T.rrdb : nullptr;
      E.res2 = E.has_res2 ? T.rrdb : nullptr;
      E.out_f32 = (t.k == 4 && r3 == 2) ? T.rrdb : nullptr;
      E.lo_in = (t.k == 4 && r3 != 0) ? T.lo : nullptr;
      E.lo_out = (t.k == 4 && r3 != 2) ? T.lo : nullptr;
      E.out_stride = 192;
      E.out_t = t.k < 4 ? T.dense[t.rdb & 1] + 64 + 32 * t.k : T.dense[(t.rdb + 1) & 1];
      const long long run_step = (vert ? (long long)T.w : 1LL) * 192;
      if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->t_full[accbuf]), acc_phase, wd)) tc_fail(Pm, 31);
      ptx::tc_fence_after();
      for (int r = r_first; r < R; r += TC_EPI_WARPS / 4) {
        const int v = t.v0 + r;
        if (v >= v_lim) break;
        const int y = vert ? u : v, x = vert ? v : u;
        const bool valid = u < u_lim;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + accbuf * 256 + r * N;
        for (int c32 = 0; c32 < N / 32; c32++) {
          uint32_t rr[32];
          ptx::tmem_ld32(taddr + c32 * 32, rr);
          ptx::tmem_ld_wait();
          float vv[32];
#pragma unroll
          for (int j = 0; j < 32; j++) vv[j] = __uint_as_float(rr[j]);
          uint16_t* px = E.out_t + (((long long)n * T.h + y) * T.w + x) * 192 + c32 * 32;
          if (t.k < 4) {
            epi_plain32(E, vv, L.bias + c32 * 32, px, run_step, u, u_lim);
          } else {
            const long long fb = valid ? f32_index(T.f32, T.h, n, y, x, c32 * 32) : 0;
            const long long lb = valid ? lo_index(T.f32, T.h, n, y, x, c32 * 32) : 0;
            epi_res32<true>(E, vv, L.bias + c32 * 32, fb, lb, valid, px, run_step, u, u_lim);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->t_empty[accbuf]));
      // publish: this warp's stores become visible at gpu scope (and to the async proxy of the consumers' TMA loads)
      __threadfence();
      ptx::fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) atomicAdd(T.counters + ((size_t)(t.rdb * 5 + t.k) * T.G + t.wgi), 1u);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) ptx::tmem_dealloc(tmem_base, 512u);
}
