// EXPERIMENTAL — round-2 work in progress, NOT on the product path and NOT yet run on hardware (written after the round's
// GPU budget was spent).  Reached only with option trunk_fuse = first fused conv (1-based: 4 = conv4 + conv5) set before the
// network is loaded; its pytest is opt-in (WOWSR_TEST_FUSED=1).
//
// rdb_fused_kernel — the last convs of ONE residual dense block (cnn_super_resolution.py:85-91) over all windows of a batch
// in one persistent launch that follows a host-built task list (sched_plan.h): a skewed wavefront in which conv k + 1 trails
// conv k by `lag` tiles.  Why: rdb.conv5 is the memory-bound layer (profiles/r01_launch_report_cfg2s.txt: 925 B of HBM per
// pixel, 384 of them the dense buffer that conv4 read and extended a moment earlier).  Launched right behind conv4 it finds
// that buffer in L2 — but only if "right behind" is a few tens of MB, not a whole layer over the batch (0.7 GB for 25
// windows).  The skew gives exactly that distance, and unlike the layer-major dataflow trunk (trunk_kernel.cuh: one
// dependency chain per CTA, every publish -> poll -> TMA latency exposed, tools/dataflow_sim.py) its consumers never wait.
//
// Per task the body is the one of rdb_trunk_kernel / conv3x3_tc_kernel in its 32-channel-chunk form: stacked-tap MMA issue,
// identity K-step, the two specialised epilogues, weights streamed per chunk.  Dependencies are band counters
// (sched_plan.h): an epilogue warp publishes its tile into the 8-row bands the tile covers; a TMA producer polls the bands
// the tile's halo touches, one lane per band.
#pragma once
#include "sched_plan.h"
#include "trunk_kernel.cuh"

struct FusedParams {
  int n_tasks, n_win, n_bands;
  int k_first;               // first fused conv (0-based); counters exist for convs k_first .. 3
  unsigned int band_target;  // publications that complete a band = tiles per band x epilogue warps
  int rdb, n_rdb;            // this launch's RDB (buffer parity, residual roles) and their number (the last one feeds the tail)
  int h, w;
  int n_stage;
  int fp16, last_fp16;
  uint32_t idesc_base;
  F32Layout f32;
  uint16_t* dense[2];
  uint16_t* lo;
  float* rrdb;
  const TrunkLayerW* layers;  // [n_rdb * 5]
  const SchedTask* tasks;
  unsigned int* counters;     // [(4 - k_first) * n_win * n_bands], zeroed before the launch
  int* err_flag;
  long long* trace;           // optional: CTA 0, first 64 tasks: [dep_wait_begin, dep_wait_end, mma_begin, mma_issued]
};

static_assert(sizeof(SchedTask) == 16, "SchedTask is read as one uint4");

__device__ __forceinline__ SchedTask fused_task(const FusedParams& T, int idx) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(T.tasks) + idx);
  SchedTask t;
  t.k = (uint8_t)(q.x & 255u); t.vert = (uint8_t)((q.x >> 8) & 255u); t.win = (uint16_t)(q.x >> 16);
  t.u0 = (uint16_t)(q.y & 0xFFFFu); t.v0 = (uint16_t)(q.y >> 16);
  t.dep_b0 = (uint16_t)(q.z & 0xFFFFu); t.dep_n = (uint16_t)(q.z >> 16);
  t.pub_b0 = (uint16_t)(q.w & 0xFFFFu); t.pub_n = (uint16_t)(q.w >> 16);
  return t;
}

// All `n` band counters starting at `ctr` must reach `target`: lane i polls counters i, i + 32, ...  Bounded like every wait.
__device__ __forceinline__ bool fused_dep_wait(const unsigned int* ctr, int n, unsigned int target, int lane, uint32_t& budget) {
  bool ok = true;
  for (int b = lane; b < n; b += 32) {
    bool got = false;
#pragma unroll 1
    for (uint32_t i = 0; i < budget; i++) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + b) : "memory");
      if (v >= target) { got = true; break; }
      __nanosleep(100);
    }
    ok &= got;
  }
  __syncwarp();  // orders every lane's acquire before the electing lane's proxy fence and TMA issue
  ok = __all_sync(0xFFFFFFFFu, ok);
  if (!ok) budget = 4;
  return ok;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
rdb_fused_kernel(const __grid_constant__ CUtensorMap tm_h0, const __grid_constant__ CUtensorMap tm_v0,
                 const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_v1, const FusedParams T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_smem = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t w_smem = a_smem + T.n_stage * TC_ASTAGE32;
  const uint32_t id_smem = w_smem + 2 * TRUNK_WBUF_BYTES;  // two 64 x 64 B identity tiles (input channels 0..31 / 32..63)
  const uint32_t ctl_addr = id_smem + 8192u;
  TcSmemCtl* ctl = reinterpret_cast<TcSmemCtl*>(smem + (ctl_addr - ptx::smem_u32(smem)));
  const int task0 = (int)blockIdx.x, task_step = (int)gridDim.x;
  const int n_my = task0 < T.n_tasks ? (T.n_tasks - task0 + task_step - 1) / task_step : 0;

  if (threadIdx.x == TC_WARP_TMA * 32) {
    ptx::prefetch_tmap((T.rdb & 1) ? &tm_h1 : &tm_h0);
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
  const bool tracing = T.trace != nullptr && blockIdx.x == 0;

  uint32_t wd = 1u << 18;
  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer =====================
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t aphase = 0, wcount = 0, dep_budget = 1u << 21;
    const CUtensorMap* tm_h = (T.rdb & 1) ? &tm_h1 : &tm_h0;
    const CUtensorMap* tm_v = (T.rdb & 1) ? &tm_v1 : &tm_v0;
    for (int i = 0; i < n_my; i++) {
      const SchedTask t = fused_task(T, task0 + i * task_step);
      const int N = t.k == 4 ? 64 : 32, R = t.k == 4 ? 4 : 8;
      const int n_chunks = (64 + 32 * t.k) / 32;
      if (tracing && lane == 0 && i < 64) T.trace[i * 4 + 0] = clock64();
      if (t.dep_n) {  // the bands of conv k - 1 this tile's halo touches must be complete
        const unsigned int* ctr = T.counters + ((size_t)(t.k - 1 - T.k_first) * T.n_win + t.win) * T.n_bands + t.dep_b0;
        if (!fused_dep_wait(ctr, t.dep_n, T.band_target, lane, dep_budget)) tc_fail(Pm, 41);
        ptx::fence_proxy_async_all();  // generic-proxy writes of the producers before this warp's async-proxy reads
      }
      if (tracing && lane == 0 && i < 64) T.trace[i * 4 + 1] = clock64();
      const TrunkLayerW L = T.layers[T.rdb * 5 + t.k];
      const uint8_t* wsrc = t.vert ? L.w32v : L.w32;
      const uint32_t wbytes = 3u * 3u * (uint32_t)N * 64u;
      const CUtensorMap* tm = t.vert ? tm_v : tm_h;
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
            ptx::tma_load_4d(a_smem + stage * TC_ASTAGE32, tm, ptx::smem_u32(&ctl->a_full[stage]), c * 32, (int)t.u0 - 1,
                             (int)t.v0 - 1 + 2 * sp, (int)t.win);
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
      const SchedTask t = fused_task(T, task0 + i * task_step);
      const int n_chunks = (64 + 32 * t.k) / 32;
      const int accbuf = i & 1;
      const uint32_t acc_phase = (i >> 1) & 1;
      if (!ptx::mbar_wait_hot(ptx::smem_u32(&ctl->t_empty[accbuf]), acc_phase ^ 1, S.wd)) tc_fail(Pm, 21);
      // the first stage of a task is waited for here, never prefetch-waited inside the previous task (trunk_kernel.cuh)
      if (!ptx::mbar_wait_hot(full0 + 8 * S.stage, S.aphase, S.wd)) tc_fail(Pm, 23);
      ptx::tc_fence_after();
      if (tracing && elected && i < 64) T.trace[i * 4 + 2] = clock64();
      const uint32_t acc_base = tmem_base + accbuf * 256;  // R * N = 256 columns for both layer kinds
      const bool ident_layer = t.k == 4 && (T.rdb % 3) != 0;  // rdb2 / rdb3 of an RRDB take the trunk's hi half through the MMA
      for (int c = 0; c < n_chunks; c++) {
        const uint32_t wb = wcount & 1;
        if (!ptx::mbar_wait_hot(ptx::smem_u32(&ctl->w_full[wb]), (wcount >> 1) & 1, S.wd)) tc_fail(Pm, 22);
        wcount++;
        ptx::tc_fence_after();
        const uint64_t bd = bdesc64 + (uint64_t)((wb * TRUNK_WBUF_BYTES) >> 4);
        const bool last_chunk = c == n_chunks - 1;
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
      if (tracing && elected && i < 64) T.trace[i * 4 + 3] = clock64();
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, r_first = warp >> 2;
    const int r3 = T.rdb % 3;
    for (int i = 0; i < n_my; i++) {
      const SchedTask t = fused_task(T, task0 + i * task_step);
      const int N = t.k == 4 ? 64 : 32, R = t.k == 4 ? 4 : 8;
      const bool vert = t.vert != 0;
      const int u_lim = vert ? T.h : T.w, v_lim = vert ? T.w : T.h;
      const int n = t.win, u = (int)t.u0 + q * 32 + lane;
      const int accbuf = i & 1;
      const uint32_t acc_phase = (i >> 1) & 1;
      const TrunkLayerW L = T.layers[T.rdb * 5 + t.k];
      EpiConst E;
      E.do_act = t.k < 4;
      E.slope = 0.2f;
      E.out_fp16 = (t.k == 4 && T.rdb == T.n_rdb - 1) ? T.last_fp16 != 0 : T.fp16 != 0;
      E.scale1 = 0.2f;
      E.scale2 = 0.2f;
      E.has_res2 = t.k == 4 && r3 == 2;
      E.res1 = (t.k == 4 && r3 == 0) ? T.rrdb : nullptr;
      E.res2 = E.has_res2 ? T.rrdb : nullptr;
      E.out_f32 = (t.k == 4 && r3 == 2) ? T.rrdb : nullptr;
      E.lo_in = (t.k == 4 && r3 != 0) ? T.lo : nullptr;
      E.lo_out = (t.k == 4 && r3 != 2) ? T.lo : nullptr;
      E.out_stride = 192;
      E.out_t = t.k < 4 ? T.dense[T.rdb & 1] + 64 + 32 * t.k : T.dense[(T.rdb + 1) & 1];
      const long long run_step = (vert ? (long long)T.w : 1LL) * 192;
      if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->t_full[accbuf]), acc_phase, wd)) tc_fail(Pm, 31);
      ptx::tc_fence_after();
      for (int r = r_first; r < R; r += TC_EPI_WARPS / 4) {
        const int v = (int)t.v0 + r;
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
      if (t.pub_n) {
        // publish: this warp's stores become visible at gpu scope (and to the async proxy of the consumers' TMA loads),
        // then one counter per covered band (a vertical strip tile covers up to 16 of them: one lane each)
        __threadfence();
        ptx::fence_proxy_async_all();
        __syncwarp();
        unsigned int* ctr = T.counters + ((size_t)(t.k - T.k_first) * T.n_win + t.win) * T.n_bands + t.pub_b0;
        for (int b = lane; b < (int)t.pub_n; b += 32) atomicAdd(ctr + b, 1u);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_MMA) ptx::tmem_dealloc(tmem_base, 512u);
}
