// conv3x3_roll_kernel — the 3x3 convolution of the RRDBNet path (cnn_super_resolution.py:73-158) as a ROLLING implicit GEMM.
//
// Same operands and arithmetic as conv3x3_tc_kernel (conv_kernels.cuh): M = a run of 128 output pixels, one input row is
// fetched once by TMA and multiplied by the stacked taps [W(ky=2) | W(ky=1) | W(ky=0)] (N_eff = 3 * Cout), the three
// run-axis taps come from row-shifted A descriptors.  What changes is the schedule:
//
//   * A CTA walks DOWN a column of the window: input row i is multiplied once and accumulated into the TMEM slots of
//     output rows i-1, i, i+1, which form a RING of accumulator rows.  There are no tiles: no halo rows are re-read
//     (the tile kernel reads R + 2 input rows for R output rows, 1.25x for R = 8) and no tile-edge MMAs with fewer stacked
//     taps exist.  Every tcgen05.mma of a launch has the same shape (N_eff = 3 * Cout, accumulate), so the whole weight set
//     can be split in halves between the two CTAs of a pair (below) with ONE image per CTA.
//   * Ring layout: SLOTS = 512 / Cout column blocks; RING = SLOTS - 2 of them hold output rows, the last two MIRROR slots 0
//     and 1: the MMA of a centre row in slot RING-1 (RING-2) writes its third (second and third) block into the mirror instead
//     of wrapping around, and the epilogue adds real + mirror for rows in slots 0 and 1.  Output row v of a column ALWAYS lives
//     in slot v % RING (v = row index inside the window), so the order in which a pixel's partial sums are added depends on the
//     window geometry only — not on the batch, the cut points of the work list or the grid: a window gives bit-identical
//     results alone, inside any batch and on any number of GPUs (tests/test_gpu_full_size.py, tools/multigpu_check.py).
//   * The epilogue drains output rows in pairs as soon as the input row below them has been multiplied, and CLEARS the slots
//     (tcgen05.st) before handing them back, so the issuer never needs an overwrite-instead-of-accumulate MMA.
//   * Work is a host-built list of column segments (RollTask): the row sequence of all columns is cut into equal parts, one
//     per CTA (or pair), so the launch is balanced to a few rows whatever the window count; a segment costs two extra input
//     rows and two junk pairs (they absorb the taps that fall outside the segment and are cleared like any other pair), which
//     keeps the issue loop free of special cases.  Pair slots are handed back and forth with per-slot parity bits because a
//     new segment starts wherever its first row index puts it in the ring.
//   * PAIR = true: clusters of two CTAs (tcgen05 cta_group::2, M = 256).  Each CTA loads its own 128-pixel runs and drains its
//     own TMEM; the leader issues every MMA; the stacked weight rows are split in halves between the two CTAs' shared memory,
//     so each SM fetches only half of B per MMA (measured 49.3 instead of 56 cycles at N_eff = 96, profiles/r02_queue_mma_2cta_bench.txt)
//     and rdb.conv5's 221 KB of weights become RESIDENT (110 KB per CTA) instead of being streamed once per 4-row tile.
//
// Weights are resident for the whole launch.  Vertical tasks cover the remainder strip (w % 128 columns) exactly like the
// tile kernel: transposed tensor map, weights with transposed taps.
#pragma once
#include "conv_kernels.cuh"

struct RollTask {
  int n[2];   // window of pair rank 0 / 1 (single-CTA launches use entry 0); -1: dummy partner, nothing is stored
  int u0[2];  // run-axis origin of the 128-pixel run
  int v0;     // first output row (row axis) of the segment
  int rows;   // output rows in the segment
  int pad[2];
};

struct RollParams {
  const RollTask* tasks;
  const int* task_off;      // [units + 1] task range of every unit (a unit = one CTA, or one CTA pair)
  int units_h;              // units [0, units_h) run horizontal tasks, the others vertical ones
  int w_bytes;              // resident weight image per CTA
  int w_chunk_bytes;        // bytes of one 64-channel chunk in that image: 3 run-axis taps x rows_b x 128
  // per pair rank: [chunk][run-axis tap][rows_b stacked rows][64 ch] swizzled smem image, horizontal tasks / vertical tasks
  // (transposed taps).  Separate fields: a runtime index into a kernel parameter would force a local-memory copy of the struct.
  const uint8_t* wimg0; const uint8_t* wimg1;
  const uint8_t* wimg_v0; const uint8_t* wimg_v1;
  long long* unit_ns;       // [gridDim.x] (may be null): nanoseconds every CTA spent in its roles; the host rebalances the next batch's
                            // work list of this layer with them (conv.cu, balance_update)
};

// Field access by pair rank without materialising the task in local memory.
struct RollTaskView {
  int n, u0, v0, rows;
};
__device__ __forceinline__ RollTaskView roll_task(const RollTask* tasks, int t, uint32_t rank) {
  const int4* p = reinterpret_cast<const int4*>(tasks + t);
  const int4 a = __ldg(p), b = __ldg(p + 1);  // {n0, n1, u0_0, u0_1}, {v0, rows, -, -}
  RollTaskView v;
  v.n = rank ? a.y : a.x;
  v.u0 = rank ? a.w : a.z;
  v.v0 = b.x;
  v.rows = b.y;
  return v;
}

struct RollCtl {
  uint64_t a_full[TC_MAX_STAGES], a_empty[TC_MAX_STAGES];
  uint64_t t_full[8], t_empty[8];
  uint64_t w_full;
  uint32_t tmem_base;
  uint32_t pad[1];
  float bias[64];  // 16-byte aligned (read as float4)
};

// NK consecutive K-steps of ONE run-axis tap as one asm block (see mma_group_raw in conv_kernels.cuh); PAIR selects cta_group::2.
template <bool PAIR, int A0, int B0, int NK>
__device__ __forceinline__ void roll_mma_group(uint32_t col, uint64_t a, uint64_t b, uint32_t idesc) {
#define ROLL_STEP1(IA, IB) \
  "add.u64 ta, %1, %" #IA ";\n\tadd.u64 tb, %2, %" #IB ";\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %3, p;\n\t"
#define ROLL_STEP2(IA, IB) \
  "add.u64 ta, %1, %" #IA ";\n\tadd.u64 tb, %2, %" #IB ";\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], ta, tb, %3, p;\n\t"
#define ROLL_HEAD "{\n\t.reg .b64 ta, tb;\n\t.reg .pred p;\n\tsetp.eq.u32 p, 0, 0;\n\t"
  static_assert(NK == 4 || NK == 2, "");
  if constexpr (PAIR) {
    if constexpr (NK == 4) {
      asm volatile(ROLL_HEAD ROLL_STEP2(4, 5) ROLL_STEP2(6, 7) ROLL_STEP2(8, 9) ROLL_STEP2(10, 11) "}\n" ::"r"(col), "l"(a), "l"(b),
                   "r"(idesc), "n"(A0), "n"(B0), "n"(A0 + 2), "n"(B0 + 2), "n"(A0 + 4), "n"(B0 + 4), "n"(A0 + 6), "n"(B0 + 6)
                   : "memory");
    } else {
      asm volatile(ROLL_HEAD ROLL_STEP2(4, 5) ROLL_STEP2(6, 7) "}\n" ::"r"(col), "l"(a), "l"(b), "r"(idesc), "n"(A0), "n"(B0), "n"(A0 + 2),
                   "n"(B0 + 2)
                   : "memory");
    }
  } else {
    if constexpr (NK == 4) {
      asm volatile(ROLL_HEAD ROLL_STEP1(4, 5) ROLL_STEP1(6, 7) ROLL_STEP1(8, 9) ROLL_STEP1(10, 11) "}\n" ::"r"(col), "l"(a), "l"(b),
                   "r"(idesc), "n"(A0), "n"(B0), "n"(A0 + 2), "n"(B0 + 2), "n"(A0 + 4), "n"(B0 + 4), "n"(A0 + 6), "n"(B0 + 6)
                   : "memory");
    } else {
      asm volatile(ROLL_HEAD ROLL_STEP1(4, 5) ROLL_STEP1(6, 7) "}\n" ::"r"(col), "l"(a), "l"(b), "r"(idesc), "n"(A0), "n"(B0), "n"(A0 + 2),
                   "n"(B0 + 2)
                   : "memory");
    }
  }
#undef ROLL_STEP1
#undef ROLL_STEP2
#undef ROLL_HEAD
}

// Run-axis taps [KX0, KX1) of both input rows of one stage (one 64-channel chunk, or the 32-channel remainder chunk when HALF).
// ROWS_B = stacked weight rows per (chunk, tap) block in THIS CTA's image.  Descriptor units are 16 bytes: an operand row is
// 128 B (64 B when HALF), the second row of a stage starts 130 operand rows after the first, a tap shifts A by one operand row.
template <int N, bool PAIR, bool HALF, int KX0, int KX1, bool UPS = false>
__device__ __forceinline__ void roll_issue_taps(bool leader, uint32_t col0, uint32_t col1, uint64_t ad, uint64_t bd, uint32_t idesc) {
  constexpr int ROWS_B = PAIR ? 3 * N / 2 : 3 * N;
  constexpr int RU = HALF ? 4 : 8;            // operand row in descriptor units
  constexpr int NKS = HALF ? 2 : 4;
  // second input row of the stage: 130 operand rows after the first; folded upsample: its own 1024-aligned 132-pixel box
  constexpr int ROW1 = UPS ? TC_UPS_ROWB / 16 : TC_AROWS * RU;
  if (leader) {
#pragma unroll
    for (int kx = KX0; kx < KX1; kx++) {
      if (kx == 0) { roll_mma_group<PAIR, 0 * RU, 0 * ROWS_B * RU, NKS>(col0, ad, bd, idesc); roll_mma_group<PAIR, ROW1 + 0 * RU, 0 * ROWS_B * RU, NKS>(col1, ad, bd, idesc); }
      if (kx == 1) { roll_mma_group<PAIR, 1 * RU, 1 * ROWS_B * RU, NKS>(col0, ad, bd, idesc); roll_mma_group<PAIR, ROW1 + 1 * RU, 1 * ROWS_B * RU, NKS>(col1, ad, bd, idesc); }
      if (kx == 2) { roll_mma_group<PAIR, 2 * RU, 2 * ROWS_B * RU, NKS>(col0, ad, bd, idesc); roll_mma_group<PAIR, ROW1 + 2 * RU, 2 * ROWS_B * RU, NKS>(col1, ad, bd, idesc); }
    }
  }
}

// UPS: the input is at HALF the layer resolution and the nearest-x2 upsample (cnn_super_resolution.py:150-153) is folded into
// the TMA address generation exactly like conv3x3_tc_ups_kernel (ups_kernel.cuh): a zero-stride 5-D tensor map replicates
// along the run axis (one 132-pixel box per stage row, A descriptors start one pixel in), the producer halves the row index.
template <int N, int MODE, bool PAIR, bool UPS = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_roll_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_h32, const __grid_constant__ CUtensorMap tmap_v32, const ConvParams P,
                    const RollParams Q) {
  constexpr int SLOTS = N == 64 ? 8 : 16;  // accumulator row slots of N columns (N = 16: 256 TMEM columns, else 512)
  constexpr int RING = SLOTS - 2;          // slots that hold output rows; the last two mirror slots 0 and 1
  constexpr int NP = RING / 2;             // row pairs in the ring
  constexpr uint32_t TMEM_COLS = SLOTS * N;
  constexpr int ASTAGE = UPS ? 2 * TC_UPS_ROWB : TC_ASTAGE;  // bytes of one stage slot (two input rows)
  static_assert(!UPS || (N == 64 && MODE == EPI_PLAIN), "folded upsample: the two 64 -> 64 upsample convs only");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  // debug (tools/roll_trace.py): phases of CTA 0 in nanoseconds — entry, barriers + TMEM ready, weights resident, roles done, exit
  const bool tr_ph = P.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  auto tr_now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  if (tr_ph) P.trace[16] = tr_now();
  // Programmatic dependent launch (conv.cu roll_launch, option pdl): the next layer's CTAs may take SMs as this launch's CTAs
  // leave and run their set-up — barriers, TMEM, resident weights, none of which depends on a previous layer — under this
  // launch's tail.  Everything a previous layer wrote is first touched after griddep_wait() below.
  ptx::griddep_launch_dependents();
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const bool vert = unit >= Q.units_h;
  const CUtensorMap& tmap = vert ? tmap_v : tmap_h;
  const CUtensorMap& tmap32 = vert ? tmap_v32 : tmap_h32;  // 32-channel box, SWIZZLE_64B: the remainder chunk when Cin % 64 == 32
  const uint8_t* wimg = vert ? (rank ? Q.wimg_v1 : Q.wimg_v0) : (rank ? Q.wimg1 : Q.wimg0);
  const int t0 = Q.task_off[unit], t1 = Q.task_off[unit + 1];
  const uint32_t smem_base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_smem = smem_base;
  const uint32_t w_smem = a_smem + P.n_stage * ASTAGE;
  const uint32_t id_smem = w_smem + Q.w_bytes;  // identity operand of the split trunk (P.ident): 64 (pair: 32) rows x 128 B
  const uint32_t id_bytes = P.ident ? (PAIR ? 4096u : 8192u) : 0u;
  const uint32_t ctl_addr = id_smem + id_bytes;
  RollCtl* ctl = reinterpret_cast<RollCtl*>(smem + (ctl_addr - ptx::smem_u32(smem)));
  uint32_t wd = 1u << 18;  // watchdog poll budget (each poll may park up to 100 us); collapses after the first timeout

  // ---------------- set-up, phase 1: barriers, TMEM, bias, identity operand ----------------
  if (threadIdx.x == TC_WARP_TMA * 32) {
    ptx::prefetch_tmap(&tmap);
    for (int i = 0; i < P.n_stage; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->a_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->a_empty[i]), 1);
    }
    for (int i = 0; i < NP; i++) {
      ptx::mbar_init(ptx::smem_u32(&ctl->t_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->t_empty[i]), TC_EPI_WARPS * (PAIR ? 2 : 1));  // the leader's copy collects both CTAs' epilogue warps
    }
    ptx::mbar_init(ptx::smem_u32(&ctl->w_full), 1);
    ptx::fence_barrier_init();
  }
  if (warp == TC_WARP_MMA) {
    if constexpr (PAIR) { ptx::tmem_alloc_pair(ptx::smem_u32(&ctl->tmem_base), TMEM_COLS); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(ptx::smem_u32(&ctl->tmem_base), TMEM_COLS); ptx::tmem_relinquish(); }
  }
  if (threadIdx.x < 64) ctl->bias[threadIdx.x] = (int)threadIdx.x < N ? P.bias[threadIdx.x] : 0.0f;
  if (P.ident) {
    // B = 5 * I (1 / 0.2, exact in bf16 and fp16), K-major SWIZZLE_128B: operand row r holds 5 at input channel co(r).
    // Pair: this CTA supplies output channels [32 rank, 32 rank + 32) — its half of the N = 64 identity MMA.
    const uint32_t five = (P.flags & CF_FP16) ? 0x4500u : 0x40A0u;
    uint32_t* idw = reinterpret_cast<uint32_t*>(smem + (id_smem - ptx::smem_u32(smem)));
    const int n_words = (int)id_bytes / 4, co0 = PAIR ? 32 * (int)rank : 0;
    for (int wd_i = threadIdx.x; wd_i < n_words; wd_i += TC_THREADS) {
      const int row = wd_i >> 5, b = (wd_i & 31) * 4;
      const int c0 = (((b >> 4) ^ (row & 7)) << 3) + ((b & 15) >> 1);  // logical input channel of the word's low half
      const int co = co0 + row;
      idw[wd_i] = (c0 == co ? five : 0u) | (c0 + 1 == co ? five << 16 : 0u);
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync();  // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xFFFFFFFFu, ctl->tmem_base, 0);
  if (tr_ph) P.trace[17] = tr_now();

  // ---------------- set-up, phase 2: resident weights, cleared accumulator ring ----------------
  if (warp == TC_WARP_TMA && ptx::elect_one()) {
    const uint32_t bar = ptx::smem_u32(&ctl->w_full);
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)Q.w_bytes);
    for (int off = 0; off < Q.w_bytes; off += 32768) {
      const int nb = Q.w_bytes - off < 32768 ? Q.w_bytes - off : 32768;
      ptx::bulk_load(w_smem + off, wimg + off, (uint32_t)nb, bar);
    }
  }
  if (warp < TC_EPI_WARPS) {
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int c_lo = (warp >> 2) * (int)(TMEM_COLS / (TC_EPI_WARPS / 4)), c_hi = c_lo + (int)(TMEM_COLS / (TC_EPI_WARPS / 4));
    for (int c = c_lo; c < c_hi; c += 32) ptx::tmem_st32_zero(lane_base + c);
    ptx::tmem_st_wait();
  }
  if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->w_full), 0, wd)) tc_fail(P, 13);
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync();  // both CTAs' weight halves are in place before the leader issues
  ptx::tc_fence_after();
  ptx::griddep_wait();  // the previous layer has completed: activations may be read, outputs written
  if (tr_ph) P.trace[18] = tr_now();
  const long long ns0 = (Q.unit_ns != nullptr && threadIdx.x == 0) ? tr_now() : 0;
  if (P.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 148) P.trace[64 + blockIdx.x] = tr_now();

  if (warp == TC_WARP_TMA) {
    // ===================== TMA producer: one 2-row box per (stage, chunk) =====================
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t aphase = 0;
    const bool tr = P.trace != nullptr && blockIdx.x == 0;  // debug: cycle attribution of CTA 0 (tools/roll_trace.py)
    long long tr_wait = 0, tr_n = 0;
    const long long tr_t0 = tr ? clock64() : 0;
    for (int t = t0; t < t1; t++) {
      const RollTaskView T = roll_task(Q.tasks, t, rank);
      const int n = T.n < 0 ? P.Nw : T.n;  // a dummy partner reads out of bounds: zero fill
      const int u0 = T.u0;
      const int ns = (T.rows + 3) >> 1;
      for (int s = 0; s < ns; s++) {
        for (int c = 0; c < P.n_chunks; c++) {
          const long long tw0 = tr ? clock64() : 0;
          if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->a_empty[stage]), aphase ^ 1, wd)) tc_fail(P, 12);
          if (tr) { tr_wait += clock64() - tw0; tr_n++; }
          if (leader) {
            const bool half_c = (P.cin - c * 64) < 64;
            const uint32_t bytes = half_c ? TC_ABYTES : 2 * TC_ABYTES;
            const uint32_t full = ptx::smem_u32(&ctl->a_full[stage]);
            if constexpr (UPS) {
              const uint32_t bar = PAIR ? ptx::mapa(full, 0) : full;
              if (!PAIR || rank == 0) ptx::mbar_arrive_expect_tx(full, (PAIR ? 2u : 1u) * 2u * TC_UPS_BOXB);
#pragma unroll
              for (int half = 0; half < 2; half++) {
                const int row_up = T.v0 - 1 + 2 * s + half;  // row of the upsampled image; (-1) >> 1 = -1 and (2h) >> 1 = h are out of bounds: zero fill
                const uint32_t dst = a_smem + stage * ASTAGE + half * TC_UPS_ROWB;
                if constexpr (PAIR) ptx::tma_load_5d_pair(dst, &tmap, bar, 0, 0, (u0 >> 1) - 1, row_up >> 1, n);
                else ptx::tma_load_5d(dst, &tmap, bar, 0, 0, (u0 >> 1) - 1, row_up >> 1, n);
              }
            } else if constexpr (PAIR) {
              if (rank == 0) ptx::mbar_arrive_expect_tx(full, 2 * bytes);  // both CTAs' boxes complete on the leader's barrier
              ptx::tma_load_4d_pair(a_smem + stage * ASTAGE, half_c ? &tmap32 : &tmap, ptx::mapa(full, 0), c * 64, u0 - 1,
                                    T.v0 - 1 + 2 * s, n);
            } else {
              ptx::mbar_arrive_expect_tx(full, bytes);
              ptx::tma_load_4d(a_smem + stage * ASTAGE, half_c ? &tmap32 : &tmap, full, c * 64, u0 - 1, T.v0 - 1 + 2 * s, n);
            }
          }
          if (++stage == P.n_stage) { stage = 0; aphase ^= 1; }
        }
      }
    }
    if (tr && leader) { P.trace[0] = clock64() - tr_t0; P.trace[1] = tr_wait; P.trace[2] = tr_n; P.trace[3] = P.n_stage; }
  } else if (warp == TC_WARP_MMA) {
    // ===================== MMA issuer (pair: the leader CTA only) =====================
    // Runs on the whole warp with warp-uniform control flow; only the issuing instructions are predicated on one elected lane
    // (see conv3x3_tc_kernel).  Every MMA: N_eff = 3N, accumulate.  Group g = the two input rows of one stage: it completes
    // ring pair g and first touches pair g + 1.
    if (!PAIR || rank == 0) {
      const bool leader = ptx::elect_one();
      const bool fp16 = (P.flags & CF_FP16) != 0;
      const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, 3 * N, fp16);
      const uint32_t idesc_id = make_idesc_f16(PAIR ? 256 : 128, 64, fp16);
      const uint64_t adesc128 = ptx::smem_desc_sw128(a_smem + (UPS ? 128u : 0u), 1024, 0), adesc64 = ptx::smem_desc_sw64(a_smem, 512);
      const uint64_t bdesc128 = ptx::smem_desc_sw128(w_smem, 1024, 0), bdesc64 = ptx::smem_desc_sw64(w_smem, 512);
      const uint64_t id_desc = ptx::smem_desc_sw128(id_smem, 1024, 0);
      const uint32_t full0 = ptx::smem_u32(&ctl->a_full[0]), empty0 = ptx::smem_u32(&ctl->a_empty[0]);
      const uint32_t tfull0 = ptx::smem_u32(&ctl->t_full[0]), tempty0 = ptx::smem_u32(&ctl->t_empty[0]);
      int stage = 0;
      uint32_t aphase = 0;
      uint32_t alloc_par = 0;  // bit p: parity of the NEXT use of ring pair p (both roles count uses of a pair slot from 0)
      const bool tr = P.trace != nullptr && blockIdx.x == 0;
      long long tr_full = 0, tr_empty = 0, tr_groups = 0;
      const long long tr_t0 = tr ? clock64() : 0;
      // The k-th use of pair slot p may start once the epilogue has drained and cleared use k - 1 (t_empty phase k - 1).
      auto wait_free = [&](int p) {
        if (!ptx::mbar_wait_hot(tempty0 + 8 * p, ((alloc_par >> p) & 1u) ^ 1u, wd)) tc_fail(P, 21);
        alloc_par ^= 1u << p;
      };
      if (t1 > t0) {
        if (!ptx::mbar_wait_hot(full0, 0, wd)) tc_fail(P, 23);
        ptx::tc_fence_after();
      }
      for (int t = t0; t < t1; t++) {
        const RollTaskView T = roll_task(Q.tasks, t, 0);
        const int ns = (T.rows + 3) >> 1;
        // rows (v0 - 2, v0 - 1) — the head junk pair — and the segment's first real pair
        int pc = ((T.v0 + RING - 2) % RING) >> 1;
        {
          const long long tw0 = tr ? clock64() : 0;
          wait_free(pc);
          wait_free(pc + 1 == NP ? 0 : pc + 1);
          ptx::tc_fence_after();
          if (tr) tr_empty += clock64() - tw0;
        }
        for (int s = 0; s < ns; s++) {
          const uint32_t col0 = tmem_base + 2 * pc * N, col1 = col0 + N;
          const bool last_group = (t == t1 - 1) && (s == ns - 1);
          for (int c = 0; c < P.n_chunks; c++) {
            const bool half_c = (P.cin - c * 64) < 64;
            const bool last_c = c == P.n_chunks - 1;
            const uint64_t ad = (half_c ? adesc64 : adesc128) + (uint64_t)(stage * (ASTAGE >> 4));
            const uint64_t bd = (half_c ? bdesc64 : bdesc128) + (uint64_t)((c * Q.w_chunk_bytes) >> 4);
            if (half_c) roll_issue_taps<N, PAIR, true, 0, 2>(leader, col0, col1, ad, bd, idesc);
            else roll_issue_taps<N, PAIR, false, 0, 2, UPS>(leader, col0, col1, ad, bd, idesc);
            // waits for the NEXT stage, hidden behind the MMAs queued above
            int ns_stage = stage + 1;
            uint32_t ns_phase = aphase;
            if (ns_stage == P.n_stage) { ns_stage = 0; ns_phase ^= 1; }
            if (!(last_group && last_c)) {
              const long long tw1 = tr ? clock64() : 0;
              if (!ptx::mbar_wait_hot(full0 + 8 * ns_stage, ns_phase, wd)) tc_fail(P, 23);
              if (tr) tr_full += clock64() - tw1;
              ptx::tc_fence_after();
            }
            if (half_c) roll_issue_taps<N, PAIR, true, 2, 3>(leader, col0, col1, ad, bd, idesc);
            else roll_issue_taps<N, PAIR, false, 2, 3, UPS>(leader, col0, col1, ad, bd, idesc);
            if constexpr (N == 64) {
              // split trunk: centre tap of input channels [0, 64) times 5 * I adds the hi half of the residual to the centre rows
              if (P.ident && c == 0 && leader) {
                roll_mma_group<PAIR, 8, 0, 4>(col0 + N, ad, id_desc, idesc_id);
                roll_mma_group<PAIR, TC_AROWS * 8 + 8, 0, 4>(col1 + N, ad, id_desc, idesc_id);
              }
            }
            if (leader) {
              if constexpr (PAIR) ptx::mma_commit_pair(empty0 + 8 * stage);
              else ptx::mma_commit(empty0 + 8 * stage);
            }
            __syncwarp();
            stage = ns_stage;
            aphase = ns_phase;
          }
          const int pnext = pc + 1 == NP ? 0 : pc + 1;
          if (leader) {
            if constexpr (PAIR) ptx::mma_commit_pair(tfull0 + 8 * pc);
            else ptx::mma_commit(tfull0 + 8 * pc);
            if (s == ns - 1) {  // the tail junk pair this group first touched: hand it to the epilogue to be cleared
              if constexpr (PAIR) ptx::mma_commit_pair(tfull0 + 8 * pnext);
              else ptx::mma_commit(tfull0 + 8 * pnext);
            }
          }
          __syncwarp();
          // The next group of this segment first touches the pair after next.  Waited for AFTER this group is committed: a wait
          // in the middle of the group would hold back its last MMAs — and with them the accumulators the epilogue is waiting
          // for — until the epilogue has drained an older pair (measured: 4 050 instead of 2 850 cycles per group on the
          // epilogue-bound 64 -> 64 layers, profiles/r02_roll_trace_cfg2s_pair_ups.txt).
          if (s + 1 < ns) {
            const long long tw0 = tr ? clock64() : 0;
            int pn = pc + 2;
            if (pn >= NP) pn -= NP;
            wait_free(pn);
            ptx::tc_fence_after();
            if (tr) tr_empty += clock64() - tw0;
          }
          pc = pnext;
          tr_groups++;
        }
      }
      if (tr && leader) { P.trace[4] = clock64() - tr_t0; P.trace[5] = tr_full; P.trace[6] = tr_empty; P.trace[7] = tr_groups; }
    }
  } else {
    // ===================== epilogue warps: drain + clear one ring pair per group =====================
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int rsel = warp >> 2;   // which row of the pair (TC_EPI_WARPS = 8: 0 / 1)
    const int u_lim = vert ? P.h : P.w, v_lim = vert ? P.w : P.h;
    const EpiConst E = make_epi_const(P);
    const long long run_step = vert ? E.out_row : E.out_stride;  // elements between pixels of a run
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tempty_leader = PAIR ? ptx::mapa(ptx::smem_u32(&ctl->t_empty[0]), 0) : ptx::smem_u32(&ctl->t_empty[0]);
    uint32_t full_par = 0;  // bit p: parity of the next use of ring pair p
    const bool tr = P.trace != nullptr && blockIdx.x == 0 && warp == 0;
    long long tr_wait = 0, tr_work = 0, tr_zero = 0, tr_n = 0, tr_math = 0, tr_store = 0;
    const long long tr_t0 = tr ? clock64() : 0;
    for (int t = t0; t < t1; t++) {
      const RollTaskView T = roll_task(Q.tasks, t, rank);
      const int n = T.n;
      const int u = T.u0 + q * 32 + lane;
      const int ns = (T.rows + 3) >> 1;
      int pc = ((T.v0 + RING - 2) % RING) >> 1;  // row v lives in slot v % RING; the first pair holds the junk rows v0 - 2, v0 - 1
      for (int s = 0; s <= ns; s++) {            // ns pairs completed by the segment's groups + its tail junk pair
        const int f = 2 * pc;
        // Residual epilogue (rdb.conv5): fetch this warp's residual-1 pieces BEFORE waiting for the accumulators — they do not
        // depend on the MMAs, and their DRAM latency was most of an iteration (profiles/r02_roll_trace_cfg2s_pair_before_prefetch.txt:
        // 9 186 cycles per pair against 7 300 of MMA).  Registers hold both 32-channel halves of the 16-bit lo residual, or the
        // first half of an fp32 residual; everything else is pulled into L2.
        uint4 pf[8];
        if constexpr (MODE == EPI_RES) {
          static_assert(TC_EPI_WARPS == 8, "one row of the pair per epilogue warp");
          const int o = 2 * s - 2 + rsel, v = T.v0 + o;
          const bool row_ok = n >= 0 && o >= 0 && o < T.rows && v < v_lim;
          const int y = vert ? u : v, x = vert ? v : u;
          const bool valid = row_ok && u < u_lim;
          const long long fb0 = valid ? f32_index(P.f32, P.h, n, y, x, 0) : 0, lb0 = valid ? lo_index(P.f32, P.h, n, y, x, 0) : 0;
          if (E.lo_in) {
            const uint4* r = reinterpret_cast<const uint4*>(E.lo_in + lb0);   // channel c of a pixel: + (c / 8) * 32 pieces
#pragma unroll
            for (int i = 0; i < 8; i++) pf[i] = valid ? r[i * 32] : make_uint4(0u, 0u, 0u, 0u);
          } else {
            const uint4* r = reinterpret_cast<const uint4*>(E.res1 + fb0);    // channel c of a pixel: + (c / 4) * 32 pieces
#pragma unroll
            for (int i = 0; i < 8; i++) pf[i] = valid ? r[i * 32] : make_uint4(0u, 0u, 0u, 0u);
            if (valid) {  // second half (channels 32..63): 8 pieces x 512 B per warp = 32 lines, one per lane
              const float* wb = E.res1 + fb0 - (u & 31) * 4 + 8 * 32 * 4;
              ptx::prefetch_l2(wb + lane * 32);
            }
          }
          if (E.has_res2 && valid) {  // 2 x 32 lines per warp
            const float* wb = E.res2 + fb0 - (u & 31) * 4;
            ptx::prefetch_l2(wb + lane * 32);
            ptx::prefetch_l2(wb + 8 * 32 * 4 + lane * 32);
          }
        }
        const long long te0 = tr ? clock64() : 0;
        if (!ptx::mbar_wait_wd(ptx::smem_u32(&ctl->t_full[pc]), (full_par >> pc) & 1u, wd)) tc_fail(P, 31);
        full_par ^= 1u << pc;
        const long long te1 = tr ? clock64() : 0;
        ptx::tc_fence_after();
        // The pair goes back to the issuer as soon as this warp's LAST accumulator block is in registers and cleared — before
        // the epilogue arithmetic and the stores: the ring then only has to cover the TMEM read, not the store latency (the
        // three-pair ring of the 64-output layers otherwise couples the MMA stream to the drain time of a pair).
        long long te2 = 0;
        auto hand_back = [&]() {
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty_leader + 8 * pc);
            else ptx::mbar_arrive_relaxed(tempty_leader + 8 * pc);
          }
          if (tr) te2 = clock64();
        };
        static_assert(TC_EPI_WARPS == 8, "one row of a ring pair per epilogue warp");
        {
          const int rr_i = rsel;
          const int o = 2 * s - 2 + rr_i;  // output row of the segment held by slot f + rr_i
          const int slot = f + rr_i;
          const int v = T.v0 + o;
          const bool row_ok = n >= 0 && o >= 0 && o < T.rows && v < v_lim && !(P.flags & CF_DBG_NO_EPI);
          const uint32_t taddr = lane_base + slot * N;
          const uint32_t maddr = lane_base + (RING + slot) * N;  // mirror (slots 0 and 1 only)
          const int y = vert ? u : v, x = vert ? v : u;
          const bool valid = u < u_lim;
          if constexpr (N == 64 && MODE == EPI_PLAIN) {
            // both 32-channel halves into registers first, hand the pair back, then the arithmetic of both (two independent
            // instruction streams for the scheduler to interleave)
            float va[32], vb[32];
            if (row_ok) {
              uint32_t ra[32], rb[32];
              ptx::tmem_ld32(taddr, ra);
              ptx::tmem_ld32(taddr + 32, rb);
              if (slot < 2) {
                uint32_t ma[32], mb[32];
                ptx::tmem_ld32(maddr, ma);
                ptx::tmem_ld32(maddr + 32, mb);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) {
                  va[i] = __fadd_rn(__uint_as_float(ra[i]), __uint_as_float(ma[i]));
                  vb[i] = __fadd_rn(__uint_as_float(rb[i]), __uint_as_float(mb[i]));
                }
              } else {
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) { va[i] = __uint_as_float(ra[i]); vb[i] = __uint_as_float(rb[i]); }
              }
            }
            ptx::tmem_st32_zero(taddr);
            ptx::tmem_st32_zero(taddr + 32);
            if (slot < 2) { ptx::tmem_st32_zero(maddr); ptx::tmem_st32_zero(maddr + 32); }
            hand_back();
            if (row_ok) {
              uint16_t* px = E.out_t + ((long long)n * P.h + y) * E.out_row + (long long)x * E.out_stride;
              epi_plain32(E, va, ctl->bias, px, run_step, u, u_lim);
              epi_plain32(E, vb, ctl->bias + 32, px + 32, run_step, u, u_lim);
            }
          } else if constexpr (N >= 32) {
#pragma unroll
            for (int c32 = 0; c32 < N / 32; c32++) {
              float vv[32];
              if (row_ok) {
                uint32_t rr[32];
                ptx::tmem_ld32(taddr + c32 * 32, rr);
                if (slot < 2) {
                  uint32_t mm[32];
                  ptx::tmem_ld32(maddr + c32 * 32, mm);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; i++) vv[i] = __fadd_rn(__uint_as_float(rr[i]), __uint_as_float(mm[i]));
                } else {
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; i++) vv[i] = __uint_as_float(rr[i]);
                }
              }
              ptx::tmem_st32_zero(taddr + c32 * 32);
              if (slot < 2) ptx::tmem_st32_zero(maddr + c32 * 32);
              if (c32 == N / 32 - 1) hand_back();
              if (row_ok) {
                if constexpr (MODE == EPI_GENERIC) {
                  epilogue_pixel<32, true>(P, n, y, x, c32 * 32, vv, ctl->bias, valid, vert ? (long long)P.w : 1LL, u, u_lim);
                } else {
                  uint16_t* px = E.out_t + ((long long)n * P.h + y) * E.out_row + (long long)x * E.out_stride + c32 * 32;
                  if constexpr (MODE == EPI_PLAIN) {
                    if (tr) {  // traced warp only: the same arithmetic with a time stamp between the math and the stores
                      epi_bias32(vv, ctl->bias + c32 * 32);
                      if (E.do_act) {
#pragma unroll
                        for (int i = 0; i < 32; i++) vv[i] = fmaxf(vv[i], __fmul_rn(vv[i], E.slope));
                      }
                      uint32_t pk[16];
                      epi_pack16(vv, E.out_fp16, pk);
                      const long long tm = clock64();
                      epi_store_quad(pk, px, run_step, u, u_lim);
                      tr_math += tm - te2;
                      tr_store += clock64() - tm;
                    } else {
                      epi_plain32(E, vv, ctl->bias + c32 * 32, px, run_step, u, u_lim);
                    }
                  } else {
                    const long long fb = valid ? f32_index(P.f32, P.h, n, y, x, c32 * 32) : 0;
                    const long long lb = valid ? lo_index(P.f32, P.h, n, y, x, c32 * 32) : 0;
                    float t[32];
                    if (E.lo_in) {
                      epi_res_unpack(E, pf + 4 * c32, t);
                    } else if (c32 == 0) {
                      epi_res_unpack(E, pf, t);
                    } else {
                      uint4 w2[8];
                      epi_res_fetch(E, fb, lb, valid, w2);
                      epi_res_unpack(E, w2, t);
                    }
                    epi_res32_t(E, vv, t, ctl->bias + c32 * 32, fb, lb, valid, px, run_step, u, u_lim);
                  }
                }
              }
            }
          } else {
            float vv[16];
            if (row_ok) {
              uint32_t rr[16];
              ptx::tmem_ld16(taddr, rr);
              if (slot < 2) {
                uint32_t mm[16];
                ptx::tmem_ld16(maddr, mm);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) vv[i] = __fadd_rn(__uint_as_float(rr[i]), __uint_as_float(mm[i]));
              } else {
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) vv[i] = __uint_as_float(rr[i]);
              }
            }
            ptx::tmem_st16_zero(taddr);
            if (slot < 2) ptx::tmem_st16_zero(maddr);
            hand_back();
            if (row_ok) epilogue_pixel<16, true>(P, n, y, x, 0, vv, ctl->bias, valid, vert ? (long long)P.w : 1LL, u, u_lim);
          }
        }
        if (tr) { tr_wait += te1 - te0; tr_work += te2 - te1; tr_zero += clock64() - te2; tr_n++; }  // work = until the hand-back, zero = the rest
        if (++pc == NP) pc = 0;
      }
    }
    if (tr && lane == 0) {
      P.trace[8] = clock64() - tr_t0; P.trace[9] = tr_wait; P.trace[10] = tr_work; P.trace[11] = tr_zero; P.trace[12] = tr_n;
      P.trace[13] = tr_math; P.trace[14] = tr_store;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tr_ph) P.trace[19] = tr_now();
  if (Q.unit_ns != nullptr && threadIdx.x == 0) Q.unit_ns[blockIdx.x] = tr_now() - ns0;
  if constexpr (PAIR) ptx::cluster_sync();  // the peer's shared memory and barriers stay alive until both CTAs are done
  if (warp == TC_WARP_MMA) {
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tr_ph) P.trace[20] = tr_now();
  // debug: when every CTA of the traced launch starts its roles and leaves (ns after CTA 0's entry is subtracted by the tool):
  // load balance of the work list, horizontal against vertical units
  if (P.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 148) {
    P.trace[64 + 148 + blockIdx.x] = tr_now();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    P.trace[64 + 296 + blockIdx.x] = smid;
  }
}
