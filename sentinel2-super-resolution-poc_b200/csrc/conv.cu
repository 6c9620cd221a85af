// Host side of the RRDBNet / EDSR path: weight repacking, TMA tensor maps, per-layer launch
// plans, window batching, and the enhance()/upsample() entry points of the C ABI.
//
// Replaces RRDBNet.forward (cnn_super_resolution.py:140-158) and RealESRGAN.enhance/_tile_process
// (cnn_super_resolution.py:217-280).  See conv_kernels.cuh for the kernels.
#include <cudaTypedefs.h>

#include <algorithm>
#include <chrono>

#include "common.h"
#include "hoststage.h"
#include "conv_kernels.cuh"
#include "ups_kernel.cuh"
#include "roll_kernel.cuh"

// Default of option `pdl` (programmatic dependent launch of the rolling kernel, roll_launch).  A build flag so that an A/B
// library (WOWSR_LIB) can run the whole test suite with the other default.
#ifndef WOWSR_PDL_DEFAULT
#define WOWSR_PDL_DEFAULT 1
#endif

namespace {

constexpr size_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per block on sm_100
constexpr size_t SMEM_SLACK = 1024 + 1024;  // alignment slack + control block

struct LayerW {
  int cin = 0, cout = 0, N = 0, n_chunks = 0;
  size_t chunk_bytes = 0;
  bool fp16 = false;  // operand type of this layer (input activations and weights)
  uint8_t* wpack = nullptr;
  uint8_t* wpack_v = nullptr;  // taps transposed, for vertical tiles
  uint8_t* wpack32 = nullptr;  // same as [chunk of 32 ch][kx][j][co][32ch] (SWIZZLE_64B rows): streamed layers under option tc_chunk32
  uint8_t* wpack32_v = nullptr;
  size_t chunk_bytes32 = 0;
  // rolling kernel, CTA pairs (roll_kernel.cuh): the stacked rows of every (chunk, tap) block split in halves,
  // [rank][chunk][tap][3N/2 rows]; `pair_bytes` = one rank's image
  uint8_t* wpair = nullptr;
  uint8_t* wpair_v = nullptr;
  size_t pair_bytes = 0;
  float* wsimple = nullptr;
  float* bias = nullptr;
};

}  // namespace

struct ConvNet {
  int kind = 0;  // 0 RRDBNet, 1 EDSR
  int num_block = 0, nf = 64, gc = 32;
  bool body_fp16 = false, tail_fp16 = false;  // operand types: RRDB trunk / the 5 tail convs
  float res_scale = 1.0f;
  float* first_w = nullptr;  // [ky][kx][ci][64] fp32
  float* first_b = nullptr;
  std::vector<LayerW> layers;
  DevBuf dense0, dense1, feat, trunk, rrdb, lo, up1, hra, hrb, wins, winxy, err;
  // rolling kernel: task lists per launch geometry (roll_plan_get)
  struct RollPlanDev { int key[9]; DevBuf tasks, off; int units, units_h; std::vector<int> unit_cost; };
  std::vector<RollPlanDev> roll_plans;
  // Measured-speed balancing (balance_update): per conv layer, the relative speed of every unit (CTA pair) on that layer as the
  // kernel timed it in the previous batches, and what the current batch launched.
  struct Balance {
    std::vector<float> speed;  // [units], mean 1 within the horizontal and within the vertical units; empty: not measured yet
    int obs = 0, version = 0;
    int plan = -1;             // index into roll_plans of the plan launched in this batch (-1: none / not a rolling launch)
    bool pair = false;
  };
  std::vector<Balance> balance;  // [layers]
  DevBuf unit_ns;                // [layers][160] nanoseconds, written by the kernels
  std::vector<long long> unit_ns_host;
};

void wowsr_net_free(ConvNet* n) {
  if (!n) return;
  for (auto& l : n->layers) {
    if (l.wpack) cudaFree(l.wpack);
    if (l.wpack_v) cudaFree(l.wpack_v);
    if (l.wpack32) cudaFree(l.wpack32);
    if (l.wpack32_v) cudaFree(l.wpack32_v);
    if (l.wpair) cudaFree(l.wpair);
    if (l.wpair_v) cudaFree(l.wpair_v);
    if (l.wsimple) cudaFree(l.wsimple);
    if (l.bias) cudaFree(l.bias);
  }
  if (n->first_w) cudaFree(n->first_w);
  if (n->first_b) cudaFree(n->first_b);
  for (auto& rp : n->roll_plans) {
    if (rp.tasks.p) cudaFree(rp.tasks.p);
    if (rp.off.p) cudaFree(rp.off.p);
  }
  DevBuf* bufs[] = {&n->dense0, &n->dense1, &n->feat, &n->trunk, &n->rrdb, &n->lo, &n->up1, &n->hra, &n->hrb, &n->wins, &n->winxy, &n->err, &n->unit_ns};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  delete n;
}

namespace {

inline uint16_t to_t(float v, bool fp16) {
  if (fp16) {
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&b);
}
inline float from_t(uint16_t u, bool fp16) {
  if (fp16) return __half2float(*reinterpret_cast<__half*>(&u));
  return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
}

int pad_n(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

// weight OIHW fp32 -> (a) smem image for the tensor-core kernel, (b) [ky][kx][ci][N] for the simple one
int upload_layer(wowsr_ctx* ctx, LayerW& L, const float* w, const float* b, int cin, int cout, bool fp16) {
  if (cin % 32 != 0 || cout > 64 || cout < 1) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "conv %d->%d unsupported (Cin must be a multiple of 32, Cout <= 64)", cin, cout);
  L.cin = cin;
  L.fp16 = fp16;
  L.cout = cout;
  L.N = pad_n(cout);
  L.n_chunks = (cin + 63) / 64;
  const int N = L.N;
  L.chunk_bytes = (size_t)3 * 3 * N * 128;
  std::vector<uint8_t> pack(L.chunk_bytes * L.n_chunks, 0), pack_v(L.chunk_bytes * L.n_chunks, 0);
  std::vector<float> simple((size_t)9 * cin * N, 0.0f);
  // `a` = tap along the run axis (selected by the shifted A descriptor), `j` = stacked tap along the row
  // axis (row-axis tap index 2-j).  Horizontal tiles: run axis x, row axis y.  Vertical tiles: swapped.
  for (int c = 0; c < L.n_chunks; c++)
    for (int a = 0; a < 3; a++)
      for (int j = 0; j < 3; j++) {
        const int b = 2 - j;
        for (int co = 0; co < cout; co++) {
          const int row = j * N + co;
          for (int ch = 0; ch < 64; ch++) {
            const int ci = c * 64 + ch;
            if (ci >= cin) break;
            // full chunk: 128-byte rows, SWIZZLE_128B (16-byte unit index ^ row%8); 32-channel remainder chunk:
            // 64-byte rows, SWIZZLE_64B (unit index ^ (row/2)%4) — the image the UMMA descriptors expect in smem
            const bool half_c = cin - c * 64 < 64;
            size_t off = half_c ? (size_t)c * L.chunk_bytes + (size_t)a * (3 * N * 64) + (size_t)row * 64 +
                                      (size_t)(((ch >> 3) ^ ((row >> 1) & 3)) << 4) + (size_t)(ch & 7) * 2
                                : (size_t)c * L.chunk_bytes + (size_t)a * (3 * N * 128) + (size_t)row * 128 +
                                      (size_t)(((ch >> 3) ^ (row & 7)) << 4) + (size_t)(ch & 7) * 2;
            const uint16_t th = to_t(w[(((size_t)co * cin + ci) * 3 + b) * 3 + a], fp16);  // ky = b, kx = a
            const uint16_t tv = to_t(w[(((size_t)co * cin + ci) * 3 + a) * 3 + b], fp16);  // ky = a, kx = b
            memcpy(&pack[off], &th, 2);
            memcpy(&pack_v[off], &tv, 2);
          }
        }
      }
  for (int ky = 0; ky < 3; ky++)
    for (int kx = 0; kx < 3; kx++)
      for (int ci = 0; ci < cin; ci++)
        for (int co = 0; co < cout; co++)
          simple[((size_t)(ky * 3 + kx) * cin + ci) * N + co] =
              from_t(to_t(w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx], fp16), fp16);
  // layers whose weights do not fit beside two activation stages are streamed per chunk; 32-channel chunks halve the
  // double buffer (run_conv picks the mode with the same test)
  // Measured (profiles/r01_epilogue_breakdown.txt, exp9): -4 % cycles per rdb.conv5 tile but no time gain under the
  // power cap, so the mode is opt-in (option tc_chunk32=1 before loading the network).
  if (wowsr_opt(ctx, "tc_chunk32", 0) && L.chunk_bytes * L.n_chunks + 2 * (size_t)TC_ASTAGE + SMEM_SLACK > SMEM_LIMIT) {
    L.chunk_bytes32 = (size_t)3 * 3 * N * 64;
    const int nc32 = cin / 32;
    std::vector<uint8_t> p32(L.chunk_bytes32 * nc32, 0), p32v(L.chunk_bytes32 * nc32, 0);
    for (int c = 0; c < nc32; c++)
      for (int a = 0; a < 3; a++)
        for (int j = 0; j < 3; j++) {
          const int bt = 2 - j;
          for (int co = 0; co < cout; co++) {
            const int row = j * N + co;
            for (int ch = 0; ch < 32; ch++) {
              const int ci = c * 32 + ch;
              const size_t off = (size_t)c * L.chunk_bytes32 + (size_t)a * (3 * N * 64) + (size_t)row * 64 +
                                 (size_t)(((ch >> 3) ^ ((row >> 1) & 3)) << 4) + (size_t)(ch & 7) * 2;
              const uint16_t th = to_t(w[(((size_t)co * cin + ci) * 3 + bt) * 3 + a], fp16);
              const uint16_t tv = to_t(w[(((size_t)co * cin + ci) * 3 + a) * 3 + bt], fp16);
              memcpy(&p32[off], &th, 2);
              memcpy(&p32v[off], &tv, 2);
            }
          }
        }
    WCUDA(ctx, cudaMalloc((void**)&L.wpack32, p32.size()));
    WCUDA(ctx, cudaMalloc((void**)&L.wpack32_v, p32v.size()));
    WCUDA(ctx, cudaMemcpy(L.wpack32, p32.data(), p32.size(), cudaMemcpyHostToDevice));
    WCUDA(ctx, cudaMemcpy(L.wpack32_v, p32v.data(), p32v.size(), cudaMemcpyHostToDevice));
  }
  {  // CTA-pair images: rank r holds stacked rows [r * 3N/2, (r + 1) * 3N/2) of every (chunk, tap) block (3N/2 is a multiple of 8,
     // so the swizzle phase of a row is the same in the split image)
    const int rows_b = 3 * N / 2;
    L.pair_bytes = L.chunk_bytes * L.n_chunks / 2;
    std::vector<uint8_t> pp(2 * L.pair_bytes, 0), ppv(2 * L.pair_bytes, 0);
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < L.n_chunks; c++) {
        const size_t rb = (cin - c * 64 < 64) ? 64 : 128;  // operand row bytes of this chunk
        for (int a = 0; a < 3; a++) {
          const size_t src = (size_t)c * L.chunk_bytes + (size_t)a * (3 * N * rb) + (size_t)r * rows_b * rb;
          const size_t dst = (size_t)r * L.pair_bytes + (size_t)c * (L.chunk_bytes / 2) + (size_t)a * (rows_b * rb);
          memcpy(&pp[dst], &pack[src], rows_b * rb);
          memcpy(&ppv[dst], &pack_v[src], rows_b * rb);
        }
      }
    WCUDA(ctx, cudaMalloc((void**)&L.wpair, pp.size()));
    WCUDA(ctx, cudaMalloc((void**)&L.wpair_v, ppv.size()));
    WCUDA(ctx, cudaMemcpy(L.wpair, pp.data(), pp.size(), cudaMemcpyHostToDevice));
    WCUDA(ctx, cudaMemcpy(L.wpair_v, ppv.data(), ppv.size(), cudaMemcpyHostToDevice));
  }
  std::vector<float> bias(64, 0.0f);
  for (int co = 0; co < cout; co++) bias[co] = b ? b[co] : 0.0f;
  WCUDA(ctx, cudaMalloc((void**)&L.wpack, pack.size()));
  WCUDA(ctx, cudaMalloc((void**)&L.wpack_v, pack_v.size()));
  WCUDA(ctx, cudaMemcpy(L.wpack_v, pack_v.data(), pack_v.size(), cudaMemcpyHostToDevice));
  WCUDA(ctx, cudaMalloc((void**)&L.wsimple, simple.size() * 4));
  WCUDA(ctx, cudaMalloc((void**)&L.bias, 64 * 4));
  WCUDA(ctx, cudaMemcpy(L.wpack, pack.data(), pack.size(), cudaMemcpyHostToDevice));
  WCUDA(ctx, cudaMemcpy(L.wsimple, simple.data(), simple.size() * 4, cudaMemcpyHostToDevice));
  WCUDA(ctx, cudaMemcpy(L.bias, bias.data(), 64 * 4, cudaMemcpyHostToDevice));
  return 0;
}

int upload_first(wowsr_ctx* ctx, ConvNet* net, const float* w, const float* b, int cin, int cout) {
  if (cin != 3 || cout != 64) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "first conv must be 3->64 (got %d->%d)", cin, cout);
  std::vector<float> ww(27 * 64);
  for (int ky = 0; ky < 3; ky++)
    for (int kx = 0; kx < 3; kx++)
      for (int ci = 0; ci < 3; ci++)
        for (int co = 0; co < 64; co++) ww[((ky * 3 + kx) * 3 + ci) * 64 + co] = w[(((size_t)co * 3 + ci) * 3 + ky) * 3 + kx];
  WCUDA(ctx, cudaMalloc((void**)&net->first_w, ww.size() * 4));
  WCUDA(ctx, cudaMalloc((void**)&net->first_b, 64 * 4));
  WCUDA(ctx, cudaMemcpy(net->first_w, ww.data(), ww.size() * 4, cudaMemcpyHostToDevice));
  WCUDA(ctx, cudaMemcpy(net->first_b, b, 64 * 4, cudaMemcpyHostToDevice));
  return 0;
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// 4-D map over an NHWC activation buffer: dims (C, W, H, N), box (64 ch, 130 px, 1 row, 1 window)
// `transposed`: dims (C, H, W, N) — the run axis (box of 130) walks y; used by the vertical tiles.
int make_tmap(wowsr_ctx* ctx, CUtensorMap* m, const void* base, int C, int W, int H, int Nw, bool fp16, bool transposed,
              bool half = false) {
  auto enc = get_encode();
  if (!enc) return wowsr_fail(ctx, WOWSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nw};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  if (transposed) {
    dims[1] = (cuuint64_t)H;
    dims[2] = (cuuint64_t)W;
    strides[0] = (cuuint64_t)C * 2 * W;
    strides[1] = (cuuint64_t)C * 2;
  }
  cuuint32_t box[4] = {half ? 32u : 64u, (cuuint32_t)TC_AROWS, 2, 1};  // one pipeline stage = two consecutive rows of the row axis
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   half ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return wowsr_fail(ctx, WOWSR_ERR_CUDA, "cuTensorMapEncodeTiled(C=%d,W=%d,H=%d,N=%d) -> CUresult %d", C, W, H, Nw, (int)r);
  return 0;
}

// Folded upsample (ups_kernel.cuh): 5-D map of the x-upsampled view of an NHWC buffer [Nw][Hs][Ws][C]: dims (C, rep = 2 with a
// ZERO stride, Ws, Hs, Nw), box (64 ch, 2, 66 source pixels, 1 row, 1 window) = 132 upsampled pixels.  `transposed`: the run
// axis (the replicated one) walks y, for the vertical tiles.  Whether the driver accepts a zero stride is what
// tools/tma_stride0_probe.cu measures; a rejection surfaces here as an error, never as a silent fallback.
int make_tmap_ups(wowsr_ctx* ctx, CUtensorMap* m, const void* base, int C, int Ws, int Hs, int Nw, bool fp16, bool transposed) {
  auto enc = get_encode();
  if (!enc) return wowsr_fail(ctx, WOWSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)Nw};
  cuuint64_t strides[4] = {0, (cuuint64_t)C * 2, (cuuint64_t)C * 2 * Ws, (cuuint64_t)C * 2 * Ws * Hs};
  if (transposed) {
    dims[2] = (cuuint64_t)Hs;
    dims[3] = (cuuint64_t)Ws;
    strides[1] = (cuuint64_t)C * 2 * Ws;
    strides[2] = (cuuint64_t)C * 2;
  }
  cuuint32_t box[5] = {64u, 2u, 66u, 1u, 1u};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled rejected the zero-stride upsample map (C=%d,Ws=%d,Hs=%d,N=%d): CUresult %d",
                      C, Ws, Hs, Nw, (int)r);
  return 0;
}

struct LayerIO {
  const void* in = nullptr;  // activation buffer (T), C_total channels per pixel
  int in_C = 0;
  int Nw = 0, h = 0, w = 0;
  int act = 0;
  float scale1 = 1.0f, scale2 = 1.0f;
  const float* res1 = nullptr;
  const float* res2 = nullptr;
  float* out_f32_a = nullptr;
  float* out_f32_b = nullptr;
  const uint16_t* lo_in = nullptr;  // split trunk (see ConvParams)
  uint16_t* lo_out = nullptr;
  int ident = 0;
  void* out_t = nullptr;
  int out_stride = 0, out_choff = 0, out_rep = 1;
  long long out_row = 0;  // elements between output rows, 0: w * out_stride (rolling kernel, plain epilogue)
  int in_ups = 0;  // `in` is at HALF the layer resolution, the nearest-x2 upsample is folded into the tensor maps
  float final_scale = 255.0f;
  float final_add[3] = {0.0f, 0.0f, 0.0f};
  int final_round = 0;
  int out_fp16 = 0;
  F32Layout f32 = {0, 0, 0, 0, 0};
  int final = 0;
  uint8_t* out_u8 = nullptr;
  long long out_u8_pitch = 0;
  float* out_img_f32 = nullptr;
  long long out_img_f32_pitch = 0;
  const WinDev* wins = nullptr;
};

// ---------------------------------------------------------------------------------------------
// rolling kernel (roll_kernel.cuh): task lists and launch
// ---------------------------------------------------------------------------------------------

// Cuts the row sequence of all columns of a launch into `units` near-equal parts.  A column = one 128-pixel run over the
// whole row axis of a window (horizontal: runs along x over the h rows of x < strip_x0; vertical: runs along y over the
// strip's w - strip_x0 columns).  Pair launches put two columns side by side (the two CTAs of a pair), a dummy (-1) when
// the count is odd.  Units [0, units_h) get the horizontal columns, the others the vertical ones, split in proportion to
// the rows (+ 4 per column for the extra input rows and the junk pair).
struct RollPlanHost {
  std::vector<RollTask> tasks;
  std::vector<int> off;
  int units = 0, units_h = 0;
};

void roll_plan_build(RollPlanHost& R, int Nw, int h, int w, int strip_x0, bool pair, int max_units, int v_weight_pm = 1000,
                     const float* speed = nullptr /* [max_units] relative speed of every unit, or null: equal shares */) {
  struct Col { int n, u0; };
  std::vector<Col> ch, cv;
  const int runs_h = (strip_x0 + TC_RUN - 1) / TC_RUN, rem = w - strip_x0, runs_v = rem > 0 ? (h + TC_RUN - 1) / TC_RUN : 0;
  for (int n = 0; n < Nw; n++) {
    for (int r = 0; r < runs_h; r++) ch.push_back({n, r * TC_RUN});
    for (int r = 0; r < runs_v; r++) cv.push_back({n, r * TC_RUN});
  }
  const int per = pair ? 2 : 1;
  const long long pc_h = ((long long)ch.size() + per - 1) / per, pc_v = ((long long)cv.size() + per - 1) / per;  // (pair-)columns
  // v_weight_pm (per mille): how much slower a row of a vertical task is than a row of a horizontal one for this layer.  Measured
  // on full cfg5 batches (profiles/r02_roll_balance_cfg5b.txt): 1.27 for conv_last (64 -> 3 at 4x: HBM-bound, and the transposed
  // runs read 128 different image rows per box), within +-8 % of 1 for every other layer.
  const long long cost_h = pc_h * (h + 4), cost_v = pc_v * (rem + 4) * v_weight_pm / 1000;
  // Do not cut segments shorter than this: each costs 2 extra input rows + a junk pair.  Small launches (fewer than 8 rows per
  // unit, e.g. one 128 x 128 tile) are bound by the LATENCY of a unit's groups (~2 us each), not by throughput, and the SMs a
  // longer minimum would leave idle do the extra rows for free: 128 rows over 64 units of 2 rows instead of 13 units of 10
  // (cfg1: 15.7 -> ~10 us per N = 32 launch).
  int kMinRows = 8;
  {
    const long long total_rows = pc_h * h + pc_v * rem;
    if (total_rows < (long long)max_units * 8) {
      long long per_unit = (total_rows + max_units - 1) / max_units;
      per_unit += per_unit & 1;
      kMinRows = (int)std::max<long long>(2, std::min<long long>(8, per_unit));
    }
  }
  int units_v = 0;
  if (pc_v > 0) {
    units_v = (int)((double)max_units * cost_v / (double)(cost_h + cost_v) + 0.5);
    if (units_v < 1) units_v = 1;
    if (pc_h > 0 && units_v > max_units - 1) units_v = max_units - 1;
    const long long cap = (pc_v * rem + kMinRows - 1) / kMinRows;
    if (units_v > cap) units_v = (int)cap;
  }
  int units_h = 0;
  if (pc_h > 0) {
    units_h = max_units - units_v;
    const long long cap = (pc_h * h + kMinRows - 1) / kMinRows;
    if (units_h > cap) units_h = (int)cap;
    if (units_h < 1) units_h = 1;
  }
  R.tasks.clear();
  R.off.assign(1, 0);
  auto cut = [&](const std::vector<Col>& cols, long long pcs, int len, int v_first, int units, int unit0) {
    if (units <= 0) return;
    const long long total = pcs * len;
    long long done = 0;
    double wsum = 0.0, wacc = 0.0;
    for (int k = 0; k < units; k++) wsum += speed ? (double)speed[unit0 + k] : 1.0;
    for (int k = 0; k < units; k++) {
      // exclusive end of this unit's share of the row sequence: equal shares, or shares in proportion to the measured speeds
      wacc += speed ? (double)speed[unit0 + k] : 1.0;
      const long long end = k + 1 == units ? total : std::min<long long>(total, (long long)((double)total * wacc / wsum + 0.5));
      while (done < end) {
        const long long pcix = done / len;
        const int r0 = (int)(done - pcix * len);
        int take = (int)std::min<long long>(len - r0, end - done);
        // never leave a sliver: a tail of fewer than kMinRows rows of this column goes to the same unit
        if (len - (r0 + take) > 0 && len - (r0 + take) < kMinRows) take = len - r0;
        if (take < kMinRows && r0 + take < len && k + 1 < units) { take = std::min(len - r0, kMinRows); }
        // segments start on even rows: row v of a column lives in ring slot v % RING and a ring pair holds rows (2j, 2j + 1)
        if (r0 + take < len && (take & 1)) take++;
        RollTask T;
        memset(&T, 0, sizeof T);
        for (int q = 0; q < 2; q++) {
          const size_t ci = (size_t)pcix * per + q;
          if (q < per && ci < cols.size()) { T.n[q] = cols[ci].n; T.u0[q] = cols[ci].u0; }
          else { T.n[q] = -1; T.u0[q] = 0; }
        }
        T.v0 = v_first + r0;
        T.rows = take;
        R.tasks.push_back(T);
        done += take;
      }
      R.off.push_back((int)R.tasks.size());
    }
  };
  cut(ch, pc_h, h, 0, units_h, 0);
  cut(cv, pc_v, rem, strip_x0, units_v, units_h);
  R.units_h = units_h;
  R.units = units_h + units_v;
}

int roll_plan_get(wowsr_ctx* ctx, ConvNet* net, int Nw, int h, int w, int strip_x0, bool pair, int max_units, int v_weight_pm,
                  int bal_layer /* -1: equal shares */, const ConvNet::RollPlanDev** out, cudaStream_t st) {
  const ConvNet::Balance* B = bal_layer >= 0 && bal_layer < (int)net->balance.size() && (int)net->balance[bal_layer].speed.size() >= max_units
                                  ? &net->balance[bal_layer] : nullptr;
  const int key[9] = {Nw, h, w, strip_x0, pair ? 1 : 0, max_units, v_weight_pm, B ? bal_layer + 1 : 0, B ? B->version : 0};
  for (const auto& rp : net->roll_plans)
    if (memcmp(rp.key, key, sizeof key) == 0) { *out = &rp; return 0; }
  RollPlanHost H;
  roll_plan_build(H, Nw, h, w, strip_x0, pair, max_units, v_weight_pm, B ? B->speed.data() : nullptr);
  if (H.units < 1 || H.tasks.empty()) return wowsr_fail(ctx, WOWSR_ERR_ARG, "rolling plan: nothing to do");
  // one plan per window size and layer resolution, or — once speeds are measured — per layer: a few hundred small lists
  if (net->roll_plans.size() >= 2048) {
    for (auto& rp : net->roll_plans) { if (rp.tasks.p) cudaFree(rp.tasks.p); if (rp.off.p) cudaFree(rp.off.p); }
    net->roll_plans.clear();
    for (auto& b : net->balance) b.plan = -1;
  }
  net->roll_plans.emplace_back();
  ConvNet::RollPlanDev& D = net->roll_plans.back();
  memcpy(D.key, key, sizeof key);
  D.units = H.units; D.units_h = H.units_h;
  D.unit_cost.assign(H.units, 0);
  for (int u = 0; u < H.units; u++)
    for (int t = H.off[u]; t < H.off[u + 1]; t++) D.unit_cost[u] += H.tasks[t].rows + 4;
  if (int e = wowsr_ensure(ctx, D.tasks, H.tasks.size() * sizeof(RollTask))) return e;
  if (int e = wowsr_ensure(ctx, D.off, H.off.size() * sizeof(int))) return e;
  WCUDA(ctx, cudaMemcpyAsync(D.tasks.p, H.tasks.data(), H.tasks.size() * sizeof(RollTask), cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaMemcpyAsync(D.off.p, H.off.data(), H.off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaStreamSynchronize(st));  // H goes out of scope
  *out = &D;
  return 0;
}

template <int N, int MODE, bool PAIR, bool UPS = false>
int roll_launch(wowsr_ctx* ctx, int grid, size_t smem, cudaStream_t st, const CUtensorMap& th, const CUtensorMap& tv, const CUtensorMap& th32,
                const CUtensorMap& tv32, const ConvParams& P, const RollParams& Q) {
  auto kern = conv3x3_roll_kernel<N, MODE, PAIR, UPS>;
  static bool attr_set[8] = {false, false, false, false, false, false, false, false};  // per device
  if (ctx->device < 8 ? !attr_set[ctx->device] : true) {
    WCUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    if (ctx->device < 8) attr_set[ctx->device] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Programmatic dependent launch: this launch may start while the previous kernel of the stream drains; the kernel reads and
  // writes activations only after griddepcontrol.wait (roll_kernel.cuh).  Small launches gain the most (cfg1: 354 dependent
  // launches of 7-17 us, each with 2.3-3 us of set-up and a launch gap).  Off while a layer is traced (the phase stamps of two
  // launches would interleave).
  if (wowsr_opt(ctx, "pdl", WOWSR_PDL_DEFAULT) != 0 && P.trace == nullptr) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  WCUDA(ctx, cudaLaunchKernelEx(&cfg, kern, th, tv, th32, tv32, P, Q));
  ctx->launches++;
  return 0;
}

// Launches one conv layer through the rolling kernel.  Returns 1 when the layer is not eligible (the caller falls back to the
// tile kernel): weights that do not fit next to two activation stages, the folded-upsample input, timing-only debug flags.
int run_conv_roll(wowsr_ctx* ctx, ConvNet* net, const LayerW& L, const LayerIO& io, ConvParams& P, int mode, cudaStream_t st) {
  const int N = L.N;
  const int64_t roll = wowsr_opt(ctx, "roll", 1);
  if (!roll || (P.flags & (CF_DBG_NO_TMA | CF_DBG_NO_MMA | CF_DBG_NO_STORE))) return 1;
  if (io.in_ups && !wowsr_opt(ctx, "roll_ups", 1)) return 1;  // A/B: the tile kernel's folded-upsample variant (ups_kernel.cuh)
  if (io.in_ups) {  // the folded-upsample instantiation has the plain epilogue only (whatever tc_generic_epilogue says)
    const bool plain_ok = !io.final && io.out_t && io.out_rep == 1 && !io.out_f32_a && !io.out_f32_b && !io.res1 && !io.res2 &&
                          !io.lo_in && !io.lo_out;
    if (L.cin != 64 || N != 64 || !plain_ok || (io.w & 1) || (io.h & 1) || io.in_C != 64)
      return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "folded upsample: 64 -> 64 layers with the plain epilogue at an even resolution only");
    mode = EPI_PLAIN;
  }
  const size_t astage = io.in_ups ? 2 * (size_t)TC_UPS_ROWB : (size_t)TC_ASTAGE;
  bool pair = wowsr_opt(ctx, "roll_pair", 1) != 0 && ctx->sm_count >= 2;
  const size_t id_bytes = P.ident ? (pair ? 4096 : 8192) : 0;
  // resident weights: full 64-channel chunks + the half-size remainder chunk
  auto w_need = [&](bool pr) {
    const size_t cb = pr ? L.chunk_bytes / 2 : L.chunk_bytes;
    return cb * (size_t)(L.cin / 64) + (L.cin % 64 ? cb / 2 : 0);
  };
  auto stages_for = [&](bool pr) {
    const size_t need = w_need(pr) + (P.ident ? (pr ? 4096 : 8192) : 0) + SMEM_SLACK;
    return need >= SMEM_LIMIT ? 0 : (int)((SMEM_LIMIT - need) / astage);
  };
  if (stages_for(pair) < 2) {
    if (!pair && stages_for(true) >= 2 && ctx->sm_count >= 2 && wowsr_opt(ctx, "roll_pair", 1) != 0) pair = true;
    else return 1;
  }
  (void)id_bytes;
  int stages = stages_for(pair);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const int64_t optS = wowsr_opt(ctx, "tc_stages", 0);
  if (optS >= 2 && optS < stages) stages = (int)optS;
  P.n_stage = stages;
  P.n_chunks = L.n_chunks;
  P.chunk_ch = 64;
  // geometry: the remainder strip (w % 128 columns) goes to vertical tasks under the same rule as the trunk layout
  const int wm = io.w / TC_RUN * TC_RUN, rem = io.w - wm;
  int max_units = pair ? ctx->sm_count / 2 : ctx->sm_count;
  const int64_t optG = wowsr_opt(ctx, "roll_grid", 0);
  if (optG > 0 && optG < max_units) max_units = (int)optG;
  // a unit is horizontal or vertical for the whole launch (its weights are resident): the strip needs a unit of its own
  const bool strip = rem > 0 && wm > 0 && io.h >= 64 && max_units >= 2 && !wowsr_opt(ctx, "tc_no_strip", 0);
  const int strip_x0 = strip ? wm : io.w;
  const ConvNet::RollPlanDev* plan = nullptr;
  // Measured-speed balancing (option roll_adapt = 1; OFF by default): big launches are timed per CTA, and the next batches cut this
  // layer's work list in proportion to what every CTA pair achieved on it (balance_update).  Results do not depend on the cuts.
  // Measured (tools/experiments/r2_adapt*.sh): mean / max of the unit finish times of a full cfg5 batch goes from 0.83-0.95 to
  // 0.97-0.99 on rdb.conv2, conv_up1 and conv_hr — worth 3-5 % at full clocks — but a full-size step runs under the 1 kW power cap,
  // where the SMs that used to wait at the end of a launch were saving the power the others ran on: cfg2 -0.85 %, cfg5 scene flat.
  const int li = (int)(&L - net->layers.data());
  const bool li_ok = li >= 0 && li < (int)net->layers.size();
  const bool adapt = li_ok && wowsr_opt(ctx, "roll_adapt", 0) != 0;
  if (adapt && net->balance.size() != net->layers.size()) {
    net->balance.assign(net->layers.size(), ConvNet::Balance());
    if (int e = wowsr_ensure(ctx, net->unit_ns, net->layers.size() * 160 * sizeof(long long))) return e;
    WCUDA(ctx, cudaMemsetAsync(net->unit_ns.p, 0, net->layers.size() * 160 * sizeof(long long), st));
  }
  if (int e = roll_plan_get(ctx, net, io.Nw, io.h, io.w, strip_x0, pair, max_units, io.final ? 1270 : 1000, adapt ? li : -1, &plan, st)) return e;
  if (adapt) {
    ConvNet::Balance& B = net->balance[li];
    B.plan = (int)(plan - net->roll_plans.data());
    B.pair = pair;
  }
  CUtensorMap tmap, tmap_v, tmap32, tmap_v32;
  const bool has_half = L.cin % 64 == 32;
  if (io.in_ups) {
    if (int e = make_tmap_ups(ctx, &tmap, io.in, 64, io.w / 2, io.h / 2, io.Nw, L.fp16, false)) return e;
  } else if (int e = make_tmap(ctx, &tmap, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, false)) return e;
  tmap_v = tmap; tmap32 = tmap;
  if (has_half)
    if (int e = make_tmap(ctx, &tmap32, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, false, true)) return e;
  tmap_v32 = tmap32;
  if (strip) {
    if (io.in_ups) {
      if (int e = make_tmap_ups(ctx, &tmap_v, io.in, 64, io.w / 2, io.h / 2, io.Nw, L.fp16, true)) return e;
    } else if (int e = make_tmap(ctx, &tmap_v, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, true)) return e;
    if (has_half)
      if (int e = make_tmap(ctx, &tmap_v32, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, true, true)) return e;
  }
  RollParams Q;
  memset(&Q, 0, sizeof Q);
  Q.tasks = (const RollTask*)plan->tasks.p;
  Q.unit_ns = adapt ? (long long*)net->unit_ns.p + (size_t)li * 160 : nullptr;
  Q.task_off = (const int*)plan->off.p;
  Q.units_h = plan->units_h;
  Q.w_bytes = (int)w_need(pair);
  Q.w_chunk_bytes = (int)(pair ? L.chunk_bytes / 2 : L.chunk_bytes);
  if (pair) {
    Q.wimg0 = L.wpair; Q.wimg1 = L.wpair + L.pair_bytes;
    Q.wimg_v0 = L.wpair_v; Q.wimg_v1 = L.wpair_v + L.pair_bytes;
  } else {
    Q.wimg0 = Q.wimg1 = L.wpack;
    Q.wimg_v0 = Q.wimg_v1 = L.wpack_v;
  }
  const size_t smem = (size_t)stages * astage + Q.w_bytes + (P.ident ? (pair ? 4096 : 8192) : 0) + SMEM_SLACK;
  const int grid = plan->units * (pair ? 2 : 1);
  if (io.in_ups)
    return pair ? roll_launch<64, EPI_PLAIN, true, true>(ctx, grid, smem, st, tmap, tmap_v, tmap32, tmap_v32, P, Q)
                : roll_launch<64, EPI_PLAIN, false, true>(ctx, grid, smem, st, tmap, tmap_v, tmap32, tmap_v32, P, Q);
#define ROLL_GO(NN, MM)                                                                                          \
  return pair ? roll_launch<NN, MM, true>(ctx, grid, smem, st, tmap, tmap_v, tmap32, tmap_v32, P, Q)             \
              : roll_launch<NN, MM, false>(ctx, grid, smem, st, tmap, tmap_v, tmap32, tmap_v32, P, Q)
  if (N == 16) { ROLL_GO(16, EPI_GENERIC); }
  if (N == 32) {
    if (mode == EPI_PLAIN) { ROLL_GO(32, EPI_PLAIN); }
    ROLL_GO(32, EPI_GENERIC);
  }
  if (mode == EPI_PLAIN) { ROLL_GO(64, EPI_PLAIN); }
  if (mode == EPI_RES) { ROLL_GO(64, EPI_RES); }
  ROLL_GO(64, EPI_GENERIC);
#undef ROLL_GO
}

int run_conv(wowsr_ctx* ctx, ConvNet* net, const LayerW& L, const LayerIO& io, cudaStream_t st) {
  ConvParams P;
  memset(&P, 0, sizeof P);
  P.Nw = io.Nw; P.h = io.h; P.w = io.w;
  P.cin = L.cin; P.n_chunks = L.n_chunks; P.N = L.N; P.cout = L.cout;
  const int N = L.N;
  const int R = N == 64 ? 4 : 8;  // must match conv3x3_tc_kernel<N>
  P.R = R;
  P.tiles_x = (io.w + TC_RUN - 1) / TC_RUN;
  P.tiles_y = (io.h + R - 1) / R;
  P.n_tiles = P.tiles_x * P.tiles_y * io.Nw;
  P.flags = (int)wowsr_opt(ctx, "tc_flags", CF_STACK) | (L.fp16 ? CF_FP16 : 0) | (io.out_fp16 ? CF_OUT_FP16 : 0);
  P.idesc_base = make_idesc_f16(128, 0, L.fp16);
  P.reverse = wowsr_opt(ctx, "tc_boustrophedon", 0) ? (int)(ctx->launches & 1) : 0;
  P.w_chunk_bytes = (uint32_t)L.chunk_bytes;
  P.lo_in = io.lo_in; P.lo_out = io.lo_out; P.ident = io.ident;
  if (P.ident && (N != 64 || L.cin < 64 || io.scale1 != 0.2f || !io.lo_in || io.res1))
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "identity K-step needs a 64-output layer, scale1 = 0.2 and a lo residual");
  const size_t id_bytes = P.ident ? 8192 : 0;
  size_t wtotal = L.chunk_bytes * L.n_chunks + id_bytes;
  size_t chunk_bytes = L.chunk_bytes;
  P.chunk_ch = 64;
  P.astage = io.in_ups ? 2 * TC_UPS_ROWB : TC_ASTAGE;  // folded-upsample stages: two 1024-aligned 132-pixel rows
  if (wtotal + 2 * (size_t)TC_ASTAGE + SMEM_SLACK <= SMEM_LIMIT && L.n_chunks <= TC_MAX_WBUF &&
      !wowsr_opt(ctx, "tc_force_stream", 0)) {
    P.w_resident = 1;
    P.n_wbuf = L.n_chunks;
  } else {
    P.w_resident = 0;
    P.n_wbuf = (int)wowsr_opt(ctx, "tc_wbuf", 2);
    if (P.n_wbuf < 2 || P.n_wbuf > TC_MAX_WBUF) P.n_wbuf = 2;
    if (L.wpack32 && wowsr_opt(ctx, "tc_chunk32", 0)) {  // streamed weights: 32-channel chunks, half-size stage slots
      P.chunk_ch = 32;
      P.astage = TC_ASTAGE32;
      P.n_chunks = L.cin / 32;
      chunk_bytes = L.chunk_bytes32;
      P.w_chunk_bytes = (uint32_t)chunk_bytes;
    }
  }
  size_t left = SMEM_LIMIT - SMEM_SLACK - (size_t)P.n_wbuf * chunk_bytes - id_bytes;
  int stages = (int)(left / P.astage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  int64_t optS = wowsr_opt(ctx, "tc_stages", 0);
  if (optS > 0 && optS < stages) stages = (int)optS;
  if (stages < 2) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "not enough shared memory for conv %d->%d", L.cin, L.cout);
  P.n_stage = stages;
  P.wpack = P.chunk_ch == 32 ? L.wpack32 : L.wpack; P.wsimple = L.wsimple; P.bias = L.bias;
  P.in = io.in; P.in_stride = io.in_C;
  P.act = io.act; P.scale1 = io.scale1; P.res1 = io.res1; P.scale2 = io.scale2; P.res2 = io.res2;
  P.out_f32_a = io.out_f32_a; P.out_f32_b = io.out_f32_b;
  P.out_t = io.out_t; P.out_stride = io.out_stride; P.out_choff = io.out_choff; P.out_rep = io.out_rep;
  P.f32 = io.f32;
  P.out_row = io.out_row; P.final_scale = io.final_scale; P.final_round = io.final_round;
  for (int i = 0; i < 3; i++) P.final_add[i] = io.final_add[i];
  P.final = io.final; P.out_u8 = io.out_u8; P.out_u8_pitch = io.out_u8_pitch;
  P.out_img_f32 = io.out_img_f32; P.out_img_f32_pitch = io.out_img_f32_pitch; P.wins = io.wins;
  P.err_flag = (int*)net->err.p;
  P.trace = nullptr;
  {  // debug tracing of one launch: option tc_trace_layer = 1-based index of the conv launch to trace
    int64_t tl = wowsr_opt(ctx, "tc_trace_layer", 0);
    if (tl > 0 && ++ctx->trace_counter == tl) {
      if (int e = wowsr_ensure(ctx, ctx->trace_buf, 2 * 64 * 4 * 8)) return e;
      cudaMemsetAsync(ctx->trace_buf.p, 0, 2 * 64 * 4 * 8, st);
      P.trace = (long long*)ctx->trace_buf.p;
    }
  }

  if (io.out_row && (wowsr_opt(ctx, "conv_impl", 0) == 1 || wowsr_opt(ctx, "tc_generic_epilogue", 0) || !wowsr_opt(ctx, "roll", 1)))
    return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "an output row pitch (EDSR upsampler) needs the rolling kernel's plain epilogue");
  if (wowsr_opt(ctx, "conv_impl", 0) == 1) {
    long long total = (long long)io.Nw * io.h * io.w;
    unsigned blocks = (unsigned)((total + 127) / 128);
#define SIMPLE(NN)                                                                              \
  conv3x3_simple_kernel<NN><<<blocks, 128, 0, st>>>(P);
    if (N == 16) { SIMPLE(16) } else if (N == 32) { SIMPLE(32) } else { SIMPLE(64) }
#undef SIMPLE
    WLAUNCH_CHECK(ctx);
    return 0;
  }
  {  // rolling kernel (roll_kernel.cuh) for every layer it can hold the weights of; the tile kernel below is the fallback
    int mode_r = EPI_GENERIC;
    if (!wowsr_opt(ctx, "tc_generic_epilogue", 0) && N >= 32 && !io.final && io.out_t && io.out_rep == 1 && !io.out_f32_b) {
      if (!io.res1 && !io.res2 && !io.out_f32_a && !io.lo_in && !io.lo_out) mode_r = EPI_PLAIN;
      else if (N == 64 && io.f32.wpb && (io.res1 != nullptr) != (io.lo_in != nullptr) && (io.out_f32_a || io.lo_out)) mode_r = EPI_RES;
    }
    ConvParams PR = P;
    const int rr = run_conv_roll(ctx, net, L, io, PR, mode_r, st);
    if (rr <= 0) return rr;
    if (io.out_row) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "an output row pitch needs the rolling kernel (layer does not fit it)");
  }
  CUtensorMap tmap, tmap_v, tmap32, tmap_v32;
  if (io.in_ups) {
    if (L.cin != 64 || N != 64 || !P.w_resident || (io.w & 1) || (io.h & 1) || io.in_C != 64 || P.ident)
      return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "folded upsample: 64 -> 64 layers at an even resolution only");
    if (int e = make_tmap_ups(ctx, &tmap, io.in, 64, io.w / 2, io.h / 2, io.Nw, L.fp16, false)) return e;
  } else
  if (int e = make_tmap(ctx, &tmap, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, false)) return e;
  tmap_v = tmap;
  const bool has_half = L.cin % 64 == 32 || P.chunk_ch == 32;  // 32-channel chunks: their own 32-channel / SWIZZLE_64B maps
  tmap32 = tmap;
  if (has_half)
    if (int e = make_tmap(ctx, &tmap32, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, false, true)) return e;
  tmap_v32 = tmap32;
  // Remainder strip (w not a multiple of 128): cover it with vertical runs when that wastes fewer MMA rows.
  const int wm = io.w / TC_RUN * TC_RUN, rem = io.w - wm;
  int max_grid = ctx->sm_count;
  int64_t optG = wowsr_opt(ctx, "tc_grid", 0);
  if (optG > 0 && optG < max_grid) max_grid = (int)optG;
  bool use_v = false;
  if (rem > 0 && !wowsr_opt(ctx, "tc_no_strip", 0)) {
    const int v_runs = (io.h + TC_RUN - 1) / TC_RUN, v_rows = (rem + R - 1) / R;
    const double eff_h = rem / (double)TC_RUN;
    const double eff_v = (io.h / (double)(v_runs * TC_RUN)) * (rem / (double)(v_rows * R));
    if (eff_v > eff_h && max_grid >= 2 &&
        (io.in_ups ? make_tmap_ups(ctx, &tmap_v, io.in, 64, io.w / 2, io.h / 2, io.Nw, L.fp16, true)
                   : make_tmap(ctx, &tmap_v, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, true)) == 0 &&
        (!has_half || make_tmap(ctx, &tmap_v32, io.in, io.in_C, io.w, io.h, io.Nw, L.fp16, true, true) == 0)) {
      use_v = true;
      P.strip_x0 = wm;
      P.v_runs = v_runs;
      P.v_rows = v_rows;
      P.n_tiles_v = v_runs * v_rows * io.Nw;
      P.tiles_x = wm / TC_RUN;
      P.n_tiles = P.tiles_x * P.tiles_y * io.Nw;
      P.wpack_v = P.chunk_ch == 32 ? L.wpack32_v : L.wpack_v;
    }
  }
  int grid = P.n_tiles + P.n_tiles_v < max_grid ? P.n_tiles + P.n_tiles_v : max_grid;
  P.grid_h = grid;
  if (use_v) {
    if (P.n_tiles == 0) {
      P.grid_h = 0;
    } else {
      // split the CTAs between the two orientations so that both finish together (tiles cost the same)
      int best = 1;
      long long best_t = -1;
      for (int gv = 1; gv < grid; gv++) {
        long long th = (P.n_tiles + (grid - gv) - 1) / (grid - gv), tv = (P.n_tiles_v + gv - 1) / gv;
        long long t = th > tv ? th : tv;
        if (best_t < 0 || t < best_t) { best_t = t; best = gv; }
      }
      P.grid_h = grid - best;
    }
  }
  size_t smem = (size_t)P.n_stage * P.astage + (size_t)P.n_wbuf * chunk_bytes + id_bytes + SMEM_SLACK;
  if (io.in_ups) {  // folded-upsample kernel (ups_kernel.cuh): plain epilogue only
    if (P.final || P.res1 || P.res2 || P.out_f32_a || P.out_f32_b || P.lo_in || P.lo_out || P.out_rep != 1 || !P.out_t)
      return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "folded upsample: plain epilogue only");
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_ups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    conv3x3_tc_ups_kernel<<<grid, TC_THREADS, smem, st>>>(tmap, tmap_v, P);
    WLAUNCH_CHECK(ctx);
    return 0;
  }
  // epilogue specialisation (conv_kernels.cuh): the generic path handles every other layer shape
  int mode = EPI_GENERIC;
  if (!wowsr_opt(ctx, "tc_generic_epilogue", 0) && N >= 32 && !P.final && P.out_t && P.out_rep == 1 && !P.out_f32_b &&
      !(P.flags & CF_DBG_NO_STORE)) {
    if (!P.res1 && !P.res2 && !P.out_f32_a && !P.lo_in && !P.lo_out) mode = EPI_PLAIN;
    else if (N == 64 && P.f32.wpb && (P.res1 != nullptr) != (P.lo_in != nullptr) && (P.out_f32_a || P.lo_out)) mode = EPI_RES;
  }
  if (!ctx->tc_attr_set) {  // per device (one handle per device)
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<16, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<32, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<64, EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<32, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<64, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    WCUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<64, EPI_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
    ctx->tc_attr_set = true;
  }
#define TC_LAUNCH(NN, MM) conv3x3_tc_kernel<NN, MM><<<grid, TC_THREADS, smem, st>>>(tmap, tmap_v, tmap32, tmap_v32, P)
  if (N == 16) TC_LAUNCH(16, EPI_GENERIC);
  else if (N == 32) { if (mode == EPI_PLAIN) TC_LAUNCH(32, EPI_PLAIN); else TC_LAUNCH(32, EPI_GENERIC); }
  else if (mode == EPI_PLAIN) TC_LAUNCH(64, EPI_PLAIN);
  else if (mode == EPI_RES) TC_LAUNCH(64, EPI_RES);
  else TC_LAUNCH(64, EPI_GENERIC);
#undef TC_LAUNCH
  WLAUNCH_CHECK(ctx);
  return 0;
}

// After a batch (the stream is idle): turns the per-CTA role times of every rolling launch of the batch into relative unit speeds of
// that layer.  The SMs do not get equal memory bandwidth — on the bandwidth-heavy layers the same 3-4 of every 8 SM pairs finish
// 15-25 % later, which pairs differs from chip to chip (profiles/r02_roll_balance_*.txt) — so equal rows are not equal time.  The
// first measurement is taken as it is, later ones are averaged in; after four batches a layer's shares are frozen.
int balance_update(wowsr_ctx* ctx, ConvNet* net) {
  if (net->balance.empty() || !net->unit_ns.p) return 0;
  bool any = false;
  for (const auto& b : net->balance) any |= b.plan >= 0 && b.obs < 4;
  if (!any) {
    for (auto& b : net->balance) b.plan = -1;
    return 0;
  }
  net->unit_ns_host.resize(net->balance.size() * 160);
  WCUDA(ctx, cudaMemcpy(net->unit_ns_host.data(), net->unit_ns.p, net->unit_ns_host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  for (size_t li = 0; li < net->balance.size(); li++) {
    ConvNet::Balance& B = net->balance[li];
    const int pi = B.plan;
    B.plan = -1;
    if (pi < 0 || pi >= (int)net->roll_plans.size() || B.obs >= 4) continue;
    const ConvNet::RollPlanDev& D = net->roll_plans[pi];
    const long long* ns = net->unit_ns_host.data() + li * 160;
    const int per = B.pair ? 2 : 1, units = D.units;
    if (units * per > 160 || units < 2) continue;
    std::vector<double> rate(units, 0.0);
    bool ok = true;
    long long cost_min = 1 << 30;
    for (int u = 0; u < units && ok; u++) {
      long long t = 0;
      for (int q = 0; q < per; q++) t = std::max(t, ns[u * per + q]);
      cost_min = std::min<long long>(cost_min, D.unit_cost[u]);
      if (t < 20000 || D.unit_cost[u] <= 0) ok = false;  // a unit without work, or a launch too short to time (< 20 us)
      else rate[u] = (double)D.unit_cost[u] / (double)t;
    }
    if (!ok || cost_min < 64) continue;
    for (int g = 0; g < 2; g++) {  // horizontal and vertical units are separate groups (their split is the cost model's)
      const int a = g ? D.units_h : 0, b = g ? units : D.units_h;
      if (b <= a) continue;
      double mean = 0.0;
      for (int u = a; u < b; u++) mean += rate[u];
      mean /= (b - a);
      for (int u = a; u < b; u++) rate[u] = std::min(1.5, std::max(0.6, rate[u] / mean));
    }
    const int max_units = D.key[5];
    if ((int)B.speed.size() != max_units) { B.speed.assign(max_units, 1.0f); B.obs = 0; }
    for (int u = 0; u < units; u++) B.speed[u] = B.obs == 0 ? (float)rate[u] : 0.5f * B.speed[u] + 0.5f * (float)rate[u];
    B.obs++;
    B.version++;
  }
  return 0;
}

int check_err_flag(wowsr_ctx* ctx, ConvNet* net, cudaStream_t st) {
  int flag = 0;
  WCUDA(ctx, cudaMemcpyAsync(&flag, net->err.p, 4, cudaMemcpyDeviceToHost, st));
  WCUDA(ctx, cudaStreamSynchronize(st));
  if (flag) {
    cudaMemsetAsync(net->err.p, 0, 4, st);
    return wowsr_fail(ctx, WOWSR_ERR_CUDA, "tensor-core conv kernel watchdog tripped (code %d)", flag);
  }
  return 0;
}

// One batch of equally sized windows through the whole RRDBNet.
int rrdbnet_batch(wowsr_ctx* ctx, ConvNet* net, const uint8_t* img, long long pitch, const wowsr_window* wins, int nb,
                  uint8_t* out, long long out_pitch, float* out_f32, long long out_f32_pitch, cudaStream_t st) {
  const int h = wins[0].y1 - wins[0].y0, w = wins[0].x1 - wins[0].x0;
  const size_t px = (size_t)nb * h * w;
  // fp32 trunk buffers: warp-blocked layout (F32Layout); the remainder strip is blocked along y when the
  // body layers will cover it with vertical tiles
  F32Layout fl;
  {
    const int wm = w / TC_RUN * TC_RUN, rem = w - wm;
    const bool strip = rem > 0 && wm > 0 && h >= 64;
    fl.x0 = strip ? wm : w;
    fl.wpb = strip ? wm / 32 : (w + 31) / 32;
    fl.rem = strip ? rem : 0;
    fl.hpb = (h + 31) / 32;
    fl.strip_off = (long long)nb * h * fl.wpb * 32 * 64;
  }
  const size_t pxb = (size_t)nb * h * fl.wpb * 32 + (size_t)nb * fl.rem * fl.hpb * 32;
  const bool body16 = net->body_fp16, tail16 = net->tail_fp16;
  if (int e = wowsr_ensure(ctx, net->dense0, px * 192 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->dense1, px * 192 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->feat, pxb * 64 * 4)) return e;
  // residual trunk inside an RRDB: split hi/lo (default) or a full fp32 copy (option trunk_hilo=0)
  const bool hilo = wowsr_opt(ctx, "trunk_hilo", 1) != 0;
  if (hilo) {
    if (int e = wowsr_ensure(ctx, net->lo, pxb * 64 * 2)) return e;
  } else {
    if (int e = wowsr_ensure(ctx, net->trunk, pxb * 64 * 4)) return e;
  }
  if (int e = wowsr_ensure(ctx, net->rrdb, pxb * 64 * 4)) return e;
  if (int e = wowsr_ensure(ctx, net->up1, px * 4 * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->hra, px * 16 * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->hrb, px * 16 * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->wins, (size_t)nb * sizeof(WinDev))) return e;
  if (int e = wowsr_ensure(ctx, net->winxy, (size_t)nb * 8)) return e;
  std::vector<WinDev> wd(nb);
  std::vector<int> wxy(2 * nb);
  for (int i = 0; i < nb; i++) {
    wd[i] = WinDev{4 * wins[i].x0, 4 * wins[i].y0, 4 * wins[i].ox0, 4 * wins[i].oy0, 4 * wins[i].ox1, 4 * wins[i].oy1};
    wxy[2 * i] = wins[i].x0;
    wxy[2 * i + 1] = wins[i].y0;
  }
  WCUDA(ctx, cudaMemcpyAsync(net->wins.p, wd.data(), nb * sizeof(WinDev), cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaMemcpyAsync(net->winxy.p, wxy.data(), nb * 8, cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaStreamSynchronize(st));  // host vectors go out of scope

  WCUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  {
    FirstParams F;
    memset(&F, 0, sizeof F);
    F.img = img; F.pitch = pitch; F.cin = 3; F.win_xy = (const int*)net->winxy.p;
    F.Nw = nb; F.h = h; F.w = w; F.weight = net->first_w; F.bias = net->first_b;
    F.f32_a = (float*)net->feat.p; F.f32_b = (float*)net->rrdb.p;  // rrdb = T0: the RRDB input x_rrdb
    F.f32 = fl;
    F.out_t = net->dense0.p; F.out_stride = 192; F.out_fp16 = body16; F.in_scale_div = 255.0f;
    dim3 grid((unsigned)((px + 127) / 128), 4);
    conv_first_kernel<<<grid, 128, 0, st>>>(F);
    WLAUNCH_CHECK(ctx);
  }
  WCUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  void* cur = net->dense0.p;
  void* nxt = net->dense1.p;
  size_t li = 0;
  for (int b = 0; b < net->num_block; b++)
    for (int r = 0; r < 3; r++) {
      for (int k = 0; k < 4; k++) {
        LayerIO io;
        io.in = cur; io.in_C = 192; io.Nw = nb; io.h = h; io.w = w;
        io.act = 1;
        io.out_t = cur; io.out_stride = 192; io.out_choff = 64 + 32 * k; io.out_fp16 = body16;
        if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
      }
      LayerIO io;
      io.in = cur; io.in_C = 192; io.Nw = nb; io.h = h; io.w = w;
      io.f32 = fl;
      // fp32 residual trunk without a separate copy of the RRDB input: T0 (`rrdb`) holds x_rrdb and stays
      // untouched while rdb1/rdb2 run on T1 (`trunk`); rdb3 reads both and writes the next RRDB's input to T0.
      // Split trunk (default): rdb1 reads x_rrdb (fp32, T0) and leaves x1 as hi (operand copy in the next dense
      // buffer) + lo (16-bit); rdb2 / rdb3 get hi through the identity K-step and read only lo; rdb3 writes T0.
      io.scale1 = 0.2f;
      if (hilo) {
        if (r == 0) io.res1 = (const float*)net->rrdb.p;
        else { io.lo_in = (const uint16_t*)net->lo.p; io.ident = 1; }
        if (r == 2) io.out_f32_a = (float*)net->rrdb.p;
        else io.lo_out = (uint16_t*)net->lo.p;
      } else {
        io.res1 = (const float*)(r == 0 ? net->rrdb.p : net->trunk.p);
        io.out_f32_a = (float*)(r == 2 ? net->rrdb.p : net->trunk.p);
      }
      if (r == 2) {
        io.scale2 = 0.2f; io.res2 = (const float*)net->rrdb.p;
      }
      io.out_t = nxt; io.out_stride = 192; io.out_choff = 0;
      // the last RDB feeds conv_body, which may run in a different operand type (mixed precision)
      io.out_fp16 = (b == net->num_block - 1 && r == 2) ? tail16 : body16;
      if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
      std::swap(cur, nxt);
    }
  WCUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  // nearest-x2 folded into the consumer's TMA address generation (ups_kernel.cuh): producers store at their own resolution, the
  // upsample convs replicate through zero-stride tensor maps.  Option tail_fold_upsample=0 (and the CUDA-core cross-check
  // path) keeps the older producer-side 2x2 replicated store; both give bit-identical output (tests/test_gpu_rrdbnet.py).
  const int fold = wowsr_opt(ctx, "tail_fold_upsample", 1) && wowsr_opt(ctx, "conv_impl", 0) == 0 ? 1 : 0;
  {  // conv_body + long skip, written nearest-x2 replicated for conv_up1
    LayerIO io;
    io.in = cur; io.in_C = 192; io.Nw = nb; io.h = h; io.w = w;
    io.f32 = fl;
    io.scale1 = 1.0f; io.res1 = (const float*)net->feat.p;
    io.out_t = net->up1.p; io.out_stride = 64; io.out_rep = fold ? 1 : 2; io.out_fp16 = tail16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  {  // conv_up1 @2x + lrelu, replicated for conv_up2
    LayerIO io;
    io.in = net->up1.p; io.in_C = 64; io.Nw = nb; io.h = 2 * h; io.w = 2 * w;
    io.in_ups = fold;
    io.act = 1; io.out_t = net->hra.p; io.out_stride = 64; io.out_rep = fold ? 1 : 2; io.out_fp16 = tail16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  {  // conv_up2 @4x + lrelu
    LayerIO io;
    io.in = net->hra.p; io.in_C = 64; io.Nw = nb; io.h = 4 * h; io.w = 4 * w;
    io.in_ups = fold;
    io.act = 1; io.out_t = net->hrb.p; io.out_stride = 64; io.out_fp16 = tail16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  {  // conv_hr + lrelu
    LayerIO io;
    io.in = net->hrb.p; io.in_C = 64; io.Nw = nb; io.h = 4 * h; io.w = 4 * w;
    io.act = 1; io.out_t = net->hra.p; io.out_stride = 64; io.out_fp16 = tail16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  {  // conv_last + quantise + stitch
    LayerIO io;
    io.in = net->hra.p; io.in_C = 64; io.Nw = nb; io.h = 4 * h; io.w = 4 * w;
    io.final = 1; io.out_u8 = out; io.out_u8_pitch = out_pitch;
    io.out_img_f32 = out_f32; io.out_img_f32_pitch = out_f32_pitch / 4;
    io.wins = (const WinDev*)net->wins.p;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  WCUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

extern "C" int wowsr_load_rrdbnet(wowsr_ctx* ctx, int32_t num_block, int32_t num_feat, int32_t num_grow,
                                  const float* const* tensors, int32_t n_tensors, int32_t precision) {
  if (!ctx || !tensors) return WOWSR_ERR_ARG;
  if (num_feat != 64 || num_grow != 32)
    return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "only num_feat=64, num_grow=32 (x4plus / anime-6B) are built");
  const int n_convs = 1 + num_block * 15 + 5;
  if (n_tensors != 2 * n_convs) return wowsr_fail(ctx, WOWSR_ERR_ARG, "expected %d tensors, got %d", 2 * n_convs, n_tensors);
  DeviceGuard g(ctx->device);
  wowsr_net_free(ctx->net);
  ctx->net = nullptr;
  ConvNet* net = new ConvNet();
  net->kind = 0;
  net->num_block = num_block;
  net->body_fp16 = precision == WOWSR_PREC_FP16;
  net->tail_fp16 = precision != WOWSR_PREC_BF16;
  int t = 0;
  int e = upload_first(ctx, net, tensors[0], tensors[1], 3, 64);
  t = 2;
  for (int b = 0; b < num_block && !e; b++)
    for (int r = 0; r < 3 && !e; r++)
      for (int k = 0; k < 5 && !e; k++) {
        net->layers.emplace_back();
        e = upload_layer(ctx, net->layers.back(), tensors[t], tensors[t + 1], 64 + 32 * k, k < 4 ? 32 : 64, net->body_fp16);
        t += 2;
      }
  for (int i = 0; i < 5 && !e; i++) {  // conv_body, conv_up1, conv_up2, conv_hr, conv_last
    net->layers.emplace_back();
    e = upload_layer(ctx, net->layers.back(), tensors[t], tensors[t + 1], 64, i < 4 ? 64 : 3, net->tail_fp16);
    t += 2;
  }
  if (!e) e = wowsr_ensure(ctx, net->err, 4);
  if (!e && cudaMemset(net->err.p, 0, 4) != cudaSuccess) e = wowsr_fail(ctx, WOWSR_ERR_CUDA, "memset");
  if (e) {
    wowsr_net_free(net);
    return e;
  }
  ctx->net = net;
  return WOWSR_OK;
}

// `sink` (optional): streams the finished output rows to a host buffer while later batches compute (wowsr_enhance_host).
static int forward_windows_impl(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, int64_t pitch, const wowsr_window* windows,
                                int32_t n, uint8_t* out_dev, int64_t out_pitch, float* out_f32, int64_t out_f32_pitch, void* stream,
                                StageOut* sink) {
  if (!ctx || !img_dev || !windows || n < 1 || !out_dev) return WOWSR_ERR_ARG;
  if (!ctx->net || ctx->net->kind != 0) return wowsr_fail(ctx, WOWSR_ERR_STATE, "wowsr_load_rrdbnet has not been called");
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int h = windows[0].y1 - windows[0].y0, w = windows[0].x1 - windows[0].x0;
  if (pitch < (int64_t)W * 3 || out_pitch < (int64_t)W * 4 * 3 || (out_f32 && out_f32_pitch < (int64_t)W * 4 * 3 * 4))
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "row pitch smaller than a row (pitch %lld, out_pitch %lld, out_f32_pitch %lld for W = %d)",
                      (long long)pitch, (long long)out_pitch, (long long)out_f32_pitch, W);
  for (int i = 0; i < n; i++) {
    const wowsr_window& q = windows[i];
    if (q.y1 - q.y0 != h || q.x1 - q.x0 != w || q.x0 < 0 || q.y0 < 0 || q.x1 > W || q.y1 > H)
      return wowsr_fail(ctx, WOWSR_ERR_ARG, "window %d has a different size or lies outside the image", i);
    // the owned rectangle (may be empty) must lie inside its window: the kernels write it without further checks
    if (q.ox1 > q.ox0 && q.oy1 > q.oy0 && (q.ox0 < q.x0 || q.ox1 > q.x1 || q.oy0 < q.y0 || q.oy1 > q.y1))
      return wowsr_fail(ctx, WOWSR_ERR_ARG, "window %d owns pixels outside itself", i);
  }
  // batch size from the workspace budget: 6144 B per LR pixel (data layout).  64 GiB of the 180 GB: a rank of an
  // 8-GPU cfg5 run (231 windows of 276 x 276) needs two batches instead of three, one GPU 13 instead of 18 — every batch costs
  // 350 launch prologues and pipeline tails.
  int64_t budget = wowsr_opt(ctx, "mem_budget_mb", 65536) << 20;
  int64_t per_win = (int64_t)h * w * 6144;
  int maxb = (int)std::max<int64_t>(1, budget / per_win);
  int nbatches = (n + maxb - 1) / maxb;
  int per = (n + nbatches - 1) / nbatches;
  float t_head = 0, t_trunk = 0, t_tail = 0;
  double t_enqueue = 0;  // host wall time inside rrdbnet_batch: buffers, window table, 350 launches — against the event times it
                         // tells whether a small forward (cfg1) is bound by the device or by the launching thread
  // suffix minimum of the first output row a window owns: after the windows before i are done, no later window writes above it
  std::vector<int> first_row;
  if (sink) {
    first_row.assign(n + 1, 4 * H);
    for (int i = n - 1; i >= 0; i--) {
      const wowsr_window& q = windows[i];
      first_row[i] = (q.ox1 > q.ox0 && q.oy1 > q.oy0) ? std::min(first_row[i + 1], 4 * q.oy0) : first_row[i + 1];
    }
  }
  int sent_rows = 0, final_rows = 0;
  for (int i0 = 0; i0 < n; i0 += per) {
    int nb = std::min(per, n - i0);
    const auto tq0 = std::chrono::steady_clock::now();
    if (int e = rrdbnet_batch(ctx, ctx->net, img_dev, pitch, windows + i0, nb, out_dev, out_pitch, out_f32, out_f32_pitch, st))
      return e;
    t_enqueue += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq0).count();
    if (sink && final_rows > sent_rows) {  // rows the PREVIOUS batches completed leave while this batch computes
      sink->enqueue(out_dev, (size_t)out_pitch, sent_rows, final_rows, nullptr);
      sent_rows = final_rows;
    }
    if (int e = check_err_flag(ctx, ctx->net, st)) return e;
    if (int e = balance_update(ctx, ctx->net)) return e;
    if (sink) final_rows = first_row[i0 + nb];
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]);
    t_head += a; t_trunk += b; t_tail += c;
  }
  if (sink) sink->enqueue(out_dev, (size_t)out_pitch, sent_rows, 4 * H, nullptr);
  ctx->timing[0] = t_head + t_trunk + t_tail;
  ctx->timing[1] = t_head; ctx->timing[2] = t_trunk; ctx->timing[3] = t_tail;
  ctx->timing[4] = (float)t_enqueue;
  return WOWSR_OK;
}

extern "C" int wowsr_rrdbnet_forward_windows(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, int64_t pitch,
                                             const wowsr_window* windows, int32_t n, uint8_t* out_dev, int64_t out_pitch,
                                             float* out_f32, int64_t out_f32_pitch, void* stream) {
  return forward_windows_impl(ctx, img_dev, H, W, pitch, windows, n, out_dev, out_pitch, out_f32, out_f32_pitch, stream, nullptr);
}

static int enhance_dev_impl(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, int32_t tile_size, uint8_t* out_dev,
                            float* out_f32_dev, void* stream, StageOut* sink) {
  if (!ctx || H < 1 || W < 1 || tile_size < 1) return WOWSR_ERR_ARG;
  int n = wowsr_plan_windows(H, W, tile_size, 10, nullptr, 0);
  if (n < 1) return wowsr_fail(ctx, WOWSR_ERR_ARG, "planner failed");
  std::vector<wowsr_window> wins(n);
  wowsr_plan_windows(H, W, tile_size, 10, wins.data(), n);
  return forward_windows_impl(ctx, img_dev, H, W, (int64_t)W * 3, wins.data(), n, out_dev, (int64_t)W * 4 * 3, out_f32_dev,
                              (int64_t)W * 4 * 3 * 4, stream, sink);
}

extern "C" int wowsr_enhance_dev(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, int32_t tile_size,
                                 uint8_t* out_dev, float* out_f32_dev, void* stream) {
  return enhance_dev_impl(ctx, img_dev, H, W, tile_size, out_dev, out_f32_dev, stream, nullptr);
}

// Drop-in for RealESRGAN.enhance(img) (cnn_super_resolution.py:217-234) with pageable host buffers.  The input goes up through
// the pinned ring; the output rows of finished window batches leave through it (copy stream + a few CPU threads) while the
// following batches compute, so only the last batch's rows are copied after the kernels end.
extern "C" int wowsr_enhance_host(wowsr_ctx* ctx, const uint8_t* img_host, int32_t H, int32_t W, int32_t tile_size,
                                  uint8_t* out_host, float* out_f32_host) {
  if (!ctx || !img_host || !out_host) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  size_t in_bytes = (size_t)H * W * 3, out_bytes = in_bytes * 16;
  if (int e = wowsr_ensure(ctx, ctx->img_in, in_bytes)) return e;
  if (int e = wowsr_ensure(ctx, ctx->img_out, out_bytes)) return e;
  if (out_f32_host)
    if (int e = wowsr_ensure(ctx, ctx->img_out_f32, out_bytes * 4)) return e;
  if (int e = stage_init(ctx)) return e;
  const size_t row = (size_t)W * 3;
  if (int e = stage_in(ctx, (uint8_t*)ctx->img_in.p, row, img_host, row, row, H, [](int, int, cudaEvent_t) { return 0; })) return e;
  WCUDA(ctx, cudaEventRecord(ctx->stage_sync, ctx->copy_stream));
  WCUDA(ctx, cudaStreamWaitEvent(0, ctx->stage_sync, 0));
  StageOut sink(ctx, out_host, row * 4, row * 4);
  if (int e = enhance_dev_impl(ctx, (const uint8_t*)ctx->img_in.p, H, W, tile_size, (uint8_t*)ctx->img_out.p,
                               out_f32_host ? (float*)ctx->img_out_f32.p : nullptr, nullptr, &sink)) {
    sink.finish();
    return e;
  }
  if (out_f32_host) WCUDA(ctx, cudaMemcpyAsync(out_f32_host, ctx->img_out_f32.p, out_bytes * 4, cudaMemcpyDeviceToHost, 0));
  if (int e = sink.finish()) return e;
  WCUDA(ctx, cudaStreamSynchronize(0));
  return WOWSR_OK;
}

extern "C" int32_t wowsr_debug_trace(wowsr_ctx* ctx, int64_t* out, int32_t cap) {
  if (!ctx || !out || !ctx->trace_buf.p) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  int n = cap < 512 ? cap : 512;
  if (cudaMemcpy(out, ctx->trace_buf.p, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return WOWSR_ERR_CUDA;
  ctx->trace_counter = 0;
  return n;
}

extern "C" int32_t wowsr_debug_roll_plan(int32_t n_win, int32_t h, int32_t w, int32_t strip_x0, int32_t pair, int32_t max_units,
                                         int32_t* tasks, int32_t cap_tasks, int32_t* off, int32_t cap_off, int32_t* info) {
  if (n_win < 1 || h < 1 || w < 1 || strip_x0 < 1 || strip_x0 > w || max_units < 1 || (strip_x0 < w && max_units < 2)) return WOWSR_ERR_ARG;
  RollPlanHost H;
  roll_plan_build(H, n_win, h, w, strip_x0, pair != 0, max_units);
  if (info) { info[0] = H.units; info[1] = H.units_h; info[2] = (int32_t)H.off.size(); info[3] = 0; }
  const int32_t n = (int32_t)H.tasks.size();
  if (tasks && cap_tasks >= n) memcpy(tasks, H.tasks.data(), (size_t)n * sizeof(RollTask));
  if (off && cap_off >= (int32_t)H.off.size()) memcpy(off, H.off.data(), H.off.size() * sizeof(int));
  return n;
}

extern "C" int32_t wowsr_get_timing(const wowsr_ctx* ctx, float* ms, int32_t cap) {
  if (!ctx || !ms) return WOWSR_ERR_ARG;
  int n = cap < 5 ? cap : 5;
  for (int i = 0; i < n; i++) ms[i] = ctx->timing[i];
  return n;
}

extern "C" int wowsr_conv3x3_host(wowsr_ctx* ctx, const float* in, int32_t n, int32_t h, int32_t w, int32_t cin,
                                  const float* weight, const float* bias, int32_t cout, int32_t act, int32_t precision,
                                  float* out) {
  if (!ctx || !in || !weight || !out || n < 1 || h < 1 || w < 1) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  const bool fp16 = precision == WOWSR_PREC_FP16;
  ConvNet* net = new ConvNet();
  net->body_fp16 = net->tail_fp16 = fp16;
  net->layers.emplace_back();
  int e = upload_layer(ctx, net->layers[0], weight, bias, cin, cout, fp16);
  const size_t px = (size_t)n * h * w;
  const int C = (cin + 63) / 64 * 64;
  DevBuf din, dout;
  if (!e) e = wowsr_ensure(ctx, din, px * C * 2);
  if (!e) e = wowsr_ensure(ctx, dout, px * 64 * 4);
  if (!e) e = wowsr_ensure(ctx, net->err, 4);
  if (!e) {
    std::vector<uint16_t> t(px * C, 0);
    for (size_t p = 0; p < px; p++)
      for (int c = 0; c < cin; c++) t[p * C + c] = to_t(in[p * cin + c], fp16);
    cudaError_t ce = cudaMemcpy(din.p, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemset(dout.p, 0, px * 64 * 4);
    if (ce == cudaSuccess) ce = cudaMemset(net->err.p, 0, 4);
    if (ce != cudaSuccess) e = wowsr_fail(ctx, WOWSR_ERR_CUDA, "conv3x3_host upload: %s", cudaGetErrorString(ce));
  }
  if (!e) {
    LayerIO io;
    io.in = din.p; io.in_C = C; io.Nw = n; io.h = h; io.w = w; io.act = act;
    io.out_f32_a = (float*)dout.p;
    e = run_conv(ctx, net, net->layers[0], io, 0);
  }
  if (!e) e = check_err_flag(ctx, net, 0);
  if (!e) {
    std::vector<float> o(px * 64);
    cudaError_t ce = cudaMemcpy(o.data(), dout.p, o.size() * 4, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) e = wowsr_fail(ctx, WOWSR_ERR_CUDA, "conv3x3_host download: %s", cudaGetErrorString(ce));
    else
      for (size_t p = 0; p < px; p++)
        for (int c = 0; c < cout; c++) out[p * cout + c] = o[p * 64 + c];
  }
  if (din.p) cudaFree(din.p);
  if (dout.p) cudaFree(dout.p);
  wowsr_net_free(net);
  return e;
}

// ---------------------------------------------------------------------------------------------
// EDSR-baseline x4 ("farm SR" variant named by BASELINE; the reference reaches it through
// cv2.dnn_superres with an external EDSR_x4.pb, super_resolution.py:92-124,196 — third-party, parity
// unpinned).  Layer list restated in oracle/edsr_ref.py; tensor order documented in include/wowsr.h.
// ---------------------------------------------------------------------------------------------

namespace {
const float kEdsrMean[3] = {103.1545782f, 111.561547f, 114.35629928f};  // BGR, 0..255 range

F32Layout plain_blocked(int nb, int h, int w) {
  F32Layout fl;
  const int wm = w / TC_RUN * TC_RUN, rem = w - wm;
  const bool strip = rem > 0 && wm > 0 && h >= 64;
  fl.x0 = strip ? wm : w;
  fl.wpb = strip ? wm / 32 : (w + 31) / 32;
  fl.rem = strip ? rem : 0;
  fl.hpb = (h + 31) / 32;
  fl.strip_off = (long long)nb * h * fl.wpb * 32 * 64;
  return fl;
}
}  // namespace

extern "C" int wowsr_load_edsr(wowsr_ctx* ctx, int32_t num_block, int32_t num_feat, float res_scale,
                               const float* const* tensors, int32_t n_tensors, int32_t precision) {
  if (!ctx || !tensors) return WOWSR_ERR_ARG;
  if (num_feat != 64) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "EDSR-baseline has 64 features");
  const int n_convs = 2 * num_block + 5;
  if (n_tensors != 2 * n_convs) return wowsr_fail(ctx, WOWSR_ERR_ARG, "expected %d tensors, got %d", 2 * n_convs, n_tensors);
  DeviceGuard g(ctx->device);
  wowsr_net_free(ctx->edsr);
  ctx->edsr = nullptr;
  ConvNet* net = new ConvNet();
  net->kind = 1;
  net->num_block = num_block;
  net->res_scale = res_scale;
  net->body_fp16 = net->tail_fp16 = precision != WOWSR_PREC_BF16;  // EDSR works on the 0..255 range: fp16 unless bf16_pure
  const bool f16 = net->body_fp16;
  int t = 0;
  int e = upload_first(ctx, net, tensors[0], tensors[1], 3, 64);
  t = 2;
  for (int i = 0; i < 2 * num_block + 1 && !e; i++) {  // resblock convs + body_end
    net->layers.emplace_back();
    e = upload_layer(ctx, net->layers.back(), tensors[t], tensors[t + 1], 64, 64, f16);
    t += 2;
  }
  for (int u = 0; u < 2 && !e; u++) {  // 64 -> 256 + depth-to-space(2): one 64 -> 64 layer per sub-pixel phase s = 2 dy + dx
    const float* w = tensors[t];
    const float* b = tensors[t + 1];
    for (int grp = 0; grp < 4 && !e; grp++) {
      std::vector<float> wg((size_t)64 * 64 * 9), bg(64);
      for (int c = 0; c < 64; c++) {
        const int src = c * 4 + grp;  // PixelShuffle: channel c*4 + s -> (c, sub-pixel s)
        memcpy(&wg[(size_t)c * 64 * 9], &w[(size_t)src * 64 * 9], sizeof(float) * 64 * 9);
        bg[c] = b[src];
      }
      net->layers.emplace_back();
      e = upload_layer(ctx, net->layers.back(), wg.data(), bg.data(), 64, 64, f16);
    }
    t += 2;
  }
  if (!e) {
    net->layers.emplace_back();
    e = upload_layer(ctx, net->layers.back(), tensors[t], tensors[t + 1], 64, 3, f16);
  }
  if (!e) e = wowsr_ensure(ctx, net->err, 4);
  if (!e && cudaMemset(net->err.p, 0, 4) != cudaSuccess) e = wowsr_fail(ctx, WOWSR_ERR_CUDA, "memset");
  if (e) {
    wowsr_net_free(net);
    return e;
  }
  ctx->edsr = net;
  return WOWSR_OK;
}

// The EDSR forward on device buffers: img [H,W,3] BGR u8 (pitch W*3) -> out [4H,4W,3] u8 (pitch 4W*3), optional fp32 copy.
static int edsr_forward(wowsr_ctx* ctx, ConvNet* net, const uint8_t* img_dev, int32_t H, int32_t W, uint8_t* out_dev, float* out_f32_dev,
                        cudaStream_t st) {
  const size_t px = (size_t)H * W;
  const bool f16 = net->body_fp16;
  const F32Layout fl = plain_blocked(1, H, W);
  const size_t pxb = (size_t)H * fl.wpb * 32 + (size_t)fl.rem * fl.hpb * 32;
  if (int e = wowsr_ensure(ctx, net->dense0, px * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->dense1, px * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->feat, pxb * 64 * 4)) return e;
  if (int e = wowsr_ensure(ctx, net->trunk, pxb * 64 * 4)) return e;
  if (int e = wowsr_ensure(ctx, net->up1, px * 4 * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->hra, px * 16 * 64 * 2)) return e;
  if (int e = wowsr_ensure(ctx, net->wins, sizeof(WinDev))) return e;
  if (int e = wowsr_ensure(ctx, net->winxy, 8)) return e;
  WinDev wd{0, 0, 0, 0, 4 * W, 4 * H};
  int wxy[2] = {0, 0};
  WCUDA(ctx, cudaMemcpyAsync(net->wins.p, &wd, sizeof wd, cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaMemcpyAsync(net->winxy.p, wxy, 8, cudaMemcpyHostToDevice, st));
  WCUDA(ctx, cudaStreamSynchronize(st));
  WCUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  {
    FirstParams F;
    memset(&F, 0, sizeof F);
    F.img = img_dev; F.pitch = (long long)W * 3; F.cin = 3; F.win_xy = (const int*)net->winxy.p;
    F.Nw = 1; F.h = H; F.w = W; F.weight = net->first_w; F.bias = net->first_b;
    F.f32 = fl;
    F.f32_a = (float*)net->feat.p;  // the first resblock reads its residual from `feat`: no second fp32 copy of the head's output
    F.out_t = net->dense0.p; F.out_stride = 64; F.out_fp16 = f16; F.in_scale_div = 1.0f;
    for (int i = 0; i < 3; i++) F.sub[i] = kEdsrMean[i];
    dim3 grid((unsigned)((px + 127) / 128), 4);
    conv_first_kernel<<<grid, 128, 0, st>>>(F);
    WLAUNCH_CHECK(ctx);
  }
  WCUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  void* a = net->dense0.p;
  void* b = net->dense1.p;
  size_t li = 0;
  for (int blk = 0; blk < net->num_block; blk++) {
    LayerIO io;
    io.in = a; io.in_C = 64; io.Nw = 1; io.h = H; io.w = W; io.act = 2;
    io.out_t = b; io.out_stride = 64; io.out_fp16 = f16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
    LayerIO io2;
    io2.in = b; io2.in_C = 64; io2.Nw = 1; io2.h = H; io2.w = W;
    io2.f32 = fl; io2.scale1 = net->res_scale; io2.res1 = (const float*)(blk == 0 ? net->feat.p : net->trunk.p); io2.out_f32_a = (float*)net->trunk.p;
    io2.out_t = a; io2.out_stride = 64; io2.out_fp16 = f16;
    if (int e = run_conv(ctx, net, net->layers[li++], io2, st)) return e;
  }
  {  // body end + global skip
    LayerIO io;
    io.in = a; io.in_C = 64; io.Nw = 1; io.h = H; io.w = W;
    io.f32 = fl; io.scale1 = 1.0f; io.res1 = (const float*)net->feat.p;
    io.out_t = b; io.out_stride = 64; io.out_fp16 = f16;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  WCUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  // Upsampler: conv 64 -> 256 + PixelShuffle(2) as FOUR 64 -> 64 launches, one per sub-pixel phase (dy, dx): each writes all 64
  // channels (one full 128-byte line) of the output pixels (2y + dy, 2x + dx) through the plain epilogue with a pixel stride
  // of two pixels and a row pitch of two rows.  (Round 1 grouped by output channel: four launches each wrote a 32-byte
  // quarter of every line through the generic epilogue: 44 % tensor pipe on these 8 launches, a third of the step.)
  for (int up = 0; up < 2; up++) {
    const int hh = up ? 2 * H : H, ww = up ? 2 * W : W;
    uint16_t* dst = (uint16_t*)(up ? net->hra.p : net->up1.p);
    for (int grp = 0; grp < 4; grp++) {
      LayerIO io;
      io.in = up ? net->up1.p : b; io.in_C = 64; io.Nw = 1; io.h = hh; io.w = ww;
      io.out_t = dst + ((long long)(grp >> 1) * (2 * ww) + (grp & 1)) * 64;
      io.out_stride = 128; io.out_row = (long long)4 * ww * 64; io.out_fp16 = f16;
      if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
    }
  }
  {  // tail 64 -> 3, + mean, round, saturate
    LayerIO io;
    io.in = net->hra.p; io.in_C = 64; io.Nw = 1; io.h = 4 * H; io.w = 4 * W;
    io.final = 1; io.final_scale = 1.0f; io.final_round = 1;
    for (int i = 0; i < 3; i++) io.final_add[i] = kEdsrMean[i];
    io.out_u8 = out_dev; io.out_u8_pitch = (long long)W * 4 * 3;
    io.out_img_f32 = out_f32_dev; io.out_img_f32_pitch = (long long)W * 4 * 3;
    io.wins = (const WinDev*)net->wins.p;
    if (int e = run_conv(ctx, net, net->layers[li++], io, st)) return e;
  }
  WCUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  if (int e = check_err_flag(ctx, net, st)) return e;  // synchronises the stream
  if (int e = balance_update(ctx, net)) return e;
  float t0 = 0, t1 = 0, t2 = 0;
  cudaEventElapsedTime(&t0, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&t1, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&t2, ctx->ev[2], ctx->ev[3]);
  ctx->timing[0] = t0 + t1 + t2; ctx->timing[1] = t0; ctx->timing[2] = t1; ctx->timing[3] = t2;
  ctx->timing[4] = 0;
  return WOWSR_OK;
}

extern "C" int wowsr_edsr_upsample_dev(wowsr_ctx* ctx, const uint8_t* img_dev, int32_t H, int32_t W, uint8_t* out_dev,
                                       float* out_f32_dev, void* stream) {
  if (!ctx || !img_dev || !out_dev || H < 1 || W < 1) return WOWSR_ERR_ARG;
  if (!ctx->edsr) return wowsr_fail(ctx, WOWSR_ERR_STATE, "wowsr_load_edsr has not been called");
  DeviceGuard g(ctx->device);
  return edsr_forward(ctx, ctx->edsr, img_dev, H, W, out_dev, out_f32_dev, (cudaStream_t)stream);
}

extern "C" int wowsr_edsr_upsample_host(wowsr_ctx* ctx, const uint8_t* img_host, int32_t H, int32_t W, uint8_t* out_host,
                                        float* out_f32_host) {
  if (!ctx || !img_host || !out_host || H < 1 || W < 1) return WOWSR_ERR_ARG;
  ConvNet* net = ctx->edsr;
  if (!net) return wowsr_fail(ctx, WOWSR_ERR_STATE, "wowsr_load_edsr has not been called");
  DeviceGuard g(ctx->device);
  cudaStream_t st = 0;
  const size_t in_bytes = (size_t)H * W * 3, out_bytes = in_bytes * 16;
  if (int e = wowsr_ensure(ctx, ctx->img_in, in_bytes)) return e;
  if (int e = wowsr_ensure(ctx, ctx->img_out, out_bytes)) return e;
  if (out_f32_host)
    if (int e = wowsr_ensure(ctx, ctx->img_out_f32, out_bytes * 4)) return e;
  WCUDA(ctx, cudaMemcpyAsync(ctx->img_in.p, img_host, in_bytes, cudaMemcpyHostToDevice, st));
  if (int e = edsr_forward(ctx, net, (const uint8_t*)ctx->img_in.p, H, W, (uint8_t*)ctx->img_out.p,
                           out_f32_host ? (float*)ctx->img_out_f32.p : nullptr, st))
    return e;
  WCUDA(ctx, cudaMemcpyAsync(out_host, ctx->img_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
  if (out_f32_host) WCUDA(ctx, cudaMemcpyAsync(out_f32_host, ctx->img_out_f32.p, out_bytes * 4, cudaMemcpyDeviceToHost, st));
  WCUDA(ctx, cudaStreamSynchronize(st));
  return WOWSR_OK;
}
