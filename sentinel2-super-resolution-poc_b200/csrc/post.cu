// WOW / farm post-process on the GPU: CLAHE (pass A histograms + LUT build) and the fused pass B
// (LUT interpolation -> Lab->RGB -> separable fixed-point Gaussian -> unsharp -> HSV green boost).
//
// Replaces wow_sr._enhance_for_crops (server/app/wow_sr.py:187-209) and the farm trio
// (server/app/farm_sr.py:61-108 as called at :170-178).  The arithmetic is OpenCV's, restated
// exactly per SURVEY.md Appendix A: integer colour conversions, fp32 CLAHE interpolation without
// FMA contraction, exact fixed-point blur with a single rounding, FMA in HSV->RGB.  All fp32
// expressions that must round step by step use __fmul_rn/__fadd_rn so -fmad cannot fuse them.
#include "common.h"
#include "hoststage.h"
#include "ptx.cuh"
#include <algorithm>
#include <cmath>

namespace {

struct ImgView {
  const uint8_t* data;
  long long pitch;
  int W, H, y0, rows;
};
struct OutView {
  uint8_t* data;
  long long pitch;
  int W, H, y0, rows;
};

__host__ __device__ inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
  }
  return i;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// App. A.1 — only L
__device__ __forceinline__ int rgb_to_L(int r, int g, int b, const uint16_t* gam, const uint16_t* cbrt) {
  int R = gam[r], G = gam[g], B = gam[b];
  int fY = cbrt[(871 * R + 2929 * G + 296 * B + 2048) >> 12];
  return clampi((296 * fY - 1336934 + 16384) >> 15, 0, 255);
}

// ---------------------------------------------------------------------------------------------
// pass A: per-tile histograms of L, warp-aggregated shared-memory atomics
// ---------------------------------------------------------------------------------------------

constexpr int HIST_THREADS = 256;

__global__ void __launch_bounds__(HIST_THREADS)
clahe_hist_kernel(ImgView img, const WowsrTables* __restrict__ tabs, int grid, int tw, int th, int prow0, int prow1,
                  int rows_per_block, int chunks_per_tile, int vec_ok, int use_match, uint32_t* __restrict__ hist) {
  __shared__ __align__(16) uint16_t s_gam[256];
  __shared__ __align__(16) uint16_t s_cbrt[3072];
  __shared__ __align__(16) uint32_t s_h[HIST_THREADS / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = blockIdx.x;
  const int ty = blockIdx.y / chunks_per_tile, chunk = blockIdx.y % chunks_per_tile;
  int r0 = ty * th + chunk * rows_per_block;
  int r1 = min(r0 + rows_per_block, (ty + 1) * th);
  r0 = max(r0, prow0);
  r1 = min(r1, prow1);
  if (r0 >= r1) return;
  {  // gam[256] and cbrt[3072] open WowsrTables back to back: 416 16-byte pieces
    const uint4* src = reinterpret_cast<const uint4*>(tabs);
    for (int i = tid; i < 32; i += HIST_THREADS) reinterpret_cast<uint4*>(s_gam)[i] = __ldg(src + i);
    for (int i = tid; i < 384; i += HIST_THREADS) reinterpret_cast<uint4*>(s_cbrt)[i] = __ldg(src + 32 + i);
    for (int i = tid; i < (HIST_THREADS / 32) * 64; i += HIST_THREADS) reinterpret_cast<uint4*>(&s_h[0][0])[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  // Thread -> 4-pixel group mapping without a division per item: `lpr` lanes (a power of two, 32 .. 256) walk along a tile
  // row, 256 / lpr rows are in flight at once.  Row and pass counters are warp-uniform, so the warp-wide histogram update
  // below always runs with 32 lanes.  Two items per iteration, the loads of both issued before either is consumed.
  const int gpr = (tw + 3) >> 2;  // 4-pixel groups per tile row
  int lpr_log = 5;
  while ((1 << lpr_log) < gpr && (1 << lpr_log) < HIST_THREADS) lpr_log++;
  const int lpr = 1 << lpr_log, rpp = HIST_THREADS >> lpr_log;
  const int g0 = tid & (lpr - 1);
  const int passes = (gpr + lpr - 1) >> lpr_log;
  int row = r0 + (tid >> lpr_log), pass = 0;

  auto issue = [&](int irow, int ipass, uint32_t (&w)[3]) -> int {  // 0: nothing, 1: three packed words loaded, 2: border item
    const int g = g0 + (ipass << lpr_log);
    if (g >= gpr) return 0;
    const int px0 = tx * tw + g * 4;
    if (vec_ok && g * 4 + 3 < tw && px0 + 3 < img.W) {
      const int sy = reflect101(irow, img.H) - img.y0;
      const uint32_t* q = reinterpret_cast<const uint32_t*>(img.data + (long long)sy * img.pitch + px0 * 3);
      w[0] = __ldg(q), w[1] = __ldg(q + 1), w[2] = __ldg(q + 2);
      return 1;
    }
    return 2;
  };
  auto finish = [&](int mode, int irow, int ipass, const uint32_t (&w)[3], uint32_t (&bins)[4]) {
    bins[0] = bins[1] = bins[2] = bins[3] = 0xFFFFFFFFu;
    if (mode == 1) {
      bins[0] = rgb_to_L(w[0] & 255, (w[0] >> 8) & 255, (w[0] >> 16) & 255, s_gam, s_cbrt);
      bins[1] = rgb_to_L(w[0] >> 24, w[1] & 255, (w[1] >> 8) & 255, s_gam, s_cbrt);
      bins[2] = rgb_to_L((w[1] >> 16) & 255, w[1] >> 24, w[2] & 255, s_gam, s_cbrt);
      bins[3] = rgb_to_L((w[2] >> 8) & 255, (w[2] >> 16) & 255, w[2] >> 24, s_gam, s_cbrt);
    } else if (mode == 2) {  // right / bottom border of the padded image, unaligned rows: byte loads through the reflection
      const int g = g0 + (ipass << lpr_log);
      const int px0 = tx * tw + g * 4;
      const int npx = min(4, tw - g * 4);
      const uint8_t* rp = img.data + (long long)(reflect101(irow, img.H) - img.y0) * img.pitch;
      for (int k = 0; k < npx; k++) {
        const uint8_t* p = rp + reflect101(px0 + k, img.W) * 3;
        bins[k] = rgb_to_L(__ldg(p), __ldg(p + 1), __ldg(p + 2), s_gam, s_cbrt);
      }
    }
  };
  auto update = [&](const uint32_t (&bins)[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      // Option hist_match: 2 (default) plain per-lane shared atomics — the compiler emits ATOMS.POPC.INC, which the
      // hardware aggregates per address; 0: a warp whose 32 lanes fall in one bin issues ONE atomic of 32 (shuffle + vote
      // per pixel: 5 us slower on the 4096 x 4096 workload); 1: full __match_any_sync grouping (2.6x slower on image-like
      // data: the match costs more than the replays it saves).
      const uint32_t bin = bins[k];
      if (use_match == 2) {
        if (bin != 0xFFFFFFFFu) atomicAdd(&s_h[warp][bin], 1u);
        continue;
      }
      const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, bin, 0);
      if (__all_sync(0xFFFFFFFFu, bin == b0)) {
        if (lane == 0 && bin != 0xFFFFFFFFu) atomicAdd(&s_h[warp][bin], 32u);
      } else if (use_match) {
        unsigned mm = __match_any_sync(0xFFFFFFFFu, bin);
        if (bin != 0xFFFFFFFFu && lane == __ffs(mm) - 1) atomicAdd(&s_h[warp][bin], (uint32_t)__popc(mm));
      } else if (bin != 0xFFFFFFFFu) {
        atomicAdd(&s_h[warp][bin], 1u);
      }
    }
  };
  while (row < r1) {
    const int rowA = row, passA = pass;
    if (++pass == passes) pass = 0, row += rpp;
    const bool hasB = row < r1;
    const int rowB = row, passB = pass;
    if (hasB && ++pass == passes) pass = 0, row += rpp;
    uint32_t wA[3], wB[3], bins[4];
    const int mA = issue(rowA, passA, wA);
    const int mB = hasB ? issue(rowB, passB, wB) : 0;
    finish(mA, rowA, passA, wA, bins);
    update(bins);
    if (hasB) {
      finish(mB, rowB, passB, wB, bins);
      update(bins);
    }
  }
  __syncthreads();
  for (int b = tid; b < 256; b += HIST_THREADS) {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < HIST_THREADS / 32; w++) sum += s_h[w][b];
    if (sum) atomicAdd(&hist[(ty * grid + tx) * 256 + b], sum);
  }
}

// ---------------------------------------------------------------------------------------------
// LUT build: clip, redistribute, cumulative sum, scale (App. A.2)
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
clahe_lut_kernel(const uint32_t* __restrict__ hist, int clip_limit, float lut_scale, uint8_t* __restrict__ luts) {
  __shared__ uint32_t s_red[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t h = hist[blockIdx.x * 256 + t];
  uint32_t excess = h > (uint32_t)clip_limit ? h - clip_limit : 0;
  if (h > (uint32_t)clip_limit) h = clip_limit;
  uint32_t e = excess;
  for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
  if (lane == 0) s_red[warp] = e;
  __syncthreads();
  uint32_t clipped = 0;
  for (int w = 0; w < 8; w++) clipped += s_red[w];
  uint32_t batch = clipped >> 8;
  uint32_t resid = clipped - (batch << 8);
  h += batch;
  if (resid) {
    uint32_t step = 256 / resid;
    if (step < 1) step = 1;
    if (t % step == 0 && t / step < resid) h++;
  }
  // inclusive scan over 256 bins
  uint32_t v = h;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if (lane >= o) v += n;
  }
  __syncthreads();
  if (lane == 31) s_red[warp] = v;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; w++) base += s_red[w];
  v += base;
  int q = __float2int_rn(__fmul_rn(__uint2float_rn(v), lut_scale));
  luts[blockIdx.x * 256 + t] = (uint8_t)clampi(q, 0, 255);
}

// ---------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------

constexpr int PB_TX = 64, PB_TY = 64, PB_THREADS = 512, PB_MAXR = 8;

struct PostK {
  int stages;
  int grid, tw, th;
  float inv_tw, inv_th;
  int r;                 // effective blur radius (outermost non-zero tap)
  int taps[2 * PB_MAXR + 1];
  float alpha, beta;
  int hue_lo, hue_hi;
  float sat;
  float inv255, hscale;
  int tail_x;            // first column handled by cv2's scalar HSV->RGB tail (W - W%32)
};

struct SmemTabs {
  uint16_t gam[256];
  uint16_t cbrt[3072];
  uint16_t lab_y[256];
  uint16_t lab_ify[256];
  uint8_t invgam[4096];
  uint32_t sdiv[256];
  uint32_t hdiv[256];
};

__device__ __forceinline__ int ab2xz(int i) {
  // App. A.3: C integer division truncates toward zero
  if (i <= 3390) return (i * 108) / 841 - 290;
  return ((i * i) / 16384 * i) / 16384;
}

__device__ __forceinline__ int ds(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// RGB -> HSV -> green boost -> HSV -> RGB (A.6 + glue wow_sr.py:200-207)
__device__ __forceinline__ uint32_t vegetation_pixel(int r, int g, int b, bool tail, const PostK& k, const SmemTabs& T) {
  int v = max(max(r, g), b), mn = min(min(r, g), b);
  int diff = v - mn;
  int s = (int)((diff * T.sdiv[v] + 2048u) >> 12);
  int h;
  if (v == r) h = g - b;
  else if (v == g) h = b - r + 2 * diff;
  else h = r - g + 4 * diff;
  h = (h * (int)T.hdiv[diff] + 2048) >> 12;
  if (h < 0) h += 180;
  if (h > k.hue_lo && h < k.hue_hi) s = (int)fminf(__fmul_rn((float)s, k.sat), 255.0f);
  // HSV -> RGB: FMA inside 1 - s*f, plain multiplies outside
  float sf = __fmul_rn((float)s, k.inv255), vf = __fmul_rn((float)v, k.inv255);
  float h6 = __fmul_rn((float)h, k.hscale);
  float secf = floorf(h6);
  float f = __fsub_rn(h6, secf);
  int sec = (int)secf;
  if (sec >= 6) sec -= 6;
  float t0 = vf;
  float t1 = __fmul_rn(vf, __fsub_rn(1.0f, sf));
  float t2 = __fmul_rn(vf, __fmaf_rn(-sf, f, 1.0f));
  float t3 = __fmul_rn(vf, __fmaf_rn(-sf, __fsub_rn(1.0f, f), 1.0f));
  float bq, gq, rq;
  switch (sec) {
    case 0: bq = t1; gq = t3; rq = t0; break;
    case 1: bq = t1; gq = t0; rq = t2; break;
    case 2: bq = t3; gq = t0; rq = t1; break;
    case 3: bq = t0; gq = t2; rq = t1; break;
    case 4: bq = t0; gq = t1; rq = t3; break;
    default: bq = t2; gq = t1; rq = t0; break;
  }
  rq = __fmul_rn(rq, 255.0f);
  gq = __fmul_rn(gq, 255.0f);
  bq = __fmul_rn(bq, 255.0f);
  int ri, gi, bi;
  if (tail) {  // cv2's scalar row tail rounds (saturate_cast) where its SIMD body truncates
    ri = __float2int_rn(rq); gi = __float2int_rn(gq); bi = __float2int_rn(bq);
  } else {
    ri = (int)rq; gi = (int)gq; bi = (int)bq;
  }
  return (uint32_t)clampi(ri, 0, 255) | ((uint32_t)clampi(gi, 0, 255) << 8) | ((uint32_t)clampi(bi, 0, 255) << 16);
}

// Stage-1 arithmetic of one halo pixel: RGB -> Lab (A.1), CLAHE bilinear LUT interpolation on L (A.2), Lab -> RGB (A.3).
__device__ __forceinline__ uint32_t enhance_px(const uint8_t* __restrict__ p, bool do_clahe, const SmemTabs& T,
                                               const uint8_t* __restrict__ lut1, const uint8_t* __restrict__ lut2, int ctx1,
                                               int ctx2, float cxa, float cxa1, float ya, float ya1) {
  const int cr = __ldg(p), cg = __ldg(p + 1), cb = __ldg(p + 2);
  if (!do_clahe) return (uint32_t)cr | ((uint32_t)cg << 8) | ((uint32_t)cb << 16);
  // RGB -> Lab (A.1)
  const int R = T.gam[cr], G = T.gam[cg], B = T.gam[cb];
  const int fX = T.cbrt[ds(1777 * R + 1541 * G + 778 * B, 12)];
  const int fY = T.cbrt[ds(871 * R + 2929 * G + 296 * B, 12)];
  const int fZ = T.cbrt[ds(73 * R + 448 * G + 3575 * B, 12)];
  int L = clampi(ds(296 * fY - 1336934, 15), 0, 255);
  const int a = clampi(ds(500 * (fX - fY) + 128 * 32768, 15), 0, 255);
  const int bb = clampi(ds(200 * (fY - fZ) + 128 * 32768, 15), 0, 255);
  // CLAHE bilinear LUT interpolation (A.2): fp32, every multiply and add rounded separately
  const float p00 = (float)__ldg(lut1 + ctx1 + L), p01 = (float)__ldg(lut1 + ctx2 + L);
  const float p10 = (float)__ldg(lut2 + ctx1 + L), p11 = (float)__ldg(lut2 + ctx2 + L);
  const float top = __fadd_rn(__fmul_rn(p00, cxa1), __fmul_rn(p01, cxa));
  const float bot = __fadd_rn(__fmul_rn(p10, cxa1), __fmul_rn(p11, cxa));
  L = clampi(__float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya))), 0, 255);
  // Lab -> RGB (A.3)
  const int yy = T.lab_y[L], ify = T.lab_ify[L];
  const int adiv = ((5 * a * 53687 + 128) >> 13) - 4194;
  const int bdiv = ((bb * 41943 + 16) >> 9) - 10485 + 1;
  const int X = ab2xz(ify + adiv), Z = ab2xz(ify - bdiv);
  const int ro = clampi(ds(12615 * X - 6296 * yy - 2223 * Z, 14), 0, 4095);
  const int go = clampi(ds(-3773 * X + 7684 * yy + 185 * Z, 14), 0, 4095);
  const int bo = clampi(ds(217 * X - 836 * yy + 4715 * Z, 14), 0, 4095);
  return (uint32_t)T.invgam[ro] | ((uint32_t)T.invgam[go] << 8) | ((uint32_t)T.invgam[bo] << 16);
}

// RAD = blur radius (compile time: the tap loops unroll and the halo geometry is constant).  Work split:
// one warp per tile row, lanes over columns, so everything that depends only on the row (source row pointer,
// CLAHE y-interpolation, LUT rows) or only on the column (reflected x, x-interpolation) is hoisted out of the
// per-pixel path; the kernel is instruction-issue bound, not HBM bound.
template <int RAD>
__global__ void __launch_bounds__(PB_THREADS)
post_apply_kernel(ImgView img, OutView out, const WowsrTables* __restrict__ tabs, const uint8_t* __restrict__ luts,
                  PostK k, int row0, int row1, int tiles_x, int n_tiles) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemTabs& T = *reinterpret_cast<SmemTabs*>(smem_raw);
  constexpr int r = RAD;
  constexpr int EW = PB_TX + 2 * r, EH = PB_TY + 2 * r;
  constexpr int NFULL = EW / 32;          // full 32-lane column blocks of the halo tile
  constexpr int TAILW = EW - 32 * NFULL;  // remaining columns (2r)
  constexpr int NWARP = PB_THREADS / 32;
  uint32_t* E = reinterpret_cast<uint32_t*>(smem_raw + ((sizeof(SmemTabs) + 15) & ~15));  // [EH][EW] packed rgb
  uint2* Hs = reinterpret_cast<uint2*>(E + EH * EW + ((EH * EW) & 1));                    // [EH][PB_TX] 3 x u16 (+pad)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tabs);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
    for (int i = tid; i < (int)(sizeof(SmemTabs) / 4); i += PB_THREADS) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const bool do_clahe = k.stages & WOWSR_STAGE_CLAHE;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int bx = (tile % tiles_x) * PB_TX;
    const int by = row0 + (tile / tiles_x) * PB_TY;
    // ---- per-column constants of this tile (registers), full 32-lane column blocks only ----
    int goff[NFULL], ctx1[NFULL], ctx2[NFULL];
    float cxa[NFULL], cxa1[NFULL];
#pragma unroll
    for (int c = 0; c < NFULL; c++) {
      const int gx = reflect101(bx - r + lane + 32 * c, img.W);
      goff[c] = gx * 3;
      const float txf = __fsub_rn(__fmul_rn((float)gx, k.inv_tw), 0.5f);
      const int t1 = (int)floorf(txf);
      cxa[c] = __fsub_rn(txf, (float)t1);
      cxa1[c] = __fsub_rn(1.0f, cxa[c]);
      ctx2[c] = min(t1 + 1, k.grid - 1) << 8;
      ctx1[c] = max(t1, 0) << 8;
    }
    // ---- stage 1: enhanced RGB for the halo tile (reflect-101 at the image borders) ----
    for (int ey = warp; ey < EH; ey += NWARP) {
      const int gy = reflect101(by - r + ey, img.H);
      const int gyb = gy - img.y0;
      uint32_t* erow = E + ey * EW;
      if (gyb < 0 || gyb >= img.rows) {  // tile overhang beyond the band: never consumed
#pragma unroll
        for (int c = 0; c < NFULL; c++) erow[lane + 32 * c] = 0;
        continue;
      }
      const uint8_t* rp = img.data + (long long)gyb * img.pitch;
      const float tyf = __fsub_rn(__fmul_rn((float)gy, k.inv_th), 0.5f);
      const int t1 = (int)floorf(tyf);
      const float ya = __fsub_rn(tyf, (float)t1), ya1 = __fsub_rn(1.0f, ya);
      const uint8_t* lut1 = luts + ((max(t1, 0) * k.grid) << 8);
      const uint8_t* lut2 = luts + ((min(t1 + 1, k.grid - 1) * k.grid) << 8);
#pragma unroll
      for (int c = 0; c < NFULL; c++)
        erow[lane + 32 * c] = enhance_px(rp + goff[c], do_clahe, T, lut1, lut2, ctx1[c], ctx2[c], cxa[c], cxa1[c], ya, ya1);
    }
    // the 2r columns right of the full blocks, all rows, packed over the whole block (a third column block per row
    // would run with 2r of 32 lanes)
    if constexpr (TAILW > 0) {
      for (int idx = tid; idx < EH * TAILW; idx += PB_THREADS) {
        const int ey = idx / TAILW, ex = 32 * NFULL + (idx - ey * TAILW);
        const int gy = reflect101(by - r + ey, img.H);
        const int gyb = gy - img.y0;
        uint32_t e = 0;
        if (gyb >= 0 && gyb < img.rows) {
          const int gx = reflect101(bx - r + ex, img.W);
          const float txf = __fsub_rn(__fmul_rn((float)gx, k.inv_tw), 0.5f);
          const int tx1 = (int)floorf(txf);
          const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
          const float tyf = __fsub_rn(__fmul_rn((float)gy, k.inv_th), 0.5f);
          const int ty1 = (int)floorf(tyf);
          const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
          e = enhance_px(img.data + (long long)gyb * img.pitch + gx * 3, do_clahe, T, luts + ((max(ty1, 0) * k.grid) << 8),
                         luts + ((min(ty1 + 1, k.grid - 1) * k.grid) << 8), max(tx1, 0) << 8, min(tx1 + 1, k.grid - 1) << 8, xa,
                         xa1, ya, ya1);
        }
        E[ey * EW + ex] = e;
      }
    }
    __syncthreads();
    if (r > 0) {
      // ---- stage 2: horizontal pass, exact in u16 ----
      for (int ey = warp; ey < EH; ey += NWARP) {
#pragma unroll
        for (int c = 0; c < PB_TX / 32; c++) {
          const int x = lane + 32 * c;
          const uint32_t* e = E + ey * EW + x;
          // R and G share one multiply-add: sum(taps) = 256, so each 16-bit half stays below 65536 and never carries
          uint32_t arg = 0, ab = 0;
#pragma unroll
          for (int t = 0; t <= 2 * r; t++) {
            const uint32_t px = e[t];
            const uint32_t w = k.taps[t];
            arg += w * __byte_perm(px, 0u, 0x4140);  // r | g << 16
            ab += w * __byte_perm(px, 0u, 0x4442);   // b
          }
          Hs[ey * PB_TX + x] = make_uint2(arg, ab);
        }
      }
      __syncthreads();
    }
    // ---- stage 3: vertical pass + unsharp + vegetation ----
    for (int y = warp; y < PB_TY; y += NWARP) {
      const int gy = by + y;
      if (gy >= row1) break;
      uint8_t* orow = out.data + (long long)(gy - out.y0) * out.pitch;
#pragma unroll
      for (int c = 0; c < PB_TX / 32; c++) {
        const int x = lane + 32 * c;
        const int gx = bx + x;
        if (gx >= img.W) continue;
        const uint32_t e = E[(y + r) * EW + x + r];
        int cr = e & 255, cg = (e >> 8) & 255, cb = (e >> 16) & 255;
        if (r > 0) {
          uint32_t ar = 0, ag = 0, ab = 0;
#pragma unroll
          for (int t = 0; t <= 2 * r; t++) {
            const uint2 hv = Hs[(y + t) * PB_TX + x];
            const uint32_t w = k.taps[t];  // < 256: a two-way dot product with one zero weight picks a 16-bit half
            ar = __dp2a_lo(hv.x, w, ar);
            ag = __dp2a_lo(hv.x, w << 8, ag);
            ab = __dp2a_lo(hv.y, w, ab);
          }
          const int br = (ar + 32768) >> 16, bg = (ag + 32768) >> 16, bb = (ab + 32768) >> 16;
          cr = clampi(__float2int_rn(__fadd_rn(__fmul_rn((float)cr, k.alpha), __fmul_rn((float)br, k.beta))), 0, 255);
          cg = clampi(__float2int_rn(__fadd_rn(__fmul_rn((float)cg, k.alpha), __fmul_rn((float)bg, k.beta))), 0, 255);
          cb = clampi(__float2int_rn(__fadd_rn(__fmul_rn((float)cb, k.alpha), __fmul_rn((float)bb, k.beta))), 0, 255);
        }
        uint32_t o = (uint32_t)cr | ((uint32_t)cg << 8) | ((uint32_t)cb << 16);
        if (k.stages & WOWSR_STAGE_VEG) o = vegetation_pixel(cr, cg, cb, gx >= k.tail_x, k, T);
        uint8_t* q = orow + gx * 3;
        q[0] = (uint8_t)(o & 255);
        q[1] = (uint8_t)((o >> 8) & 255);
        q[2] = (uint8_t)(o >> 16);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// pass B, strip-march kernel (the default): the same arithmetic with ~40 % fewer instructions per pixel.
//
// A CTA owns a strip of 4 * nt columns (hl halo columns on each side) and marches down `seg` output rows (+ RAD warm-up
// rows above and below).  A thread owns FOUR adjacent columns: one row of them is three aligned 32-bit loads and three
// 32-bit stores.  Per row: stage 1 on the four pixels -> ring of 2 RAD + 1 enhanced rows, thread-private, in shared memory
// as 16-bit pairs (r0|r1<<16, r2|r3<<16, g.., b..) -> VERTICAL pass first (one multiply-add per pair: taps sum to 256, a
// 16-bit half never carries) -> the row of vertical sums is the only data exchanged between threads (double-buffered, one
// __syncthreads per row) -> HORIZONTAL pass with two-way dot products on the pairs (weights of two neighbouring taps in one
// constant-bank operand: 4 instead of 7 multiply-adds per channel for RAD = 3) -> single rounding, unsharp, vegetation.
// Compared with post_apply_kernel: no (64 + 2r)^2 / 64^2 halo recompute of stage 1 (1.20x at r = 3, 1.56x for farm's r = 8
// template) but (seg + 2r) / seg, packed loads / stores, CLAHE LUTs in shared memory, saturating float -> u8 conversions
// instead of convert + clamp, branch-free HSV sector selection.
// ---------------------------------------------------------------------------------------------

constexpr int PM_MAXNT = 256, PM_MAXCTA = 512, PM_MAXSEG = 256, PM_MAXPAIR = 10;
constexpr int PM_MISC_BYTES = (2 * (2 * PB_MAXR + 1) * 4 + 24 + 15) & ~15;  // doubled taps + three mbarriers

struct MarchK {
  int nt;        // threads per strip = 4-column groups per strip, halo groups included
  int groups;    // strips (work items) a CTA processes side by side: blockDim.x = groups * nt; tables are shared
  int n_items;   // n_strips * segments
  int hl;        // halo columns on each side: 0 (no blur), 4 (RAD <= 4) or 8
  int swv;       // output columns per strip = 4 * nt - 2 * hl
  int seg;       // output rows per CTA
  int n_strips;
  int in_vec, out_vec;  // 32-bit loads / stores allowed (pitch and base multiples of 4)
  int lut_bytes;        // grid * grid * 256 when the LUTs are staged in shared memory
  uint32_t zero;        // 0, unknown to the compiler (see the first row load in the kernel)
  int* err;             // mapped host word, set to 1 when a wait on the exchange barrier runs out of polls (never in a correct run)
  uint32_t hw[4][PM_MAXPAIR];  // horizontal pass: for pixel j of the group and window pair k, tap of the even column | tap of
                               // the odd column << 8
  int taps2[2 * (2 * PB_MAXR + 1)];  // taps twice in a row: the vertical pass walks the ring slots in storage order
};

__device__ __forceinline__ uint32_t sat_rn_u8(float x) {
  uint32_t r;
  asm("cvt.rni.u8.f32 %0, %1;" : "=r"(r) : "f"(x));  // round to nearest even, saturate to [0, 255]: one F2IP
  return r;
}
// unsigned -> float through the 32-bit conversion (I2FP, ALU pipe); the compiler turns conversions of values it knows to be
// 16-bit into I2F.U16 on the quarter-rate XU pipe (12 per pixel in this kernel)
__device__ __forceinline__ float u2f(uint32_t x) {
  float r;
  asm("cvt.rn.f32.u32 %0, %1;" : "=f"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ uint32_t sat_rz_u8(float x) {
  uint32_t r;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(x));  // truncate, saturate
  return r;
}

// Stage 1 of one pixel in place (A.1 -> A.2 -> A.3); `lut` is the shared-memory copy or the global array.
template <bool LUTS>
__device__ __forceinline__ void enhance_rgb(int& cr, int& cg, int& cb, const SmemTabs& T, const uint8_t* __restrict__ lut,
                                            int l1, int l2, int ctx1, int ctx2, float cxa, float ya, float ya1) {
  const int R = T.gam[cr], G = T.gam[cg], B = T.gam[cb];
  const int fX = T.cbrt[ds(1777 * R + 1541 * G + 778 * B, 12)];
  const int fY = T.cbrt[ds(871 * R + 2929 * G + 296 * B, 12)];
  const int fZ = T.cbrt[ds(73 * R + 448 * G + 3575 * B, 12)];
  int L = clampi(ds(296 * fY - 1336934, 15), 0, 255);
  const int a = clampi(ds(500 * (fX - fY) + 128 * 32768, 15), 0, 255);
  const int bb = clampi(ds(200 * (fY - fZ) + 128 * 32768, 15), 0, 255);
  float p00, p01, p10, p11;
  if constexpr (LUTS) {
    p00 = u2f(lut[l1 + ctx1 + L]), p01 = u2f(lut[l1 + ctx2 + L]);
    p10 = u2f(lut[l2 + ctx1 + L]), p11 = u2f(lut[l2 + ctx2 + L]);
  } else {
    p00 = u2f(__ldg(lut + l1 + ctx1 + L)), p01 = u2f(__ldg(lut + l1 + ctx2 + L));
    p10 = u2f(__ldg(lut + l2 + ctx1 + L)), p11 = u2f(__ldg(lut + l2 + ctx2 + L));
  }
  const float cxa1 = __fsub_rn(1.0f, cxa);
  const float top = __fadd_rn(__fmul_rn(p00, cxa1), __fmul_rn(p01, cxa));
  const float bot = __fadd_rn(__fmul_rn(p10, cxa1), __fmul_rn(p11, cxa));
  L = (int)sat_rn_u8(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
  const int yy = T.lab_y[L], ify = T.lab_ify[L];
  const int adiv = ((5 * a * 53687 + 128) >> 13) - 4194;
  const int bdiv = ((bb * 41943 + 16) >> 9) - 10485 + 1;
  const int X = ab2xz(ify + adiv), Z = ab2xz(ify - bdiv);
  cr = T.invgam[clampi(ds(12615 * X - 6296 * yy - 2223 * Z, 14), 0, 4095)];
  cg = T.invgam[clampi(ds(-3773 * X + 7684 * yy + 185 * Z, 14), 0, 4095)];
  cb = T.invgam[clampi(ds(217 * X - 836 * yy + 4715 * Z, 14), 0, 4095)];
}

// RGB -> HSV -> green boost -> HSV -> RGB (A.6), branch-free; returns r | g << 8 | b << 16.  TAIL: cv2's scalar row tail
// (rounds) instead of its SIMD body (truncates).
template <bool TAIL>
__device__ __forceinline__ uint32_t vegetation_px(int r, int g, int b, const PostK& k, const SmemTabs& T) {
  const int v = max(max(r, g), b), mn = min(min(r, g), b);
  const int diff = v - mn;
  int s = (int)((diff * T.sdiv[v] + 2048u) >> 12);
  // hue numerator without a (divergent) branch: minuend, subtrahend and sector offset selected separately
  const bool mr = v == r, mg = v == g;
  const int hx = mr ? g : (mg ? b : r), hy = mr ? b : (mg ? r : g), hk = mr ? 0 : (mg ? 2 : 4);
  int h = hk * diff + hx - hy;
  h = (h * (int)T.hdiv[diff] + 2048) >> 12;
  if (h < 0) h += 180;
  const int sb = (int)sat_rz_u8(__fmul_rn((float)s, k.sat));  // (int)min(s * sat, 255)
  s = (h > k.hue_lo && h < k.hue_hi) ? sb : s;
  const float sf = __fmul_rn((float)s, k.inv255), vf = __fmul_rn((float)v, k.inv255);
  const float h6 = __fmul_rn((float)h, k.hscale);
  int sec = __float2int_rd(h6);  // h6 >= 0
  const float f = __fsub_rn(h6, (float)sec);
  if (sec >= 6) sec -= 6;
  // sector -> (b, g, r): 0 (t1, t3, t0)  1 (t1, t0, t2)  2 (t3, t0, t1)  3 (t0, t2, t1)  4 (t0, t1, t3)  5 (t2, t1, t0)
  // with t0 = v, t1 = v (1 - s), t2 = v (1 - s f), t3 = v (1 - s (1 - f)): odd sectors use t2, even sectors t3 -> one of them
  const float fo = (sec & 1) ? f : __fsub_rn(1.0f, f);
  const float t1 = __fmul_rn(vf, __fsub_rn(1.0f, sf));
  const float tv = __fmul_rn(vf, __fmaf_rn(-sf, fo, 1.0f));
  const float rq = (sec == 0 || sec == 5) ? vf : ((sec == 2 || sec == 3) ? t1 : tv);
  const float gq = (sec == 1 || sec == 2) ? vf : ((sec == 4 || sec == 5) ? t1 : tv);
  const float bq = (sec == 3 || sec == 4) ? vf : ((sec == 0 || sec == 1) ? t1 : tv);
  if constexpr (TAIL)
    return sat_rn_u8(__fmul_rn(rq, 255.0f)) | (sat_rn_u8(__fmul_rn(gq, 255.0f)) << 8) | (sat_rn_u8(__fmul_rn(bq, 255.0f)) << 16);
  else
    return sat_rz_u8(__fmul_rn(rq, 255.0f)) | (sat_rz_u8(__fmul_rn(gq, 255.0f)) << 8) | (sat_rz_u8(__fmul_rn(bq, 255.0f)) << 16);
}

// shared memory of a CTA: tables + LUTs once, then per group of nt threads: row descriptors, doubled taps + mbarriers,
// ring of enhanced rows and four exchange rows (24 B per thread and row)
__host__ __device__ inline size_t march_group_bytes(int rad, int nt) {
  size_t s = (size_t)(PM_MAXSEG + 2 * PB_MAXR) * 16 + PM_MISC_BYTES;
  if (rad > 0) s += (size_t)(2 * rad + 1 + 4) * nt * 24;
  return s;
}
__host__ __device__ inline size_t march_smem_bytes(int rad, int nt, int groups, int lut_bytes) {
  return ((sizeof(SmemTabs) + 15) & ~(size_t)15) + (size_t)((lut_bytes + 15) & ~15) + groups * march_group_bytes(rad, nt);
}

template <int RAD, bool LUTS>
__global__ void __launch_bounds__(PM_MAXCTA, 1)
post_march_kernel(ImgView img, OutView out, const WowsrTables* __restrict__ tabs, const uint8_t* __restrict__ luts,
                  const __grid_constant__ PostK k, const __grid_constant__ MarchK m, int row0, int row1) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int N = 2 * RAD + 1;
  constexpr int HL = RAD == 0 ? 0 : (RAD <= 4 ? 4 : 8);
  constexpr int HG = HL / 4;          // halo groups (threads) on each side
  constexpr int NPAIR = 2 + HL;       // 16-bit pairs in a thread's horizontal window: (4 + 2 HL) / 2
  SmemTabs& T = *reinterpret_cast<SmemTabs*>(smem_raw);
  uint8_t* sp = smem_raw + ((sizeof(SmemTabs) + 15) & ~(size_t)15);
  const uint8_t* s_lut = sp;
  sp += (m.lut_bytes + 15) & ~15;
  const int nt = m.nt, grp = threadIdx.x / nt, tid = threadIdx.x - grp * nt, lane = tid & 31;
  sp += grp * march_group_bytes(RAD, nt);
  int4* s_rows = reinterpret_cast<int4*>(sp);
  sp += (PM_MAXSEG + 2 * PB_MAXR) * 16;
  int* s_taps2 = reinterpret_cast<int*>(sp);                   // 2 * (2 PB_MAXR + 1) ints, then the group's two mbarriers
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sp + 2 * (2 * PB_MAXR + 1) * 4);
  sp += PM_MISC_BYTES;
  uint4* ringq = reinterpret_cast<uint4*>(sp);                 // [N][nt]  r01 r23 g01 g23
  uint4* xq = ringq + N * nt;                                  // [4][nt]
  uint2* ringd = reinterpret_cast<uint2*>(xq + 4 * nt);        // [N][nt]  b01 b23
  uint2* xd = ringd + N * nt;                                  // [4][nt]
  __shared__ uint64_t s_bar_tab;

  const int item = blockIdx.x * m.groups + grp;
  const bool has_item = item < m.n_items;
  const int strip = item % m.n_strips, segi = item / m.n_strips;
  const int ya = row0 + segi * m.seg, yb = has_item ? min(ya + m.seg, row1) : ya;
  const int nrows = yb - ya + 2 * RAD;
  const bool do_clahe = k.stages & WOWSR_STAGE_CLAHE, do_veg = k.stages & WOWSR_STAGE_VEG;
  // warps whose columns all lie right of the image (+ halo) leave after the prologue
  const int xs = strip * m.swv - HL;
  const int warps_active = has_item ? min(nt >> 5, (img.W + HL - xs + 127) >> 7) : 0;
  const uint32_t bar_tab = ptx::smem_u32(&s_bar_tab), bar_x = ptx::smem_u32(&s_bar[0]);  // bar_x, bar_x + 8: even / odd rows
  if (tid == 0 && has_item) {
    ptx::mbar_init(bar_x, warps_active);
    ptx::mbar_init(bar_x + 8, warps_active);
  }
  if (threadIdx.x == 0) {
    // tables (and the CLAHE LUTs) arrive by bulk copies while the threads set up their constants
    ptx::mbar_init(bar_tab, 1);
    ptx::fence_barrier_init();
    const uint32_t lb = (LUTS && do_clahe) ? (uint32_t)m.lut_bytes : 0u;
    ptx::mbar_arrive_expect_tx(bar_tab, (uint32_t)sizeof(SmemTabs) + lb);
    ptx::bulk_load(ptx::smem_u32(&T), tabs, (uint32_t)sizeof(SmemTabs), bar_tab);
    if (lb) ptx::bulk_load(ptx::smem_u32(s_lut), luts, lb, bar_tab);
  }
  for (int i = tid; i < 2 * N; i += nt) s_taps2[i] = m.taps2[i];
  for (int i = tid; i < nrows; i += nt) {
    const int gy = reflect101(ya - RAD + i, img.H);
    const int gyb = gy - img.y0;
    const float tyf = __fsub_rn(__fmul_rn((float)gy, k.inv_th), 0.5f);
    const int t1 = (int)floorf(tyf);
    const float yaf = __fsub_rn(tyf, (float)t1);
    s_rows[i] = make_int4((gyb >= 0 && gyb < img.rows) ? gyb : -1, (max(t1, 0) * k.grid) << 8,
                          (min(t1 + 1, k.grid - 1) * k.grid) << 8, __float_as_int(yaf));
  }
  // ---- per-thread column constants ----
  const int x0 = xs + 4 * tid;
  const bool inner = tid >= HG && tid < nt - HG;
  const bool vec_in = m.in_vec && x0 >= 0 && x0 + 3 < img.W;
  const bool vec_out = m.out_vec && x0 + 3 < img.W;
  int goff[4], ctx1[4], ctx2[4];
  float cxa[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int gx = reflect101(min(max(x0 + j, -HL), img.W - 1 + HL), img.W);
    goff[j] = gx * 3;
    const float txf = __fsub_rn(__fmul_rn((float)gx, k.inv_tw), 0.5f);
    const int t1 = (int)floorf(txf);
    cxa[j] = __fsub_rn(txf, (float)t1);
    ctx1[j] = max(t1, 0) << 8;
    ctx2[j] = min(t1 + 1, k.grid - 1) << 8;
  }
  const bool any_tail = do_veg && x0 + 3 >= k.tail_x;
  const bool emit_ok = inner && x0 < img.W;
  const uint8_t* lut = LUTS ? s_lut : luts;
  __syncthreads();  // barriers initialised, row descriptors and taps written
  if ((tid >> 5) >= warps_active) return;
  if (!ptx::mbar_wait(bar_tab, 0)) *m.err = 1;

  // one row of this thread's four pixels as three packed words (r g b r | g b r g | b r g b)
  auto load_row = [&](int i, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    w0 = w1 = w2 = 0;
    if (i >= nrows) return;
    const int gyb = s_rows[i].x;
    if (gyb < 0) return;  // outside the band: meets zero taps only
    const uint8_t* rp = img.data + (long long)gyb * img.pitch;
    if (vec_in) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(rp + goff[0]);
      w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    } else {
      uint32_t c[12];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint8_t* p = rp + goff[j];
        c[3 * j] = __ldg(p), c[3 * j + 1] = __ldg(p + 1), c[3 * j + 2] = __ldg(p + 2);
      }
      w0 = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
      w1 = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
      w2 = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
    }
  };
  // stage 1 of one row -> r, g, b of the four pixels
  auto stage1 = [&](int i, uint32_t w0, uint32_t w1, uint32_t w2, int (&cr)[4], int (&cg)[4], int (&cb)[4]) {
    cr[0] = w0 & 255, cr[1] = w0 >> 24, cr[2] = (w1 >> 16) & 255, cr[3] = (w2 >> 8) & 255;
    cg[0] = (w0 >> 8) & 255, cg[1] = w1 & 255, cg[2] = w1 >> 24, cg[3] = (w2 >> 16) & 255;
    cb[0] = (w0 >> 16) & 255, cb[1] = (w1 >> 8) & 255, cb[2] = w2 & 255, cb[3] = w2 >> 24;
    if (do_clahe) {
      const int4 ri = s_rows[i];
      const float yaf = __int_as_float(ri.w), ya1 = __fsub_rn(1.0f, yaf);
#pragma unroll
      for (int j = 0; j < 4; j++) enhance_rgb<LUTS>(cr[j], cg[j], cb[j], T, lut, ri.y, ri.z, ctx1[j], ctx2[j], cxa[j], yaf, ya1);
    }
  };
  // vegetation + store of one finished row
  auto finish_row = [&](int gy, uint32_t (&o)[4]) {
    if (do_veg) {
      if (any_tail) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          o[j] = (x0 + j >= k.tail_x) ? vegetation_px<true>(o[j] & 255, (o[j] >> 8) & 255, o[j] >> 16, k, T)
                                      : vegetation_px<false>(o[j] & 255, (o[j] >> 8) & 255, o[j] >> 16, k, T);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = vegetation_px<false>(o[j] & 255, (o[j] >> 8) & 255, o[j] >> 16, k, T);
      }
    }
    uint8_t* op = out.data + (long long)(gy - out.y0) * out.pitch + (long long)x0 * 3;
    if (vec_out) {
      uint32_t* q = reinterpret_cast<uint32_t*>(op);
      q[0] = o[0] | (o[1] << 24);
      q[1] = (o[1] >> 8) | (o[2] << 16);
      q[2] = (o[2] >> 16) | (o[3] << 8);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (x0 + j < img.W) {
          op[3 * j] = (uint8_t)(o[j] & 255);
          op[3 * j + 1] = (uint8_t)((o[j] >> 8) & 255);
          op[3 * j + 2] = (uint8_t)(o[j] >> 16);
        }
    }
  };

  uint32_t n0, n1, n2;
  load_row(0, n0, n1, n2);
  // The loads of row 0 must be CONSUMED before the loop: otherwise the first use of the row words inside the loop carries
  // the scoreboard wait of these loads, and the in-loop prefetch (same scoreboard) made every iteration wait for the
  // loads it had just issued (17 % of all stall samples in the first capture).
  n0 += m.zero, n1 += m.zero, n2 += m.zero;
  if constexpr (RAD == 0) {
    for (int i = 0; i < nrows; i++) {
      const uint32_t w0 = n0, w1 = n1, w2 = n2;
      load_row(i + 1, n0, n1, n2);  // prefetch
      int cr[4], cg[4], cb[4];
      stage1(i, w0, w1, w2, cr, cg, cb);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) o[j] = (uint32_t)cr[j] | ((uint32_t)cg[j] << 8) | ((uint32_t)cb[j] << 16);
      if (emit_ok) finish_row(ya + i, o);
    }
  } else {
    // Software pipeline over rows; iteration i:
    //   A  stage 1 of row i -> ring                      (thread-private)
    //   B  wait for the exchange of row i - 2, horizontal pass, unsharp, vegetation, store   (reads the neighbours' sums)
    //   C  vertical pass ending at row i -> exchange row, arrive
    // The arrival of C(i - 2) and the wait of B(i) are more than an iteration apart, so warps drift instead of meeting
    // at a barrier.  Four exchange rows: a thread that passed the wait of B(i) knows every warp finished C(i - 2), hence
    // B(i - 2), the last reader of the exchange row of row i - 4 that C(i) overwrites.
    static_assert(RAD >= 2, "the centre row i - 2 - RAD must still be in the ring");
    int slot = 0;       // ring slot of row i
    int cslot = N - RAD - 2;  // ring slot of row i - 2 - RAD, the centre row of the output finished in B(i)
    for (int i = 0; i <= nrows + 1; i++) {
      if (i < nrows) {
        const uint32_t w0 = n0, w1 = n1, w2 = n2;
        load_row(i + 1, n0, n1, n2);  // prefetch
        int cr[4], cg[4], cb[4];
        stage1(i, w0, w1, w2, cr, cg, cb);
        ringq[slot * nt + tid] = make_uint4(cr[0] | (cr[1] << 16), cr[2] | (cr[3] << 16), cg[0] | (cg[1] << 16), cg[2] | (cg[3] << 16));
        ringd[slot * nt + tid] = make_uint2(cb[0] | (cb[1] << 16), cb[2] | (cb[3] << 16));
      }
      if (i - 2 >= 2 * RAD) {
        // exchange e = i - 2 - 2 RAD completes on barrier e & 1 (this thread has already arrived for e + 1: on ONE barrier
        // that phase could complete too and the parity wait for e would never return)
        const uint32_t e = (uint32_t)(i - 2 - 2 * RAD);
        if (!ptx::mbar_wait(bar_x + 8 * (e & 1u), (e >> 1) & 1u)) *m.err = 1;
        if (emit_ok) {
          const int xb = ((i - 2) & 3) * nt;
          uint32_t P[3][NPAIR];
#pragma unroll
          for (int dt = -HG; dt <= HG; dt++) {
            const uint4 q = xq[xb + tid + dt];
            const uint2 d = xd[xb + tid + dt];
            const int kb = 2 * (dt + HG);
            P[0][kb] = q.x, P[0][kb + 1] = q.y, P[1][kb] = q.z, P[1][kb + 1] = q.w, P[2][kb] = d.x, P[2][kb + 1] = d.y;
          }
          const uint4 cq = ringq[cslot * nt + tid];
          const uint2 cd = ringd[cslot * nt + tid];
          const uint32_t cen[3][2] = {{cq.x, cq.y}, {cq.z, cq.w}, {cd.x, cd.y}};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int k_lo = (j - RAD + HL) >> 1, k_hi = (j + RAD + HL) >> 1;
            uint32_t res[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
              uint32_t acc = 32768u;
#pragma unroll
              for (int kk = 0; kk < NPAIR; kk++)
                if (kk >= k_lo && kk <= k_hi) acc = __dp2a_lo(P[c][kk], m.hw[j][kk], acc);
              const float blur = u2f(acc >> 16);
              const uint32_t cw = cen[c][j >> 1];
              const float cv = u2f((j & 1) ? (cw >> 16) : (cw & 0xFFFFu));
              res[c] = sat_rn_u8(__fadd_rn(__fmul_rn(cv, k.alpha), __fmul_rn(blur, k.beta)));
            }
            o[j] = res[0] | (res[1] << 8) | (res[2] << 16);
          }
          finish_row(ya + i - 2 - 2 * RAD, o);
        }
      }
      const int phase = slot + 1 == N ? 0 : slot + 1;  // slot of the oldest row i - 2 RAD
      if (i < nrows && i >= 2 * RAD) {
        const int* wb = s_taps2 + N - phase;
        uint32_t v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int s = 0; s < N; s++) {
          const uint32_t w = wb[s];
          const uint4 q = ringq[s * nt + tid];
          const uint2 d = ringd[s * nt + tid];
          v[0] += w * q.x, v[1] += w * q.y, v[2] += w * q.z, v[3] += w * q.w, v[4] += w * d.x, v[5] += w * d.y;
        }
        const int xb = (i & 3) * nt;
        xq[xb + tid] = make_uint4(v[0], v[1], v[2], v[3]);
        xd[xb + tid] = make_uint2(v[4], v[5]);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_x + 8 * ((uint32_t)(i - 2 * RAD) & 1u));
      }
      slot = phase;
      cslot = cslot + 1 == N ? 0 : cslot + 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// HSV vegetation mask: vector_extraction.compute_green_mask_hsv (server/app/vector_extraction.py:251-270)
// = cv2.cvtColor(RGB2HSV) -> cv2.inRange per colour range -> bitwise_or -> (> 0).astype(float32).
// HBM-bound: 3 B read + 4 B written per pixel; a thread converts 4 pixels (three 32-bit loads, one float4 store).
// The RGB -> HSV arithmetic is the one of vegetation_pixel above (App. A.6).
// ---------------------------------------------------------------------------------------------

constexpr int GM_THREADS = 256, GM_MAX_RANGES = 4;

struct MaskK {
  int n;
  int lo[GM_MAX_RANGES][3], hi[GM_MAX_RANGES][3];  // inclusive bounds on (H, S, V), like cv2.inRange
};

__device__ __forceinline__ float hsv_in_ranges(int r, int g, int b, const MaskK& k, const uint32_t* sdiv, const uint32_t* hdiv) {
  const int v = max(max(r, g), b), mn = min(min(r, g), b);
  const int diff = v - mn;
  const int s = (int)((diff * sdiv[v] + 2048u) >> 12);
  int h;
  if (v == r) h = g - b;
  else if (v == g) h = b - r + 2 * diff;
  else h = r - g + 4 * diff;
  h = (h * (int)hdiv[diff] + 2048) >> 12;
  if (h < 0) h += 180;
  bool in = false;
#pragma unroll
  for (int i = 0; i < GM_MAX_RANGES; i++)  // fully unrolled: the bounds stay in the constant bank (no local copy of k)
    in |= i < k.n && h >= k.lo[i][0] && h <= k.hi[i][0] && s >= k.lo[i][1] && s <= k.hi[i][1] && v >= k.lo[i][2] && v <= k.hi[i][2];
  return in ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(GM_THREADS)
green_mask_kernel(ImgView img, const WowsrTables* __restrict__ tabs, MaskK k, int vec_ok, float* __restrict__ mask,
                  long long mask_pitch /* floats */) {
  __shared__ uint32_t s_sdiv[256], s_hdiv[256];
  for (int i = threadIdx.x; i < 256; i += GM_THREADS) {
    s_sdiv[i] = tabs->sdiv[i];
    s_hdiv[i] = tabs->hdiv[i];
  }
  __syncthreads();
  const int gpr = (img.W + 3) >> 2;  // 4-pixel groups per row
  const long long total = (long long)img.rows * gpr;
  for (long long it = (long long)blockIdx.x * GM_THREADS + threadIdx.x; it < total; it += (long long)gridDim.x * GM_THREADS) {
    const int row = (int)(it / gpr), x0 = (int)(it % gpr) * 4;
    const uint8_t* rp = img.data + (long long)row * img.pitch;
    float* mp = mask + (long long)row * mask_pitch + x0;
    if (vec_ok && x0 + 3 < img.W) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(rp + x0 * 3);
      const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
      float4 o;
      o.x = hsv_in_ranges(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, k, s_sdiv, s_hdiv);
      o.y = hsv_in_ranges(w0 >> 24, w1 & 255, (w1 >> 8) & 255, k, s_sdiv, s_hdiv);
      o.z = hsv_in_ranges((w1 >> 16) & 255, w1 >> 24, w2 & 255, k, s_sdiv, s_hdiv);
      o.w = hsv_in_ranges((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24, k, s_sdiv, s_hdiv);
      *reinterpret_cast<float4*>(mp) = o;
    } else {
      for (int j = 0; j < 4 && x0 + j < img.W; j++) {
        const uint8_t* p = rp + (x0 + j) * 3;
        mp[j] = hsv_in_ranges(__ldg(p), __ldg(p + 1), __ldg(p + 2), k, s_sdiv, s_hdiv);
      }
    }
  }
}

size_t post_smem_bytes(int r) {
  int EW = PB_TX + 2 * r, EH = PB_TY + 2 * r;
  size_t s = (sizeof(SmemTabs) + 15) & ~(size_t)15;
  s += (size_t)(EH * EW + ((EH * EW) & 1)) * 4;
  s += (size_t)EH * PB_TX * 8;
  return s;
}

// Strip width (threads per strip, multiples of 32), strips per CTA and segment height of the strip-march kernel: fewest
// waves of work items times rows per wave, counting idle columns of the last strip, the warm-up rows of a segment and
// SMs left with few warps (the kernel is latency-bound below ~16 warps per SM).
void march_plan(int sm_count, int W, int rows, int rad, int hl, int lut_bytes, int force_nt, int force_seg, int force_groups,
                int* nt_out, int* groups_out, int* seg_out) {
  double best = 1e300;
  *nt_out = 128, *groups_out = 1, *seg_out = std::min(64, rows);
  for (int nt = 64; nt <= PM_MAXNT; nt += 32) {
    if (force_nt && nt != force_nt) continue;
    for (int groups = 1; groups * nt <= PM_MAXCTA; groups++) {
      if (force_groups && groups != force_groups) continue;
      const size_t smem = march_smem_bytes(rad, nt, groups, lut_bytes) + 1024;
      if (smem > (size_t)227 * 1024) continue;
      int cps = (int)((size_t)228 * 1024 / smem);
      cps = std::min(cps, std::min(2048 / (groups * nt), 65536 / (128 * groups * nt)));
      if (cps < 1) continue;
      const int swv = 4 * nt - 2 * hl;
      const int strips = (W + swv - 1) / swv;
      const double resident = (double)sm_count * cps * groups;  // work items in flight
      const double warps = cps * groups * nt / 32.0;
      const double starve = warps >= 16 ? 1.0 : std::pow(16.0 / warps, 0.6);
      auto consider = [&](int n_segs) {
        if (n_segs < 1) n_segs = 1;
        if (n_segs > rows) n_segs = rows;
        const int seg = (rows + n_segs - 1) / n_segs;
        if (seg > PM_MAXSEG) return false;
        // (the warps of the last strip that lie right of the image leave early, but their group still holds its slot)
        const double waves = std::ceil(strips * (double)n_segs / resident);
        const double cost = waves * cps * groups * nt * (seg + 1.2 * rad + 2.0) * starve;
        if (cost < best) best = cost, *nt_out = nt, *groups_out = groups, *seg_out = seg;
        return true;
      };
      if (force_seg) {
        const int seg = std::max(1, std::min(std::min(force_seg, rows), PM_MAXSEG));
        consider((rows + seg - 1) / seg);
        continue;
      }
      // Only segment counts that fill whole waves can be optimal (more segments in the same number of waves only shorten
      // them): try the first few wave counts whose segments fit the row-descriptor array; beyond them the warm-up rows
      // of ever shorter segments only add cost.
      int tried = 0;
      for (int w = 1; tried < 3; w++) {
        const int n_segs = (int)std::min<double>(rows, std::floor(w * resident / strips));
        if (n_segs >= 1 && consider(n_segs)) tried++;
        if (n_segs >= rows) break;
      }
      consider((rows + PM_MAXSEG - 1) / PM_MAXSEG);  // the fewest, longest segments
    }
  }
}

int check_image(wowsr_ctx* ctx, const wowsr_image* im, const char* what) {
  if (!im || !im->data || im->W <= 0 || im->H <= 0 || im->rows <= 0 || im->y0 < 0 || im->y0 + im->rows > im->H ||
      im->pitch < (int64_t)im->W * 3)
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad %s image descriptor", what);
  return 0;
}

// A strip-march launch whose exchange barrier ran out of polls (a protocol bug, never seen) has produced garbage: the kernel
// sets the mapped host word, the next post-process call and the host-buffer entry point (after its final synchronise) fail.
int post_check_err(wowsr_ctx* ctx) {
  if (ctx->post_err && *(volatile int*)ctx->post_err) {
    *ctx->post_err = 0;
    return wowsr_fail(ctx, WOWSR_ERR_CUDA, "a post-process kernel timed out on its exchange barrier; its output is invalid");
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

extern "C" int wowsr_clahe_hist(wowsr_ctx* ctx, const wowsr_image* rgb, int32_t grid, int32_t prow0, int32_t prow1,
                                uint32_t* hist_dev, void* stream) {
  if (!ctx) return WOWSR_ERR_ARG;
  if (int e = check_image(ctx, rgb, "input")) return e;
  if (grid < 1 || grid > 64 || !hist_dev) return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad grid/hist");
  DeviceGuard g(ctx->device);
  int tw, th, pw, ph;
  wowsr_clahe_geometry(rgb->H, rgb->W, grid, &tw, &th, &pw, &ph);
  if (prow0 < 0) prow0 = 0;
  if (prow1 > ph) prow1 = ph;
  if (prow0 >= prow1) return WOWSR_OK;
  // the band must hold every source row the padded range maps to
  for (int pr : {prow0, prow1 - 1}) {
    int sy = reflect101(pr, rgb->H);
    if (sy < rgb->y0 || sy >= rgb->y0 + rgb->rows)
      return wowsr_fail(ctx, WOWSR_ERR_ARG, "band [%d,%d) does not hold source row %d", rgb->y0, rgb->y0 + rgb->rows, sy);
  }
  int target_chunks = (4 * ctx->sm_count + grid * grid - 1) / (grid * grid);
  int rows_per_block = (th + target_chunks - 1) / target_chunks;
  if (rows_per_block < 1) rows_per_block = 1;
  int chunks = (th + rows_per_block - 1) / rows_per_block;
  ImgView v{(const uint8_t*)rgb->data, rgb->pitch, rgb->W, rgb->H, rgb->y0, rgb->rows};
  int vec_ok = (rgb->pitch % 4 == 0) && (((uintptr_t)rgb->data) % 4 == 0) && (tw % 4 == 0);
  dim3 gr(grid, grid * chunks);
  clahe_hist_kernel<<<gr, HIST_THREADS, 0, (cudaStream_t)stream>>>(v, ctx->d_tables, grid, tw, th, prow0, prow1,
                                                                    rows_per_block, chunks, vec_ok, (int)wowsr_opt(ctx, "hist_match", 2), hist_dev);
  WLAUNCH_CHECK(ctx);
  return WOWSR_OK;
}

extern "C" int wowsr_clahe_luts(wowsr_ctx* ctx, const uint32_t* hist_dev, int32_t grid, int32_t tile_w, int32_t tile_h,
                                double clip_limit, uint8_t* luts_dev, void* stream) {
  if (!ctx || !hist_dev || !luts_dev || grid < 1 || tile_w < 1 || tile_h < 1) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  int area = tile_w * tile_h;
  // cv2: clipLimit = max(int(clip * tileSizeTotal / histSize), 1), evaluated in double
  int clip = 0x7FFFFFFF;
  if (clip_limit > 0.0) {
    clip = (int)(clip_limit * area / 256.0);
    if (clip < 1) clip = 1;
  }
  float lut_scale = 255.0f / (float)area;
  clahe_lut_kernel<<<grid * grid, 256, 0, (cudaStream_t)stream>>>(hist_dev, clip, lut_scale, luts_dev);
  WLAUNCH_CHECK(ctx);
  return WOWSR_OK;
}

extern "C" int wowsr_post_apply(wowsr_ctx* ctx, const wowsr_image* rgb, const uint8_t* luts_dev,
                                const wowsr_post_params* p, int32_t row0, int32_t row1, const wowsr_image* out,
                                void* stream) {
  if (!ctx || !p) return WOWSR_ERR_ARG;
  if (int e = post_check_err(ctx)) return e;
  if (int e = check_image(ctx, rgb, "input")) return e;
  if (int e = check_image(ctx, out, "output")) return e;
  if (out->W != rgb->W || out->H != rgb->H) return wowsr_fail(ctx, WOWSR_ERR_ARG, "input/output size mismatch");
  if ((p->stages & WOWSR_STAGE_CLAHE) && !luts_dev) return wowsr_fail(ctx, WOWSR_ERR_ARG, "CLAHE stage needs luts");
  DeviceGuard g(ctx->device);
  PostK k;
  memset(&k, 0, sizeof k);
  k.stages = p->stages;
  k.grid = p->grid;
  int pw, ph;
  wowsr_clahe_geometry(rgb->H, rgb->W, p->grid, &k.tw, &k.th, &pw, &ph);
  k.inv_tw = 1.0f / (float)k.tw;
  k.inv_th = 1.0f / (float)k.th;
  k.r = 0;
  if (p->stages & WOWSR_STAGE_UNSHARP) {
    int taps[32];
    int ks = wowsr_gaussian_taps(p->sigma, taps, 32);
    if (ks < 0) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "sigma %f gives a kernel wider than 31", p->sigma);
    int half = ks / 2, r = half;
    while (r > 0 && taps[half - r] == 0) r--;
    if (r > PB_MAXR) return wowsr_fail(ctx, WOWSR_ERR_UNSUPPORTED, "blur radius %d > %d", r, PB_MAXR);
    k.r = r;
    for (int t = 0; t <= 2 * r; t++) k.taps[t] = taps[half - r + t];
  }
  k.alpha = p->alpha;
  k.beta = p->beta;
  k.hue_lo = p->hue_lo;
  k.hue_hi = p->hue_hi;
  k.sat = p->sat_boost;
  k.inv255 = (float)(1.0 / 255.0);
  k.hscale = (float)(6.0 / 180.0);
  // cv2's HSV2RGB_b converts rows in SIMD groups that truncate and finishes each row with a scalar tail that rounds; the
  // group width is a property of the cv2 BUILD: 32 pixels on the AVX2 builds the goldens come from (opencv-python wheels),
  // 16 on SSE / NEON builds, 64 on AVX-512 builds.  Option hsv_simd_width selects it (0: no scalar tail).
  const int simd_w = (int)wowsr_opt(ctx, "hsv_simd_width", 32);
  k.tail_x = simd_w > 0 ? rgb->W - rgb->W % simd_w : rgb->W;
  if (row0 < 0) row0 = 0;
  if (row1 > rgb->H) row1 = rgb->H;
  if (row0 >= row1) return WOWSR_OK;
  int need0 = row0 - k.r < 0 ? 0 : row0 - k.r, need1 = row1 + k.r > rgb->H ? rgb->H : row1 + k.r;
  if (need0 < rgb->y0 || need1 > rgb->y0 + rgb->rows)
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "input band [%d,%d) lacks halo rows [%d,%d)", rgb->y0, rgb->y0 + rgb->rows, need0, need1);
  if (row0 < out->y0 || row1 > out->y0 + out->rows) return wowsr_fail(ctx, WOWSR_ERR_ARG, "output band too small");
  ImgView iv{(const uint8_t*)rgb->data, rgb->pitch, rgb->W, rgb->H, rgb->y0, rgb->rows};
  OutView ov{(uint8_t*)out->data, out->pitch, out->W, out->H, out->y0, out->rows};
  // the kernels are specialised on the halo radius; taps beyond the true radius are zero, so rounding r up is exact
  const int true_r = k.r;
  const int rr = true_r == 0 ? 0 : (true_r <= 3 ? 3 : (true_r <= 4 ? 4 : PB_MAXR));
  if (rr != true_r) {  // re-centre the taps in the wider window
    int tmp[2 * PB_MAXR + 1] = {0};
    for (int t = 0; t <= 2 * true_r; t++) tmp[t + (rr - true_r)] = k.taps[t];
    for (int t = 0; t <= 2 * rr; t++) k.taps[t] = tmp[t];
    k.r = rr;  // rows the wider halo reads outside the band are zero-filled in the kernel and meet zero taps
  }
  static_assert(sizeof(SmemTabs) == sizeof(WowsrTables), "table layouts must match");
  if (wowsr_opt(ctx, "post_kernel", 1) == 0) {  // round-1 tile kernel (kept as the A/B yardstick)
    int tiles_x = (rgb->W + PB_TX - 1) / PB_TX, tiles_y = (row1 - row0 + PB_TY - 1) / PB_TY;
    int n_tiles = tiles_x * tiles_y;
    size_t smem = post_smem_bytes(rr);
    int per_sm = (int)(200 * 1024 / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / PB_THREADS) per_sm = 2048 / PB_THREADS;
    int blocks = ctx->sm_count * per_sm;
    if (blocks > n_tiles) blocks = n_tiles;
#define WOWSR_POST_LAUNCH(RR)                                                                                          \
  {                                                                                                                    \
    WCUDA(ctx, cudaFuncSetAttribute(post_apply_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    post_apply_kernel<RR><<<blocks, PB_THREADS, smem, (cudaStream_t)stream>>>(iv, ov, ctx->d_tables, luts_dev, k, row0, \
                                                                              row1, tiles_x, n_tiles);                 \
  }
    if (rr == 0) WOWSR_POST_LAUNCH(0) else if (rr == 3) WOWSR_POST_LAUNCH(3) else if (rr == 4) WOWSR_POST_LAUNCH(4) else WOWSR_POST_LAUNCH(PB_MAXR)
#undef WOWSR_POST_LAUNCH
    WLAUNCH_CHECK(ctx);
    return WOWSR_OK;
  }
  // ---- strip-march kernel: pick the strip width (threads per CTA) and the segment height ----
  MarchK m;
  memset(&m, 0, sizeof m);
  m.hl = rr == 0 ? 0 : (rr <= 4 ? 4 : 8);
  m.err = ctx->post_err;  // unified addressing: the mapped host pointer is valid on the device
  m.in_vec = rgb->pitch % 4 == 0 && ((uintptr_t)rgb->data) % 4 == 0;
  m.out_vec = out->pitch % 4 == 0 && ((uintptr_t)out->data) % 4 == 0;
  const bool lut_smem = (p->stages & WOWSR_STAGE_CLAHE) && p->grid <= 8 && ((uintptr_t)luts_dev) % 16 == 0;  // bulk-copy alignment
  m.lut_bytes = lut_smem ? p->grid * p->grid * 256 : 0;
  const int n = 2 * rr + 1;
  for (int u = 0; u < 2 * n; u++) m.taps2[u] = k.taps[u % n];
  for (int j = 0; j < 4; j++)
    for (int kk = 0; kk < 2 + m.hl; kk++) {
      uint32_t w = 0;
      for (int half = 0; half < 2; half++) {
        const int t = 2 * kk - m.hl + half - (j - rr);  // tap that meets this column for pixel j of the group
        if (t >= 0 && t <= 2 * rr) w |= (uint32_t)k.taps[t] << (8 * half);
      }
      m.hw[j][kk] = w;
    }
  march_plan(ctx->sm_count, rgb->W, row1 - row0, rr, m.hl, m.lut_bytes, (int)wowsr_opt(ctx, "post_nt", 0),
             (int)wowsr_opt(ctx, "post_seg", 0), (int)wowsr_opt(ctx, "post_groups", 0), &m.nt, &m.groups, &m.seg);
  m.swv = 4 * m.nt - 2 * m.hl;
  m.n_strips = (rgb->W + m.swv - 1) / m.swv;
  const int n_segs = (row1 - row0 + m.seg - 1) / m.seg;
  m.n_items = m.n_strips * n_segs;
  const size_t smem = march_smem_bytes(rr, m.nt, m.groups, m.lut_bytes);
  const unsigned blocks = (unsigned)((m.n_items + m.groups - 1) / m.groups);
#define WOWSR_MARCH_LAUNCH(RR, LS)                                                                                      \
  {                                                                                                                    \
    WCUDA(ctx, cudaFuncSetAttribute(post_march_kernel<RR, LS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    post_march_kernel<RR, LS><<<blocks, m.groups * m.nt, smem, (cudaStream_t)stream>>>(iv, ov, ctx->d_tables, luts_dev, k, m, row0, row1); \
  }
#define WOWSR_MARCH_RR(LS)                                                                                              \
  if (rr == 0) WOWSR_MARCH_LAUNCH(0, LS) else if (rr == 3) WOWSR_MARCH_LAUNCH(3, LS) else if (rr == 4) WOWSR_MARCH_LAUNCH(4, LS) else WOWSR_MARCH_LAUNCH(PB_MAXR, LS)
  if (lut_smem) { WOWSR_MARCH_RR(true) } else { WOWSR_MARCH_RR(false) }
#undef WOWSR_MARCH_RR
#undef WOWSR_MARCH_LAUNCH
  WLAUNCH_CHECK(ctx);
  return WOWSR_OK;
}

extern "C" int wowsr_post_process_dev(wowsr_ctx* ctx, const wowsr_image* rgb, const wowsr_post_params* p,
                                      const wowsr_image* out, void* stream) {
  if (!ctx || !p) return WOWSR_ERR_ARG;
  if (int e = check_image(ctx, rgb, "input")) return e;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* luts = nullptr;
  if (p->stages & WOWSR_STAGE_CLAHE) {
    int n = p->grid * p->grid * 256;
    if (int e = wowsr_ensure(ctx, ctx->hist, (size_t)n * 4)) return e;
    if (int e = wowsr_ensure(ctx, ctx->luts, (size_t)n)) return e;
    WCUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)n * 4, st));
    int tw, th, pw, ph;
    wowsr_clahe_geometry(rgb->H, rgb->W, p->grid, &tw, &th, &pw, &ph);
    if (int e = wowsr_clahe_hist(ctx, rgb, p->grid, 0, ph, (uint32_t*)ctx->hist.p, stream)) return e;
    if (int e = wowsr_clahe_luts(ctx, (const uint32_t*)ctx->hist.p, p->grid, tw, th, p->clip_limit, (uint8_t*)ctx->luts.p, stream))
      return e;
    luts = (const uint8_t*)ctx->luts.p;
  }
  return wowsr_post_apply(ctx, rgb, luts, p, 0, rgb->H, out, stream);
}

// Drop-in for _enhance_for_crops(img) (wow_sr.py:187-209) with pageable host buffers, pipelined through the pinned ring
// (hoststage.h): the histogram pass runs on each chunk as it arrives, the apply pass runs per row block and each finished
// block leaves while the next one is computed.
extern "C" int wowsr_post_process_host(wowsr_ctx* ctx, const uint8_t* rgb_host, int32_t H, int32_t W,
                                       const wowsr_post_params* p, uint8_t* out_host) {
  if (!ctx || !rgb_host || !out_host || !p || H <= 0 || W <= 0) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  const size_t row = (size_t)W * 3;
  size_t pitch = (row + 15) & ~(size_t)15;
  if (int e = wowsr_ensure(ctx, ctx->post_in, pitch * H)) return e;
  if (int e = wowsr_ensure(ctx, ctx->post_out, pitch * H)) return e;
  if (int e = stage_init(ctx)) return e;
  wowsr_image in{ctx->post_in.p, (int64_t)pitch, W, H, 0, H};
  wowsr_image out{ctx->post_out.p, (int64_t)pitch, W, H, 0, H};
  const bool clahe = (p->stages & WOWSR_STAGE_CLAHE) != 0;
  int tw = 0, th = 0, pw = 0, ph = 0;
  if (clahe) {
    if (p->grid < 1 || p->grid > 64) return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad grid");
    const int n = p->grid * p->grid * 256;
    if (int e = wowsr_ensure(ctx, ctx->hist, (size_t)n * 4)) return e;
    if (int e = wowsr_ensure(ctx, ctx->luts, (size_t)n)) return e;
    WCUDA(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)n * 4, 0));
    wowsr_clahe_geometry(H, W, p->grid, &tw, &th, &pw, &ph);
  }
  // upload in chunks; pass A on the rows of each chunk as soon as they are on the device (the padded rows below the image
  // are reflections of rows near the bottom: they go with the last chunk)
  if (int e = stage_in(ctx, (uint8_t*)ctx->post_in.p, pitch, rgb_host, row, row, H, [&](int r0, int r1, cudaEvent_t ev) -> int {
        if (cudaStreamWaitEvent(0, ev, 0) != cudaSuccess) return wowsr_fail(ctx, WOWSR_ERR_CUDA, "cudaStreamWaitEvent");
        if (!clahe) return 0;
        return wowsr_clahe_hist(ctx, &in, p->grid, r0, r1 == H ? ph : r1, (uint32_t*)ctx->hist.p, nullptr);
      }))
    return e;
  const uint8_t* luts = nullptr;
  if (clahe) {
    if (int e = wowsr_clahe_luts(ctx, (const uint32_t*)ctx->hist.p, p->grid, tw, th, p->clip_limit, (uint8_t*)ctx->luts.p, nullptr)) return e;
    luts = (const uint8_t*)ctx->luts.p;
  }
  StageOut sink(ctx, out_host, row, row);
  int per = (int)std::max<size_t>(1, STAGE_BYTES / row);
  if (per >= 64) per -= per % 64;  // whole rows of 64 x 64 apply tiles
  for (int r0 = 0; r0 < H; r0 += per) {
    const int r1 = std::min(H, r0 + per);
    if (int e = wowsr_post_apply(ctx, &in, luts, p, r0, r1, &out, nullptr)) {
      sink.finish();
      return e;
    }
    WCUDA(ctx, cudaEventRecord(ctx->stage_sync, 0));
    sink.enqueue((const uint8_t*)ctx->post_out.p, pitch, r0, r1, ctx->stage_sync);
  }
  if (int e = sink.finish()) return e;
  WCUDA(ctx, cudaStreamSynchronize(0));
  return post_check_err(ctx);
}

extern "C" int wowsr_green_mask(wowsr_ctx* ctx, const wowsr_image* rgb, const wowsr_hsv_range* ranges, int32_t n_ranges,
                                float* mask_dev, int64_t mask_pitch, void* stream) {
  if (!ctx || !ranges || !mask_dev) return WOWSR_ERR_ARG;
  if (int e = check_image(ctx, rgb, "input")) return e;
  if (n_ranges < 1 || n_ranges > GM_MAX_RANGES) return wowsr_fail(ctx, WOWSR_ERR_ARG, "1..%d colour ranges", GM_MAX_RANGES);
  if (mask_pitch < (int64_t)rgb->W * 4 || mask_pitch % 4) return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad mask pitch");
  DeviceGuard g(ctx->device);
  MaskK k;
  k.n = n_ranges;
  for (int i = 0; i < n_ranges; i++)
    for (int c = 0; c < 3; c++) {
      k.lo[i][c] = ranges[i].lo[c];
      k.hi[i][c] = ranges[i].hi[c];
    }
  ImgView v{(const uint8_t*)rgb->data, rgb->pitch, rgb->W, rgb->H, rgb->y0, rgb->rows};
  const int vec_ok = rgb->pitch % 4 == 0 && ((uintptr_t)rgb->data) % 4 == 0 && mask_pitch % 16 == 0 && ((uintptr_t)mask_dev) % 16 == 0;
  const long long groups = (long long)rgb->rows * ((rgb->W + 3) / 4);
  long long blocks = (groups + GM_THREADS - 1) / GM_THREADS;
  const long long cap = (long long)ctx->sm_count * (2048 / GM_THREADS);  // one resident wave, grid-stride beyond it
  if (blocks > cap) blocks = cap;
  green_mask_kernel<<<(unsigned)blocks, GM_THREADS, 0, (cudaStream_t)stream>>>(v, ctx->d_tables, k, vec_ok, mask_dev, mask_pitch / 4);
  WLAUNCH_CHECK(ctx);
  return WOWSR_OK;
}

extern "C" int wowsr_green_mask_host(wowsr_ctx* ctx, const uint8_t* rgb_host, int32_t H, int32_t W, const wowsr_hsv_range* ranges,
                                     int32_t n_ranges, float* mask_host) {
  if (!ctx || !rgb_host || !mask_host || !ranges || H <= 0 || W <= 0) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  const size_t pitch = ((size_t)W * 3 + 15) & ~(size_t)15, mpitch = ((size_t)W * 4 + 15) & ~(size_t)15;
  if (int e = wowsr_ensure(ctx, ctx->post_in, pitch * H)) return e;
  if (int e = wowsr_ensure(ctx, ctx->post_out, mpitch * H)) return e;
  WCUDA(ctx, cudaMemcpy2DAsync(ctx->post_in.p, pitch, rgb_host, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyHostToDevice, 0));
  wowsr_image in{ctx->post_in.p, (int64_t)pitch, W, H, 0, H};
  if (int e = wowsr_green_mask(ctx, &in, ranges, n_ranges, (float*)ctx->post_out.p, (int64_t)mpitch, nullptr)) return e;
  WCUDA(ctx, cudaMemcpy2DAsync(mask_host, (size_t)W * 4, ctx->post_out.p, mpitch, (size_t)W * 4, H, cudaMemcpyDeviceToHost, 0));
  WCUDA(ctx, cudaStreamSynchronize(0));
  return WOWSR_OK;
}
