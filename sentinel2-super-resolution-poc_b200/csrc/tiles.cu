// XYZ tile pyramid from a device-resident RGB image: the resampling step of what the reference does by shelling
// out to gdal2tiles.py --xyz --resampling average --tilesize 256 (server/app/tiling.py:147-186) after the SR image has been
// written to disk.  Here the finished image never leaves the GPU before it is cut: one launch resamples a strip of a zoom
// level's tile mosaic (RGBA, alpha = 0 outside the raster) straight from the source image.
//
// PARITY UNPINNED: gdal2tiles / GDAL are not available (neither here nor in the reference's tests), so the arithmetic below is a
// documented restatement of "average" resampling — the area-weighted mean of the source pixels a mosaic pixel covers, rounded
// half up; nearest source pixel when the mosaic is finer than the source; a mosaic pixel is opaque when its centre lies inside
// the raster — checked against a numpy restatement of the same definition (oracle/tiling_np.py), not against GDAL.  Every zoom
// level is resampled from the SOURCE (gdal2tiles averages overview tiles from their four children, which rounds once per
// level).
#include "common.h"

namespace {

struct TileK {
  const uint8_t* src;
  long long pitch;
  int W, H, y0, rows;      // source image / stored band
  double sx0, sy0;         // source coordinates of the mosaic strip's top-left corner
  double sxp, syp;         // source pixels per mosaic pixel
  uint8_t* out;
  long long out_pitch;
  int OW, OH;
};

__global__ void __launch_bounds__(256) tiles_resample_kernel(TileK k) {
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (X >= k.OW || Y >= k.OH) return;
  const double ax = k.sx0 + X * k.sxp, bx = ax + k.sxp;
  const double ay = k.sy0 + Y * k.syp, by = ay + k.syp;
  const double cx = 0.5 * (ax + bx), cy = 0.5 * (ay + by);
  const bool inside = cx >= 0.0 && cx < (double)k.W && cy >= 0.0 && cy < (double)k.H;
  uint32_t rgba = 0;
  if (inside) {
    double r = 0, g = 0, b = 0, wsum = 0;
    if (k.sxp <= 1.0 && k.syp <= 1.0) {  // mosaic finer than the source: nearest source pixel
      const int ix = min(max((int)floor(cx), 0), k.W - 1), iy = min(max((int)floor(cy), 0), k.H - 1);
      const uint8_t* p = k.src + (long long)(iy - k.y0) * k.pitch + (long long)ix * 3;
      r = p[0]; g = p[1]; b = p[2];
      wsum = 1.0;
    } else {
      const int x0 = max((int)floor(ax), 0), x1 = min((int)ceil(bx), k.W);
      const int y0 = max((int)floor(ay), 0), y1 = min((int)ceil(by), k.H);
      for (int y = y0; y < y1; y++) {
        const double wy = fmin((double)(y + 1), by) - fmax((double)y, ay);
        if (wy <= 0.0) continue;
        const uint8_t* row = k.src + (long long)(y - k.y0) * k.pitch;
        for (int x = x0; x < x1; x++) {
          const double wx = fmin((double)(x + 1), bx) - fmax((double)x, ax);
          if (wx <= 0.0) continue;
          const double w = wx * wy;
          const uint8_t* p = row + (long long)x * 3;
          r += w * p[0]; g += w * p[1]; b += w * p[2];
          wsum += w;
        }
      }
    }
    if (wsum > 0.0) {
      const int ri = min(255, (int)floor(r / wsum + 0.5)), gi = min(255, (int)floor(g / wsum + 0.5)), bi = min(255, (int)floor(b / wsum + 0.5));
      rgba = (uint32_t)ri | ((uint32_t)gi << 8) | ((uint32_t)bi << 16) | 0xFF000000u;
    }
  }
  *reinterpret_cast<uint32_t*>(k.out + (long long)Y * k.out_pitch + (long long)X * 4) = rgba;
}

}  // namespace

extern "C" int wowsr_tiles_resample(wowsr_ctx* ctx, const wowsr_image* rgb, double sx0, double sy0, double sxp, double syp,
                                    uint8_t* out_rgba_dev, int64_t out_pitch, int32_t OW, int32_t OH, void* stream) {
  if (!ctx || !rgb || !rgb->data || !out_rgba_dev) return WOWSR_ERR_ARG;
  if (rgb->W < 1 || rgb->H < 1 || rgb->rows < 1 || rgb->y0 < 0 || rgb->y0 + rgb->rows > rgb->H || rgb->pitch < (int64_t)rgb->W * 3)
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad source image descriptor");
  if (OW < 1 || OH < 1 || out_pitch < (int64_t)OW * 4 || (out_pitch & 3) || ((uintptr_t)out_rgba_dev & 3) || !(sxp > 0.0) || !(syp > 0.0))
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "bad mosaic geometry");
  // the strip must only read stored rows
  const double need0 = sy0, need1 = sy0 + OH * syp;
  const int r0 = need0 < 0.0 ? 0 : (int)need0, r1 = need1 > rgb->H ? rgb->H : (int)ceil(need1);
  if (r1 > r0 && (r0 < rgb->y0 || r1 > rgb->y0 + rgb->rows))
    return wowsr_fail(ctx, WOWSR_ERR_ARG, "band [%d,%d) lacks source rows [%d,%d)", rgb->y0, rgb->y0 + rgb->rows, r0, r1);
  DeviceGuard g(ctx->device);
  TileK k{(const uint8_t*)rgb->data, rgb->pitch, rgb->W, rgb->H, rgb->y0, rgb->rows, sx0, sy0, sxp, syp, out_rgba_dev, out_pitch, OW, OH};
  dim3 grid((OW + 31) / 32, (OH + 7) / 8);
  tiles_resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k);
  WLAUNCH_CHECK(ctx);
  return WOWSR_OK;
}
