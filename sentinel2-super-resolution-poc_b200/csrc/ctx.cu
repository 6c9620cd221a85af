// Handle management, integer tables, window planner (host-side pieces of libwowsr.so).
#include <math.h>
#include <stdarg.h>

#include <mutex>

#include "common.h"
#include "hoststage.h"

static thread_local std::string g_create_err;  // last wowsr_create() failure of the calling thread (no ctx to hold it)

int wowsr_fail(wowsr_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_err = buf;
  return code;
}

int wowsr_ensure(wowsr_ctx* ctx, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return wowsr_fail(ctx, WOWSR_ERR_NOMEM, "cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
  }
  b.bytes = bytes;
  return 0;
}

int64_t wowsr_opt(const wowsr_ctx* ctx, const char* key, int64_t dflt) {
  auto it = ctx->opts.find(key);
  return it == ctx->opts.end() ? dflt : it->second;
}

// ---------------------------------------------------------------------------------------------
// tables — built in float32 like cv2 does at start-up
// ---------------------------------------------------------------------------------------------

static WowsrTables build_tables() {
  WowsrTables t;
  for (int i = 0; i < 256; i++) {
    float x = (float)i / 255.0f;
    float lin = x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
    t.gam[i] = (uint16_t)lrintf(lin * 2040.0f);
  }
  // cv2 fills this table with its own (not correctly rounded) float cube root, so it ships as data
  static const uint16_t kCbrt[3072] = {
#include "cbrt_table.inc"
  };
  memcpy(t.cbrt, kCbrt, sizeof kCbrt);
  const double BASE = 16384.0;
  for (int l = 0; l < 256; l++) {
    if (l <= 20) {
      t.lab_y[l] = (uint16_t)lrintf((float)(l * BASE * 180.0 / (17.0 * 29.0 * 29.0 * 29.0)));
      t.lab_ify[l] = (uint16_t)lrintf((float)(BASE * (16.0 / 116.0 + 5.0 * l / (3.0 * 17.0 * 29.0))));
    } else {
      float fy = (float)(l * 100.0 * BASE / (255.0 * 116.0) + 16.0 * BASE / 116.0);
      t.lab_ify[l] = (uint16_t)lrintf(fy);
      t.lab_y[l] = (uint16_t)lrintf(fy * fy * fy / (float)(BASE * BASE));
    }
  }
  for (int k = 0; k < 4096; k++) {
    float x = (float)k / 4096.0f;
    float g = x <= 0.0031308f ? x * 12.92f : 1.055f * powf(x, (float)(1.0 / 2.4)) - 0.055f;
    long v = lrintf(255.0f * g);
    t.invgam[k] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
  }
  t.sdiv[0] = t.hdiv[0] = 0;
  for (int i = 1; i < 256; i++) {
    t.sdiv[i] = (uint32_t)lrint((double)(255 << 12) / (double)i);
    t.hdiv[i] = (uint32_t)lrint((double)(180 << 12) / (6.0 * i));
  }
  return t;
}

const WowsrTables& wowsr_host_tables() {
  static WowsrTables t = build_tables();
  return t;
}

extern "C" int64_t wowsr_get_table(int32_t id, void* out, int64_t cap) {
  const WowsrTables& t = wowsr_host_tables();
  const void* src = nullptr;
  int64_t n = 0;
  switch (id) {
    case 0: src = t.gam; n = sizeof t.gam; break;
    case 1: src = t.cbrt; n = sizeof t.cbrt; break;
    case 2: src = t.lab_y; n = sizeof t.lab_y; break;
    case 3: src = t.lab_ify; n = sizeof t.lab_ify; break;
    case 4: src = t.invgam; n = sizeof t.invgam; break;
    case 5: src = t.sdiv; n = sizeof t.sdiv; break;
    case 6: src = t.hdiv; n = sizeof t.hdiv; break;
    default: return WOWSR_ERR_ARG;
  }
  if (cap < n) return WOWSR_ERR_ARG;
  memcpy(out, src, n);
  return n;
}

extern "C" int32_t wowsr_gaussian_taps(double sigma, int32_t* taps, int32_t cap) {
  // cv2.GaussianBlur ksize=(0,0) on CV_8U: ksize = round(6*sigma+1)|1, 8-bit fixed-point taps by
  // error diffusion from getGaussianKernel (float64) — SURVEY.md App. A.4.
  int ksize = ((int)lrint(sigma * 6.0 + 1.0)) | 1;
  if (ksize > cap || ksize > 31) return WOWSR_ERR_ARG;
  double k[32], sum = 0;
  for (int i = 0; i < ksize; i++) {
    double x = i - (ksize - 1) * 0.5;
    k[i] = exp(-(x * x) / (2.0 * sigma * sigma));
    sum += k[i];
  }
  for (int i = 0; i < ksize; i++) k[i] /= sum;
  double err = 0;
  int half = ksize / 2, acc = 0;
  for (int i = 0; i < half; i++) {
    double adj = k[i] * 256.0 + err;
    double v = nearbyint(adj);
    err = adj - v;
    taps[i] = taps[ksize - 1 - i] = (int)v;
    acc += (int)v;
  }
  taps[half] = 256 - 2 * acc;
  return ksize;
}

extern "C" void wowsr_clahe_geometry(int32_t H, int32_t W, int32_t grid, int32_t* tile_w, int32_t* tile_h,
                                     int32_t* padded_w, int32_t* padded_h) {
  int pw = W, ph = H;
  if (W % grid != 0 || H % grid != 0) {
    pw = W + (grid - W % grid);
    ph = H + (grid - H % grid);
  }
  if (tile_w) *tile_w = pw / grid;
  if (tile_h) *tile_h = ph / grid;
  if (padded_w) *padded_w = pw;
  if (padded_h) *padded_h = ph;
}

extern "C" void wowsr_post_params_wow(wowsr_post_params* p) {
  *p = wowsr_post_params{2.5, 1.2, 1.4f, -0.4f, 1.2f, 8, 35, 85, WOWSR_STAGE_ALL, 0};
}
extern "C" void wowsr_post_params_farm(wowsr_post_params* p) {
  *p = wowsr_post_params{2.5, 1.5, 2.2f, -1.2f, 1.3f, 8, 35, 85, WOWSR_STAGE_ALL, 0};
}

// ---------------------------------------------------------------------------------------------
// window planner (cnn_super_resolution.py:226,244-278)
// ---------------------------------------------------------------------------------------------

namespace {
struct Axis {
  int a, b, lo, hi;
};
std::vector<Axis> plan_axis(int L, int T, int P) {
  int n = (L + T - 1) / T;
  std::vector<Axis> v;
  for (int i = 0; i < n; i++) {
    int a = i * T;
    int b = a + T + 2 * P < L ? a + T + 2 * P : L;
    a = b - T - 2 * P > 0 ? b - T - 2 * P : 0;
    int lo = a + (i > 0 ? P : 0);
    int hi = b - (i < n - 1 ? P : 0);
    v.push_back({a, b, lo, hi});
  }
  // last-writer-wins: a later window's kept interval overrides earlier ones where they overlap.
  // Along one axis window j>i is written after window i in both loop orders, so window i owns
  // its kept interval minus the union of later kept intervals.  Kept intervals are ordered, so
  // the owned part is [lo, min(hi, min_{j>i} lo_j)) (empty if that is <= lo).
  int next_lo = 1 << 30;
  for (int i = n - 1; i >= 0; i--) {
    if (v[i].hi > next_lo) v[i].hi = next_lo;
    if (v[i].hi < v[i].lo) v[i].hi = v[i].lo;
    if (v[i].lo < next_lo) next_lo = v[i].lo;
  }
  return v;
}
}  // namespace

extern "C" int32_t wowsr_plan_windows(int32_t H, int32_t W, int32_t tile, int32_t pad, wowsr_window* out,
                                      int32_t cap) {
  if (H <= 0 || W <= 0 || tile <= 0 || pad < 0) return WOWSR_ERR_ARG;
  if ((int64_t)H * W <= (int64_t)4 * tile * tile) {
    if (out && cap >= 1) out[0] = wowsr_window{0, 0, W, H, 0, 0, W, H};
    return 1;
  }
  std::vector<Axis> ys = plan_axis(H, tile, pad), xs = plan_axis(W, tile, pad);
  int n = 0;
  for (auto& y : ys)
    for (auto& x : xs) {
      if (out && n < cap) out[n] = wowsr_window{x.a, y.a, x.b, y.b, x.lo, y.lo, x.hi, y.hi};
      n++;
    }
  return n;
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------

extern "C" int wowsr_abi_version(void) { return WOWSR_ABI_VERSION; }

extern "C" int wowsr_create(int device, wowsr_ctx** out) {
  if (!out) return WOWSR_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return wowsr_fail(nullptr, WOWSR_ERR_CUDA, "no CUDA device (%s); libwowsr has no CPU fallback",
                      cudaGetErrorString(e));
  if (device < 0 || device >= n) return wowsr_fail(nullptr, WOWSR_ERR_ARG, "device %d out of range (%d)", device, n);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return wowsr_fail(nullptr, WOWSR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return wowsr_fail(nullptr, WOWSR_ERR_UNSUPPORTED, "device %d is sm_%d%d; libwowsr is built for sm_100a only",
                      device, prop.major, prop.minor);
  wowsr_ctx* ctx = new wowsr_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  DeviceGuard g(device);
  e = cudaMalloc((void**)&ctx->d_tables, sizeof(WowsrTables));
  if (e == cudaSuccess)
    e = cudaMemcpy(ctx->d_tables, &wowsr_host_tables(), sizeof(WowsrTables), cudaMemcpyHostToDevice);
  for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev[i]);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->post_err, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) *ctx->post_err = 0;
  if (e != cudaSuccess) {
    wowsr_fail(nullptr, WOWSR_ERR_CUDA, "context setup: %s", cudaGetErrorString(e));
    delete ctx;
    return WOWSR_ERR_CUDA;
  }
  *out = ctx;
  return WOWSR_OK;
}

extern "C" void wowsr_destroy(wowsr_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->hist, &ctx->luts, &ctx->post_in, &ctx->post_out, &ctx->img_in, &ctx->img_out, &ctx->img_out_f32, &ctx->trace_buf};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->post_err) cudaFreeHost(ctx->post_err);
  for (int i = 0; i < 8; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  wowsr_net_free(ctx->net);
  wowsr_net_free(ctx->edsr);
  stage_free(ctx);
  delete ctx;
}

// Device image -> pageable host buffer through the pinned staging ring (hoststage.h): what the file entry points use to
// bring the finished image down before encoding it (wow_sr.py:126-164 writes files from a host array).
extern "C" int wowsr_download(wowsr_ctx* ctx, const void* dev, int64_t dev_pitch, int64_t row_bytes, int32_t rows, void* host,
                              int64_t host_pitch, void* stream) {
  if (!ctx || !dev || !host || rows < 1 || row_bytes < 1 || dev_pitch < row_bytes || host_pitch < row_bytes) return WOWSR_ERR_ARG;
  DeviceGuard g(ctx->device);
  if (int e = stage_init(ctx)) return e;
  WCUDA(ctx, cudaEventRecord(ctx->stage_sync, (cudaStream_t)stream));  // everything queued on the producer's stream so far
  StageOut sink(ctx, (uint8_t*)host, (size_t)host_pitch, (size_t)row_bytes);
  sink.enqueue((const uint8_t*)dev, (size_t)dev_pitch, 0, rows, ctx->stage_sync);
  return sink.finish();
}

extern "C" const char* wowsr_last_error(const wowsr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" uint64_t wowsr_launch_count(const wowsr_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Every key some wowsr_opt() call reads; anything else is a typo and is rejected (include/wowsr.h).
static const char* const kOptionKeys[] = {
    "conv_impl",   "hist_match", "mem_budget_mb",       "tc_boustrophedon", "tc_chunk32", "tc_flags",
    "tc_force_stream", "tc_generic_epilogue", "tc_grid", "tc_no_strip", "tc_stages", "tc_trace_layer",
    "tc_wbuf",     "trunk_hilo", "tail_fold_upsample",  "roll", "roll_pair", "roll_grid", "roll_ups", "hsv_simd_width", "post_kernel", "post_nt", "post_seg", "post_groups", "roll_adapt", "pdl"};

extern "C" int wowsr_set_option(wowsr_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return WOWSR_ERR_ARG;
  bool known = false;
  for (const char* k : kOptionKeys) known |= strcmp(k, key) == 0;
  if (!known) return wowsr_fail(ctx, WOWSR_ERR_ARG, "unknown option '%s'", key);
  ctx->opts[key] = value;
  return WOWSR_OK;
}
extern "C" int wowsr_get_option(const wowsr_ctx* ctx, const char* key, int64_t* value) {
  if (!ctx || !key || !value) return WOWSR_ERR_ARG;
  auto it = ctx->opts.find(key);
  if (it == ctx->opts.end()) return WOWSR_ERR_ARG;
  *value = it->second;
  return WOWSR_OK;
}
