// Host <-> device staging for the HOST-buffer entry points (wowsr_enhance_host, wowsr_post_process_host): the literal drop-in
// calls of the reference (RealESRGAN.enhance(ndarray), _enhance_for_crops(ndarray); wow_sr.py:85-110) hand over pageable numpy
// arrays.  A pageable cudaMemcpy is staged by the driver through its own bounce buffers on ONE CPU thread and does not overlap
// with anything (measured round 1: ~450 ms of copies around 810 ms of compute on BASELINE config 2).  Here a ring of pinned
// buffers is filled / drained by a few CPU threads while the copy engine moves the previous chunk on its own stream, and the
// kernels run on the chunks that have arrived.  One cudaMemcpyAsync per chunk.
#pragma once
#include <thread>

#include "common.h"

constexpr int STAGE_BUFS = 4;
constexpr size_t STAGE_BYTES = 32u << 20;
constexpr int STAGE_THREADS = 4;

inline int stage_init(wowsr_ctx* ctx) {
  if (ctx->stage[0]) return 0;
  for (int i = 0; i < STAGE_BUFS; i++) {
    WCUDA(ctx, cudaHostAlloc(&ctx->stage[i], STAGE_BYTES, cudaHostAllocDefault));
    WCUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
  }
  WCUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  WCUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_sync, cudaEventDisableTiming));
  return 0;
}

inline void stage_free(wowsr_ctx* ctx) {
  for (int i = 0; i < STAGE_BUFS; i++) {
    if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    ctx->stage[i] = nullptr;
    ctx->stage_ev[i] = nullptr;
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stage_sync) cudaEventDestroy(ctx->stage_sync);
  ctx->copy_stream = nullptr;
  ctx->stage_sync = nullptr;
}

// `rows` rows of `row_bytes` bytes between two host buffers with different pitches, split over a few threads.
inline void stage_memcpy_rows(uint8_t* dst, size_t dpitch, const uint8_t* src, size_t spitch, size_t row_bytes, int rows) {
  auto work = [=](int r0, int r1) {
    if (dpitch == row_bytes && spitch == row_bytes) {
      memcpy(dst + (size_t)r0 * dpitch, src + (size_t)r0 * spitch, (size_t)(r1 - r0) * row_bytes);
    } else {
      for (int r = r0; r < r1; r++) memcpy(dst + (size_t)r * dpitch, src + (size_t)r * spitch, row_bytes);
    }
  };
  const size_t total = (size_t)rows * row_bytes;
  int nt = total < (4u << 20) ? 1 : STAGE_THREADS;
  if (nt > rows) nt = rows;
  if (nt <= 1) {
    work(0, rows);
    return;
  }
  std::thread th[STAGE_THREADS];
  for (int t = 1; t < nt; t++) th[t] = std::thread(work, (int)((long long)rows * t / nt), (int)((long long)rows * (t + 1) / nt));
  work(0, rows / nt);
  for (int t = 1; t < nt; t++) th[t].join();
}

// Device rows -> host rows through the pinned ring.  enqueue() may be called while kernels for later rows are still being
// launched: the copies run on ctx->copy_stream.  A chunk whose pinned buffer is needed again is drained first; finish() drains
// everything.  The device rows must be complete (or ordered before `after`, an event on the producing stream).
struct StageOut {
  wowsr_ctx* ctx;
  uint8_t* host;
  size_t host_pitch, row_bytes;
  struct Pending { int buf, row0, rows; bool live; } pend[STAGE_BUFS];
  int next = 0;
  int err = 0;

  StageOut(wowsr_ctx* c, uint8_t* h, size_t hp, size_t rb) : ctx(c), host(h), host_pitch(hp), row_bytes(rb) {
    for (auto& p : pend) p.live = false;
  }
  void drain_one(int b) {
    Pending& p = pend[b];
    if (!p.live) return;
    if (cudaEventSynchronize(ctx->stage_ev[b]) != cudaSuccess && !err) err = wowsr_fail(ctx, WOWSR_ERR_CUDA, "staged device-to-host copy failed");
    stage_memcpy_rows(host + (size_t)p.row0 * host_pitch, host_pitch, (const uint8_t*)ctx->stage[b], row_bytes, row_bytes, p.rows);
    p.live = false;
  }
  // rows [row0, row1) of a device image whose row `y` starts at dev + y * dev_pitch
  void enqueue(const uint8_t* dev, size_t dev_pitch, int row0, int row1, cudaEvent_t after) {
    if (err || row1 <= row0) return;
    if (after && cudaStreamWaitEvent(ctx->copy_stream, after, 0) != cudaSuccess) { err = wowsr_fail(ctx, WOWSR_ERR_CUDA, "cudaStreamWaitEvent"); return; }
    const int per = (int)std::max<size_t>(1, STAGE_BYTES / row_bytes);
    for (int r = row0; r < row1 && !err; r += per) {
      const int n = std::min(per, row1 - r);
      const int b = next;
      next = (next + 1) % STAGE_BUFS;
      drain_one(b);
      cudaError_t e = cudaMemcpy2DAsync(ctx->stage[b], row_bytes, dev + (size_t)r * dev_pitch, dev_pitch, row_bytes, n,
                                        cudaMemcpyDeviceToHost, ctx->copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_ev[b], ctx->copy_stream);
      if (e != cudaSuccess) { err = wowsr_fail(ctx, WOWSR_ERR_CUDA, "staged device-to-host copy: %s", cudaGetErrorString(e)); return; }
      pend[b] = Pending{b, r, n, true};
    }
  }
  // copies that have already landed, without blocking on the ones still in flight
  void poll() {
    for (int i = 0; i < STAGE_BUFS; i++) {
      const int b = (next + i) % STAGE_BUFS;  // oldest first
      if (pend[b].live && cudaEventQuery(ctx->stage_ev[b]) == cudaSuccess) drain_one(b);
    }
  }
  int finish() {
    for (int i = 0; i < STAGE_BUFS; i++) drain_one((next + i) % STAGE_BUFS);
    return err;
  }
};

// Host rows -> device rows through the pinned ring; `on_chunk(row0, row1, ev)` is called after each chunk's copy has been
// enqueued, `ev` being recorded behind it on ctx->copy_stream (the caller orders its kernels with cudaStreamWaitEvent).
template <class F>
int stage_in(wowsr_ctx* ctx, uint8_t* dev, size_t dev_pitch, const uint8_t* host, size_t host_pitch, size_t row_bytes, int rows, F on_chunk) {
  const int per = (int)std::max<size_t>(1, STAGE_BYTES / row_bytes);
  bool used[STAGE_BUFS] = {false, false, false, false};
  int next = 0;
  for (int r = 0; r < rows; r += per) {
    const int n = std::min(per, rows - r);
    const int b = next;
    next = (next + 1) % STAGE_BUFS;
    if (used[b]) WCUDA(ctx, cudaEventSynchronize(ctx->stage_ev[b]));  // its previous upload has left the pinned buffer
    stage_memcpy_rows((uint8_t*)ctx->stage[b], row_bytes, host + (size_t)r * host_pitch, host_pitch, row_bytes, n);
    WCUDA(ctx, cudaMemcpy2DAsync(dev + (size_t)r * dev_pitch, dev_pitch, ctx->stage[b], row_bytes, row_bytes, n, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    WCUDA(ctx, cudaEventRecord(ctx->stage_ev[b], ctx->copy_stream));
    used[b] = true;
    if (int e = on_chunk(r, r + n, ctx->stage_ev[b])) return e;
  }
  return 0;
}
