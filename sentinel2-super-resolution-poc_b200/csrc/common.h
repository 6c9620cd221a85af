// Internal header of libwowsr.so (not part of the ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/wowsr.h"

// Integer tables shared by the post-process kernels.
struct WowsrTables {
  uint16_t gam[256];     // A.1  rint(2040 * srgb_to_linear(i/255))
  uint16_t cbrt[3072];   // A.1  rint(32768 * labf(j/2040))
  uint16_t lab_y[256];   // A.3
  uint16_t lab_ify[256]; // A.3
  uint8_t invgam[4096];  // A.3  rint(255 * linear_to_srgb(k/4096))
  uint32_t sdiv[256];    // A.6
  uint32_t hdiv[256];    // A.6
};
const WowsrTables& wowsr_host_tables();

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ConvNet;  // conv.cu

struct wowsr_ctx {
  int device = 0;
  int sm_count = 148;
  std::string err;
  uint64_t launches = 0;
  bool tc_attr_set = false;
  WowsrTables* d_tables = nullptr;
  std::map<std::string, int64_t> opts;
  // post-process scratch
  DevBuf hist, luts, post_in, post_out;
  int* post_err = nullptr;  // mapped pinned host word: a post-process kernel whose exchange barrier timed out sets it (post.cu)
  // network state
  ConvNet* net = nullptr;
  ConvNet* edsr = nullptr;
  DevBuf img_in, img_out, img_out_f32;
  DevBuf trace_buf;
  int64_t trace_counter = 0;
  float timing[8] = {0};
  cudaEvent_t ev[8] = {nullptr};
  // pinned staging ring of the host-buffer entry points (hoststage.h), allocated on first use
  void* stage[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_sync = nullptr;
  cudaStream_t copy_stream = nullptr;
};

int wowsr_fail(wowsr_ctx* ctx, int code, const char* fmt, ...);
int wowsr_ensure(wowsr_ctx* ctx, DevBuf& b, size_t bytes);
int64_t wowsr_opt(const wowsr_ctx* ctx, const char* key, int64_t dflt);
void wowsr_net_free(ConvNet* n);

#define WCUDA(ctx, call)                                                                    \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return wowsr_fail(ctx, WOWSR_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
                        cudaGetErrorString(e__));                                           \
  } while (0)

#define WLAUNCH_CHECK(ctx)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return wowsr_fail(ctx, WOWSR_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,      \
                        cudaGetErrorString(e__));                                           \
    (ctx)->launches++;                                                                      \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
