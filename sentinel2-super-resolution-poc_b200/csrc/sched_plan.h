// EXPERIMENTAL — round-2 work in progress, reached only through option trunk_fuse (sched_kernel.cuh).
//
// Host-side schedule of a "fused tail" launch: the last few convs of ONE residual dense block (cnn_super_resolution.py:85-91)
// over ALL windows of a batch, as one persistent kernel whose task list is a skewed wavefront.  At step s the list holds
// tile s of the first fused conv, tile s - lag of the next, ... (tile = index into that conv's window-major tile list), so
//   * a consumer trails its producers by `lag` steps: their results were published long before it polls for them, which
//     hides the publish -> poll -> TMA latency that a layer-major order exposes on every task (tools/dataflow_sim.py), and
//   * what it re-reads was touched at most (layers - 1) * lag steps earlier: the re-read distance, not the batch size,
//     decides what has to stay in L2 (conv4 + conv5, lag 120: ~47 MB of a 276 x 276 batch).
// Dependencies are per 8-row band of a window: every tile publishes into the bands it covers, a consumer waits until the
// bands its halo touches hold `band_target` publications.  The plan is plain C++ (no CUDA) so that the CPU tests can check it:
// every producer of a band precedes its consumers in the list (deadlock freedom of in-order CTAs), every tile appears once.
#pragma once
#include <stdint.h>

#include <vector>

struct SchedTask {  // 16 bytes, read by the three roles of the kernel
  uint8_t k;        // conv index inside the RDB, 0 = conv1 .. 4 = conv5
  uint8_t vert;     // 1: vertical tile of the remainder strip
  uint16_t win;     // window of the batch
  uint16_t u0, v0;  // tile origin on the run axis / row axis (pixels)
  uint16_t dep_b0, dep_n;  // bands [dep_b0, dep_b0 + dep_n) of conv k - 1 must be complete (dep_n = 0: no in-launch dependency)
  uint16_t pub_b0, pub_n;  // bands of conv k this tile publishes into (pub_n = 0: nobody in this launch waits for it)
};

struct SchedGeom {  // tile geometry of one conv kind over one window (mirrors TrunkKind)
  int R, tiles_x, tiles_y, v_runs, v_rows;
};

struct SchedPlan {
  int h = 0, w = 0, n_win = 0, strip_x0 = 0;
  int k_first = 0, n_layers = 0;  // fused convs k_first .. 4
  int band_rows = 8, n_bands = 0;
  int lag = 0;                    // steps (tiles of the first fused conv) between consecutive convs
  int band_target_tiles = 0;      // publishing tiles per band of an N = 32 conv (x epilogue warps = counter target)
  SchedGeom kind[2];
  std::vector<SchedTask> tasks;
};

namespace sched_detail {

inline SchedGeom geom(int h, int strip_x0, int rem, int R) {
  SchedGeom g;
  g.R = R;
  g.tiles_x = (strip_x0 + 127) / 128;
  g.tiles_y = (h + R - 1) / R;
  g.v_runs = rem ? (h + 127) / 128 : 0;
  g.v_rows = rem ? (rem + R - 1) / R : 0;
  return g;
}

// Tiles of conv k over one window: vertical strip tiles of y-run u right before the horizontal row block that starts at row 128 u.
inline void window_tiles(const SchedPlan& P, int k, int win, std::vector<SchedTask>& out) {
  const SchedGeom& g = P.kind[k == 4];
  const int br = P.band_rows;
  auto band_range = [&](int r0, int r1, uint16_t& b0, uint16_t& n) {  // rows [r0, r1] clipped to the window
    if (r0 < 0) r0 = 0;
    if (r1 > P.h - 1) r1 = P.h - 1;
    b0 = (uint16_t)(r0 / br);
    n = (uint16_t)(r1 / br - r0 / br + 1);
  };
  int next_vrun = 0;
  for (int vb = 0; vb < g.tiles_y; vb++) {
    const int row0 = vb * g.R;
    while (next_vrun < g.v_runs && row0 >= next_vrun * 128) {
      for (int c = 0; c < g.v_rows; c++) {
        SchedTask t{};
        t.k = (uint8_t)k; t.vert = 1; t.win = (uint16_t)win;
        t.u0 = (uint16_t)(next_vrun * 128); t.v0 = (uint16_t)(P.strip_x0 + c * g.R);
        if (k > P.k_first) band_range(t.u0 - 1, t.u0 + 128, t.dep_b0, t.dep_n);
        if (k < 4) band_range(t.u0, t.u0 + 127, t.pub_b0, t.pub_n);
        out.push_back(t);
      }
      next_vrun++;
    }
    for (int ur = 0; ur < g.tiles_x; ur++) {
      SchedTask t{};
      t.k = (uint8_t)k; t.vert = 0; t.win = (uint16_t)win;
      t.u0 = (uint16_t)(ur * 128); t.v0 = (uint16_t)row0;
      if (k > P.k_first) band_range(row0 - 1, row0 + g.R, t.dep_b0, t.dep_n);
      if (k < 4) band_range(row0, row0 + g.R - 1, t.pub_b0, t.pub_n);
      out.push_back(t);
    }
  }
}

// Walks the list in order and checks that every dependency is complete when its consumer is reached.
inline bool producers_precede_consumers(const SchedPlan& P) {
  std::vector<int> cnt((size_t)P.n_layers * P.n_win * P.n_bands, 0);
  auto at = [&](int k, int win, int b) -> int& { return cnt[((size_t)(k - P.k_first) * P.n_win + win) * P.n_bands + b]; };
  for (const SchedTask& t : P.tasks) {
    for (int b = t.dep_b0; b < t.dep_b0 + t.dep_n; b++)
      if (at(t.k - 1, t.win, b) != P.band_target_tiles) return false;
    for (int b = t.pub_b0; b < t.pub_b0 + t.pub_n; b++) at(t.k, t.win, b)++;
  }
  return true;
}

}  // namespace sched_detail

// Builds the plan.  `lag` > 0 is taken literally (false if it is not legal).  `lag` <= 0 is automatic: the smallest legal
// multiple of 8 steps (producers precede consumers; >= `min_lag`) plus `slack_tasks` tasks' worth of steps, so that a consumer's
// last producer is more than one machine-full of tasks (148 CTAs) behind it and the publish -> poll -> TMA latency is hidden;
// slack_tasks = 0 returns the legal minimum itself.  Returns false when the shape cannot be scheduled (no band structure: h < 8).
inline bool sched_build(SchedPlan& P, int h, int w, int n_win, int k_first, int lag, int min_lag = 48, int slack_tasks = 240) {
  using namespace sched_detail;
  if (h < 8 || w < 1 || n_win < 1 || n_win > 65535 || h > 65000 || w > 65000 || k_first < 0 || k_first > 3) return false;
  P.h = h; P.w = w; P.n_win = n_win; P.k_first = k_first; P.n_layers = 5 - k_first;
  const int wm = w / 128 * 128, rem_raw = w - wm;
  const bool strip = rem_raw > 0 && wm > 0 && h >= 64;  // same rule as F32Layout in rrdbnet_batch (conv.cu)
  P.strip_x0 = strip ? wm : w;
  const int rem = strip ? rem_raw : 0;
  P.kind[0] = geom(h, P.strip_x0, rem, 8);
  P.kind[1] = geom(h, P.strip_x0, rem, 4);
  P.band_rows = 8;
  P.n_bands = (h + 7) / 8;
  P.band_target_tiles = P.kind[0].tiles_x + P.kind[0].v_rows;  // every band: all horizontal runs + all column blocks of one strip run
  std::vector<std::vector<SchedTask>> per(P.n_layers);
  for (int i = 0; i < P.n_layers; i++)
    for (int win = 0; win < n_win; win++) window_tiles(P, k_first + i, win, per[i]);
  const size_t n0 = per[0].size();
  // lag = n0 is plain layer-major order (a conv starts when its producer has finished everywhere): always legal
  for (int try_lag = lag > 0 ? lag : min_lag;; try_lag += 8) {
    if (try_lag > (int)n0) try_lag = (int)n0;
    P.lag = try_lag;
    P.tasks.clear();
    for (size_t s = 0; s < n0 + (size_t)(P.n_layers - 1) * try_lag; s++)
      for (int i = 0; i < P.n_layers; i++) {
        const long long u = (long long)s - (long long)i * try_lag;
        if (u < 0 || u >= (long long)n0) continue;
        const size_t nk = per[i].size();
        for (size_t j = (size_t)u * nk / n0; j < (size_t)(u + 1) * nk / n0; j++) P.tasks.push_back(per[i][j]);
      }
    if (producers_precede_consumers(P)) {
      if (lag > 0 || slack_tasks <= 0 || try_lag >= (int)n0) return true;
      const int tasks_per_step = P.n_layers + 1;  // one tile of every N = 32 conv + two 4-row tiles of conv5
      int want = try_lag + (slack_tasks + tasks_per_step - 1) / tasks_per_step;
      if (want > (int)n0) want = (int)n0;
      return sched_build(P, h, w, n_win, k_first, want, min_lag, 0);  // a longer lag only delays consumers: still legal
    }
    if (lag > 0 || try_lag >= (int)n0) return false;  // an explicit lag is taken literally
  }
}
