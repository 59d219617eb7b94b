// thr_net.cuh - host-side interface of k_thr_net (thr_net.cu) for thresholds_launch (threshold.cu).
#pragma once

#include <stdint.h>
#include <vector>

#include "common.cuh"

namespace hdp {

// Positions and weights of the requested percentiles, counted from the TOP of a window (index 0 = largest sample).
struct NetSel {
    int idx_lo[HDP_B200_MAX_PERCENTILES], idx_hi[HDP_B200_MAX_PERCENTILES], is_max[HDP_B200_MAX_PERCENTILES];
    double w_lo[HDP_B200_MAX_PERCENTILES], w_hi[HDP_B200_MAX_PERCENTILES];
};

// Where a cell with NaN / +-inf samples goes: onto the hand-over list of k_thr_seg (threshold.cu), block numbering included.
struct NetHandOver {
    uint32_t *list;                     // [count, flags[n_blocks], blocks[n_blocks]]
    int n_blocks, n_seg, gc, n_groups, group_cells;
};

struct NetGeom {
    int NY, K, M;                       // template instance: samples per row (padded), list length, blocks per window (1 or 3)
    int s;                              // rows per block: W == M * s
    int W, n_y, n_doy, P;
    int unit;                           // input unit of the samples (common.cuh to_celsius_f), converted as they are loaded
    int n_seq;                          // rows of the linear row sequence (n_doy + r); row n_seq of seq_time is the all-pad row
    int n_win;                          // windows 0 .. n_win-1 of W consecutive sequence rows; window k belongs to day win_day[k] (or -1)
    int n_steps, steps_per_chunk, n_chunks;
    int n_irr;                          // days whose window is not a run of W consecutive rows (the mirrored year end)
    int n_irr_steps;                    // row steps of the program that works them out (irr_day / irr_time): k_thr_net_irr, one warp per tile
    int64_t n_tiles;                    // 32-cell tiles
    size_t smem;                        // k_thr_net: bytes per one-warp CTA
    // k_thr_net_tm (persistent CTAs, suffix lists in tensor memory): warps per CTA (0 = not used), suffix lists per warp kept in
    // TMEM, TMEM columns per warp, shared-memory slots per warp
    int tm_warps, tm_lists, tm_cols, tm_smem_slots;
};

struct NetPlan {
    bool usable = false;
    NetGeom geo{};
    NetSel sel{};
    std::vector<int> seq_time;          // [(n_seq + 1) * NY] time index of every sample of every sequence row, -1 = pad
    std::vector<int> win_day;           // [n_steps * s]
    std::vector<int> irr_day;           // the irregular days' PROGRAM, 4 ints per row step: {flags (1 = the running list starts over, 2 = the
                                        // row is merged into a COPY of the running list), (other slot + 1) | (slot the running list is
                                        // stored to + 1) << 8, day to emit or -1, 0}
    std::vector<int> irr_time;          // [n_irr_steps * NY] time indices of the row of every step
};

struct NetTables {                      // device copies (workspace)
    int *seq_time = nullptr, *win_day = nullptr, *irr_day = nullptr, *irr_time = nullptr;
    unsigned long long *next_item = nullptr;   // k_thr_net_tm's work counter
};

// Decides whether the tables and quantiles fit k_thr_net and fills the plan (host only, no CUDA calls).
void net_plan(const int32_t *time_index, const int32_t *win_rows, int64_t T_b, int n_doy, int n_y, int W,
              const int *pos_lo, const int *pos_hi, const int *mode_is_max, const int *mode_is_interp, const double *w_lo, const double *w_hi,
              int P, int64_t C, NetPlan &pl);
void net_set_cells(NetPlan &pl, int64_t C);
int net_launch(const NetPlan &pl, const NetTables &tb, const float *x, int64_t C, int64_t ld_t, double *out, const NetHandOver &hand, cudaStream_t st,
               bool allow_tmem);

}  // namespace hdp
