// seams_host.cpp - TEST INFRASTRUCTURE: replays the lane loops of k_index_heatwaves / k_season_metrics
// (hdp_b200/csrc/seams.cu) on the CPU around the very same per-run / per-season functions (hdp_b200/csrc/seams.h), so that
// tests/test_seams_host.py can check that logic against the oracle and the reference's known answers without a GPU.
// Not part of the product: the library computes on the device only.
//     g++ -O2 -shared -fPIC -I hdp_b200/csrc tests/seams_host.cpp -o <tmp>/libseams_host.so
#include <stdint.h>
#include <limits.h>

#include "seams.h"

using namespace hdp;

static int ffs32(uint32_t v) { return __builtin_ffs((int)v); }

extern "C" {

// one warp of k_index_heatwaves: series `hot` u8 [T], one definition -> hw i64 [T] (zero-filled here, as the entry point does)
void seams_host_index_heatwaves(const uint8_t *hot, int64_t T, int64_t min_duration, int64_t max_break, int64_t max_subs, int64_t *hw)
{
    for (int64_t t = 0; t < T; t++) hw[t] = 0;
    IndexState st;
    int64_t run_start = 0;
    uint32_t carry = 0u;
    for (int64_t t0 = 0; t0 < T; t0 += 32) {
        uint32_t w = 0u;                                                   // __ballot_sync
        for (int lane = 0; lane < 32; lane++) {
            const int64_t t = t0 + lane;
            if (t < T && hot[t] != 0) w |= 1u << lane;
        }
        uint32_t trans = w ^ ((w << 1) | carry);
        carry = w >> 31;
        while (trans != 0u) {
            const int b = ffs32(trans) - 1;
            trans &= trans - 1u;
            if ((w >> b) & 1u) {
                run_start = t0 + b;
            } else {
                const int64_t e = t0 + b;
                const int64_t id = index_run(st, run_start, e, min_duration, max_break, max_subs);
                if (id != 0)
                    for (int lane = 0; lane < 32; lane++)
                        for (int64_t j = run_start + lane; j < e; j += 32) hw[j] = id;
            }
        }
    }
    if (carry) {
        const int64_t id = index_run(st, run_start, T, min_duration, max_break, max_subs);
        if (id != 0)
            for (int lane = 0; lane < 32; lane++)
                for (int64_t j = run_start + lane; j < T; j += 32) hw[j] = id;
    }
}

// the warps of k_season_metrics for one series: hw i64 [T], seasons i64 [Y, 2] -> hwf / hwn / hwd i64 [Y], hwa f64 [Y]
void seams_host_season_metrics(const int64_t *hw, int64_t T, const int64_t *seasons, int Y,
                               int64_t *hwf, int64_t *hwn, int64_t *hwd, double *hwa)
{
    for (int y = 0; y < Y; y++) {
        int64_t lo, hi;
        season_slice(seasons[2 * y], seasons[2 * y + 1], T, lo, hi);
        const int64_t n = hi - lo;
        const int64_t *v = hw + lo;
        int64_t vmin = INT64_MAX, vmax = INT64_MIN;
        for (int64_t i = 0; i < n; i++) {
            vmin = v[i] < vmin ? v[i] : vmin;
            vmax = v[i] > vmax ? v[i] : vmax;
        }
        const bool multi = n > 0 && vmin != vmax;
        SeasonAcc total;
        for (int lane = 0; lane < 32; lane++) {
            SeasonAcc acc;
            season_lane(v, n, lane, 32, multi, vmin, acc);
            season_merge(total, acc);
        }
        hwf[y] = total.hot;
        hwn[y] = total.uniq_nz;
        hwd[y] = total.longest;
        hwa[y] = season_average(total, n, multi);
    }
}

}  // extern "C"
