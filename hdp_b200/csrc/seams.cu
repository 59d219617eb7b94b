// seams.cu - the BUILDING BLOCKS of path 2 under the reference's own names, on the device.
//
// The hot path (metric.cu) never materialises what these functions exchange - the per-day heatwave id series - but the
// reference exposes them as public functions and its unit tests call them directly (hdp/tests/test_index_heatwaves.py,
// test_heatwave_{frequency,number,duration,average}.py).  A user who switches libraries finds them here, with the reference's
// semantics for ARBITRARY inputs (any id series, not only one that index_heatwaves produced):
//
//   k_index_heatwaves   index_heatwaves                                   hdp/metric.py:11-60
//   k_season_metrics    heatwave_frequency / number / duration / average  hdp/metric.py:63-172
//
// One warp per series (x definition / x season); the per-run and per-season logic lives in seams.h, shared with the host
// replay of tests/seams_host.cpp.  Integer work on a few bytes per day: nothing here is shaped for tensor cores.
#include "common.cuh"
#include "seams.h"

namespace hdp {

constexpr int kSeamWarps = 4;

struct SeamDefs {                      // [min_duration, max_break, max_subs] per definition (by value: at most 32 x 3 ints)
    int32_t v[3 * HDP_B200_MAX_DEFINITIONS];
};

// hot u8 [S, T] (non-zero = hot day) -> hw i64 [S, D, T], zero-filled by the caller (only labelled runs are written).
// Warp = one (series, definition).  32 days per step: one coalesced load, a ballot turns them into a word of hot bits, the
// word's transitions are visited in time order by every lane alike (the state machine is warp-uniform), and a labelled run is
// filled by the lanes together, 32 ids per store instruction.
__global__ void __launch_bounds__(kSeamWarps * 32)
k_index_heatwaves(const uint8_t *__restrict__ hot, int64_t S, int64_t T, const __grid_constant__ SeamDefs defs, int D,
                  int64_t *__restrict__ hw)
{
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * kSeamWarps + (threadIdx.x >> 5);
    if (wg >= S * D) return;                                               // warp-uniform
    const int64_t s = wg / D;
    const int d = (int)(wg - s * D);
    const int64_t min_duration = defs.v[3 * d], max_break = defs.v[3 * d + 1], max_subs = defs.v[3 * d + 2];
    const uint8_t *h = hot + s * T;
    int64_t *o = hw + (s * D + d) * T;

    IndexState st;
    int64_t run_start = 0;
    uint32_t carry = 0u;                                                   // yesterday was hot
    for (int64_t t0 = 0; t0 < T; t0 += 32) {
        const int64_t t = t0 + lane;
        const uint32_t w = __ballot_sync(0xffffffffu, t < T && h[t] != 0);
        uint32_t trans = w ^ ((w << 1) | carry);                           // bit b: day t0 + b differs from the day before
        carry = w >> 31;
        while (trans != 0u) {
            const int b = __ffs(trans) - 1;
            trans &= trans - 1u;
            if ((w >> b) & 1u) {
                run_start = t0 + b;
            } else {
                const int64_t e = t0 + b;
                const int64_t id = index_run(st, run_start, e, min_duration, max_break, max_subs);
                if (id != 0)
                    for (int64_t j = run_start + lane; j < e; j += 32) o[j] = id;
            }
        }
    }
    if (carry) {                                                           // the series ends hot: the pad day closes the run
        const int64_t id = index_run(st, run_start, T, min_duration, max_break, max_subs);
        if (id != 0)
            for (int64_t j = run_start + lane; j < T; j += 32) o[j] = id;
    }
}

__device__ __forceinline__ int64_t warp_sum(int64_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int64_t warp_max(int64_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int64_t u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
    return v;
}
__device__ __forceinline__ int64_t warp_min(int64_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int64_t u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
    return v;
}

// hw i64 [S, T], seasons i64 [Y, 2] (raw: Python slice rules are applied here) -> hwf / hwn / hwd i64 [S, Y], hwa f64 [S, Y]
// (any of them may be null).  Warp = one (series, season); the lanes split the days of the slice, each day looks at the whole
// slice (np.unique's job without a sort: seasons are a few hundred days), a butterfly adds the lanes up.
__global__ void __launch_bounds__(kSeamWarps * 32)
k_season_metrics(const int64_t *__restrict__ hw, int64_t S, int64_t T, const int64_t *__restrict__ seasons, int Y,
                 int64_t *__restrict__ hwf, int64_t *__restrict__ hwn, int64_t *__restrict__ hwd, double *__restrict__ hwa)
{
    const int lane = threadIdx.x & 31;
    const int64_t wg = (int64_t)blockIdx.x * kSeamWarps + (threadIdx.x >> 5);
    if (wg >= S * Y) return;                                               // warp-uniform
    const int64_t s = wg / Y;
    const int y = (int)(wg - s * Y);
    int64_t lo, hi;
    season_slice(seasons[2 * y], seasons[2 * y + 1], T, lo, hi);
    const int64_t n = hi - lo;
    const int64_t *v = hw + s * T + lo;

    int64_t vmin = INT64_MAX, vmax = INT64_MIN;
    for (int64_t i = lane; i < n; i += 32) {
        const int64_t x = v[i];
        vmin = x < vmin ? x : vmin;
        vmax = x > vmax ? x : vmax;
    }
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    const bool multi = n > 0 && vmin != vmax;

    SeasonAcc acc;
    season_lane(v, n, lane, 32, multi, vmin, acc);
    acc.hot = warp_sum(acc.hot);
    acc.uniq = warp_sum(acc.uniq);
    acc.uniq_nz = warp_sum(acc.uniq_nz);
    acc.sum = warp_sum(acc.sum);
    acc.longest = warp_max(acc.longest);
    if (lane == 0) {
        const int64_t o = s * Y + y;
        if (hwf) hwf[o] = acc.hot;
        if (hwn) hwn[o] = acc.uniq_nz;
        if (hwd) hwd[o] = acc.longest;
        if (hwa) hwa[o] = season_average(acc, n, multi);
    }
}

}  // namespace hdp

using namespace hdp;

extern "C" {

int hdp_b200_index_heatwaves(const uint8_t *d_hot, int64_t S, int64_t T, const int32_t *h_defs, int D, int64_t *d_hw, void *stream)
{
    if (S < 0 || T < 0 || D <= 0 || !h_defs) return HDP_B200_ERR_INVALID;
    if (D > HDP_B200_MAX_DEFINITIONS) return HDP_B200_ERR_UNSUPPORTED;
    if (S == 0 || T == 0) return HDP_B200_OK;
    if (!d_hot || !d_hw) return HDP_B200_ERR_INVALID;
    const int64_t warps = S * D, blocks = (warps + kSeamWarps - 1) / kSeamWarps;
    if (blocks > 0x7fffffffLL) return HDP_B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    SeamDefs defs;
    for (int i = 0; i < 3 * HDP_B200_MAX_DEFINITIONS; i++) defs.v[i] = i < 3 * D ? h_defs[i] : 0;
    HDP_CUDA_TRY(cudaMemsetAsync(d_hw, 0, sizeof(int64_t) * (size_t)S * (size_t)D * (size_t)T, st));
    KernelTimer timer(kSeam, st);
    k_index_heatwaves<<<(unsigned)blocks, kSeamWarps * 32, 0, st>>>(d_hot, S, T, defs, D, d_hw);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

int hdp_b200_season_metrics(const int64_t *d_hw, int64_t S, int64_t T, const int64_t *d_seasons, int Y,
                            int64_t *d_hwf, int64_t *d_hwn, int64_t *d_hwd, double *d_hwa, void *stream)
{
    if (S < 0 || T < 0 || Y < 0) return HDP_B200_ERR_INVALID;
    if (S == 0 || Y == 0) return HDP_B200_OK;
    if (!d_seasons || (T > 0 && !d_hw)) return HDP_B200_ERR_INVALID;
    const int64_t warps = S * Y, blocks = (warps + kSeamWarps - 1) / kSeamWarps;
    if (blocks > 0x7fffffffLL) return HDP_B200_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    KernelTimer timer(kSeam, st);
    k_season_metrics<<<(unsigned)blocks, kSeamWarps * 32, 0, st>>>(d_hw, S, T, d_seasons, Y, d_hwf, d_hwn, d_hwd, d_hwa);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

}  // extern "C"
