// threshold.cu - path 1: per cell and day of year, percentiles over the pooled day-of-year window.
//
// Replaces (reference = AgentOxygen/HDP v1.0.2):
//   compute_percentiles gufunc        hdp/threshold.py:52-78
//   compute_percentiles_wrapper       hdp/threshold.py:81-93   (the loop over cells)
//   np.quantile as compiled by Numba  numba/np/arraymath.py:1655-1704, 1754-1768
//
// The quantile arithmetic is reproduced operation by operation in double precision with explicitly
// rounded intrinsics (__dmul_rn/__dadd_rn/...), so nvcc cannot contract the interpolation into an FMA:
//     rank = 1 + (n-1) * ((q*100)/100);  f = floor(rank);  m = rank - f
//     val  = sorted[f-1] * (1-m) + sorted[f] * m
// which is bit-identical to the reference (tests/test_gpu_parity.py compares at 0 ulp).
#include <math.h>
#include <vector>

#include "common.cuh"

namespace hdp {

struct QTable {
    double q[HDP_B200_MAX_PERCENTILES];
};

// One output value from an ascending-sorted window (sorted[i], i < n, NaN-free: NaNs are counted separately).
// n_nan / n_pinf / n_ninf: how many NaN, +inf, -inf samples the window holds.
__device__ __forceinline__ double quantile_from_sorted(const float *sorted, int n, double q, int n_nan, int n_pinf, int n_ninf)
{
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    // _can_collect_percentiles, arraymath.py:1707-1721
    if (n_nan > 0 || n == 0) return nan;
    if (n == 1) return (n_pinf + n_ninf) ? nan : (double)sorted[0];           // arraymath.py:1661-1663 / :1719
    const double pct = __dmul_rn(q, 100.0);                                   // arraymath.py:1757
    const bool all_finite = (n_pinf + n_ninf) == 0;
    if (pct == 100.0) {                                                       // arraymath.py:1669-1675
        const double v = (double)sorted[n - 1];
        return (!all_finite && isinf(v)) ? nan : v;
    }
    if (pct == 0.0) {                                                         // arraymath.py:1678-1695
        double v = (double)sorted[0];
        if (!all_finite) {
            const int n_fin = n - (n_pinf + n_ninf);
            if (n_fin == 0) v = nan;
            if (n_pinf == 1 && n == 2) v = nan;
            if (n_ninf > 1) v = nan;
            if (n_fin == 1 && n_pinf > 1 && n_ninf != 1) v = nan;
        }
        return v;
    }
    // arraymath.py:1697-1701
    const double rank = __dadd_rn(1.0, __dmul_rn((double)(n - 1), __ddiv_rn(pct, 100.0)));
    const double f = floor(rank);
    const double m = __dsub_rn(rank, f);
    int k = (int)f - 1;
    double lower, upper;
    if (k >= n - 1) { lower = upper = (double)sorted[n - 1]; }                // rank == n: q rounded up to the maximum
    else { if (k < 0) k = 0; lower = (double)sorted[k]; upper = (double)sorted[k + 1]; }
    return __dadd_rn(__dmul_rn(lower, __dsub_rn(1.0, m)), __dmul_rn(upper, m));
}

// ----------------------------------------------------------------------------------------------------
// k_thr_generic: gather + bitonic sort per (cell, day of year).  Handles every table the reference can
// produce (mirrored upper wrap, -1 pads, duplicated rows, any window size up to 32768 samples).
// A CTA sorts NC cells' windows for one day of year side by side.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_thr_generic(const float *__restrict__ temps, int64_t C, int64_t T_b, int64_t ld_t,
              const int *__restrict__ time_index, const int *__restrict__ win_rows, int n_doy, int n_y, int W,
              const __grid_constant__ QTable qt, int P, int NC, int b_pad_log2, double *__restrict__ out)
{
    extern __shared__ float keys[];                       // [NC][b_pad], then int counts[NC][3]
    const int b = W * n_y, b_pad = 1 << b_pad_log2;
    int *counts = (int *)(keys + (size_t)NC * b_pad);
    const int tid = threadIdx.x;
    const int d = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * NC;

    for (int i = tid; i < NC * 3; i += 256) counts[i] = 0;
    __syncthreads();

    // gather: consecutive threads -> consecutive cells of the same sample (one 32-byte sector for 8 cells)
    const float pinf = __int_as_float(0x7f800000);
    for (int idx = tid; idx < NC * b_pad; idx += 256) {
        const int cell = idx % NC, i = idx / NC;
        float v = pinf;                                   // padding sorts to the end
        if (i < b && c0 + cell < C) {
            const int row = win_rows[d * W + i / n_y];
            int64_t t = time_index[row * n_y + i % n_y];
            if (t < 0) t += T_b;                          // -1 pads read the LAST sample (threshold.py:35,77)
            v = temps[t * ld_t + (c0 + cell)];
            if (v != v) { atomicAdd(&counts[cell * 3 + 0], 1); v = pinf; }
            else if (v == pinf) atomicAdd(&counts[cell * 3 + 1], 1);
            else if (v == -pinf) atomicAdd(&counts[cell * 3 + 2], 1);
        }
        keys[(size_t)cell * b_pad + i] = v;
    }
    __syncthreads();

    // bitonic sort, ascending, all NC arrays in lock step
    const int half = b_pad >> 1;
    for (int k = 2; k <= b_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int pi = tid; pi < NC * half; pi += 256) {
                const int arr = pi >> (b_pad_log2 - 1), l = pi & (half - 1);
                const int i = ((l & ~(j - 1)) << 1) | (l & (j - 1));
                float *a = keys + (size_t)arr * b_pad;
                const float x = a[i], y = a[i | j];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[i | j] = x; }
            }
            __syncthreads();
        }
    }

    for (int idx = tid; idx < NC * P; idx += 256) {
        const int cell = idx / P, p = idx - cell * P;
        if (c0 + cell >= C) continue;
        const int *cn = counts + cell * 3;
        out[((c0 + cell) * n_doy + d) * (int64_t)P + p] =
            quantile_from_sorted(keys + (size_t)cell * b_pad, b, qt.q[p], cn[0], cn[1], cn[2]);
    }
}

static bool bad_dims(int64_t C, int64_t T_b, int n_doy, int n_y, int W, int P)
{
    return C < 0 || T_b < 0 || n_doy <= 0 || n_y <= 0 || W <= 0 || P <= 0 || T_b > 0x3fffffff;
}

struct ThrLayout {
    size_t total = 0;
    float *xn = nullptr;
    int *time_index = nullptr, *win_rows = nullptr;
};

static ThrLayout carve_thr(void *ws, size_t ws_bytes, int64_t C, int64_t T_b, bool need_norm, int n_doy, int n_y, int W)
{
    ThrLayout L;
    Carver cv(ws, ws_bytes);
    if (need_norm) L.xn = cv.take<float>((size_t)C * T_b);
    L.time_index = cv.take<int>((size_t)n_doy * n_y);
    L.win_rows = cv.take<int>((size_t)n_doy * W);
    L.total = cv.off;
    return L;
}

}  // namespace hdp

using namespace hdp;

extern "C" {

size_t hdp_b200_thresholds_workspace_bytes(int64_t C, int64_t T_b, int64_t ld_t, int64_t ld_c,
                                           int n_doy, int n_y, int W, int P)
{
    (void)ld_t;
    if (bad_dims(C, T_b, n_doy, n_y, W, P)) return 0;
    return carve_thr(nullptr, 0, C, T_b, ld_c != 1, n_doy, n_y, W).total;
}

int hdp_b200_thresholds(const float *d_temps, int64_t C, int64_t T_b, int64_t ld_t, int64_t ld_c,
                        const int32_t *h_time_index, const int32_t *h_win_rows, int n_doy, int n_y, int W,
                        const double *h_q, int P, double *d_out,
                        void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (bad_dims(C, T_b, n_doy, n_y, W, P) || !h_time_index || !h_win_rows || !h_q) return HDP_B200_ERR_INVALID;
    if (C > 0 && (!d_temps || !d_out)) return HDP_B200_ERR_INVALID;
    if (C > 0 && T_b == 0) return HDP_B200_ERR_INVALID;                      // nothing to index into
    if (P > HDP_B200_MAX_PERCENTILES) return HDP_B200_ERR_UNSUPPORTED;
    const int64_t b = (int64_t)W * n_y;
    if (b > HDP_B200_MAX_WINDOW) return HDP_B200_ERR_UNSUPPORTED;
    QTable qt;
    for (int p = 0; p < HDP_B200_MAX_PERCENTILES; p++) {
        qt.q[p] = p < P ? h_q[p] : 0.0;
        if (p < P && !(h_q[p] >= 0.0 && h_q[p] <= 1.0)) return HDP_B200_ERR_INVALID;   // quantile_is_valid, arraymath.py:1747
    }
    for (int64_t i = 0; i < (int64_t)n_doy * n_y; i++)
        if (h_time_index[i] < -T_b || h_time_index[i] >= T_b) return HDP_B200_ERR_INVALID;
    for (int64_t i = 0; i < (int64_t)n_doy * W; i++)
        if (h_win_rows[i] < 0 || h_win_rows[i] >= n_doy) return HDP_B200_ERR_INVALID;
    if (C == 0) return HDP_B200_OK;

    cudaStream_t st = (cudaStream_t)stream;
    const bool need_norm = ld_c != 1;
    ThrLayout L = carve_thr(d_workspace, workspace_bytes, C, T_b, need_norm, n_doy, n_y, W);
    if (!d_workspace || L.total > workspace_bytes) return HDP_B200_ERR_WORKSPACE;
    const float *x = d_temps;
    if (need_norm) {
        int rc = normalize_layout(d_temps, C, T_b, ld_t, ld_c, L.xn, st);
        if (rc != HDP_B200_OK) return rc;
        x = L.xn;
        ld_t = C;
    }
    HDP_CUDA_TRY(cudaMemcpyAsync(L.time_index, h_time_index, sizeof(int) * (size_t)n_doy * n_y, cudaMemcpyHostToDevice, st));
    HDP_CUDA_TRY(cudaMemcpyAsync(L.win_rows, h_win_rows, sizeof(int) * (size_t)n_doy * W, cudaMemcpyHostToDevice, st));

    int b_pad_log2 = 1;
    while ((1 << b_pad_log2) < b) b_pad_log2++;
    const int b_pad = 1 << b_pad_log2;
    int NC = 16384 / b_pad;                                                  // <= 64 KB of keys per CTA
    if (NC < 1) NC = 1;
    if (NC > 8) NC = 8;
    const size_t smem = (size_t)NC * b_pad * sizeof(float) + (size_t)NC * 3 * sizeof(int);
    HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((C + NC - 1) / NC), (unsigned)n_doy);
    KernelTimer timer(kThrGeneric, st);
    k_thr_generic<<<grid, 256, smem, st>>>(x, C, T_b, ld_t, L.time_index, L.win_rows, n_doy, n_y, W, qt, P, NC, b_pad_log2, d_out);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

}  // extern "C"
