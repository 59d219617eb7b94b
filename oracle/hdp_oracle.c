/*
 * hdp_oracle.c - CPU restatement of the two HDP hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker and the reported CPU baseline.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product path (hdp_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
 * against (a) the known-answer vectors of the reference's own unit tests
 * (hdp/tests/test_index_heatwaves.py, test_heatwave_{frequency,number,duration,average}.py)
 * and (b) fixtures under tests/golden/ produced by running the UNMODIFIED reference Numba
 * kernels in the build container (tests/golden/make_golden.py).
 *
 * The percentile arithmetic of the reference lives in a third-party dependency that is
 * not vendored under /root/reference: Numba (pyproject.toml:22 pins numba>=0.60.0, no lock
 * file; the fixtures were produced with numba 0.65.0 / llvmlite 0.47.0).  Its published
 * algorithm is restated in quantile_row() from numba/np/arraymath.py:1655-1704 and
 * :1754-1768; the reference call site is hdp/threshold.py:78.
 *
 * All "file:line" citations are relative to the reference tree (AgentOxygen/HDP v1.0.2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HDP_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* Path 1: thresholds                                                                    */
/* ------------------------------------------------------------------------------------ */

/* Hoare-style quickselect: after the call a[k] holds the k-th smallest of a[lo..hi] and the
 * array is partitioned around it.  The result value depends only on the multiset, so any
 * correct selection reproduces Numba's _select/_select_two (arraymath.py:1568-1617). */
static void select_kth(double *a, int64_t lo, int64_t hi, int64_t k)
{
    while (lo < hi) {
        double pivot = a[lo + ((hi - lo) >> 1)];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (a[i] < pivot) i++;
            while (a[j] > pivot) j--;
            if (i <= j) {
                double t = a[i]; a[i] = a[j]; a[j] = t;
                i++; j--;
            }
        }
        if (k <= j) hi = j;
        else if (k >= i) lo = i;
        else return;
    }
}

/* numba/np/arraymath.py:1655-1704 (_collect_percentiles_inner) and :1754-1768
 * (_collect_percentiles with skip_nan=False, factor=100.0).  `a` is scratch of n doubles
 * (mutated).  Each floating-point operation below is a separately rounded IEEE double op;
 * build with -ffp-contract=off so the compiler does not fuse the interpolation. */
static void quantile_row(double *a, int64_t n, const double *q, int P, double *out)
{
    int64_t i;
    int p;
    int has_nan = 0;
    for (i = 0; i < n; i++) if (a[i] != a[i]) { has_nan = 1; break; }
    /* _can_collect_percentiles, arraymath.py:1707-1721 */
    if (has_nan || n == 0 || (n == 1 && !isfinite(a[0]))) {
        for (p = 0; p < P; p++) out[p] = NAN;
        return;
    }
    if (n == 1) {                                   /* arraymath.py:1661-1663 */
        for (p = 0; p < P; p++) out[p] = a[0];
        return;
    }
    for (p = 0; p < P; p++) {
        volatile double percentile = q[p] * 100.0;  /* arraymath.py:1757, q = q * factor */
        double val;
        if (percentile == 100.0) {                  /* arraymath.py:1669-1675 */
            int all_finite = 1;
            val = a[0];
            for (i = 0; i < n; i++) { if (a[i] > val) val = a[i]; if (!isfinite(a[i])) all_finite = 0; }
            if (!all_finite && !isfinite(val)) val = NAN;
        } else if (percentile == 0.0) {             /* arraymath.py:1678-1695 */
            int64_t num_pos_inf = 0, num_neg_inf = 0, num_finite;
            val = a[0];
            for (i = 0; i < n; i++) {
                if (a[i] < val) val = a[i];
                if (a[i] == INFINITY) num_pos_inf++;
                if (a[i] == -INFINITY) num_neg_inf++;
            }
            if (num_pos_inf + num_neg_inf > 0) {
                num_finite = n - (num_neg_inf + num_pos_inf);
                if (num_finite == 0) val = NAN;
                if (num_pos_inf == 1 && n == 2) val = NAN;
                if (num_neg_inf > 1) val = NAN;
                if (num_finite == 1 && num_pos_inf > 1 && num_neg_inf != 1) val = NAN;
            }
        } else {                                    /* arraymath.py:1697-1701 */
            volatile double frac = percentile / 100.0;
            volatile double scaled = (double)(n - 1) * frac;
            volatile double rank = 1.0 + scaled;
            double f = floor(rank);
            volatile double m = rank - f;
            int64_t k = (int64_t)(f - 1.0);
            double lower, upper;
            volatile double w0, t0, t1;
            if (k < 0) k = 0;
            if (k >= n - 1) {
                /* rank == n can only arise from rounding of a q just below 1; Numba would read one
                 * element past the end with weight m == 0.  We define upper = lower = max. */
                k = n - 1;
                select_kth(a, 0, n - 1, k);
                lower = upper = a[k];
            } else {
                select_kth(a, 0, n - 1, k);
                lower = a[k];
                upper = a[k + 1];
                for (i = k + 2; i < n; i++) if (a[i] < upper) upper = a[i];
            }
            w0 = 1.0 - m;
            t0 = lower * w0;
            t1 = upper * m;
            val = t0 + t1;
        }
        out[p] = val;
    }
}

/* hdp/threshold.py:52-78 (compute_percentiles gufunc core '(t),(d,b),(p)->(d,p)') for ONE cell.
 * temps: the cell's series, element t at temps[t*stride].  window_samples: int64[n_doy*b]
 * time indices; negative entries index from the end like NumPy/Numba (the -1 pads written by
 * datetimes_to_windows, threshold.py:35, read the last sample).  scratch: b doubles. */
static void percentiles_cell(const float *temps, int64_t T, int64_t stride,
                             const int64_t *window_samples, int64_t n_doy, int64_t b,
                             const double *q, int P, double *out, double *scratch)
{
    int64_t d, i;
    for (d = 0; d < n_doy; d++) {
        const int64_t *win = window_samples + d * b;
        for (i = 0; i < b; i++) {
            int64_t t = win[i];
            if (t < 0) t += T;
            scratch[i] = (double)temps[t * stride];   /* threshold.py:75-77, f32 -> f64 */
        }
        quantile_row(scratch, b, q, P, out + d * P);  /* threshold.py:78 */
    }
}

HDP_API int hdp_oracle_percentiles(const float *temps, int64_t T,
                                   const int64_t *window_samples, int64_t n_doy, int64_t b,
                                   const double *q, int P, double *out)
{
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(b > 0 ? b : 1));
    if (!scratch) return -1;
    percentiles_cell(temps, T, 1, window_samples, n_doy, b, q, P, out, scratch);
    free(scratch);
    return 0;
}

/* Batch driver: what xarray.apply_ufunc does around the gufunc (threshold.py:81-93) - loop
 * cells.  temps element (t, c) at temps[t*ld_t + c*ld_c]; out is [C, n_doy, P].
 * threads <= 0 uses every core (the reference's Dask LocalCluster stand-in). */
HDP_API int hdp_oracle_thresholds_batch(const float *temps, int64_t C, int64_t T, int64_t ld_t, int64_t ld_c,
                                        const int64_t *window_samples, int64_t n_doy, int64_t b,
                                        const double *q, int P, double *out, int threads)
{
    int64_t c;
    int failed = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
    {
        double *scratch = (double *)malloc(sizeof(double) * (size_t)(b > 0 ? b : 1));
        float *series = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
        if (!scratch || !series) {
            failed = 1;
        } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
            for (c = 0; c < C; c++) {
                int64_t t;
                for (t = 0; t < T; t++) series[t] = temps[t * ld_t + c * ld_c];
                percentiles_cell(series, T, 1, window_samples, n_doy, b, q, P, out + c * n_doy * P, scratch);
            }
        }
        free(scratch);
        free(series);
    }
    return failed ? -1 : 0;
}

/* ------------------------------------------------------------------------------------ */
/* Path 2: heatwave metrics                                                              */
/* ------------------------------------------------------------------------------------ */

/* hdp/metric.py:280-301.  f32 measure against f64 threshold: compared in double; NaN -> 0. */
HDP_API void hdp_oracle_indicate_hot_days(const float *measure, const double *threshold,
                                          const int64_t *doy_map, int64_t T, uint8_t *hot)
{
    int64_t t;
    for (t = 0; t < T; t++) hot[t] = ((double)measure[t] > threshold[doy_map[t]]) ? 1 : 0;
}

/* hdp/metric.py:11-60.  hw is int64[T]. */
HDP_API int hdp_oracle_index_heatwaves(const uint8_t *hot, int64_t T, int64_t min_duration,
                                       int64_t max_break, int64_t max_subs, int64_t *hw)
{
    /* ts = zeros(T+2); ts[i+1] = hot[i]; diff_ts = diff(ts) has T+1 entries (metric.py:27-31) */
    int64_t n_diff = T + 1, i, n_tr = 0, j;
    int8_t *diff_ts = (int8_t *)malloc((size_t)n_diff);
    int64_t *diff_indices = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_diff);
    int64_t *hw_indices = (int64_t *)calloc((size_t)n_diff, sizeof(int64_t));
    int in_heatwave = 0;
    int64_t current_hw_index = 0, sub_events = 0;
    if (!diff_ts || !diff_indices || !hw_indices) { free(diff_ts); free(diff_indices); free(hw_indices); return -1; }
    for (i = 0; i < n_diff; i++) {
        int prev = (i == 0) ? 0 : (hot[i - 1] ? 1 : 0);
        int cur = (i == T) ? 0 : (hot[i] ? 1 : 0);
        diff_ts[i] = (int8_t)(cur - prev);
        if (diff_ts[i] != 0) diff_indices[n_tr++] = i;       /* metric.py:32 */
    }
    for (i = 0; i + 1 < n_tr; i++) {                          /* metric.py:39 */
        int64_t index = diff_indices[i], next_index = diff_indices[i + 1];
        if (diff_ts[index] == 1 && next_index - index >= min_duration && !in_heatwave) {   /* :43 */
            current_hw_index += 1;
            in_heatwave = 1;
            for (j = index; j < next_index; j++) hw_indices[j] = current_hw_index;
        } else if (diff_ts[index] == -1 && next_index - index > max_break) {                /* :47 */
            in_heatwave = 0;
        } else if (diff_ts[index] == 1 && in_heatwave && sub_events < max_subs) {           /* :49 */
            sub_events += 1;
            for (j = index; j < next_index; j++) hw_indices[j] = current_hw_index;
        } else if (diff_ts[index] == 1 && in_heatwave && sub_events >= max_subs) {          /* :52 */
            if (next_index - index >= min_duration) {
                current_hw_index += 1;
                for (j = index; j < next_index; j++) hw_indices[j] = current_hw_index;
            } else {
                in_heatwave = 0;
            }
            sub_events = 0;                                                                 /* :58 */
        }
    }
    memcpy(hw, hw_indices, sizeof(int64_t) * (size_t)T);     /* metric.py:60 */
    free(diff_ts); free(diff_indices); free(hw_indices);
    return 0;
}

/* Python slice semantics for hw_ts[a:b] (metric.py:79,100,122,157). */
static void py_slice(int64_t a, int64_t b, int64_t T, int64_t *lo, int64_t *hi)
{
    if (a < 0) { a += T; if (a < 0) a = 0; }
    if (b < 0) { b += T; if (b < 0) b = 0; }
    if (a > T) a = T;
    if (b > T) b = T;
    if (b < a) b = a;
    *lo = a; *hi = b;
}

static int cmp_i64(const void *x, const void *y)
{
    int64_t a = *(const int64_t *)x, b = *(const int64_t *)y;
    return (a > b) - (a < b);
}

/* np.unique of a slice: sorted distinct values.  Returns count. */
static int64_t unique_sorted(const int64_t *v, int64_t n, int64_t *out)
{
    int64_t i, m = 0;
    if (n == 0) return 0;
    memcpy(out, v, sizeof(int64_t) * (size_t)n);
    qsort(out, (size_t)n, sizeof(int64_t), cmp_i64);
    for (i = 0; i < n; i++) if (m == 0 || out[i] != out[m - 1]) out[m++] = out[i];
    return m;
}

/* hdp/metric.py:85-102 */
HDP_API void hdp_oracle_heatwave_frequency(const int64_t *hw, int64_t T, const int64_t *ranges, int64_t Y, int64_t *out)
{
    int64_t y, t, lo, hi;
    for (y = 0; y < Y; y++) {
        int64_t s = 0;
        py_slice(ranges[2 * y], ranges[2 * y + 1], T, &lo, &hi);
        for (t = lo; t < hi; t++) s += hw[t] > 0;
        out[y] = s;
    }
}

/* hdp/metric.py:63-82 */
HDP_API int hdp_oracle_heatwave_number(const int64_t *hw, int64_t T, const int64_t *ranges, int64_t Y, int64_t *out)
{
    int64_t y, i, lo, hi;
    int64_t *u = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    if (!u) return -1;
    for (y = 0; y < Y; y++) {
        int64_t m, c = 0;
        py_slice(ranges[2 * y], ranges[2 * y + 1], T, &lo, &hi);
        m = unique_sorted(hw + lo, hi - lo, u);
        for (i = 0; i < m; i++) c += u[i] != 0;
        out[y] = c;
    }
    free(u);
    return 0;
}

/* Shared body of heatwave_duration (metric.py:105-137) and heatwave_average (:140-172):
 * lengths of every id present in the season slice, with the reference's handling of the
 * unique() result (drop the first unique value unless it is the only one, :124-128). */
static int64_t season_lengths(const int64_t *slice, int64_t n, int64_t *u, int64_t *lengths)
{
    int64_t m = unique_sorted(slice, n, u), i, t, first = 0;
    if (m != 1) first = 1;                 /* unique_indices = unique_indices[1:] */
    if (m == 0) return 0;                  /* empty slice: np.max/np.mean of empty - reference raises; we yield 0 */
    for (i = first; i < m; i++) {
        int64_t c = 0;
        if (u[i] != 0) for (t = 0; t < n; t++) c += slice[t] == u[i];
        lengths[i - first] = c;
    }
    return m - first;
}

HDP_API int hdp_oracle_heatwave_duration(const int64_t *hw, int64_t T, const int64_t *ranges, int64_t Y, int64_t *out)
{
    int64_t y, i, lo, hi;
    int64_t *u = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    int64_t *len = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    if (!u || !len) { free(u); free(len); return -1; }
    for (y = 0; y < Y; y++) {
        int64_t m, mx = 0;
        py_slice(ranges[2 * y], ranges[2 * y + 1], T, &lo, &hi);
        m = season_lengths(hw + lo, hi - lo, u, len);
        for (i = 0; i < m; i++) if (i == 0 || len[i] > mx) mx = len[i];
        out[y] = mx;                        /* np.max(hw_lengths), metric.py:136 */
    }
    free(u); free(len);
    return 0;
}

HDP_API int hdp_oracle_heatwave_average(const int64_t *hw, int64_t T, const int64_t *ranges, int64_t Y, double *out)
{
    int64_t y, i, lo, hi;
    int64_t *u = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    int64_t *len = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    if (!u || !len) { free(u); free(len); return -1; }
    for (y = 0; y < Y; y++) {
        int64_t m, s = 0;
        py_slice(ranges[2 * y], ranges[2 * y + 1], T, &lo, &hi);
        m = season_lengths(hw + lo, hi - lo, u, len);
        for (i = 0; i < m; i++) s += len[i];
        out[y] = m > 0 ? (double)s / (double)m : 0.0;   /* np.mean(hw_lengths), metric.py:171 */
    }
    free(u); free(len);
    return 0;
}

/* hdp/metric.py:304-341: out is int64[4*Y] in the order HWF, HWN, HWD, HWA; the float64
 * average is truncated toward zero when stored into the int64 array (:340). */
HDP_API int hdp_oracle_heatwave_metrics(const float *measure, const double *threshold, const int64_t *doy_map,
                                        int64_t T, int64_t min_duration, int64_t max_break, int64_t max_subs,
                                        const int64_t *season_ranges, int64_t Y, int64_t *out)
{
    uint8_t *hot = (uint8_t *)malloc((size_t)(T > 0 ? T : 1));
    int64_t *hw = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
    double *avg = (double *)malloc(sizeof(double) * (size_t)(Y > 0 ? Y : 1));
    int rc = -1;
    int64_t y;
    if (hot && hw && avg) {
        hdp_oracle_indicate_hot_days(measure, threshold, doy_map, T, hot);
        rc = hdp_oracle_index_heatwaves(hot, T, min_duration, max_break, max_subs, hw);
        if (rc == 0) {
            hdp_oracle_heatwave_frequency(hw, T, season_ranges, Y, out);
            rc = hdp_oracle_heatwave_number(hw, T, season_ranges, Y, out + Y);
        }
        if (rc == 0) rc = hdp_oracle_heatwave_duration(hw, T, season_ranges, Y, out + 2 * Y);
        if (rc == 0) rc = hdp_oracle_heatwave_average(hw, T, season_ranges, Y, avg);
        if (rc == 0) for (y = 0; y < Y; y++) out[3 * Y + y] = (int64_t)avg[y];
    }
    free(hot); free(hw); free(avg);
    return rc;
}

/* Batch driver: what compute_heatwave_metrics_wrapper does (metric.py:344-369) - for every
 * percentile, for every definition, for every cell call compute_heatwave_metrics.
 *   measure (t, c) at measure[t*ld_t + c*ld_c];  thresholds f64 [C, n_doy, P];
 *   defs int64 [D,3];  seasons_north/south int64 [Y,2];  is_south uint8 [C];
 *   out int64 [P, D, C, 4, Y]  (the reference's (percentile, definition, <cells>, metric, year)). */
HDP_API int hdp_oracle_metrics_batch(const float *measure, int64_t C, int64_t T, int64_t ld_t, int64_t ld_c,
                                     const double *thresholds, int64_t n_doy, int P,
                                     const int64_t *doy_map, const int64_t *defs, int D,
                                     const int64_t *seasons_north, const int64_t *seasons_south, int64_t Y,
                                     const uint8_t *is_south, int64_t *out, int threads)
{
    int64_t c;
    int failed = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
    {
        float *series = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
        double *thr = (double *)malloc(sizeof(double) * (size_t)(n_doy > 0 ? n_doy : 1));
        if (!series || !thr) {
            failed = 1;
        } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
            for (c = 0; c < C; c++) {
                int64_t t, d;
                int p, k;
                const int64_t *seasons = is_south[c] ? seasons_south : seasons_north;
                for (t = 0; t < T; t++) series[t] = measure[t * ld_t + c * ld_c];
                for (p = 0; p < P; p++) {
                    for (d = 0; d < n_doy; d++) thr[d] = thresholds[(c * n_doy + d) * P + p];
                    for (k = 0; k < D; k++) {
                        int64_t *o = out + ((((int64_t)p * D + k) * C + c) * 4) * Y;
                        if (hdp_oracle_heatwave_metrics(series, thr, doy_map, T, defs[3 * k], defs[3 * k + 1],
                                                        defs[3 * k + 2], seasons, Y, o) != 0)
                            failed = 1;
                    }
                }
            }
        }
        free(series);
        free(thr);
    }
    return failed ? -1 : 0;
}

HDP_API int hdp_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
