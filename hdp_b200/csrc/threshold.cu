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
#include <algorithm>

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

// ----------------------------------------------------------------------------------------------------
// k_thr_ranked: the fast path.  One CTA per cell.
//
//   1. gather the cell's E = n_doy * n_y window elements (every slot of the reference's time_index table,
//      -1 pads included: they are ordinary elements holding the last sample) into shared memory;
//   2. order them ONCE with a stable LSD radix sort (4-bit digits, per-thread private counters) on the
//      order-preserving integer image of the f32 samples, and record rank_of[element];
//   3. every warp owns a contiguous range of days of year and keeps that window as a bitmap over RANKS
//      (plane A: row present, plane B: row present twice - the reference's mirrored year-end wrap pools
//      some rows twice).  Moving to the next day of year flips the bits of the rows that leave / enter;
//   4. every requested percentile is read from that single ordering: the k-th and (k+1)-th window members
//      are found by a popcount prefix scan across lanes and an in-word select, and interpolated in double.
// ----------------------------------------------------------------------------------------------------
constexpr int kRankedThreads = 1024;
constexpr int kRankedWarps = kRankedThreads / 32;
constexpr int kRadixBits = 4, kRadixBins = 1 << kRadixBits, kRadixPasses = 32 / kRadixBits;

enum SelMode { kSelInterp = 0, kSelMax = 1, kSelMin = 2 };

struct SelTable {                    // per percentile, identical for every cell and day of year (n is fixed)
    int pos_lo[HDP_B200_MAX_PERCENTILES];     // 0-based positions in the sorted window
    int pos_hi[HDP_B200_MAX_PERCENTILES];
    int mode[HDP_B200_MAX_PERCENTILES];
    double w_lo[HDP_B200_MAX_PERCENTILES];    // 1 - m
    double w_hi[HDP_B200_MAX_PERCENTILES];    // m
    int8_t b_slot[64];                        // per warp: index of its plane B (rows pooled twice), -1 = none needed
};

__device__ __forceinline__ uint32_t f32_to_key(float v)
{
    const uint32_t u = __float_as_uint(v);
    if (v != v) return 0xffffffffu;                               // every NaN sorts last
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float key_to_f32(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// position (0..31) of the n-th (0-based) member of the multiset {bit i of a} + {bit i of b} in rank order
template <bool kTwoPlanes>
__device__ __forceinline__ int select_in_word(uint32_t a, uint32_t b, int n)
{
    int pos = 0;
#pragma unroll
    for (int width = 16; width >= 1; width >>= 1) {
        const uint32_t mask = (1u << width) - 1u;
        int c = __popc(a & mask);
        if (kTwoPlanes) c += __popc(b & mask);
        if (n >= c) { n -= c; pos += width; a >>= width; if (kTwoPlanes) b >>= width; }
    }
    return pos;
}

// members of the window whose rank lies in [lo, hi) (warp-cooperative; non-finite bookkeeping only)
__device__ __forceinline__ int range_count(const uint32_t *A, const uint32_t *B, bool dup, int wpl, int lane, int lo, int hi)
{
    int c = 0;
    for (int i = 0; i < wpl; i++) {
        const int w = lane * wpl + i, r0 = w * 32;
        if (r0 + 32 <= lo || r0 >= hi) continue;
        uint32_t m = 0xffffffffu;
        if (lo > r0) m &= 0xffffffffu << (lo - r0);
        if (hi < r0 + 32) m &= (1u << (hi - r0)) - 1u;
        c += __popc(A[w] & m) + (dup ? __popc(B[w] & m) : 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

__global__ void __launch_bounds__(kRankedThreads, 1)
k_thr_ranked(const float *__restrict__ temps, int64_t T_b, int64_t ld_t,
             const int *__restrict__ time_index, int E, int n_y, int n_doy, int n,
             const int *__restrict__ op_off, const int *__restrict__ ops, const uint8_t *__restrict__ doy_dup,
             int dpw, int ept, int nwords_pad, const __grid_constant__ SelTable sel, int P, double *__restrict__ out,
             const int *__restrict__ cell_count, const int *__restrict__ cell_list)
{
    constexpr int NT = kRankedThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Epad = (E + 63) & ~63;
    uint32_t *keyA = (uint32_t *)smem_raw;                        // sorted keys after the last pass
    uint16_t *rank_of = (uint16_t *)(keyA + Epad);                // also the second index buffer of the sort
    uint32_t *keyB = (uint32_t *)(rank_of + Epad);
    uint16_t *idxA = (uint16_t *)(keyB + Epad);
    uint16_t *cnt = idxA + Epad;                                  // [16][NT] private digit counters
    uint32_t *planes = keyB;                                      // after the sort: [warp][2][nwords_pad] (over keyB, idxA, cnt)
    __shared__ int s_nonfinite[3];                                // NaN, +inf, -inf elements of this cell
    __shared__ int s_warp_tot[kRankedWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // cells: blockIdx.x itself, or (as the hand-over target of k_thr_cell) the entries of a device-side list
    const int n_cells = cell_list ? *cell_count : (int)gridDim.x;
    for (int ci = blockIdx.x; ci < n_cells; ci += gridDim.x) {
    const int64_t c = cell_list ? cell_list[ci] : ci;

    __syncthreads();
    if (tid < 3) s_nonfinite[tid] = 0;
    __syncthreads();

    // ---- 1. gather ----
    const float pinf = __int_as_float(0x7f800000);
    for (int e = tid; e < E; e += NT) {
        int64_t t = time_index[e];
        if (t < 0) t += T_b;                                      // -1 pads read the LAST sample (threshold.py:35,77)
        const float v = temps[t * ld_t + c];
        if (v != v) atomicAdd(&s_nonfinite[0], 1);
        else if (v == pinf) atomicAdd(&s_nonfinite[1], 1);
        else if (v == -pinf) atomicAdd(&s_nonfinite[2], 1);
        keyA[e] = f32_to_key(v);
        idxA[e] = (uint16_t)e;
    }
    __syncthreads();

    // ---- 2. stable LSD radix sort of (key, element) ----
    {
        uint32_t *sk = keyA, *dk = keyB;
        uint16_t *si = idxA, *di = rank_of;
        const int e0 = min(tid * ept, E), e1 = min(e0 + ept, E);
        for (int pass = 0; pass < kRadixPasses; pass++) {
            const int shift = pass * kRadixBits;
#pragma unroll
            for (int d = 0; d < kRadixBins; d++) cnt[d * NT + tid] = 0;
            for (int e = e0; e < e1; e++) cnt[((sk[e] >> shift) & (kRadixBins - 1)) * NT + tid]++;
            __syncthreads();
            // exclusive scan of the 16*NT counters in (digit, thread) order; thread t owns entries [16t, 16t+16)
            uint16_t *mine = cnt + tid * kRadixBins;
            int local[kRadixBins], tot = 0;
#pragma unroll
            for (int i = 0; i < kRadixBins; i++) { local[i] = tot; tot += mine[i]; }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            int base = incl - tot;
            for (int w = 0; w < warp; w++) base += s_warp_tot[w];
#pragma unroll
            for (int i = 0; i < kRadixBins; i++) mine[i] = (uint16_t)(base + local[i]);
            __syncthreads();
            for (int e = e0; e < e1; e++) {
                const uint32_t k = sk[e];
                const int slot = ((k >> shift) & (kRadixBins - 1)) * NT + tid;
                const int pos = cnt[slot];
                cnt[slot] = (uint16_t)(pos + 1);
                dk[pos] = k;
                di[pos] = si[e];
            }
            __syncthreads();
            uint32_t *tk = sk; sk = dk; dk = tk;
            uint16_t *ti = si; si = di; di = ti;
        }
        // an even number of passes: sorted keys are back in keyA, their elements in idxA
    }
    for (int r = tid; r < E; r += NT) rank_of[idxA[r]] = (uint16_t)r;
    __syncthreads();

    // ---- 3./4. sliding rank bitmaps, one day-of-year range per warp ----
    const int wpl = nwords_pad >> 5;                              // bitmap words per lane
    // layout: [warp] plane A | [warp] members-in-front-of-word (u16) | plane B of the few warps that need one
    uint32_t *A = planes + (size_t)warp * nwords_pad;
    uint16_t *pre = (uint16_t *)(planes + (size_t)kRankedWarps * nwords_pad) + (size_t)warp * nwords_pad;
    const bool range_dup = sel.b_slot[warp] >= 0;
    uint32_t *B = planes + (size_t)kRankedWarps * nwords_pad * 3 / 2 + (size_t)(range_dup ? sel.b_slot[warp] : 0) * nwords_pad;
    for (int i = lane; i < nwords_pad; i += 32) { A[i] = 0u; if (range_dup) B[i] = 0u; }
    const int d_begin = warp * dpw, d_end = min(n_doy, d_begin + dpw);
    const int n_nan = s_nonfinite[0], n_pinf = s_nonfinite[1], n_ninf = s_nonfinite[2];
    const bool nonfinite = (n_nan | n_pinf | n_ninf) != 0;
    __syncwarp();

    for (int d = d_begin; d < d_end; d++) {
        // rows leaving / entering the window (multiset difference to the previous day; full build on the first)
        for (int o = op_off[d]; o < op_off[d + 1]; o++) {
            const int op = ops[o], row = op >> 1;
            for (int j = lane; j < n_y; j += 32) {
                const int r = rank_of[row * n_y + j];
                const uint32_t bit = 1u << (r & 31);
                if (op & 1) {
                    const uint32_t old = atomicOr(&A[r >> 5], bit);
                    if (range_dup && (old & bit)) atomicOr(&B[r >> 5], bit);
                } else {
                    if (range_dup && (B[r >> 5] & bit)) atomicAnd(&B[r >> 5], ~bit);
                    else atomicAnd(&A[r >> 5], ~bit);
                }
            }
            __syncwarp();
        }
        const bool dup = doy_dup[d] != 0;

        // members per lane slice (and the running count in front of every word), inclusive scan across lanes
        int s = 0;
        if (!dup) {
#pragma unroll 4
            for (int i = 0; i < wpl; i++) { pre[lane * wpl + i] = (uint16_t)s; s += __popc(A[lane * wpl + i]); }
        } else {
            for (int i = 0; i < wpl; i++) { pre[lane * wpl + i] = (uint16_t)s; s += __popc(A[lane * wpl + i]) + __popc(B[lane * wpl + i]); }
        }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncwarp();

        int w_nan = 0, w_pinf = 0, w_ninf = 0;
        if (nonfinite) {                                          // rare: count the window's non-finite members by rank range
            w_ninf = range_count(A, B, dup, wpl, lane, 0, n_ninf);
            w_pinf = range_count(A, B, dup, wpl, lane, E - n_nan - n_pinf, E - n_nan);
            w_nan = range_count(A, B, dup, wpl, lane, E - n_nan, E);
        }

        for (int p0 = 0; p0 < P; p0 += 16) {
            // lane 2i -> lower pick of percentile p0+i, lane 2i+1 -> upper pick
            const int p = min(p0 + (lane >> 1), P - 1);
            const int target = (lane & 1) ? sel.pos_hi[p] : sel.pos_lo[p];
            int lo = 0, hi = 31;                                  // first lane whose inclusive count exceeds target
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi) >> 1;
                const int v = __shfl_sync(0xffffffffu, incl, mid);
                if (v > target) hi = mid; else lo = mid + 1;
            }
            const int owner = lo;
            int rem = target - (__shfl_sync(0xffffffffu, incl, owner) - __shfl_sync(0xffffffffu, s, owner));
            // last word of the owner's slice whose running count is <= rem (binary search over <= 16 words)
            const uint16_t *pw = pre + owner * wpl;
            int wl = 0, wh = wpl - 1;
            while (wl < wh) {
                const int mid = (wl + wh + 1) >> 1;
                if ((int)pw[mid] <= rem) wl = mid; else wh = mid - 1;
            }
            const int w = owner * wpl + wl;
            rem -= pw[wl];
            const uint32_t a = A[w], b = dup ? B[w] : 0u;
            const int bitpos = dup ? select_in_word<true>(a, b, rem) : select_in_word<false>(a, 0u, rem);
            const int r = min(w * 32 + bitpos, E - 1);
            const double val = (double)key_to_f32(keyA[r]);
            const double lower = __shfl_sync(0xffffffffu, val, (lane & 15) * 2);
            const double upper = __shfl_sync(0xffffffffu, val, (lane & 15) * 2 + 1);
            if (lane < 16 && p0 + lane < P) {
                const int pp = p0 + lane, mode = sel.mode[pp];
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                double v;
                if (mode == kSelInterp) {                         // arraymath.py:1697-1701
                    v = __dadd_rn(__dmul_rn(lower, sel.w_lo[pp]), __dmul_rn(upper, sel.w_hi[pp]));
                } else if (mode == kSelMax) {                     // arraymath.py:1669-1675
                    v = upper;
                    if ((w_pinf | w_ninf) && isinf(v)) v = nan;
                } else {                                          // arraymath.py:1678-1695
                    v = lower;
                    if (w_pinf | w_ninf) {
                        const int n_fin = n - (w_pinf + w_ninf);
                        if (n_fin == 0) v = nan;
                        if (w_pinf == 1 && n == 2) v = nan;
                        if (w_ninf > 1) v = nan;
                        if (n_fin == 1 && w_pinf > 1 && w_ninf != 1) v = nan;
                    }
                }
                if (w_nan > 0) v = nan;                           // _can_collect_percentiles, arraymath.py:1714
                out[(c * n_doy + d) * (int64_t)P + pp] = v;
            }
        }
    }
}   // cells
}

// ----------------------------------------------------------------------------------------------------
// k_thr_cell: the current fast path.  One CTA per cell, same plan as k_thr_ranked (order the cell's samples once,
// slide a rank bitmap over the days of year) with a cheaper ordering stage:
//
//   1. gather the E = n_doy * n_y elements into shared memory (batched independent loads), find the finite
//      min / max and count NaN / +inf / -inf;
//   2. quantise every sample to a MONOTONE 16-bit bucket  b = 1 + trunc((v - vmin) * 65532 / (vmax - vmin))
//      (float subtraction, multiplication and truncation are all monotone, so bucket order never contradicts
//      value order; -inf -> 0, +inf -> 65534, NaN -> 65535) and pack (bucket << 16 | element) into ONE word;
//   3. stable LSD radix sort of the packed words on the bucket: 4 passes of 4 bits instead of 8 passes over
//      (key, element) pairs.  Digit counters are private per thread, two digits per 32-bit word, in an
//      XOR-swizzled layout so that both the per-thread updates and the 64-byte-per-thread scan are free of
//      bank conflicts;
//   4. samples that share a bucket are adjacent now; each run is put in true order by its first owner thread
//      (insertion sort on the exact float keys; runs are 1-3 long for real temperature data).  A cell whose
//      samples pile into one bucket (an outlier stretching the range) is handed to k_thr_ranked through a
//      device-side list instead;
//   5./6. rank bitmaps and percentile selection exactly as in k_thr_ranked.
// ----------------------------------------------------------------------------------------------------
// 32-bit shared-state-space accesses: no generic -> shared conversion in front of every atomic
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_and(uint32_t a, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_or(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

constexpr int kCellThreads = 1024;
constexpr int kCellWarps = kCellThreads / 32;
constexpr int kMaxTieRun = 48;

__global__ void __launch_bounds__(kCellThreads, 1)
k_thr_cell(const float *__restrict__ temps, int64_t T_b, int64_t ld_t,
           const int *__restrict__ time_index, int E, int n_y, int n_doy, int n,
           const int *__restrict__ op_off, const int *__restrict__ ops, const uint8_t *__restrict__ doy_dup,
           int n_ops, int dpw, int ept, int nwords_pad, const __grid_constant__ SelTable sel, int P, double *__restrict__ out,
           int *__restrict__ fallback_count, int *__restrict__ fallback_cells)
{
    constexpr int NT = kCellThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Epad = (E + 63) & ~63;
    float *x = (float *)smem_raw;                                 // samples by element
    uint32_t *PA = (uint32_t *)(x + Epad);                        // packed (bucket << 16 | element), sorted at the end
    uint32_t *PB = PA + Epad;                                     // second sort buffer; sorted VALUES afterwards
    uint32_t *Wc = PB + Epad;                                     // [8][NT] digit counters (two digits per word), swizzled
    float *V = (float *)PB;
    uint16_t *rank_of = (uint16_t *)Wc;                           // after the sort
    uint32_t *planes = (uint32_t *)x;                             // after the sort: bitmaps over x + PA
    __shared__ int s_nonfinite[3];                                // NaN, +inf, -inf elements of this cell
    __shared__ float s_min[kCellWarps], s_max[kCellWarps];
    __shared__ uint32_t s_tot[kCellWarps];
    __shared__ int s_fallback;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c = blockIdx.x;

    if (tid < 3) s_nonfinite[tid] = 0;
    if (tid == 0) s_fallback = 0;
    __syncthreads();

    // ---- 1. gather (loads issued four at a time, independent of each other) ----
    const float pinf = __int_as_float(0x7f800000);
    float vmin = pinf, vmax = -pinf;
    int c_nan = 0, c_pinf = 0, c_ninf = 0;
    for (int e4 = tid; e4 < E; e4 += 4 * NT) {
        int64_t t[4];
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e4 + i * NT;
            t[i] = e < E ? time_index[e] : 0;
            if (t[i] < 0) t[i] += T_b;                            // -1 pads read the LAST sample (threshold.py:35,77)
        }
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = temps[t[i] * ld_t + c];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e4 + i * NT;
            if (e < E) {
                x[e] = v[i];
                if (v[i] != v[i]) c_nan++;
                else if (v[i] == pinf) c_pinf++;
                else if (v[i] == -pinf) c_ninf++;
                else { vmin = fminf(vmin, v[i]); vmax = fmaxf(vmax, v[i]); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    c_nan = __reduce_add_sync(0xffffffffu, c_nan);
    c_pinf = __reduce_add_sync(0xffffffffu, c_pinf);
    c_ninf = __reduce_add_sync(0xffffffffu, c_ninf);
    if (lane == 0) {
        s_min[warp] = vmin; s_max[warp] = vmax;
        if (c_nan) atomicAdd(&s_nonfinite[0], c_nan);
        if (c_pinf) atomicAdd(&s_nonfinite[1], c_pinf);
        if (c_ninf) atomicAdd(&s_nonfinite[2], c_ninf);
    }
    __syncthreads();
    vmin = s_min[lane]; vmax = s_max[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }

    // ---- 2. monotone 16-bit buckets, packed with the element index ----
    const float range = vmax - vmin;
    const float scale = (range > 0.0f && range < pinf) ? 65532.0f / range : 0.0f;
    for (int e = tid; e < E; e += NT) {
        const float v = x[e];
        uint32_t b;
        if (v != v) b = 65535u;
        else if (v == pinf) b = 65534u;
        else if (v == -pinf) b = 0u;
        else b = 1u + (uint32_t)min(65532, max(0, __float2int_rz((v - vmin) * scale)));
        PA[e] = (b << 16) | (uint32_t)e;
    }
    __syncthreads();

    // ---- 3. stable LSD radix sort on the bucket: 4 passes x 4 bits ----
    {
        uint32_t *src = PA, *dst = PB;
        const int e0 = min(tid * ept, E), e1 = min(e0 + ept, E);
        // word of digit pair k of this thread: k * NT + tsw (16-byte pieces XOR-swizzled inside groups of 8)
        const int tsw = ((((tid >> 2) ^ (warp & 7)) << 2) | (tid & 3));
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 16 + 4 * pass;
#pragma unroll
            for (int k = 0; k < 8; k++) Wc[k * NT + tsw] = 0u;
            // (no barrier needed: every thread only touches its own 8 words until the scan)
            for (int e = e0; e < e1; e++) {
                const uint32_t d = (src[e] >> shift) & 15u;
                Wc[(d >> 1) * NT + tsw] += 1u << ((d & 1u) << 4);
            }
            __syncthreads();
            // exclusive scan in (digit, thread) order.  Thread u < 512 owns 16 consecutive words of one digit pair
            // (both halves): chunk u -> pair k = u >> 6, threads 16 * (u & 63) ... + 15.
            uint32_t ex[16], run = 0u, incl = 0u;
            uint4 *W4 = reinterpret_cast<uint4 *>(Wc);
            const int sw = (tid >> 1) & 7;
            if (tid < 512) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint4 q = W4[(4 * tid + j) ^ sw];
                    ex[4 * j + 0] = q.x; ex[4 * j + 1] = q.y; ex[4 * j + 2] = q.z; ex[4 * j + 3] = q.w;
                }
#pragma unroll
                for (int j = 0; j < 16; j++) { const uint32_t w = ex[j]; ex[j] = run; run += w; }   // packed lo | hi << 16
                incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (lane == 31) s_tot[warp] = incl;
            }
            __syncthreads();
            if (tid < 512) {
                const int g0 = warp & ~1;                         // first warp of this digit pair
                const uint32_t t = lane < 16 ? s_tot[lane] : 0u;
                const uint32_t before = __reduce_add_sync(0xffffffffu, lane < g0 ? t : 0u);
                const uint32_t t0 = __shfl_sync(0xffffffffu, t, g0), t1 = __shfl_sync(0xffffffffu, t, g0 + 1);
                const uint32_t before_all = (before & 0xffffu) + (before >> 16);
                const uint32_t grp = warp == g0 ? 0u : t0;        // warps of this pair in front of this one
                const uint32_t exw = incl - run;                  // threads of this warp in front of this one
                const uint32_t base_lo = before_all + (grp & 0xffffu) + (exw & 0xffffu);
                const uint32_t base_hi = before_all + ((t0 + t1) & 0xffffu) + (grp >> 16) + (exw >> 16);
#pragma unroll
                for (int j = 0; j < 16; j++) ex[j] = (base_lo + (ex[j] & 0xffffu)) | ((base_hi + (ex[j] >> 16)) << 16);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    W4[(4 * tid + j) ^ sw] = make_uint4(ex[4 * j + 0], ex[4 * j + 1], ex[4 * j + 2], ex[4 * j + 3]);
            }
            __syncthreads();
            for (int e = e0; e < e1; e++) {
                const uint32_t w = src[e];
                const uint32_t d = (w >> shift) & 15u, sh = (d & 1u) << 4;
                uint32_t *cw = &Wc[(d >> 1) * NT + tsw];
                const uint32_t cv = *cw;
                *cw = cv + (1u << sh);
                dst[(cv >> sh) & 0xffffu] = w;
            }
            __syncthreads();
            uint32_t *tp = src; src = dst; dst = tp;
        }
        // four passes: the sorted words are back in PA

        // ---- 4. exact order inside runs of equal bucket ----
        if (e0 < e1) {
            uint32_t prev_b = e0 > 0 ? PA[e0 - 1] >> 16 : 0xffffffffu;
            uint32_t w_cur = PA[e0];
            for (int i = e0; i < e1; i++) {
                const uint32_t b = w_cur >> 16;
                const uint32_t w_next = i + 1 < E ? PA[i + 1] : 0xffffffffu;
                const bool start = b != prev_b;
                prev_b = b;
                w_cur = w_next;
                if (!start || (w_next >> 16) != b || i + 1 >= E) continue;   // not a run start, or a run of one
                if (b == 0u || b >= 65534u) continue;             // -inf / +inf / NaN: identical keys
                int j = i + 2;
                while (j < E && (PA[j] >> 16) == b) j++;
                if (j - i > kMaxTieRun) {                         // long run: fine if already ordered (ties), else hand over
                    bool ordered = true;
                    uint32_t kp = f32_to_key(x[PA[i] & 0xffffu]);
                    for (int a = i + 1; a < j && ordered; a++) {
                        const uint32_t ka = f32_to_key(x[PA[a] & 0xffffu]);
                        ordered = ka >= kp;
                        kp = ka;
                    }
                    if (!ordered) s_fallback = 1;
                    continue;
                }
                for (int a = i + 1; a < j; a++) {
                    const uint32_t wa = PA[a];
                    const uint32_t ka = f32_to_key(x[wa & 0xffffu]);
                    int bpos = a;
                    while (bpos > i) {
                        const uint32_t wb = PA[bpos - 1];
                        if (f32_to_key(x[wb & 0xffffu]) <= ka) break;
                        PA[bpos] = wb;
                        bpos--;
                    }
                    PA[bpos] = wa;
                }
                if (i + 1 < e1) w_cur = PA[i + 1];                // the run was permuted: re-read the next word
            }
        }
    }
    __syncthreads();
    if (s_fallback) {                                             // block-uniform
        if (tid == 0) fallback_cells[atomicAdd(fallback_count, 1)] = (int)c;
        return;
    }
    uint16_t *s_opoff = rank_of + Epad;                           // [n_doy + 1] op offsets, then the ops (u16: row << 1 | enter)
    uint16_t *s_ops = s_opoff + ((n_doy + 2) & ~1);
    for (int r = tid; r < E; r += NT) {
        const uint32_t idx = PA[r] & 0xffffu;
        V[r] = x[idx];
        rank_of[idx] = (uint16_t)r;
    }
    for (int i = tid; i <= n_doy; i += NT) s_opoff[i] = (uint16_t)op_off[i];
    for (int i = tid; i < n_ops; i += NT) s_ops[i] = (uint16_t)ops[i];
    __syncthreads();

    // ---- 5./6. sliding rank bitmaps, one day-of-year range per warp (algorithm of k_thr_ranked; the tables live
    //      in shared memory, shared-space addresses are 32-bit, per-lane selection constants are hoisted) ----
    const int wpl = nwords_pad >> 5;
    const bool range_dup = sel.b_slot[warp] >= 0;
    const uint32_t sA = smem_u32(planes + (size_t)warp * nwords_pad);
    const uint32_t sPre = smem_u32((uint16_t *)(planes + (size_t)kCellWarps * nwords_pad) + (size_t)warp * nwords_pad);
    const uint32_t sB = smem_u32(planes + (size_t)kCellWarps * nwords_pad * 3 / 2 + (size_t)(range_dup ? sel.b_slot[warp] : 0) * nwords_pad);
    const uint32_t sRank = smem_u32(rank_of);
    for (int i = lane; i < nwords_pad; i += 32) { sts_u32(sA + 4 * i, 0u); if (range_dup) sts_u32(sB + 4 * i, 0u); }
    const int d_begin = warp * dpw, d_end = min(n_doy, d_begin + dpw);
    const int n_nan = s_nonfinite[0], n_pinf = s_nonfinite[1], n_ninf = s_nonfinite[2];
    const bool nonfinite = (n_nan | n_pinf | n_ninf) != 0;
    const uint32_t *A = planes + (size_t)warp * nwords_pad;       // generic views for the rare non-finite bookkeeping
    const uint32_t *B = planes + (size_t)kCellWarps * nwords_pad * 3 / 2 + (size_t)(range_dup ? sel.b_slot[warp] : 0) * nwords_pad;

    // lane 2i -> lower pick of percentile i, lane 2i+1 -> upper pick; lanes 0..15 finish percentile `lane`
    const int n_rounds = (P + 15) >> 4;
    int tgt[2], md[2];
    double w_lo[2], w_hi[2];
#pragma unroll
    for (int rnd = 0; rnd < 2; rnd++) {
        const int p = min(rnd * 16 + (lane >> 1), P - 1), pp = min(rnd * 16 + (lane & 15), P - 1);
        tgt[rnd] = (lane & 1) ? sel.pos_hi[p] : sel.pos_lo[p];
        md[rnd] = sel.mode[pp]; w_lo[rnd] = sel.w_lo[pp]; w_hi[rnd] = sel.w_hi[pp];
    }
    __syncwarp();

    int o1 = d_begin < d_end ? s_opoff[d_begin] : 0;
    for (int d = d_begin; d < d_end; d++) {
        // rows leaving / entering the window (multiset difference to the previous day; full build on the first)
        const int o0 = o1;
        o1 = s_opoff[d + 1];
        for (int o = o0; o < o1; o++) {
            const uint32_t op = s_ops[o], row = op >> 1;
            for (int j = lane; j < n_y; j += 32) {
                const uint32_t r = lds_u16(sRank + 2u * (row * n_y + j));
                const uint32_t bit = 1u << (r & 31u), wo = (r >> 5) << 2;
                if (!range_dup) {
                    if (op & 1u) reds_or(sA + wo, bit); else reds_and(sA + wo, ~bit);
                } else if (op & 1u) {
                    if (atoms_or(sA + wo, bit) & bit) reds_or(sB + wo, bit);
                } else {
                    if (lds_u32(sB + wo) & bit) reds_and(sB + wo, ~bit); else reds_and(sA + wo, ~bit);
                }
            }
            __syncwarp();
        }
        const bool dup = range_dup && doy_dup[d] != 0;

        // members per lane slice (and the running count in front of every word), inclusive scan across lanes
        int s = 0;
        {
            const uint32_t a0 = sA + 4u * (lane * wpl), b0 = sB + 4u * (lane * wpl), p0a = sPre + 2u * (lane * wpl);
            if (!dup) {
#pragma unroll 11
                for (int i = 0; i < wpl; i++) { sts_u16(p0a + 2 * i, (uint32_t)s); s += __popc(lds_u32(a0 + 4 * i)); }
            } else {
                for (int i = 0; i < wpl; i++) { sts_u16(p0a + 2 * i, (uint32_t)s); s += __popc(lds_u32(a0 + 4 * i)) + __popc(lds_u32(b0 + 4 * i)); }
            }
        }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        __syncwarp();

        int w_nan = 0, w_pinf = 0, w_ninf = 0;
        if (nonfinite) {                                          // rare: count the window's non-finite members by rank range
            w_ninf = range_count(A, B, dup, wpl, lane, 0, n_ninf);
            w_pinf = range_count(A, B, dup, wpl, lane, E - n_nan - n_pinf, E - n_nan);
            w_nan = range_count(A, B, dup, wpl, lane, E - n_nan, E);
        }

        for (int rnd = 0; rnd < n_rounds; rnd++) {
            const int target = rnd ? tgt[1] : tgt[0];
            int lo = 0, hi = 31;                                  // first lane whose inclusive count exceeds target
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi) >> 1;
                const int v = __shfl_sync(0xffffffffu, incl, mid);
                if (v > target) hi = mid; else lo = mid + 1;
            }
            const int owner = lo;
            int rem = target - (__shfl_sync(0xffffffffu, incl, owner) - __shfl_sync(0xffffffffu, s, owner));
            // last word of the owner's slice whose running count is <= rem
            const uint32_t pw = sPre + 2u * (owner * wpl);
            int wl = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cand = wl + step;
                if (step < 2 * wpl && cand < wpl && (int)lds_u16(pw + 2 * cand) <= rem) wl = cand;
            }
            const int w = owner * wpl + wl;
            rem -= (int)lds_u16(pw + 2 * wl);
            const uint32_t a = lds_u32(sA + 4u * w), b = dup ? lds_u32(sB + 4u * w) : 0u;
            const int bitpos = dup ? select_in_word<true>(a, b, rem) : select_in_word<false>(a, 0u, rem);
            const int r = min(w * 32 + bitpos, E - 1);
            const float valf = V[r];
            const double lower = (double)__shfl_sync(0xffffffffu, valf, (lane & 15) * 2);
            const double upper = (double)__shfl_sync(0xffffffffu, valf, (lane & 15) * 2 + 1);
            if (lane < 16 && rnd * 16 + lane < P) {
                const int pp = rnd * 16 + lane, mode = rnd ? md[1] : md[0];
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                double v;
                if (mode == kSelInterp) {                         // arraymath.py:1697-1701
                    v = __dadd_rn(__dmul_rn(lower, rnd ? w_lo[1] : w_lo[0]), __dmul_rn(upper, rnd ? w_hi[1] : w_hi[0]));
                } else if (mode == kSelMax) {                     // arraymath.py:1669-1675
                    v = upper;
                    if ((w_pinf | w_ninf) && isinf(v)) v = nan;
                } else {                                          // arraymath.py:1678-1695
                    v = lower;
                    if (w_pinf | w_ninf) {
                        const int n_fin = n - (w_pinf + w_ninf);
                        if (n_fin == 0) v = nan;
                        if (w_pinf == 1 && n == 2) v = nan;
                        if (w_ninf > 1) v = nan;
                        if (n_fin == 1 && w_pinf > 1 && w_ninf != 1) v = nan;
                    }
                }
                if (w_nan > 0) v = nan;                           // _can_collect_percentiles, arraymath.py:1714
                out[(c * n_doy + d) * (int64_t)P + pp] = v;
            }
        }
    }
}

static bool bad_dims(int64_t C, int64_t T_b, int n_doy, int n_y, int W, int P)
{
    return C < 0 || T_b < 0 || n_doy <= 0 || n_y <= 0 || W <= 0 || P <= 0 || T_b > 0x3fffffff;
}

// What the fast path needs besides the reference's two tables.
struct RankedPlan {
    bool usable = false;
    std::vector<int> op_off, ops;       // per day of year: (row << 1 | enter) ops relative to the previous day of the warp's range
    std::vector<uint8_t> doy_dup;       // window pools some row twice
    int dpw = 0, ept = 0, nwords_pad = 0;
    size_t smem = 0;
    size_t smem_cell = 0;               // k_thr_cell (0 = its bitmaps do not fit: use k_thr_ranked for every cell)
    SelTable sel;
};

static size_t ranked_smem(int E)
{
    const size_t Epad = ((size_t)E + 63) & ~(size_t)63;
    return Epad * (4 + 2 + 4 + 2) + (size_t)kRadixBins * kRankedThreads * 2;
}

static void plan_ranked(const int32_t *win_rows, int n_doy, int n_y, int W, const double *q, int P, RankedPlan &pl)
{
    const int64_t E = (int64_t)n_doy * n_y, n = (int64_t)W * n_y;
    pl.usable = false;
    if (E > 65535 || n < 2) return;
    pl.smem = ranked_smem((int)E);
    if (pl.smem > 227 * 1024 - 256) return;
    const int nwords = (int)((E + 31) / 32);
    pl.nwords_pad = (nwords + 31) / 32 * 32;
    pl.dpw = (n_doy + kRankedWarps - 1) / kRankedWarps;
    pl.ept = (int)((E + kRankedThreads - 1) / kRankedThreads) | 1;          // odd: conflict-free strided key reads
    pl.doy_dup.assign(n_doy, 0);
    pl.op_off.assign(n_doy + 1, 0);
    pl.ops.clear();
    std::vector<int> cur(n_doy), prev(n_doy);
    for (int d = 0; d < n_doy; d++) {
        std::fill(cur.begin(), cur.end(), 0);
        for (int k = 0; k < W; k++) cur[win_rows[d * W + k]]++;
        for (int r = 0; r < n_doy; r++) {
            if (cur[r] > 2) return;                                         // more than twice: generic kernel
            if (cur[r] == 2) pl.doy_dup[d] = 1;
        }
        const bool first = d % pl.dpw == 0;
        pl.op_off[d] = (int)pl.ops.size();
        for (int r = 0; r < n_doy; r++) {
            const int before = first ? 0 : prev[r];
            for (int i = cur[r]; i < before; i++) pl.ops.push_back(r << 1);        // leaves first ...
        }
        for (int r = 0; r < n_doy; r++) {
            const int before = first ? 0 : prev[r];
            for (int i = before; i < cur[r]; i++) pl.ops.push_back((r << 1) | 1);  // ... then enters
        }
        prev.swap(cur);
    }
    pl.op_off[n_doy] = (int)pl.ops.size();
    int n_b = 0;
    for (int w = 0; w < 64; w++) {
        pl.sel.b_slot[w] = -1;
        bool need = false;
        for (int d = w * pl.dpw; d < std::min(n_doy, (w + 1) * pl.dpw); d++) need |= pl.doy_dup[d] != 0;
        if (need && w < kRankedWarps) pl.sel.b_slot[w] = (int8_t)n_b++;
    }
    if ((size_t)pl.nwords_pad * (kRankedWarps * 6 + n_b * 4) > pl.smem - (((size_t)E + 63) & ~(size_t)63) * 6) return;
    {
        const size_t Epad = ((size_t)E + 63) & ~(size_t)63;
        const size_t need = Epad * 12 + (size_t)8 * kCellThreads * 4;
        const bool planes_fit = (size_t)pl.nwords_pad * (kCellWarps * 6 + n_b * 4) <= Epad * 8;
        const bool ranks_fit = Epad * 2 + ((size_t)n_doy + 2 + pl.ops.size()) * 2 <= (size_t)8 * kCellThreads * 4 && pl.ops.size() < 65536;
        pl.smem_cell = (need <= 227 * 1024 - 1024 && planes_fit && ranks_fit) ? need : 0;
    }
    // positions and weights: numba/np/arraymath.py:1655-1704 with n fixed (every window pools W * n_y samples)
    for (int p = 0; p < HDP_B200_MAX_PERCENTILES; p++) {
        pl.sel.pos_lo[p] = pl.sel.pos_hi[p] = 0;
        pl.sel.mode[p] = kSelInterp;
        pl.sel.w_lo[p] = pl.sel.w_hi[p] = 0.0;
        if (p >= P) continue;
        volatile double pct = q[p] * 100.0;
        if (pct == 100.0) { pl.sel.mode[p] = kSelMax; pl.sel.pos_lo[p] = pl.sel.pos_hi[p] = (int)n - 1; continue; }
        if (pct == 0.0) { pl.sel.mode[p] = kSelMin; continue; }
        volatile double frac = pct / 100.0;
        volatile double scaled = (double)(n - 1) * frac;
        volatile double rank = 1.0 + scaled;
        const double f = floor(rank);
        volatile double m = rank - f;
        volatile double w0 = 1.0 - m;
        int64_t k = (int64_t)f - 1;
        if (k < 0) k = 0;
        if (k >= n - 1) { pl.sel.pos_lo[p] = pl.sel.pos_hi[p] = (int)n - 1; }
        else { pl.sel.pos_lo[p] = (int)k; pl.sel.pos_hi[p] = (int)k + 1; }
        pl.sel.w_lo[p] = w0;
        pl.sel.w_hi[p] = m;
    }
    pl.usable = true;
}

struct ThrLayout {
    size_t total = 0;
    float *xn = nullptr;
    int *time_index = nullptr, *win_rows = nullptr, *op_off = nullptr, *ops = nullptr;
    int *fallback = nullptr;            // [1 + C]: count, then the cells k_thr_cell hands over to k_thr_ranked
    uint8_t *doy_dup = nullptr;
};

static ThrLayout carve_thr(void *ws, size_t ws_bytes, int64_t C, int64_t T_b, bool need_norm, int n_doy, int n_y, int W)
{
    ThrLayout L;
    Carver cv(ws, ws_bytes);
    if (need_norm) L.xn = cv.take<float>((size_t)C * T_b);
    L.time_index = cv.take<int>((size_t)n_doy * n_y);
    L.win_rows = cv.take<int>((size_t)n_doy * W);
    L.op_off = cv.take<int>((size_t)n_doy + 1);
    L.ops = cv.take<int>((size_t)2 * W * (n_doy + kRankedWarps));           // <= 2W changes per day, W per range start
    L.doy_dup = cv.take<uint8_t>((size_t)n_doy);
    L.fallback = cv.take<int>((size_t)C + 1);
    L.total = cv.off;
    return L;
}

static int g_force_generic = 0;     // 1: k_thr_generic for everything; 2: k_thr_ranked instead of k_thr_cell
static int g_force_ranked = 0;

}  // namespace hdp

using namespace hdp;

extern "C" {

void hdp_b200_thresholds_force_generic(int on) { g_force_generic = on == 1; g_force_ranked = on == 2; }

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

    RankedPlan plan;
    if (!g_force_generic) plan_ranked(h_win_rows, n_doy, n_y, W, h_q, P, plan);
    if (plan.usable) {
        HDP_CUDA_TRY(cudaMemcpyAsync(L.op_off, plan.op_off.data(), sizeof(int) * plan.op_off.size(), cudaMemcpyHostToDevice, st));
        HDP_CUDA_TRY(cudaMemcpyAsync(L.ops, plan.ops.data(), sizeof(int) * plan.ops.size(), cudaMemcpyHostToDevice, st));
        HDP_CUDA_TRY(cudaMemcpyAsync(L.doy_dup, plan.doy_dup.data(), plan.doy_dup.size(), cudaMemcpyHostToDevice, st));
        HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_ranked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        if (plan.smem_cell && !g_force_ranked) {
            HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_cell));
            HDP_CUDA_TRY(cudaMemsetAsync(L.fallback, 0, sizeof(int), st));
            {
                KernelTimer timer(kThrSort, st);
                k_thr_cell<<<(unsigned)C, kCellThreads, plan.smem_cell, st>>>(x, T_b, ld_t, L.time_index, n_doy * n_y, n_y, n_doy, (int)b,
                                                                            L.op_off, L.ops, L.doy_dup, (int)plan.ops.size(), plan.dpw, plan.ept,
                                                                            plan.nwords_pad, plan.sel, P, d_out, L.fallback, L.fallback + 1);
                HDP_LAUNCH_CHECK();
            }
            // cells whose samples pile into one bucket (rare): a few persistent CTAs walk the device-side list
            KernelTimer timer(kThrSelect, st);
            const unsigned grid = (unsigned)std::min<int64_t>(C, 148);
            k_thr_ranked<<<grid, kRankedThreads, plan.smem, st>>>(x, T_b, ld_t, L.time_index, n_doy * n_y, n_y, n_doy, (int)b,
                                                                 L.op_off, L.ops, L.doy_dup, plan.dpw, plan.ept, plan.nwords_pad,
                                                                 plan.sel, P, d_out, L.fallback, L.fallback + 1);
            HDP_LAUNCH_CHECK();
            return HDP_B200_OK;
        }
        KernelTimer timer(kThrSort, st);
        k_thr_ranked<<<(unsigned)C, kRankedThreads, plan.smem, st>>>(x, T_b, ld_t, L.time_index, n_doy * n_y, n_y, n_doy, (int)b,
                                                                    L.op_off, L.ops, L.doy_dup, plan.dpw, plan.ept, plan.nwords_pad,
                                                                    plan.sel, P, d_out, nullptr, nullptr);
        HDP_LAUNCH_CHECK();
        return HDP_B200_OK;
    }

    // generic path: any table, any multiplicity
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
