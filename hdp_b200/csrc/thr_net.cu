// thr_net.cu - k_thr_net: path 1 (thresholds) for high quantiles, one grid cell per LANE, everything in registers.
//
// Replaces (reference = AgentOxygen/HDP v1.0.2) the same code as threshold.cu: compute_percentiles (hdp/threshold.py:52-78)
// looped over cells (:81-93) with Numba's np.quantile arithmetic (numba/np/arraymath.py:1655-1704).
//
// Idea.  Every requested position lies among the K largest samples of its window (K = 46 of 450 for q >= 0.9, 15-day window,
// 30 years), and a window is W consecutive day-of-year rows of n_y samples each.  So:
//   * a row is ordered ONCE per use by a sorting network (compare-exchange = one FMNMX pair, no branches, no shared memory),
//   * "the K largest of rows i..j" are kept as descending K-lists and combined with merge networks (K largest of two lists),
//   * the sliding window is decomposed so that no list ever has to forget a row: rows are cut into blocks of s = W / 3 rows and
//         window(b, i) = suffix_i(block b)  u  block b+1  u  block b+2  u  prefix_i(block b+3)        (van Herk / Gil-Werman)
//     where suffix lists grow backwards through block b, prefix lists grow forwards through block b+3, and the pair of full
//     blocks in the middle is one merge per block.  (Windows with W not divisible by 3 use one block per window: suffix u prefix.)
// A lane owns one cell, a warp 32 neighbouring cells: every load of a day is one full 128-byte line, there is no gather tile, no
// atomics, no bank conflict, no divergence - the instruction stream is identical for all cells.  Lists that must outlive a row
// step wait in shared memory ([element][lane], conflict-free); s + 1 lists of K floats per cell.
// The networks come from tools/gen_networks.py (thr_net_gen.cuh): pruned odd-even merges and Knuth's merge exchange.
//
// Results are the reference's, bit for bit: selecting order statistics exactly is order-independent, and the interpolation is the
// same separately rounded f64 expression as in threshold.cu.  Cells with NaN / +-inf samples (the reference has special rules for
// them) are detected on the fly and handed to k_thr_seg through its hand-over list; days whose window is not W consecutive rows
// (the reference's mirrored year end, threshold.py:45-47) are pooled row by row from the window table.
#include <math.h>
#include <algorithm>
#include <map>

#include "thr_net.cuh"
#include "thr_net_gen.cuh"

// Template instances that exist (samples per row NY x list length K x blocks per window M).  net_plan picks the smallest
// NY >= n_y (30 exactly for 30-year baselines: the only instance without the pad test per load - NY = 30 is chosen for n_y == 30
// only) and the smallest K >= the deepest requested position; every combination of the two menus must be listed.  Each
// instance is two kernels of ~3 000 unrolled instructions: the menu is what keeps the build at a minute.
#ifdef HDP_NET_DEV                       /* quick kernel iterations: the bench shape and one padded shape */
#define HDP_NET_NY_MENU 32
#define HDP_NET_K_MENU 48
#define HDP_NET_INSTANCES(X) X(30, 48, 3) X(30, 48, 1) X(32, 48, 3) X(32, 48, 1)
#else
#define HDP_NET_NY_MENU 16, 32
#define HDP_NET_K_MENU 16, 32, 48, 64
#define HDP_NET_ROW(X, ny) X(ny, 16, 3) X(ny, 16, 1) X(ny, 32, 3) X(ny, 32, 1) X(ny, 48, 3) X(ny, 48, 1) X(ny, 64, 3) X(ny, 64, 1)
#define HDP_NET_INSTANCES(X) HDP_NET_ROW(X, 16) HDP_NET_ROW(X, 30) HDP_NET_ROW(X, 32)
#endif

namespace hdp {

// Pads (rows shorter than NY, lists shorter than K, rows past the end of the sequence) are the most negative FINITE float: they
// sort below every sample, never reach a requested position (a window holds at least K real samples), and - unlike -inf - do not
// trip the non-finite census.
#define HDP_NET_PAD (-3.402823466e+38f)

struct NetStepDesc {            // what one row step does besides "order the row and merge it into the running list" (warp-uniform)
    int start;                  // the running list starts over: -1 = no, kStartEmpty = with this row alone, else = from that slot
    int other;                  // slot of the stored list to combine the running list with (-1: none)
    int store_dst;              // slot that receives the combination (-1: none)
    int store_run;              // slot that receives the running list itself (-1: none)
    int emit_day;               // day of year whose thresholds this step yields (-1: none) - from the combination if `other` >= 0,
};                              // else from the running list.  (The rank lookup goes through slot 0, which is free whenever a window ends.)
constexpr int kStartEmpty = 1 << 20;

// The row's time indices are warp-uniform: lane y fetches entry y (ONE coalesced load per row, issued two rows ahead) and the
// data loads of the next row take them by shuffle - no load depends on a load that was issued in the same row step.
// `pitch` = bytes between consecutive time steps (< 4 GB): the address is base + t * pitch, ONE 32 x 32 -> 64 bit multiply-add.
// kPads: rows are shorter than NY (entries -1); a row past the end of the sequence is all pads and skipped as a whole.
template <int NY, bool kPads>
__device__ __forceinline__ void net_issue_loads(float (&raw)[NY], const float *xc, uint32_t pitch, int t_mine)
{
    const int t0 = __shfl_sync(0xffffffffu, t_mine, 0);
    if (t0 < 0) {                                                 // warp-uniform: entry 0 is a pad only in the all-pad row
#pragma unroll
        for (int y = 0; y < NY; y++) raw[y] = HDP_NET_PAD;
        return;
    }
#pragma unroll
    for (int y = 0; y < NY; y++) {
        const int t = y ? __shfl_sync(0xffffffffu, t_mine, y) : t0;
        const float *p = (const float *)((const char *)xc + (uint64_t)(uint32_t)t * pitch);
        if (kPads) raw[y] = t >= 0 ? __ldg(p) : HDP_NET_PAD;
        else raw[y] = __ldg(p);
    }
}

// The requested ranks of a finished window: the list goes through a scratch slot (the positions are run-time values).  Five
// percentiles at a time: their ten lookups are in flight together, then the f64 arithmetic, then the stores.
__device__ __forceinline__ double net_value(float lower, float upper, int p, const NetSel &sel)
{
    // numba/np/arraymath.py:1697-1701 (separately rounded), :1669-1675 for q == 1
    return sel.is_max[p] ? (double)upper : __dadd_rn(__dmul_rn((double)lower, sel.w_lo[p]), __dmul_rn((double)upper, sel.w_hi[p]));
}

template <int K>
__device__ __forceinline__ void net_emit(const float (&list)[K], float *o, bool valid, double *__restrict__ dst, int P, const NetSel &sel)
{
#pragma unroll
    for (int e = 0; e < K; e++) o[e * 32] = list[e];
    if (!valid) return;
    int p = 0;
    for (; p + 5 <= P; p += 5) {
        float lo[5], hi[5];
#pragma unroll
        for (int j = 0; j < 5; j++) { lo[j] = o[sel.idx_lo[p + j] * 32]; hi[j] = o[sel.idx_hi[p + j] * 32]; }
        double v[5];
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] = net_value(lo[j], hi[j], p + j, sel);
#pragma unroll
        for (int j = 0; j < 5; j++) dst[p + j] = v[j];
    }
    for (; p < P; p++) dst[p] = net_value(o[sel.idx_lo[p] * 32], o[sel.idx_hi[p] * 32], p, sel);
}

// ---- where the lists that outlive a row step wait ----
// Slot numbers: 0 = the pair of full blocks (M == 3) and the scratch for the rank lookup, 1 .. s-1 = suffix lists, s = last full block.
template <int K>
struct SlotsSmem {                                                // every slot in shared memory: element e of slot j at my[(j * K + e) * 32]
    float *my;
    __device__ __forceinline__ float *scratch() const { return my; }
    __device__ __forceinline__ void load(int slot, float (&a)[K]) const
    {
        const float *o = my + (size_t)slot * K * 32;
#pragma unroll
        for (int e = 0; e < K; e++) a[e] = o[e * 32];
    }
    __device__ __forceinline__ void store(int slot, const float (&a)[K]) const
    {
        float *w = my + (size_t)slot * K * 32;
#pragma unroll
        for (int e = 0; e < K; e++) w[e * 32] = a[e];
    }
};

// Tensor memory as list storage.  TMEM is [128 lanes][512 columns] of 32 bits; a warp reaches the 32 lanes of its quarter
// (warp % 4) with tcgen05.ld / tcgen05.st, thread i <-> lane i, N consecutive columns <-> N registers: exactly the
// [element][cell] layout of a list, with no address arithmetic per element and none of the shared memory that limits how many
// warps an SM can hold.  (No tensor-core instruction is involved: the memory is used as a 256 KB scratchpad.)
#define HDP_R16(a, o) "=f"(a[o]), "=f"(a[o + 1]), "=f"(a[o + 2]), "=f"(a[o + 3]), "=f"(a[o + 4]), "=f"(a[o + 5]), "=f"(a[o + 6]), "=f"(a[o + 7]), \
                      "=f"(a[o + 8]), "=f"(a[o + 9]), "=f"(a[o + 10]), "=f"(a[o + 11]), "=f"(a[o + 12]), "=f"(a[o + 13]), "=f"(a[o + 14]), "=f"(a[o + 15])
#define HDP_W16(a, o) "f"(a[o]), "f"(a[o + 1]), "f"(a[o + 2]), "f"(a[o + 3]), "f"(a[o + 4]), "f"(a[o + 5]), "f"(a[o + 6]), "f"(a[o + 7]), \
                      "f"(a[o + 8]), "f"(a[o + 9]), "f"(a[o + 10]), "f"(a[o + 11]), "f"(a[o + 12]), "f"(a[o + 13]), "f"(a[o + 14]), "f"(a[o + 15])
#define HDP_RW16(a, o) "+f"(a[o]), "+f"(a[o + 1]), "+f"(a[o + 2]), "+f"(a[o + 3]), "+f"(a[o + 4]), "+f"(a[o + 5]), "+f"(a[o + 6]), "+f"(a[o + 7]), \
                       "+f"(a[o + 8]), "+f"(a[o + 9]), "+f"(a[o + 10]), "+f"(a[o + 11]), "+f"(a[o + 12]), "+f"(a[o + 13]), "+f"(a[o + 14]), "+f"(a[o + 15])

template <int K, int O>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&a)[K])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : HDP_R16(a, O) : "r"(taddr + O) : "memory");
}
template <int K, int O>
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&a)[K])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr + O), HDP_W16(a, O) : "memory");
}
// The loaded registers are valid after tcgen05.wait::ld; naming them as in/out operands keeps every use of them behind the wait.
template <int K, int O>
__device__ __forceinline__ void tmem_wait_ld16(float (&a)[K])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : HDP_RW16(a, O) :: "memory");
}

template <int K>
struct SlotsMixed {                                               // suffix lists 1 .. n_tm in tensor memory, everything else in shared memory
    float *my;                                                    // shared memory: slot 0, the last full block, suffix lists n_tm + 1 ..
    uint32_t taddr;                                               // this warp's lanes and first column
    int n_tm, slot_f;
    __device__ __forceinline__ float *scratch() const { return my; }
    __device__ __forceinline__ int smem_index(int slot) const { return slot == 0 ? 0 : slot == slot_f ? 1 : 1 + slot - n_tm; }
    __device__ __forceinline__ bool in_tmem(int slot) const { return slot >= 1 && slot <= n_tm && slot != slot_f; }
    __device__ __forceinline__ void load(int slot, float (&a)[K]) const
    {
        if (in_tmem(slot)) {                                      // warp-uniform
            const uint32_t t = taddr + (uint32_t)(slot - 1) * K;
            tmem_ld16<K, 0>(t, a);
            if constexpr (K > 16) tmem_ld16<K, 16>(t, a);
            if constexpr (K > 32) tmem_ld16<K, 32>(t, a);
            if constexpr (K > 48) tmem_ld16<K, 48>(t, a);
            tmem_wait_ld16<K, 0>(a);
            if constexpr (K > 16) tmem_wait_ld16<K, 16>(a);
            if constexpr (K > 32) tmem_wait_ld16<K, 32>(a);
            if constexpr (K > 48) tmem_wait_ld16<K, 48>(a);
        } else {
            const float *o = my + (size_t)smem_index(slot) * K * 32;
#pragma unroll
            for (int e = 0; e < K; e++) a[e] = o[e * 32];
        }
    }
    __device__ __forceinline__ void store(int slot, const float (&a)[K]) const
    {
        if (in_tmem(slot)) {
            const uint32_t t = taddr + (uint32_t)(slot - 1) * K;
            tmem_st16<K, 0>(t, a);
            if constexpr (K > 16) tmem_st16<K, 16>(t, a);
            if constexpr (K > 32) tmem_st16<K, 32>(t, a);
            if constexpr (K > 48) tmem_st16<K, 48>(t, a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            float *w = my + (size_t)smem_index(slot) * K * 32;
#pragma unroll
            for (int e = 0; e < K; e++) w[e * 32] = a[e];
        }
    }
};

// One work item, worked through by ONE warp: (32-cell tile, chunk of super-steps) of the block schedule, or - kProg, a kernel of its
// own (k_thr_net_irr), so that the hot loop of the regular items carries none of its branches - the tile's irregular days, which
// follow a PROGRAM built by net_plan (item = tile).
template <int NY, int K, int M, bool kPads, bool kProg, class Slots>
__device__ __forceinline__ void net_item(int64_t item, const Slots &slots, const float *__restrict__ x, int64_t C, uint32_t ld_t,
                                         const int *__restrict__ seq_time, const int *__restrict__ win_day, const int *__restrict__ irr_day,
                                         const int *__restrict__ irr_time, const NetGeom &g, const NetSel &sel, double *__restrict__ out,
                                         const NetHandOver &hand)
{
    static_assert(NY <= 32, "one lane per time index of a row");
    const int lane = threadIdx.x & 31;
    constexpr bool irregular = kProg;
    const int64_t tile = irregular ? item : item / g.n_chunks;
    const int chunk = irregular ? 0 : (int)(item - tile * g.n_chunks);
    const int64_t c = tile * 32 + lane;
    const bool valid = c < C;
    const float *xc = x + (valid ? c : C - 1);
    const int s = g.s;
    constexpr int kPro = M == 3 ? 2 : 0;                          // prologue blocks of a chunk (the two full blocks after its first one)
    const int slot_f = s;
    const int ty = lane < NY ? lane : NY - 1;

    const int b0 = chunk * g.steps_per_chunk, b1 = min(b0 + g.steps_per_chunk, g.n_steps);
    const int n_rows = irregular ? g.n_irr_steps : (kPro + 2 * (b1 - b0)) * s;
    const int per = irregular ? 0x7fffffff : s;
    const int4 *prog = (const int4 *)irr_day;                     // kProg: what every row step does

    // the row table of step n = (pb, it): block phase and row inside it
    auto row_table = [&](int n, int pb, int it) -> const int * {
        if (irregular) return irr_time + (size_t)n * NY;
        int row;
        if (pb < kPro) row = (b0 + 1 + pb) * s + it;
        else {
            const int q = pb - kPro, b = b0 + (q >> 1);
            row = (q & 1) ? (b + M) * s + it : b * s + (s - 1 - it);
        }
        return seq_time + (size_t)min(row, g.n_seq) * NY;        // rows past the sequence: the all-pad row
    };
    auto advance = [&](int &pb, int &it) { if (++it == per) { it = 0; pb++; } };

    float raw[NY];
    float run[K];
    float bad_acc = 0.0f;
    int pb = 0, it = 0;                                           // irregular items: pb = day, it = row of its window
    int pb1 = 0, it1 = 0, pb2, it2;                               // steps n + 1 and n + 2
    int t_next = 0;                                               // time indices of row n + 1 (entry `lane`)
    if (n_rows > 0) {
        net_issue_loads<NY, kPads>(raw, xc, ld_t, __ldg(row_table(0, 0, 0) + ty));
        advance(pb1, it1);
        if (n_rows > 1) t_next = __ldg(row_table(1, pb1, it1) + ty);
    }
    pb2 = pb1; it2 = it1;
    advance(pb2, it2);
#pragma unroll
    for (int e = 0; e < K; e++) run[e] = HDP_NET_PAD;

    for (int n = 0; n < n_rows; n++) {
        float v[NY];
        if (g.unit == 0) {                                        // (warp-uniform; one branch per row, not per sample)
#pragma unroll
            for (int y = 0; y < NY; y++) v[y] = raw[y];
        } else {                                                  // Kelvin input: converted as it arrives, one FADD per sample (pads stay
#pragma unroll                                                    // the most negative values); Fahrenheit input is converted beforehand
            for (int y = 0; y < NY; y++) v[y] = to_celsius_f(raw[y], 1);
        }
#pragma unroll
        for (int y = 0; y < NY; y++) bad_acc = __fmaf_rn(v[y], 0.0f, bad_acc);   // non-finite census on the otherwise idle FMA pipe
        if (n + 1 < n_rows) net_issue_loads<NY, kPads>(raw, xc, ld_t, t_next);       // row n + 1: in flight while this row is worked on
        if (n + 2 < n_rows) t_next = __ldg(row_table(n + 2, pb2, it2) + ty);         // its indices were fetched a row earlier

        // ---- what this step does (the day a finished window belongs to is fetched now and looked at after the networks)
        NetStepDesc d{it == 0 ? kStartEmpty : -1, -1, -1, -1, -1};
        const int *emit_from = nullptr;
        bool window_of_prefix = false;
        int4 pg = make_int4(0, 0, -1, 0);
        if (irregular) {
            pg = __ldg(prog + n);
        } else if (pb < kPro) {
            if (it == s - 1) {
                d.store_run = slot_f;
                if (pb == 1) { d.other = slot_f; d.store_dst = 0; }
            }
        } else {
            const int q = pb - kPro, b = b0 + (q >> 1);
            if (!(q & 1)) {                                       // block b, last row first: suffix lists, seeded with the two full blocks
                const int i = s - 1 - it;                         //   behind it (M == 3) so that no extra merge is needed per window
                if (M == 3 && it == 0) d.start = 0;
                if (i > 0) d.store_run = i;
                else emit_from = win_day + b * s;
            } else if (it < s - 1) {                              // prefix lists of block b + M: window (b, it + 1)
                emit_from = win_day + b * s + it + 1;
                window_of_prefix = true;
            } else if (M == 3) {                                  // block b + 3 is complete: next pair of full blocks
                d.other = slot_f; d.store_dst = 0; d.store_run = slot_f;
            }
        }
        const int emit_day = emit_from ? __ldg(emit_from) : -1;

        // ---- order the row, merge it into the running list
        net::Sort<NY>::run(v);
        if (irregular) {
            d.start = (pg.x & 1) ? kStartEmpty : -1;
            d.other = (pg.y & 0xff) - 1;
            d.store_run = ((pg.y >> 8) & 0xff) - 1;
            if (pg.x & 2) {                                       // a row that joins this window only: merged into a copy, the chain goes on without it
                float tmp[K];
#pragma unroll
                for (int e = 0; e < K; e++) tmp[e] = run[e];
                net::Merge<K, NY>::run(tmp, v);
                if (pg.z >= 0) net_emit<K>(tmp, slots.scratch(), valid, out + ((size_t)(valid ? c : 0) * g.n_doy + pg.z) * g.P, g.P, sel);
                pb = pb1; it = it1; pb1 = pb2; it1 = it2;
                advance(pb2, it2);
                continue;
            }
        }
        if (d.start == kStartEmpty) {
#pragma unroll
            for (int e = 0; e < K; e++) run[e] = e < NY ? v[e < NY ? e : 0] : HDP_NET_PAD;
        } else {
            if (d.start >= 0) slots.load(d.start, run);
            net::Merge<K, NY>::run(run, v);
        }

        // ---- combine with a stored list; store; look up the requested ranks
        d.emit_day = irregular ? pg.z : emit_day;
        if (window_of_prefix && emit_day >= 0) d.other = it + 1;
        double *dst = out + ((size_t)(valid ? c : 0) * g.n_doy + max(d.emit_day, 0)) * g.P;
        if (d.other >= 0) {
            float tmp[K];
            slots.load(d.other, tmp);
            net::Merge<K, K>::run(tmp, run);
            if (d.store_dst >= 0) slots.store(d.store_dst, tmp);
            if (d.emit_day >= 0) net_emit<K>(tmp, slots.scratch(), valid, dst, g.P, sel);
        } else if (d.emit_day >= 0) {
            net_emit<K>(run, slots.scratch(), valid, dst, g.P, sel);
        }
        if (d.store_run >= 0) slots.store(d.store_run, run);
        pb = pb1; it = it1; pb1 = pb2; it1 = it2;
        advance(pb2, it2);
    }

    // ---- cells with NaN / +-inf samples: all their segments go onto k_thr_seg's hand-over list (it runs behind this kernel)
    if (valid && bad_acc != bad_acc && hand.list) {
        const int group = (int)(c / hand.group_cells), w = (int)(c - (int64_t)group * hand.group_cells);
        const int chunk_g = group / hand.gc, gcc = min(hand.gc, hand.n_groups - chunk_g * hand.gc);
        for (int sg = 0; sg < hand.n_seg; sg++) {
            const unsigned bid = (unsigned)(chunk_g * (hand.n_seg * hand.gc) + sg * gcc + (group - chunk_g * hand.gc));
            if (atomicOr(&hand.list[1 + bid], 1u << w) == 0u) hand.list[1 + hand.n_blocks + atomicAdd(&hand.list[0], 1u)] = bid;
        }
    }
}

// One warp per CTA, one item per CTA, every list in shared memory (s + 1 lists: 6 warps per SM for the 15-day window at K = 48).
template <int NY, int K, int M, bool kPads>
__global__ void __launch_bounds__(32)
k_thr_net(const float *__restrict__ x, int64_t C, uint32_t ld_t,
          const int *__restrict__ seq_time, const int *__restrict__ win_day, const int *__restrict__ irr_day, const int *__restrict__ irr_time,
          const __grid_constant__ NetGeom g, const __grid_constant__ NetSel sel, double *__restrict__ out, const __grid_constant__ NetHandOver hand)
{
    extern __shared__ __align__(16) float sm[];
    const SlotsSmem<K> slots{sm + threadIdx.x};
    net_item<NY, K, M, kPads, false>(blockIdx.x, slots, x, C, ld_t, seq_time, win_day, irr_day, irr_time, g, sel, out, hand);
}

// The irregular days of one tile (the mirrored year end): one warp per CTA, lists in shared memory, the program of net_plan.
template <int NY, int K, int M, bool kPads>
__global__ void __launch_bounds__(32)
k_thr_net_irr(const float *__restrict__ x, int64_t C, uint32_t ld_t,
              const int *__restrict__ seq_time, const int *__restrict__ win_day, const int *__restrict__ irr_day, const int *__restrict__ irr_time,
              const __grid_constant__ NetGeom g, const __grid_constant__ NetSel sel, double *__restrict__ out, const __grid_constant__ NetHandOver hand)
{
    extern __shared__ __align__(16) float sm[];
    const SlotsSmem<K> slots{sm + threadIdx.x};
    net_item<NY, K, M, kPads, true>(blockIdx.x, slots, x, C, ld_t, seq_time, win_day, irr_day, irr_time, g, sel, out, hand);
}

// The same items on persistent CTAs of 4, 8 or 12 warps (one CTA per SM) that keep most suffix lists in TENSOR MEMORY: twice the
// warps per SM, which is what the ALU pipe needs to stay busy through the networks' dependency chains.  Warps take items from a
// global counter; warp w owns the TMEM lanes of quarter w % 4 and the column range (w / 4) * g.tm_cols.
constexpr int kNetTmWarpsMax = 12;
template <int NY, int K, int M, bool kPads>
__global__ void __launch_bounds__(kNetTmWarpsMax * 32, 1)
k_thr_net_tm(const float *__restrict__ x, int64_t C, uint32_t ld_t,
             const int *__restrict__ seq_time, const int *__restrict__ win_day, const int *__restrict__ irr_day, const int *__restrict__ irr_time,
             const __grid_constant__ NetGeom g, const __grid_constant__ NetSel sel, double *__restrict__ out, const __grid_constant__ NetHandOver hand,
             unsigned long long *__restrict__ next_item)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {                                              // the whole tensor memory of this SM (the only CTA on it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    SlotsMixed<K> slots;
    slots.my = sm + (size_t)warp * g.tm_smem_slots * K * 32 + lane;
    slots.taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * g.tm_cols);
    slots.n_tm = g.tm_lists;
    slots.slot_f = g.M == 3 ? g.s : -1;
    const unsigned long long n_items = (unsigned long long)(g.n_tiles * g.n_chunks);
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1ULL);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        net_item<NY, K, M, kPads, false>((int64_t)item, slots, x, C, ld_t, seq_time, win_day, irr_day, irr_time, g, sel, out, hand);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// --------------------------------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------------------------------
static int pick(const int *menu, int n_menu, int need)
{
    for (int i = 0; i < n_menu; i++) if (menu[i] >= need) return menu[i];
    return 0;
}

void net_set_cells(NetPlan &pl, int64_t C)
{
    NetGeom &g = pl.geo;
    g.n_tiles = (C + 31) / 32;
    // chunks of super-steps: enough items for ~8 waves of resident warps, but a chunk's two prologue blocks stay a small part of it
    const int64_t want = 148 * 6 * 8;
    int chunks = (int)std::min<int64_t>((want + g.n_tiles - 1) / std::max<int64_t>(g.n_tiles, 1), std::max(1, g.n_steps / 4));
    chunks = std::max(1, std::min(chunks, g.n_steps));
    g.steps_per_chunk = (g.n_steps + chunks - 1) / chunks;
    g.n_chunks = (g.n_steps + g.steps_per_chunk - 1) / g.steps_per_chunk;
}

void net_plan(const int32_t *time_index, const int32_t *win_rows, int64_t T_b, int n_doy, int n_y, int W,
              const int *pos_lo, const int *pos_hi, const int *mode_is_max, const int *mode_is_interp, const double *w_lo, const double *w_hi,
              int P, int64_t C, NetPlan &pl)
{
    pl.usable = false;
    NetGeom &g = pl.geo;
    if (!(W & 1) || n_doy < 2 * W || n_y < 1) return;
    const int r = W / 2;
    static const int ny_menu[] = {HDP_NET_NY_MENU}, k_menu[] = {HDP_NET_K_MENU};
    g.NY = n_y == 30 ? 30 : pick(ny_menu, (int)(sizeof(ny_menu) / sizeof(int)), n_y);
    if (!g.NY) return;
    g.M = W % 3 == 0 ? 3 : 1;
    g.s = W / g.M;
    if (g.M == 1 && W > 13) return;
    // requested positions, from the top of the window
    const int64_t n = (int64_t)W * n_y;
    int k_top = 1;
    for (int p = 0; p < P; p++) {
        if (!mode_is_max[p] && !mode_is_interp[p]) return;                      // q == 0: the minimum is not among the largest
        const int64_t lo = n - 1 - pos_lo[p], hi = n - 1 - pos_hi[p];
        if (lo < 0 || hi < 0 || lo > 1000 || hi > lo) return;
        k_top = std::max<int>(k_top, (int)lo + 1);
        pl.sel.idx_lo[p] = (int)lo; pl.sel.idx_hi[p] = (int)hi; pl.sel.is_max[p] = mode_is_max[p];
        pl.sel.w_lo[p] = w_lo[p]; pl.sel.w_hi[p] = w_hi[p];
        if (mode_is_max[p]) pl.sel.idx_lo[p] = pl.sel.idx_hi[p] = 0;
    }
    g.K = pick(k_menu, (int)(sizeof(k_menu) / sizeof(int)), k_top);
    if (!g.K || k_top > n) return;
    g.W = W; g.n_y = n_y; g.n_doy = n_doy; g.P = P;
    g.n_seq = n_doy + r;
    g.n_win = n_doy - r;                                                         // windows 0 .. n_doy-r-1 = sequence rows [k, k + W)
    g.n_steps = (g.n_win + g.s - 1) / g.s;
    const int n_slots = g.M == 3 ? g.s + 1 : std::max(g.s, 1);
    const size_t list_bytes = (size_t)g.K * 32 * sizeof(float);
    g.smem = n_slots * list_bytes;
    if (g.smem > 227 * 1024) return;
    // tensor-memory variant: the largest CTA (12, 8 or 4 warps; warps of one lane quarter split the 512 columns) whose remaining
    // shared-memory slots fit one SM - used when that is more warps per SM than the one-warp CTAs reach
    g.tm_warps = 0;
    const int warps_smem_only = (int)std::min<size_t>(32, (227 * 1024) / (g.smem + 1024));
    for (int warps = kNetTmWarpsMax; warps >= 4; warps -= 4) {
        const int cols = 512 / (warps / 4), lists = std::min(g.s - 1, cols / g.K);
        const int smem_slots = n_slots - lists + (g.M == 1 ? 1 : 0);
        if (lists < 1 || warps <= warps_smem_only) break;
        if ((size_t)warps * smem_slots * list_bytes > 227 * 1024 - 256) continue;
        g.tm_warps = warps; g.tm_lists = lists; g.tm_cols = cols; g.tm_smem_slots = smem_slots;
        break;
    }

    auto time_of = [&](int row, int y) -> int {
        if (y >= n_y) return -1;
        int64_t t = time_index[(size_t)row * n_y + y];
        if (t < 0) t += T_b;                                                     // -1 pads read the LAST sample (threshold.py:35,77)
        return (int)t;
    };
    // the linear row sequence -r .. n_doy-1 (rows before 0 wrap to the end of the year, threshold.py:44-48)
    pl.seq_time.assign((size_t)(g.n_seq + 1) * g.NY, -1);
    for (int k = 0; k < g.n_seq; k++) {
        const int row = ((k - r) % n_doy + n_doy) % n_doy;
        for (int y = 0; y < g.NY; y++) pl.seq_time[(size_t)k * g.NY + y] = time_of(row, y);
    }
    // day d is regular if its window is exactly the sequence rows [d, d + W)
    pl.win_day.assign((size_t)g.n_steps * g.s, -1);
    pl.irr_day.clear(); pl.irr_time.clear();
    std::vector<int> want(W), have(W), irr;
    for (int d = 0; d < n_doy; d++) {
        bool regular = d < g.n_win;
        if (regular) {
            for (int j = 0; j < W; j++) { want[j] = ((d - r + j) % n_doy + n_doy) % n_doy; have[j] = win_rows[(size_t)d * W + j]; }
            std::sort(want.begin(), want.end()); std::sort(have.begin(), have.end());
            regular = want == have;
        }
        if (regular) pl.win_day[d] = d; else irr.push_back(d);
    }
    g.n_irr = (int)irr.size();
    if (g.n_irr * 4 > n_doy) return;                                             // mostly irregular tables: not this kernel

    // The irregular days' program (k_thr_net_irr runs it, one warp per tile).  A step orders one row and merges it into the running
    // list; flags, slots and the day to emit as documented at NetPlan::irr_day.
    auto step = [&](int row, int flags, int other, int store_run, int emit_day) {
        pl.irr_day.push_back(flags);
        pl.irr_day.push_back((other + 1) | ((store_run + 1) << 8));
        pl.irr_day.push_back(emit_day);
        pl.irr_day.push_back(0);
        for (int y = 0; y < g.NY; y++) pl.irr_time.push_back(time_of(row, y));
    };
    // (a) The reference's mirrored year end (hdp/threshold.py:45-47): day d >= n_doy - r pools rows [d - r, n_doy), rows
    // [2 n_doy - d - r, n_doy) a second time, and row 0.  Both runs are SUFFIXES of the year, so one chain down from the last row
    // yields them all: the short suffixes (second runs) are stored, row 0 joins the chain once they are, and every further row
    // finishes one day: 2 r + 2 row steps instead of r W.
    const int avail = g.M == 3 ? g.s : g.s - 1;                                  // slots besides the scratch slot 0
    bool mirrored = g.n_irr == r && r >= 1 && r - 1 <= avail + 1;
    for (int i = 0; mirrored && i < g.n_irr; i++) {
        const int d = irr[i];
        mirrored = d == n_doy - r + i;
        std::vector<int> w2;
        for (int j = d - r; j < n_doy; j++) w2.push_back(j);
        for (int j = 2 * n_doy - d - r; j < n_doy; j++) w2.push_back(j);
        w2.push_back(0);
        if ((int)w2.size() != W) { mirrored = false; break; }
        for (int j = 0; j < W; j++) have[j] = win_rows[(size_t)d * W + j];
        std::sort(w2.begin(), w2.end()); std::sort(have.begin(), have.end());
        mirrored = mirrored && w2 == have;
    }
    if (mirrored) {
        // second runs start at rows n_doy-1 (day n_doy-r+1) .. n_doy-r+1 (day n_doy-1): r - 1 lists; the one-row list of the last row
        // is redone as a side step when there is one slot too few
        const bool side = r - 1 > avail;
        auto slot_of = [&](int j) { return n_doy - j - (side ? 1 : 0); };        // suffix starting at row j -> slot (1 ..)
        for (int j = n_doy - 1; j >= n_doy - r + 1; j--)
            step(j, j == n_doy - 1 ? 1 : 0, -1, (side && j == n_doy - 1) ? -1 : slot_of(j), -1);
        step(n_doy - r, r == 1 ? 1 : 0, -1, -1, -1);                             // (no day starts or ends its second run here)
        step(0, 0, -1, -1, -1);                                                  // row 0 joins every irregular window
        for (int j = n_doy - r - 1; j >= n_doy - 2 * r; j--) {                   // first runs: day d = j + r
            const int d = j + r, j2 = 2 * n_doy - d - r;                         // its second run starts at row j2 (n_doy: none)
            if (j2 >= n_doy) step(j, 0, -1, -1, d);
            else if (side && j2 == n_doy - 1) { step(j, 0, -1, -1, -1); step(n_doy - 1, 2, -1, -1, d); }
            else step(j, 0, slot_of(j2), -1, d);
        }
    } else {
        // (b) any other irregular window: its rows one by one from the window table
        for (int d : irr)
            for (int j = 0; j < W; j++) step(win_rows[(size_t)d * W + j], j == 0 ? 1 : 0, -1, -1, j == W - 1 ? d : -1);
    }
    g.n_irr_steps = (int)(pl.irr_day.size() / 4);
    net_set_cells(pl, C);
    pl.usable = true;
}

// The irregular days: one warp per tile behind (independent of) the regular items.
template <int NY, int K, int M, bool kPads>
static int net_launch_irr(const NetPlan &pl, const NetTables &tb, const float *x, int64_t C, int64_t ld_t, double *out,
                          const NetHandOver &hand, cudaStream_t st)
{
    const NetGeom &g = pl.geo;
    if (g.n_irr_steps == 0 || g.n_tiles == 0) return HDP_B200_OK;
    HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_net_irr<NY, K, M, kPads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    k_thr_net_irr<NY, K, M, kPads><<<(unsigned)g.n_tiles, 32, g.smem, st>>>(x, C, (uint32_t)(ld_t * sizeof(float)), tb.seq_time, tb.win_day,
                                                                           tb.irr_day, tb.irr_time, g, pl.sel, out, hand);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

template <int NY, int K, int M, bool kPads>
static int net_launch_t(const NetPlan &pl, const NetTables &tb, const float *x, int64_t C, int64_t ld_t, double *out,
                        const NetHandOver &hand, cudaStream_t st, bool allow_tmem)
{
    const NetGeom &g = pl.geo;
    if (allow_tmem && g.tm_warps > 0 && tb.next_item) {
        const int64_t items = g.n_tiles * g.n_chunks;
        if (ld_t <= 0 || ld_t >= (1LL << 30) || g.n_tiles > 0x7fffffffLL) return HDP_B200_ERR_UNSUPPORTED;
        int dev = 0, sms = 0;
        HDP_CUDA_TRY(cudaGetDevice(&dev));
        HDP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t smem = (size_t)g.tm_warps * g.tm_smem_slots * g.K * 32 * sizeof(float);
        HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_net_tm<NY, K, M, kPads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HDP_CUDA_TRY(cudaMemsetAsync(tb.next_item, 0, sizeof(unsigned long long), st));
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(sms, (items + g.tm_warps - 1) / g.tm_warps));
        k_thr_net_tm<NY, K, M, kPads><<<grid, g.tm_warps * 32, smem, st>>>(x, C, (uint32_t)(ld_t * sizeof(float)), tb.seq_time, tb.win_day,
                                                                          tb.irr_day, tb.irr_time, g, pl.sel, out, hand, tb.next_item);
        HDP_LAUNCH_CHECK();
        return net_launch_irr<NY, K, M, kPads>(pl, tb, x, C, ld_t, out, hand, st);
    }
    HDP_CUDA_TRY(cudaFuncSetAttribute(k_thr_net<NY, K, M, kPads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    const int64_t items = g.n_tiles * g.n_chunks;
    if (items > 0x7fffffffLL || ld_t <= 0 || ld_t >= (1LL << 30)) return HDP_B200_ERR_UNSUPPORTED;
    k_thr_net<NY, K, M, kPads><<<(unsigned)items, 32, g.smem, st>>>(x, C, (uint32_t)(ld_t * sizeof(float)), tb.seq_time, tb.win_day, tb.irr_day,
                                                                     tb.irr_time, g, pl.sel, out, hand);
    HDP_LAUNCH_CHECK();
    return net_launch_irr<NY, K, M, kPads>(pl, tb, x, C, ld_t, out, hand, st);
}

int net_launch(const NetPlan &pl, const NetTables &tb, const float *x, int64_t C, int64_t ld_t, double *out, const NetHandOver &hand, cudaStream_t st,
               bool allow_tmem)
{
    const NetGeom &g = pl.geo;
#define HDP_NET_CASE(ny, k, m)                                                                             \
    if (g.NY == ny && g.K == k && g.M == m)                                                                \
        return net_launch_t<ny, k, m, ny != 30>(pl, tb, x, C, ld_t, out, hand, st, allow_tmem);
    HDP_NET_INSTANCES(HDP_NET_CASE)
#undef HDP_NET_CASE
    return HDP_B200_ERR_UNSUPPORTED;
}

}  // namespace hdp
