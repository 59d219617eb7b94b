// host.cu - the *_host entry points: the same two paths with HOST buffers.
//
// Cells are independent in both paths (the reference parallelises over spatial Dask blocks,
// hdp/threshold.py:161, hdp/metric.py:444), so the host variants cut the cell axis into chunks and run
// H2D copy -> kernels -> D2H copy of successive chunks on alternating streams: the copies of one chunk
// overlap the kernels of the other.  Pinned host buffers get full PCIe rate; pageable ones work too.
#include <algorithm>

#include "common.cuh"

namespace hdp {

constexpr int kSlots = 2;

struct DeviceBuf {
    void *p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return cuda_status(cudaMalloc(&p, bytes ? bytes : 1)); }
};

struct Streams {
    cudaStream_t s[kSlots] = {};
    int n = 0;
    ~Streams() { for (int i = 0; i < n; i++) cudaStreamDestroy(s[i]); }
    int create() {
        for (; n < kSlots; n++) HDP_CUDA_TRY(cudaStreamCreateWithFlags(&s[n], cudaStreamNonBlocking));
        return HDP_B200_OK;
    }
};

// Copies cells [c0, c0+nc) of a host measure array to a dense device buffer and reports the strides of
// the device copy.  Supported host layouts: cell-contiguous (ld_c == 1) and time-contiguous (ld_t == 1).
static int upload_cells(const float *h, int64_t T, int64_t ld_t, int64_t ld_c, int64_t c0, int64_t nc,
                        float *d, int64_t *d_ld_t, int64_t *d_ld_c, cudaStream_t st)
{
    if (ld_c == 1) {
        HDP_CUDA_TRY(cudaMemcpy2DAsync(d, nc * sizeof(float), h + c0, ld_t * sizeof(float), nc * sizeof(float), T,
                                       cudaMemcpyHostToDevice, st));
        *d_ld_t = nc; *d_ld_c = 1;
    } else if (ld_t == 1) {
        HDP_CUDA_TRY(cudaMemcpy2DAsync(d, T * sizeof(float), h + c0 * ld_c, ld_c * sizeof(float), T * sizeof(float), nc,
                                       cudaMemcpyHostToDevice, st));
        *d_ld_t = 1; *d_ld_c = T;
    } else {
        return HDP_B200_ERR_UNSUPPORTED;
    }
    return HDP_B200_OK;
}

static int64_t pick_chunk(int64_t C, size_t bytes_per_cell)
{
    // ~1.5 GB of device buffers per slot, a multiple of 32 cells, at least 2 chunks when there is enough work
    int64_t chunk = (int64_t)((size_t)1536 << 20) / (int64_t)std::max<size_t>(bytes_per_cell, 1);
    chunk = std::max<int64_t>(32, chunk / 32 * 32);
    if (C > 4096) chunk = std::min(chunk, ((C + 1) / 2 + 31) / 32 * 32);
    return std::min(chunk, std::max<int64_t>(C, 1));
}

}  // namespace hdp

using namespace hdp;

extern "C" {

int hdp_b200_thresholds_host(const float *h_temps, int64_t C, int64_t T_b, int64_t ld_t, int64_t ld_c,
                             const int32_t *h_time_index, const int32_t *h_win_rows, int n_doy, int n_y, int W,
                             const double *h_q, int P, double *h_out)
{
    if (C < 0 || T_b <= 0 || n_doy <= 0 || n_y <= 0 || W <= 0 || P <= 0) return HDP_B200_ERR_INVALID;
    if (C == 0) return HDP_B200_OK;
    if (!h_temps || !h_out) return HDP_B200_ERR_INVALID;
    if (ld_c != 1 && ld_t != 1) return HDP_B200_ERR_UNSUPPORTED;
    const size_t out_per_cell = (size_t)n_doy * P * sizeof(double);
    const int64_t chunk = pick_chunk(C, (size_t)T_b * 4 * 2 + out_per_cell);
    const int64_t dl_t = ld_c == 1 ? chunk : 1, dl_c = ld_c == 1 ? 1 : T_b;
    const size_t ws_bytes = hdp_b200_thresholds_workspace_bytes(chunk, T_b, dl_t, dl_c, n_doy, n_y, W, P);
    Streams ss;
    int rc = ss.create();
    if (rc) return rc;
    DeviceBuf x[kSlots], out[kSlots], ws[kSlots];
    for (int i = 0; i < kSlots; i++) {
        if ((rc = x[i].alloc((size_t)chunk * T_b * sizeof(float)))) return rc;
        if ((rc = out[i].alloc((size_t)chunk * out_per_cell))) return rc;
        if ((rc = ws[i].alloc(ws_bytes))) return rc;
    }
    int slot = 0;
    for (int64_t c0 = 0; c0 < C; c0 += chunk, slot = (slot + 1) % kSlots) {
        const int64_t nc = std::min(chunk, C - c0);
        cudaStream_t st = ss.s[slot];
        int64_t a, b;
        if ((rc = upload_cells(h_temps, T_b, ld_t, ld_c, c0, nc, (float *)x[slot].p, &a, &b, st))) return rc;
        rc = hdp_b200_thresholds((const float *)x[slot].p, nc, T_b, a, b, h_time_index, h_win_rows, n_doy, n_y, W, h_q, P,
                                 (double *)out[slot].p, ws[slot].p, ws_bytes, st);
        if (rc) return rc;
        HDP_CUDA_TRY(cudaMemcpyAsync(h_out + (size_t)c0 * n_doy * P, out[slot].p, (size_t)nc * out_per_cell,
                                     cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kSlots; i++) HDP_CUDA_TRY(cudaStreamSynchronize(ss.s[i]));
    return HDP_B200_OK;
}

int hdp_b200_metrics_host(const float *h_measure, int64_t C, int64_t T, int64_t ld_t, int64_t ld_c,
                          const double *h_thr, int n_doy, int P, const int32_t *h_doy_map,
                          const int32_t *h_defs, int D,
                          const int32_t *h_season_north, const int32_t *h_season_south, int Y,
                          const uint8_t *h_is_south, uint16_t *h_out)
{
    if (C < 0 || T <= 0 || n_doy <= 0 || P <= 0 || D <= 0 || Y < 0 || !h_doy_map) return HDP_B200_ERR_INVALID;
    if (C == 0 || Y == 0) return HDP_B200_OK;
    if (!h_measure || !h_thr || !h_out) return HDP_B200_ERR_INVALID;
    if (ld_c != 1 && ld_t != 1) return HDP_B200_ERR_UNSUPPORTED;
    const size_t thr_per_cell = (size_t)n_doy * P * sizeof(double);
    const size_t rows = (size_t)4 * P * D * Y;                                 // output rows of C cells each
    const int64_t chunk = pick_chunk(C, (size_t)T * 4 * 2 + (size_t)T / 8 * P * 2 + thr_per_cell + rows * 2);
    const int64_t dl_t = ld_c == 1 ? chunk : 1, dl_c = ld_c == 1 ? 1 : T;
    const size_t ws_bytes = hdp_b200_metrics_workspace_bytes(chunk, T, dl_t, dl_c, n_doy, P, D, Y, h_doy_map);
    Streams ss;
    int rc = ss.create();
    if (rc) return rc;
    DeviceBuf x[kSlots], thr[kSlots], south[kSlots], out[kSlots], ws[kSlots];
    for (int i = 0; i < kSlots; i++) {
        if ((rc = x[i].alloc((size_t)chunk * T * sizeof(float)))) return rc;
        if ((rc = thr[i].alloc((size_t)chunk * thr_per_cell))) return rc;
        if ((rc = south[i].alloc((size_t)chunk))) return rc;
        if ((rc = out[i].alloc(rows * chunk * sizeof(uint16_t)))) return rc;
        if ((rc = ws[i].alloc(ws_bytes))) return rc;
    }
    int slot = 0;
    for (int64_t c0 = 0; c0 < C; c0 += chunk, slot = (slot + 1) % kSlots) {
        const int64_t nc = std::min(chunk, C - c0);
        cudaStream_t st = ss.s[slot];
        int64_t a, b;
        if ((rc = upload_cells(h_measure, T, ld_t, ld_c, c0, nc, (float *)x[slot].p, &a, &b, st))) return rc;
        HDP_CUDA_TRY(cudaMemcpyAsync(thr[slot].p, h_thr + (size_t)c0 * n_doy * P, (size_t)nc * thr_per_cell,
                                     cudaMemcpyHostToDevice, st));
        if (h_is_south) HDP_CUDA_TRY(cudaMemcpyAsync(south[slot].p, h_is_south + c0, (size_t)nc, cudaMemcpyHostToDevice, st));
        rc = hdp_b200_metrics((const float *)x[slot].p, nc, T, a, b, (const double *)thr[slot].p, n_doy, P, h_doy_map,
                              h_defs, D, h_season_north, h_season_south, Y,
                              h_is_south ? (const uint8_t *)south[slot].p : nullptr,
                              (uint16_t *)out[slot].p, ws[slot].p, ws_bytes, st);
        if (rc) return rc;
        // device chunk is [rows, nc]; host array is [rows, C]
        HDP_CUDA_TRY(cudaMemcpy2DAsync(h_out + c0, (size_t)C * sizeof(uint16_t), out[slot].p, (size_t)nc * sizeof(uint16_t),
                                       (size_t)nc * sizeof(uint16_t), rows, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kSlots; i++) HDP_CUDA_TRY(cudaStreamSynchronize(ss.s[i]));
    return HDP_B200_OK;
}

}  // extern "C"
