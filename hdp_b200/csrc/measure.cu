// measure.cu - the elementwise pre-pass of hdp.measure on the device ("next" row 2).
//
// Replaces (reference = AgentOxygen/HDP v1.0.2):
//   heat_index                 hdp/measure.py:61-94    @nb.vectorize([float32(float32, float32)])
//   kelvin_to_celsius / fahrenheit_to_celsius / celsius_to_fahrenheit      hdp/measure.py:10-58 (float32 array arithmetic)
//   the heat-index branch of format_standard_measures                      hdp/measure.py:183-194
//
// One thread per 4 consecutive elements (128-bit loads and stores); HBM-streaming, 12 bytes per element.
// The arithmetic follows the compiled reference operation by operation.  Inside the Numba kernel the float32 arguments meet
// float64 literals, so sums and products with a literal are float64, but temp**2, rel_humid**2 and rel_humid*temp involve
// float32 operands only and are rounded to float32 first ((rel_humid*temp)**2 twice); the result is rounded to float32 on
// return.  The unit conversions around it are float32 array arithmetic in the reference (NumPy keeps float32 when the other
// operand is a Python scalar).  Every operation below is an explicitly rounded intrinsic: nothing may be contracted.
#include "common.cuh"

namespace hdp {

__device__ __forceinline__ float heat_index_f(float temp, float rel_humid)
{
    const double t = (double)temp, rh = (double)rel_humid;
    // hi = 0.5 * (temp + 61.0 + ((temp - 68.0)*1.2) + (rel_humid*0.094))
    double hi = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dadd_rn(t, 61.0), __dmul_rn(__dsub_rn(t, 68.0), 1.2)), __dmul_rn(rh, 0.094)));
    if (hi > 80.0) {
        const double t_sq = (double)__fmul_rn(temp, temp), rh_sq = (double)__fmul_rn(rel_humid, rel_humid);
        const float rt = __fmul_rn(rel_humid, temp);
        const double rt_sq = (double)__fmul_rn(rt, rt);
        hi = -42.379;
        hi = __dadd_rn(hi, __dmul_rn(2.04901523, t));
        hi = __dadd_rn(hi, __dmul_rn(10.14333127, rh));
        hi = __dadd_rn(hi, __dmul_rn(__dmul_rn(-0.22475541, t), rh));
        hi = __dadd_rn(hi, __dmul_rn(-0.00683783, t_sq));
        hi = __dadd_rn(hi, __dmul_rn(-0.05481717, rh_sq));
        hi = __dadd_rn(hi, __dmul_rn(__dmul_rn(0.00122874, t_sq), rh));
        hi = __dadd_rn(hi, __dmul_rn(__dmul_rn(0.00085282, t), rh_sq));
        hi = __dadd_rn(hi, __dmul_rn(-0.00000199, rt_sq));
        if (rh < 13.0 && t >= 80.0 && t <= 112.0)
            hi = __dsub_rn(hi, __dmul_rn(__ddiv_rn(__dsub_rn(13.0, rh), 4.0),
                                         __dsqrt_rn(__ddiv_rn(fabs(__dsub_rn(17.0, fabs(__dsub_rn(t, 95.0)))), 17.0))));
        else if (rh > 85.0 && t >= 80.0 && t <= 87.0)
            hi = __dadd_rn(hi, __dmul_rn(__ddiv_rn(__dsub_rn(rh, 85.0), 10.0), __ddiv_rn(__dsub_rn(87.0, t), 5.0)));
    }
    return __double2float_rn(hi);
}

// mode 0: out = heat_index(temp [F], rh [%])                                              (the reference's ufunc)
// mode 1: out = F->C(heat_index(C->F(temp [C]), rh [%] or rh [g/g] * 100))                 (format_standard_measures' branch)
// mode 2: out = to_celsius(temp, unit)                                                     (convert_temp_units)
template <int kMode>
__global__ void __launch_bounds__(256)
k_measure(const float *__restrict__ temp, const float *__restrict__ rh, int64_t n, int flag, float *__restrict__ out)
{
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= n) return;
    float t[4], r[4] = {0.0f, 0.0f, 0.0f, 0.0f}, o[4];
    const bool vec = i0 + 4 <= n && (((uintptr_t)temp | (uintptr_t)out | (kMode == 2 ? 0 : (uintptr_t)rh)) & 15) == 0;
    if (vec) {
        const float4 tv = __ldg((const float4 *)(temp + i0));
        t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w;
        if (kMode != 2) { const float4 rv = __ldg((const float4 *)(rh + i0)); r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w; }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            t[j] = i0 + j < n ? temp[i0 + j] : 0.0f;
            if (kMode != 2) r[j] = i0 + j < n ? rh[i0 + j] : 0.0f;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (kMode == 0) o[j] = heat_index_f(t[j], r[j]);
        else if (kMode == 1) {
            const float pct = flag ? __fmul_rn(r[j], 100.0f) : r[j];                        // g/g -> %, measure.py:183-185
            const float tf = __fadd_rn(__fmul_rn(t[j], 1.8f), 32.0f);                        // measure.py:55
            o[j] = __fdiv_rn(__fsub_rn(heat_index_f(tf, pct), 32.0f), 1.8f);                 // measure.py:37
        } else o[j] = to_celsius_f(t[j], flag);
    }
    if (vec) *(float4 *)(out + i0) = make_float4(o[0], o[1], o[2], o[3]);
    else
#pragma unroll
        for (int j = 0; j < 4; j++) if (i0 + j < n) out[i0 + j] = o[j];
}

template <int kMode>
static int launch_measure(const float *temp, const float *rh, int64_t n, int flag, float *out, cudaStream_t st)
{
    if (n < 0) return HDP_B200_ERR_INVALID;
    if (n == 0) return HDP_B200_OK;
    if (!temp || !out || (kMode != 2 && !rh)) return HDP_B200_ERR_INVALID;
    const int64_t blocks = (n + 1023) / 1024;
    if (blocks > 0x7fffffffLL) return HDP_B200_ERR_UNSUPPORTED;
    KernelTimer timer(kMeasure, st);
    k_measure<kMode><<<(unsigned)blocks, 256, 0, st>>>(temp, rh, n, flag, out);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

// Weighted mean over the cells of every row of a uint16 [rows, C] array (the metric planes [4, P, D, Y] x C):
//     out[row] = sum_c w[c] x[row, c] / sum_c w[c]
// which is what the reference's figure deck reduces its maps with (hdp/graphics/figure.py:14-15, weights cos(lat)).  One CTA per
// row; float64 accumulation in a fixed order (thread-strided partial sums, then a tree), so results are reproducible run to run.
__global__ void __launch_bounds__(256)
k_weighted_mean(const uint16_t *__restrict__ x, const double *__restrict__ w, int64_t C, double w_sum, double *__restrict__ out)
{
    __shared__ double part[256];
    const uint16_t *row = x + (int64_t)blockIdx.x * C;
    double acc = 0.0;
    for (int64_t c = threadIdx.x; c < C; c += 256) acc = __dadd_rn(acc, __dmul_rn(w[c], (double)row[c]));
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] = __dadd_rn(part[threadIdx.x], part[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = __ddiv_rn(part[0], w_sum);
}

int to_celsius_launch(const float *src, int64_t n, int unit, float *dst, cudaStream_t st)
{
    if (unit < 0 || unit > 2) return HDP_B200_ERR_INVALID;
    return launch_measure<2>(src, nullptr, n, unit, dst, st);
}

}  // namespace hdp

using namespace hdp;

extern "C" {

int hdp_b200_heat_index(const float *d_temp_f, const float *d_rh_pct, int64_t n, float *d_out_f, void *stream)
{
    return launch_measure<0>(d_temp_f, d_rh_pct, n, 0, d_out_f, (cudaStream_t)stream);
}

int hdp_b200_heat_index_measure(const float *d_temp_c, const float *d_rh, int64_t n, int rh_is_fraction, float *d_out_c, void *stream)
{
    return launch_measure<1>(d_temp_c, d_rh, n, rh_is_fraction ? 1 : 0, d_out_c, (cudaStream_t)stream);
}

int hdp_b200_weighted_mean(const uint16_t *d_x, int64_t rows, int64_t C, const double *d_w, double w_sum, double *d_out, void *stream)
{
    if (rows < 0 || C <= 0 || rows > 0x7fffffffLL || !(w_sum > 0.0)) return HDP_B200_ERR_INVALID;
    if (rows == 0) return HDP_B200_OK;
    if (!d_x || !d_w || !d_out) return HDP_B200_ERR_INVALID;
    KernelTimer timer(kMeasure, (cudaStream_t)stream);
    k_weighted_mean<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(d_x, d_w, C, w_sum, d_out);
    HDP_LAUNCH_CHECK();
    return HDP_B200_OK;
}

int hdp_b200_to_celsius(const float *d_temp, int64_t n, int unit, float *d_out_c, void *stream)
{
    if (unit < 0 || unit > 2) return HDP_B200_ERR_INVALID;
    return launch_measure<2>(d_temp, nullptr, n, unit, d_out_c, (cudaStream_t)stream);
}

}  // extern "C"
