// AHI HSD, SEVIRI and VIIRS SDR calibration on sm_100a ("next" rows of SURVEY.md section 8f; same elementwise,
// HBM-bound family as the ABI kernel in calibrate.cu).  Replaces the per-pixel arithmetic of
//   AHIHSDFileHandler._mask_invalid / convert_to_radiance / _vis_calibrate / _ir_calibrate
//                                                             satpy/readers/ahi_hsd.py:607-611,688-750
//   SEVIRICalibrationHandler.calibrate + SEVIRICalibrationAlgorithm   satpy/readers/core/seviri.py:574-633,660-686
//   mask_fill_values + _apply_factors (per granule)           satpy/readers/core/viirs_atms_sdr.py:197-214,328-361
// keeping numpy's float32 operation order (explicit round-to-nearest, no FMA contraction).  Golden vectors:
// satpy/tests/reader_tests/test_ahi_hsd.py:455-506, test_seviri_l1b_calibration.py:29-149.
#include <new>
#include "b200_common.cuh"
#include "b200_reduce.cuh"

#define NANF __int_as_float(0x7fc00000)

struct AhiArg {
    b200_ahi_cal_t c;
};

__device__ __forceinline__ float ahi_one(unsigned int cnt, const b200_ahi_cal_t &c) {
    float f = (float)cnt;
    if (c.has_invalid && ((int)cnt == c.count_outside_scan || (int)cnt == c.count_error)) f = NANF;
    float rad = __fadd_rn(__fmul_rn(f, c.gain), c.offset);
    if (c.mode == 0) return rad;
    if (c.mode == 1) {
        float r = __fmul_rn(__fmul_rn(rad, c.coeff), 100.0f);
        return r < 0.0f ? 0.0f : r;  // clip(0) keeps NaN
    }
    if (rad == 0.0f) rad = NANF;
    if (c.bt_in_f64) {
        // header constants are numpy float64 scalars in a real file: numpy evaluates the Planck inversion in float64
        double d = (double)rad;
        double b = c.num / (d * 1.0e6 * c.cwl5) + 1.0;
        double te = c.a / log(b);
        double r = c.c0 + c.c1 * te + c.c2 * (te * te);
        return (float)(r < 0.0 ? 0.0 : r);
    }
    float b = __fadd_rn(__fdiv_rn((float)c.num, __fmul_rn(__fmul_rn(rad, 1.0e6f), (float)c.cwl5)), 1.0f);
    float te = __fdiv_rn((float)c.a, logf(b));
    float r = __fadd_rn(__fadd_rn((float)c.c0, __fmul_rn((float)c.c1, te)), __fmul_rn((float)c.c2, __fmul_rn(te, te)));
    return r < 0.0f ? 0.0f : r;
}

// 8 counts (one 128-bit load) per thread per iteration, two 128-bit stores
__global__ void __launch_bounds__(256)
ahi_calibrate_kernel(const uint16_t *__restrict__ counts, float *__restrict__ out, int64_t n, int vec,
                     const __grid_constant__ AhiArg P) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = vec ? n >> 3 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        int4 a = ld_stream_v4(reinterpret_cast<const int4 *>(counts) + i);
        unsigned w[4] = {(unsigned)a.x, (unsigned)a.y, (unsigned)a.z, (unsigned)a.w};
        float r[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[2 * k] = ahi_one(w[k] & 0xffffu, P.c);
            r[2 * k + 1] = ahi_one(w[k] >> 16, P.c);
        }
        st_stream_f4(out + 8 * i, r[0], r[1], r[2], r[3]);
        st_stream_f4(out + 8 * i + 4, r[4], r[5], r[6], r[7]);
    }
    for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = ahi_one(counts[i], P.c);
}

extern "C" int b200_ahi_calibrate(const uint16_t *counts, float *out, int64_t n, const b200_ahi_cal_t *cal, void *stream) {
    B200_CHECK_ARG(cal != nullptr, "cal is NULL");
    B200_CHECK_ARG(n >= 0, "negative size");
    B200_CHECK_ARG(cal->mode >= 0 && cal->mode <= 2, "mode must be 0 (radiance), 1 (reflectance) or 2 (BT)");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(counts != nullptr && out != nullptr, "NULL data pointer");
    AhiArg P;
    P.c = *cal;
    int vec = (((uintptr_t)counts & 15) == 0) && (((uintptr_t)out & 15) == 0);
    ahi_calibrate_kernel<<<b200_grid_for((n + 7) / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(counts, out, n, vec, P);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

struct SevArg {
    b200_seviri_cal_t c;
};

__device__ __forceinline__ float seviri_one(float f, const b200_seviri_cal_t &c) {
    if (!(f > 0.0f)) f = NANF;  // data.where(data > 0)
    float rad = __fadd_rn(__fmul_rn(f, c.gain), c.offset);
    rad = rad < 0.0f ? 0.0f : rad;  // clip(0.0, None) keeps NaN
    if (c.mode == 0) return rad;
    if (c.mode == 1) {
        // np.pi * data * 100.0 / F, then * float32(d * d)
        float r = __fdiv_rn(__fmul_rn(__fmul_rn(c.pi_f, rad), 100.0f), c.solar_irradiance);
        return __fmul_rn(r, c.sun_earth_dist2);
    }
    // _tl15: (C2 * nu) / log((1.0 / data) * C1 * nu**3 + 1.0)
    float arg = __fadd_rn(__fmul_rn(__fmul_rn(__fdiv_rn(1.0f, rad), c.c1), c.nu3), 1.0f);
    float t = __fdiv_rn(c.c2nu, logf(arg));
    if (c.mode == 2) return __fdiv_rn(__fsub_rn(t, c.beta), c.alpha);                               // effective
    return __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(c.fit_a, t), t), __fmul_rn(c.fit_b, t)), c.fit_c);  // spectral
}

template <typename T>
__global__ void __launch_bounds__(256)
seviri_calibrate_kernel(const T *__restrict__ counts, float *__restrict__ out, int64_t n, const __grid_constant__ SevArg P) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        st_stream_f1(out + i, seviri_one((float)counts[i], P.c));
}

extern "C" int b200_seviri_calibrate(const void *counts, int32_t counts_are_f32, float *out, int64_t n,
                                     const b200_seviri_cal_t *cal, void *stream) {
    B200_CHECK_ARG(cal != nullptr, "cal is NULL");
    B200_CHECK_ARG(n >= 0, "negative size");
    B200_CHECK_ARG(cal->mode >= 0 && cal->mode <= 3, "mode must be 0 (radiance), 1 (reflectance), 2 (effective BT), 3 (spectral BT)");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(counts != nullptr && out != nullptr, "NULL data pointer");
    SevArg P;
    P.c = *cal;
    cudaStream_t s = (cudaStream_t)stream;
    if (counts_are_f32)
        seviri_calibrate_kernel<float><<<b200_grid_for(n, 256, 8), 256, 0, s>>>((const float *)counts, out, n, P);
    else
        seviri_calibrate_kernel<uint16_t><<<b200_grid_for(n, 256, 8), 256, 0, s>>>((const uint16_t *)counts, out, n, P);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// VIIRS SDR: fill masking, then data * factor + offset with one (factor, offset) pair per granule of rows
template <typename T>
__global__ void __launch_bounds__(256)
viirs_scale_kernel(const T *__restrict__ data, float *__restrict__ out, int64_t rows, int64_t cols,
                   const float2 *__restrict__ factors, const int32_t *__restrict__ row_granule, int fill_min_int,
                   float fill_max_float) {
    const int64_t n = rows * cols;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / cols;
        float2 f = __ldg(factors + __ldg(row_granule + r));
        if (!(f.x > -999.0f)) f.x = NANF;  // _mask_and_reshape_factors
        if (!(f.y > -999.0f)) f.y = NANF;
        float v;
        if (sizeof(T) == 2) {
            int c = (int)data[i];
            v = c < fill_min_int ? (float)c : NANF;
        } else {
            v = (float)data[i];
            if (!(v > fill_max_float)) v = NANF;
        }
        st_stream_f1(out + i, __fadd_rn(__fmul_rn(v, f.x), f.y));
    }
}

extern "C" int b200_viirs_scale(const void *data, int32_t data_is_f32, float *out, int64_t rows, int64_t cols,
                                const float *factors, const int32_t *rows_per_granule, int32_t n_granules,
                                int32_t fill_min_int, float fill_max_float, void *stream) {
    B200_CHECK_ARG(rows >= 0 && cols >= 0 && n_granules >= 1, "bad shape");
    B200_CHECK_ARG(factors != nullptr && rows_per_granule != nullptr, "NULL factor table");
    int64_t total = 0;
    for (int g = 0; g < n_granules; ++g) {
        B200_CHECK_ARG(rows_per_granule[g] >= 0, "negative granule size");
        total += rows_per_granule[g];
    }
    B200_CHECK_ARG(total == rows, "rows_per_granule does not add up to the number of rows");
    if (rows * cols == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *h_map = new (std::nothrow) int32_t[rows];
    if (!h_map) {
        b200_set_error("b200_viirs_scale: out of host memory");
        return B200_ENOMEM;
    }
    int64_t r = 0;
    for (int g = 0; g < n_granules; ++g)
        for (int k = 0; k < rows_per_granule[g]; ++k) h_map[r++] = g;
    float2 *d_f = nullptr;
    int32_t *d_map = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_f, (size_t)n_granules * sizeof(float2), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_map, (size_t)rows * sizeof(int32_t), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, factors, (size_t)n_granules * sizeof(float2), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_map, h_map, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // host tables are pageable
    delete[] h_map;
    if (e == cudaSuccess) {
        const int grid = b200_grid_for(rows * cols, 256, 8);
        if (data_is_f32)
            viirs_scale_kernel<float><<<grid, 256, 0, s>>>((const float *)data, out, rows, cols, d_f, d_map, fill_min_int, fill_max_float);
        else
            viirs_scale_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)data, out, rows, cols, d_f, d_map, fill_min_int, fill_max_float);
        e = cudaGetLastError();
    }
    if (d_f) cudaFreeAsync(d_f, s);
    if (d_map) cudaFreeAsync(d_map, s);
    if (e != cudaSuccess) {
        b200_set_error("b200_viirs_scale: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
    }
    return B200_OK;
}

// SunZenithReducer: _sunzen_reduction_ndarray, satpy/modifiers/angles.py:594-624 (float32 data, float32 SZA [deg])
__global__ void __launch_bounds__(256)
sunz_reduce_kernel(const float *__restrict__ data, const float *__restrict__ sunz, float *__restrict__ out, int64_t n,
                   int nbands, int64_t band_stride, float limit, float max_sza, float strength) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float corr = b200_sunz_reduction(sunz[i], limit, max_sza, strength);
        for (int b = 0; b < nbands; ++b) st_stream_f1(out + b * band_stride + i, __fmul_rn(data[b * band_stride + i], corr));
    }
}

extern "C" int b200_sunz_reduce(const float *data, const float *sunz_deg, float *out, int32_t nbands, int64_t band_stride,
                                int64_t n, float limit_deg, float max_sza_deg, float strength, void *stream) {
    B200_CHECK_ARG(nbands >= 1 && n >= 0, "bad shape");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data && sunz_deg && out, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    sunz_reduce_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, sunz_deg, out, n, nbands,
                                                                                  band_stride, limit_deg, max_sza_deg, strength);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
