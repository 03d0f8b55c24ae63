// NativeResampler (integer-factor replicate / nan-mean aggregate) and the SimulatedGreen band combination on sm_100a
// ("next" rows of SURVEY.md section 8f: the steps either side of the sun-zenith correction in the ABI true-colour chain).
// Replaces
//   NativeResampler._expand_reduce -> _replicate / _repeat_by_factor   satpy/resample/native.py:41-56,129-159
//                                   -> _aggregate / _mean (np.nanmean)  satpy/resample/native.py:101-127
//   SimulatedGreen.__call__: c01 * f1 + c02 * f2 + c03 * f3            satpy/composites/abi.py:56-61
#include "b200_common.cuh"

template <typename T>
__global__ void __launch_bounds__(256)
native_replicate_kernel(const T *__restrict__ data, T *__restrict__ out, int64_t nbands, int64_t rows, int64_t cols,
                        int fy, int fx) {
    const int64_t ocols = cols * fx, orows = rows * fy;
    const int64_t n = nbands * orows * ocols;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t oc = i % ocols, t = i / ocols, orow = t % orows, b = t / orows;
        out[i] = __ldg(data + (b * rows + orow / fy) * cols + oc / fx);
    }
}

extern "C" int b200_native_replicate(const void *data, void *out, int64_t nbands, int64_t rows, int64_t cols,
                                     int32_t fy, int32_t fx, int32_t elem_size, void *stream) {
    B200_CHECK_ARG(nbands >= 1 && rows >= 0 && cols >= 0, "bad shape");
    B200_CHECK_ARG(fy >= 1 && fx >= 1, "Expand factor must be a whole number");
    B200_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8, "elem_size must be 1, 2, 4 or 8");
    const int64_t n = nbands * rows * fy * cols * fx;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = b200_grid_for(n, 256, 8);
    switch (elem_size) {
    case 1: native_replicate_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t *)data, (uint8_t *)out, nbands, rows, cols, fy, fx); break;
    case 2: native_replicate_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)data, (uint16_t *)out, nbands, rows, cols, fy, fx); break;
    case 4: native_replicate_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t *)data, (uint32_t *)out, nbands, rows, cols, fy, fx); break;
    default: native_replicate_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t *)data, (uint64_t *)out, nbands, rows, cols, fy, fx); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// np.nanmean over fy x fx blocks of a 2-D float32 image (an all-NaN block gives NaN)
__global__ void __launch_bounds__(256)
native_aggregate_kernel(const float *__restrict__ data, float *__restrict__ out, int64_t orows, int64_t ocols, int fy,
                        int fx) {
    const int64_t n = orows * ocols, cols = ocols * fx;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t oc = i % ocols, orow = i / ocols;
        float sum = 0.0f;
        int cnt = 0;
        for (int a = 0; a < fy; ++a) {
            const float *row = data + (orow * fy + a) * cols + oc * fx;
            for (int b = 0; b < fx; ++b) {
                const float v = __ldg(row + b);
                if (!isnan(v)) {
                    sum = __fadd_rn(sum, v);
                    ++cnt;
                }
            }
        }
        st_stream_f1(out + i, cnt ? __fdiv_rn(sum, (float)cnt) : __int_as_float(0x7fc00000));
    }
}

extern "C" int b200_native_aggregate_f32(const float *data, float *out, int64_t out_rows, int64_t out_cols, int32_t fy,
                                         int32_t fx, void *stream) {
    B200_CHECK_ARG(out_rows >= 0 && out_cols >= 0, "bad shape");
    B200_CHECK_ARG(fy >= 1 && fx >= 1, "Aggregation factors are not integers");
    if (out_rows * out_cols == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    native_aggregate_kernel<<<b200_grid_for(out_rows * out_cols, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        data, out, out_rows, out_cols, fy, fx);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// c01 * f1 + c02 * f2 + c03 * f3 in float32, numpy's order of operations
__global__ void __launch_bounds__(256)
simulated_green_kernel(const float *__restrict__ c01, const float *__restrict__ c02, const float *__restrict__ c03,
                       float *__restrict__ out, int64_t n, float f1, float f2, float f3) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        st_stream_f1(out + i, __fadd_rn(__fadd_rn(__fmul_rn(c01[i], f1), __fmul_rn(c02[i], f2)), __fmul_rn(c03[i], f3)));
}

extern "C" int b200_simulated_green(const float *c01, const float *c02, const float *c03, float *out, int64_t n,
                                    float f1, float f2, float f3, void *stream) {
    B200_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(c01 && c02 && c03 && out, "NULL data pointer");
    simulated_green_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(c01, c02, c03, out, n, f1, f2, f3);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
