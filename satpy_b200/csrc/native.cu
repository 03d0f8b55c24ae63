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

// ---------------------------------------------------------------------------------------------------------------------
// RatioSharpenedRGB / SelfSharpenedRGB (satpy/composites/resolution.py:32-233) + the common channel mask of
// GenericCompositor (satpy/composites/core.py:489-490) in one pass:
//   ratio = high / low;  ratio = 1 where it is not finite or negative;  ratio = min(ratio, 1.5)       (:157-165)
//   out[hi] = high;  out[c] = band[c] * ratio for the other colours except the neutral one           (:143-154)
//   SelfSharpenedRGB: low = four_element_average(high): nanmean over 2 x 2 blocks anchored at the crop offset's
//   parity, edge-padded, repeated back to full resolution (_mean4, :168-192)
//   common mask: a pixel that is NaN in any of the three bands is NaN in all of them.
// One thread = one 2 x 2 block (the four pixels share the mean): two 8-byte accesses per band and row.
template <typename T>
struct SharpenArg {
    const T *band[3];
    const T *high;   // the high-resolution band (NULL: no sharpening, the bands are only stacked and masked)
    T *out;          // 3 x H x W
    int64_t H, W;
    int32_t hi, neutral, self_mean, row_off, col_off, common_mask;
};

template <typename T>
__global__ void __launch_bounds__(256) ratio_sharpen_kernel(const __grid_constant__ SharpenArg<T> A) {
    const int64_t bw = (A.W + A.col_off + 1) / 2, bh = (A.H + A.row_off + 1) / 2;
    const int64_t n = bw * bh;
    const T nan = (T)__int_as_float(0x7fc00000);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t bi = i / bw, bj = i - bi * bw;
        const int64_t r0 = 2 * bi - A.row_off, c0 = 2 * bj - A.col_off;
        // four-element average of the block's in-image pixels (edge padding repeats them: the mean is the same)
        T mean = nan;
        if (A.high && A.self_mean) {
            T sum = 0;
            int cnt = 0;
#pragma unroll
            for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {
                    const int64_t r = r0 + dr, c = c0 + dc;
                    if (r < 0 || r >= A.H || c < 0 || c >= A.W) continue;
                    const T v = A.high[r * A.W + c];
                    if (v == v) {
                        sum += v;
                        ++cnt;
                    }
                }
            if (cnt) mean = sum / (T)cnt;
        }
#pragma unroll
        for (int dr = 0; dr < 2; ++dr)
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int64_t r = r0 + dr, c = c0 + dc;
                if (r < 0 || r >= A.H || c < 0 || c >= A.W) continue;
                const int64_t p = r * A.W + c;
                T v[3] = {A.band[0][p], A.band[1][p], A.band[2][p]};
                if (A.high) {
                    const T h = A.high[p];
                    const T low = A.self_mean ? mean : v[A.hi];
                    T ratio = h / low;
                    const bool finite = sizeof(T) == 4 ? (fabsf((float)ratio) <= 3.4028235e38f)
                                                       : (fabs((double)ratio) <= 1.7976931348623157e308);   // NaN fails too
                    if (!finite || ratio < (T)0) ratio = (T)1;
                    ratio = ratio > (T)1.5 ? (T)1.5 : ratio;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (k == A.hi) v[k] = h;
                        else if (k != A.neutral) v[k] = v[k] * ratio;
                    }
                }
                if (A.common_mask && (v[0] != v[0] || v[1] != v[1] || v[2] != v[2])) v[0] = v[1] = v[2] = nan;
#pragma unroll
                for (int k = 0; k < 3; ++k) A.out[k * A.H * A.W + p] = v[k];
            }
    }
}

template <typename T>
static int ratio_sharpen(const T *red, const T *green, const T *blue, const T *high, T *out, int64_t H, int64_t W,
                         int32_t hi, int32_t neutral, int32_t self_mean, int32_t row_off, int32_t col_off,
                         int32_t common_mask, void *stream) {
    B200_CHECK_ARG(H >= 0 && W >= 0, "bad shape");
    B200_CHECK_ARG(hi >= 0 && hi <= 2 && neutral >= -1 && neutral <= 2, "colour index outside 0..2");
    B200_CHECK_ARG((row_off == 0 || row_off == 1) && (col_off == 0 || col_off == 1), "offsets are parities (0 or 1)");
    if (H * W == 0) return B200_OK;
    B200_CHECK_ARG(red && green && blue && out, "NULL data pointer");
    SharpenArg<T> A;
    A.band[0] = red, A.band[1] = green, A.band[2] = blue;
    A.high = high;
    A.out = out;
    A.H = H, A.W = W;
    A.hi = hi, A.neutral = neutral, A.self_mean = self_mean, A.row_off = row_off, A.col_off = col_off;
    A.common_mask = common_mask;
    const int64_t blocks = ((W + col_off + 1) / 2) * ((H + row_off + 1) / 2);
    ratio_sharpen_kernel<T><<<b200_grid_for(blocks, 256, 8), 256, 0, (cudaStream_t)stream>>>(A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_ratio_sharpen_f32(const float *red, const float *green, const float *blue, const float *high,
                                      float *out, int64_t H, int64_t W, int32_t hi, int32_t neutral, int32_t self_mean,
                                      int32_t row_off, int32_t col_off, int32_t common_mask, void *stream) {
    return ratio_sharpen<float>(red, green, blue, high, out, H, W, hi, neutral, self_mean, row_off, col_off, common_mask, stream);
}

extern "C" int b200_ratio_sharpen_f64(const double *red, const double *green, const double *blue, const double *high,
                                      double *out, int64_t H, int64_t W, int32_t hi, int32_t neutral, int32_t self_mean,
                                      int32_t row_off, int32_t col_off, int32_t common_mask, void *stream) {
    return ratio_sharpen<double>(red, green, blue, high, out, H, W, hi, neutral, self_mean, row_off, col_off, common_mask, stream);
}
