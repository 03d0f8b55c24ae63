// Elliptical weighted averaging (EWA) resampling on sm_100a: ll2cr + fornav.
//
// Replaces pyresample.ewa (Cython `_ll2cr.pyx`, `_fornav.pyx` + C++ `_fornav_templates.cpp`, ms2gt lineage), which satpy
// exposes unchanged as the "ewa" resampler (satpy/resample/ewa.py:3-11); `rows_per_scan` comes from the reader
// (satpy/readers/core/viirs_atms_sdr.py:247-255,283).  PARITY UNPINNED in the reference (no test exercises "ewa");
// the algorithm follows the library's published one (restated in oracle/ewa.py):
//   ll2cr   swath lon/lat -> fractional (col, row) of the target area, NaN where not projectable;
//   fornav  per scan of `rows_per_scan` rows, per swath column: ellipse parameters (a, b, c, u_del, v_del) from the
//           col/row gradients; every swath pixel adds weight * value to the grid cells inside its ellipse,
//           weight = wtab[int(q * qfactor)], q = a du^2 + b du dv + c dv^2 evaluated incrementally along a grid row in
//           float32 exactly as the C++ loop does (so the table indices are the same); finally out = sum(w v) / sum(w)
//           where sum(w) >= weight_sum_min.
// Weights, ellipse parameters and accumulators are float32 like the library's typedefs.
//
// One thread = one swath pixel; neighbouring pixels of a warp overlap heavily in the grid, and their contributions go
// to the accumulation grids as float32 atomics (RED.ADD.F32 resolved in L2).  The accumulation order differs from the
// sequential C++ loop, which changes the float32 sums in their last bits only (north_star tolerance for EWA: 1e-4).
#include <new>
#include "b200_common.cuh"
#include "b200_proj.cuh"

int b200_check_geo(const b200_geo_t *geo, const char *who);

#define EWA_EPSILON 1e-8f

struct Ll2crArg {
    b200_geo_t swath;
    b200_proj_t proj;
    double ox, oy, cw, ch;
    double width, height;
};

template <typename T>
__global__ void __launch_bounds__(256)
ll2cr_kernel(const __grid_constant__ Ll2crArg A, T *__restrict__ cols, T *__restrict__ rows,
             unsigned long long *__restrict__ points_in_grid) {
    const int64_t W = A.swath.width;
    const int64_t n = A.swath.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int inside = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat, x, y;
        b200_geo_lonlat(A.swath, A.swath.row0 + r, i - r * W, lon, lat);
        bool ok = b200_proj_forward(A.proj, lon, lat, x, y);
        // the projected coordinates come back in the dtype of the lon/lat arrays; the grid arithmetic is in that dtype
        T xt = (T)x, yt = (T)y;
        T c = (T)NAN, rr = (T)NAN;
        if (ok && !(xt >= (T)1e30) && !(yt >= (T)1e30) && xt == xt && yt == yt) {
            c = (xt - (T)A.ox) / (T)A.cw;
            rr = (yt - (T)A.oy) / (T)A.ch;
            if (c >= (T)-1 && c <= (T)(A.width + 1) && rr >= (T)-1 && rr <= (T)(A.height + 1)) ++inside;
        }
        cols[i] = c;
        rows[i] = rr;
    }
    for (int o = 16; o > 0; o >>= 1) inside += __shfl_xor_sync(0xffffffffu, inside, o);
    if ((threadIdx.x & 31) == 0 && inside) atomicAdd(points_in_grid, (unsigned long long)inside);
}

extern "C" int b200_ll2cr(const b200_geo_t *swath, const b200_geo_t *dst, void *cols, void *rows,
                          int64_t *points_in_grid, void *stream) {
    int rc = b200_check_geo(swath, "b200_ll2cr");
    if (rc) return rc;
    rc = b200_check_geo(dst, "b200_ll2cr");
    if (rc) return rc;
    if (dst->is_swath) {
        b200_set_error("b200_ll2cr: the target must be an area");
        return B200_EUNSUPPORTED;
    }
    const int64_t n = swath->nrows * swath->width;
    if (points_in_grid) *points_in_grid = 0;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(cols != nullptr && rows != nullptr, "NULL output");
    cudaStream_t s = (cudaStream_t)stream;
    Ll2crArg A;
    A.swath = *swath;
    A.proj = dst->area.proj;
    A.cw = dst->area.pixel_size_x;
    A.ch = -fabs(dst->area.pixel_size_y);
    A.ox = dst->area.x_ul;  // centre of the top-left cell = extent[0] + cw / 2
    A.oy = dst->area.y_ul;
    A.width = (double)dst->area.width;
    A.height = (double)dst->area.height;
    unsigned long long *d_count = nullptr;
    B200_CUDA(cudaMallocAsync((void **)&d_count, sizeof(unsigned long long), s));
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s);
    if (e == cudaSuccess) {
        const bool f32 = swath->is_swath && swath->lonlat_is_f32;
        if (f32)
            ll2cr_kernel<float><<<b200_grid_for(n, 256, 8), 256, 0, s>>>(A, (float *)cols, (float *)rows, d_count);
        else
            ll2cr_kernel<double><<<b200_grid_for(n, 256, 8), 256, 0, s>>>(A, (double *)cols, (double *)rows, d_count);
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess && points_in_grid) {
        e = cudaMemcpyAsync(&h, d_count, sizeof(h), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        *points_in_grid = (int64_t)h;
    }
    cudaFreeAsync(d_count, s);
    if (e != cudaSuccess) {
        b200_set_error("b200_ll2cr: %s", cudaGetErrorString(e));
        return B200_ECUDA;
    }
    return B200_OK;
}

// ----------------------------------------------------------------------------------------------- fornav
struct EwaPix {  // per (scan, swath column)
    float a, b, c, u_del, v_del;
};

struct FornavArg {
    const void *cols, *rows;  // float32 or float64 images (swath_rows x swath_cols)
    int32_t cr_is_f32;
    int64_t swath_rows, swath_cols;
    int32_t rows_per_scan;
    float qmax, qfactor, distance_max, delta_max;
    int32_t weight_count;
    const float *wtab;
    int64_t grid_rows, grid_cols;
};

__device__ __forceinline__ float ewa_cr(const FornavArg &A, const void *img, int64_t i) {
    return A.cr_is_f32 ? ((const float *)img)[i] : (float)((const double *)img)[i];
}

// compute_ewa_parameters: one thread per (scan, column); edge columns take their neighbour's parameters
__global__ void __launch_bounds__(256) ewa_params_kernel(const __grid_constant__ FornavArg A, EwaPix *__restrict__ ewap) {
    const int64_t nscans = A.swath_rows / A.rows_per_scan;
    const int64_t n = nscans * A.swath_cols;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t scan = i / A.swath_cols;
        int64_t col = i - scan * A.swath_cols;
        col = min(max(col, (int64_t)1), A.swath_cols - 2);
        const int64_t base = scan * A.rows_per_scan * A.swath_cols;
        const int64_t mid = base + (int64_t)(A.rows_per_scan / 2) * A.swath_cols;
        const int64_t last = base + (int64_t)(A.rows_per_scan - 1) * A.swath_cols;
        const float rowsm1 = (float)(A.rows_per_scan - 1);
        const float dm = A.distance_max;
        float ux = __fmul_rn(__fdiv_rn(__fsub_rn(ewa_cr(A, A.cols, mid + col + 1), ewa_cr(A, A.cols, mid + col - 1)), 2.0f), dm);
        float vx = __fmul_rn(__fdiv_rn(__fsub_rn(ewa_cr(A, A.rows, mid + col + 1), ewa_cr(A, A.rows, mid + col - 1)), 2.0f), dm);
        float uy = __fmul_rn(__fdiv_rn(__fsub_rn(ewa_cr(A, A.cols, last + col), ewa_cr(A, A.cols, base + col)), rowsm1), dm);
        float vy = __fmul_rn(__fdiv_rn(__fsub_rn(ewa_cr(A, A.rows, last + col), ewa_cr(A, A.rows, base + col)), rowsm1), dm);
        EwaPix p;
        if (isnan(ux) || isnan(vx) || isnan(uy) || isnan(vy)) {
            p.a = p.b = p.c = 0.0f;
            p.u_del = p.v_del = dm;
        } else {
            float fs = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
            fs = __fmul_rn(fs, fs);
            if (fs < EWA_EPSILON) fs = EWA_EPSILON;
            fs = __fdiv_rn(A.qmax, fs);
            p.a = __fmul_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), fs);
            p.b = __fmul_rn(__fmul_rn(-2.0f, __fadd_rn(__fmul_rn(ux, vx), __fmul_rn(uy, vy))), fs);
            p.c = __fmul_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), fs);
            float d = __fsub_rn(__fmul_rn(__fmul_rn(4.0f, p.a), p.c), __fmul_rn(p.b, p.b));
            if (d < EWA_EPSILON) d = EWA_EPSILON;
            d = __fdiv_rn(__fmul_rn(4.0f, A.qmax), d);
            p.u_del = fminf(__fsqrt_rn(__fmul_rn(p.c, d)), A.delta_max);
            p.v_del = fminf(__fsqrt_rn(__fmul_rn(p.a, d)), A.delta_max);
        }
        ewap[i] = p;
    }
}

// compute_ewa: MODE 0 accumulate sum(w), sum(w v); MODE 1 maximum weight per cell; MODE 2 write the value whose weight
// equals the cell's maximum (second pass of maximum_weight_mode)
template <int MODE>
__global__ void __launch_bounds__(256)
fornav_kernel(const __grid_constant__ FornavArg A, const EwaPix *__restrict__ ewap, const float *__restrict__ data,
              int nbands, int64_t band_stride, float img_fill, float *__restrict__ grid_weights,
              float *__restrict__ grid_accums) {
    const int64_t n = A.swath_rows * A.swath_cols;
    const int64_t gsize = A.grid_rows * A.grid_cols;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t row = i / A.swath_cols, col = i - row * A.swath_cols;
        const EwaPix p = ewap[(row / A.rows_per_scan) * A.swath_cols + col];
        const float u0 = ewa_cr(A, A.cols, i), v0 = ewa_cr(A, A.rows, i);
        if (u0 < -p.u_del || v0 < -p.v_del || isnan(u0) || isnan(v0)) continue;
        // (int) truncates towards zero, like the C++ casts; clamp before converting so huge values stay defined
        const float big = 2.0e9f;
        long long iu1 = (long long)fminf(fmaxf(__fsub_rn(u0, p.u_del), -big), big);
        long long iu2 = (long long)fminf(fmaxf(__fadd_rn(u0, p.u_del), -big), big);
        long long iv1 = (long long)fminf(fmaxf(__fsub_rn(v0, p.v_del), -big), big);
        long long iv2 = (long long)fminf(fmaxf(__fadd_rn(v0, p.v_del), -big), big);
        if (iu1 < 0) iu1 = 0;
        if (iu2 >= A.grid_cols) iu2 = A.grid_cols - 1;
        if (iv1 < 0) iv1 = 0;
        if (iv2 >= A.grid_rows) iv2 = A.grid_rows - 1;
        if (!(iu1 < A.grid_cols && iu2 >= 0 && iv1 < A.grid_rows && iv2 >= 0)) continue;
        const float ddq = __fmul_rn(2.0f, p.a);
        const float u = __fsub_rn((float)iu1, u0);
        const float a2up1 = __fmul_rn(p.a, __fadd_rn(__fmul_rn(2.0f, u), 1.0f));
        const float bu = __fmul_rn(p.b, u);
        const float au2 = __fmul_rn(__fmul_rn(p.a, u), u);
        for (long long iv = iv1; iv <= iv2; ++iv) {
            const float v = __fsub_rn((float)iv, v0);
            float dq = __fadd_rn(a2up1, __fmul_rn(p.b, v));
            float q = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p.c, v), bu), v), au2);
            for (long long iu = iu1; iu <= iu2; ++iu) {
                if (q >= 0.0f && q < A.qmax) {
                    int iw = (int)__fmul_rn(q, A.qfactor);
                    if (iw >= A.weight_count) iw = A.weight_count - 1;
                    const float weight = __ldg(A.wtab + iw);
                    const int64_t cell = iv * A.grid_cols + iu;
                    for (int b = 0; b < nbands; ++b) {
                        const float val = __ldg(data + b * band_stride + i);
                        const bool bad = (val == img_fill) || isnan(val);
                        if (MODE == 0) {
                            if (!bad) {
                                atomicAdd(grid_weights + b * gsize + cell, weight);
                                atomicAdd(grid_accums + b * gsize + cell, __fmul_rn(val, weight));
                            }
                        } else if (MODE == 1) {
                            // weights are positive: their bit patterns order like the values
                            atomicMax((int *)(grid_weights + b * gsize + cell), __float_as_int(weight));
                        } else {
                            if (weight == grid_weights[b * gsize + cell])
                                grid_accums[b * gsize + cell] = bad ? __int_as_float(0x7fc00000) : val;
                        }
                    }
                }
                q = __fadd_rn(q, dq);
                dq = __fadd_rn(dq, ddq);
            }
        }
    }
}

extern "C" int b200_fornav_accum(const void *cols, const void *rows, int32_t cr_is_f32, int64_t swath_rows,
                                 int64_t swath_cols, int32_t rows_per_scan, const float *data, int32_t nbands,
                                 int64_t band_stride, float img_fill, int64_t grid_rows, int64_t grid_cols,
                                 float *grid_weights, float *grid_accums, const b200_ewa_params_t *params,
                                 void *stream) {
    B200_CHECK_ARG(params != nullptr, "params is NULL");
    B200_CHECK_ARG(swath_rows >= 0 && swath_cols >= 0 && grid_rows > 0 && grid_cols > 0 && nbands >= 1, "bad shape");
    B200_CHECK_ARG(params->weight_count >= 2, "weight_count must be at least 2");
    B200_CHECK_ARG(params->weight_min > 0.0f, "weight_min must be greater than 0");
    B200_CHECK_ARG(params->weight_distance_max > 0.0f, "weight_distance_max must be greater than 0");
    B200_CHECK_ARG(params->maximum_weight_mode >= 0 && params->maximum_weight_mode <= 3, "maximum_weight_mode must be 0..3");
    if (swath_rows * swath_cols == 0) return B200_OK;
    B200_CHECK_ARG(swath_cols >= 3, "fornav needs at least 3 swath columns");
    if (rows_per_scan <= 0) rows_per_scan = (int32_t)swath_rows;
    B200_CHECK_ARG(swath_rows % rows_per_scan == 0, "swath rows must be a multiple of rows_per_scan");
    B200_CHECK_ARG(rows_per_scan >= 2, "rows_per_scan must be at least 2");
    B200_CHECK_ARG(cols && rows && data && grid_weights && grid_accums, "NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    FornavArg A;
    A.cols = cols;
    A.rows = rows;
    A.cr_is_f32 = cr_is_f32;
    A.swath_rows = swath_rows;
    A.swath_cols = swath_cols;
    A.rows_per_scan = rows_per_scan;
    A.distance_max = params->weight_distance_max;
    A.delta_max = params->weight_delta_max;
    A.weight_count = params->weight_count;
    A.qmax = A.distance_max * A.distance_max;
    A.qfactor = (float)A.weight_count / A.qmax;
    A.grid_rows = grid_rows;
    A.grid_cols = grid_cols;
    // initialize_weight: wtab[i] = exp(-alpha * qmax * i / (count - 1)), float32 parameters, libm exp
    const float alpha = -logf(params->weight_min) / A.qmax;
    float *h_wtab = new (std::nothrow) float[A.weight_count];
    if (!h_wtab) {
        b200_set_error("b200_fornav_accum: out of host memory");
        return B200_ENOMEM;
    }
    for (int i = 0; i < A.weight_count; ++i) h_wtab[i] = (float)exp(-alpha * A.qmax * i / (A.weight_count - 1));
    float *wtab = nullptr;
    EwaPix *ewap = nullptr;
    const int64_t nscans = swath_rows / rows_per_scan;
    cudaError_t e = cudaMallocAsync((void **)&wtab, (size_t)A.weight_count * sizeof(float), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&ewap, (size_t)(nscans * swath_cols) * sizeof(EwaPix), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(wtab, h_wtab, (size_t)A.weight_count * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // h_wtab is pageable
    delete[] h_wtab;
    if (e == cudaSuccess) {
        A.wtab = wtab;
        const int64_t n = swath_rows * swath_cols;
        ewa_params_kernel<<<b200_grid_for(nscans * swath_cols, 256, 8), 256, 0, s>>>(A, ewap);
        const int grid = b200_grid_for(n, 256, 8);
        // maximum_weight_mode: 1 = both passes; 2 = only the maxima of the weights, 3 = only the samples whose weight
        // equals the maximum already in grid_weights -- so that several ranks can take the maximum of their weight
        // grids in between (max-by-weight reduction)
        const int mwm = params->maximum_weight_mode;
        if (mwm == 0)
            fornav_kernel<0><<<grid, 256, 0, s>>>(A, ewap, data, nbands, band_stride, img_fill, grid_weights, grid_accums);
        if (mwm == 1 || mwm == 2)
            fornav_kernel<1><<<grid, 256, 0, s>>>(A, ewap, data, nbands, band_stride, img_fill, grid_weights, grid_accums);
        if (mwm == 1 || mwm == 3)
            fornav_kernel<2><<<grid, 256, 0, s>>>(A, ewap, data, nbands, band_stride, img_fill, grid_weights, grid_accums);
        e = cudaGetLastError();
    }
    if (wtab) cudaFreeAsync(wtab, s);
    if (ewap) cudaFreeAsync(ewap, s);
    if (e != cudaSuccess) {
        b200_set_error("b200_fornav_accum: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
    }
    return B200_OK;
}

// write_grid_image: out = sum(w v) / sum(w) (or the kept value in maximum_weight_mode) where sum(w) >= weight_sum_min
__global__ void __launch_bounds__(256)
fornav_finalize_kernel(const float *__restrict__ grid_weights, const float *__restrict__ grid_accums,
                       float *__restrict__ out, int64_t n, float weight_sum_min, int max_mode, float fill,
                       unsigned long long *__restrict__ valid_counts, int64_t gsize) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < ((n + 31) & ~31LL); i0 += stride) {
        bool valid = false;
        if (i0 < n) {
            const float w = grid_weights[i0], acc = grid_accums[i0];
            float v;
            if (w < weight_sum_min || isnan(acc))
                v = __int_as_float(0x7fc00000);
            else
                v = max_mode ? acc : __fdiv_rn(acc, w);
            valid = !isnan(v);
            st_stream_f1(out + i0, valid ? v : fill);
        }
        // all lanes of a warp belong to the same band when gsize is a multiple of 32; count per lane otherwise
        unsigned m = __ballot_sync(0xffffffffu, valid);
        if ((gsize & 31) == 0) {
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(valid_counts + i0 / gsize, (unsigned long long)__popc(m));
        } else if (valid) {
            atomicAdd(valid_counts + i0 / gsize, 1ULL);
        }
    }
}

extern "C" int b200_fornav_finalize(const float *grid_weights, const float *grid_accums, float *out, int32_t nbands,
                                    int64_t grid_size, const b200_ewa_params_t *params, float fill,
                                    int64_t *valid_counts, void *stream) {
    B200_CHECK_ARG(params != nullptr, "params is NULL");
    B200_CHECK_ARG(nbands >= 1 && grid_size >= 0, "bad shape");
    if (valid_counts)
        for (int b = 0; b < nbands; ++b) valid_counts[b] = 0;
    if (grid_size == 0) return B200_OK;
    B200_CHECK_ARG(grid_weights && grid_accums && out, "NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    float wsm = params->weight_sum_min;
    if (wsm <= 0.0f) wsm = params->weight_min;  // documented: "if weight_sum_min <= 0 it is set to weight_min"
    unsigned long long *d_counts = nullptr;
    B200_CUDA(cudaMallocAsync((void **)&d_counts, (size_t)nbands * sizeof(unsigned long long), s));
    cudaError_t e = cudaMemsetAsync(d_counts, 0, (size_t)nbands * sizeof(unsigned long long), s);
    const int64_t n = (int64_t)nbands * grid_size;
    if (e == cudaSuccess) {
        fornav_finalize_kernel<<<b200_grid_for(n, 256, 8), 256, 0, s>>>(grid_weights, grid_accums, out, n, wsm,
                                                                       params->maximum_weight_mode, fill, d_counts,
                                                                       grid_size);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && valid_counts) {
        unsigned long long *h = new (std::nothrow) unsigned long long[nbands];
        if (h) {
            e = cudaMemcpyAsync(h, d_counts, (size_t)nbands * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            for (int b = 0; b < nbands; ++b) valid_counts[b] = (int64_t)h[b];
            delete[] h;
        }
    }
    cudaFreeAsync(d_counts, s);
    if (e != cudaSuccess) {
        b200_set_error("b200_fornav_finalize: %s", cudaGetErrorString(e));
        return B200_ECUDA;
    }
    return B200_OK;
}
