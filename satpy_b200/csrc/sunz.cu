// Sun-zenith angle computation and correction on sm_100a.
//
// Replaces, per pixel and in ONE pass (no lon/lat arrays ever touch HBM):
//   get_cos_sza / _get_valid_lonlats / _cos_zen_ndarray   satpy/modifiers/angles.py:418-469
//     (+ pyorbital astronomy.cos_zen, un-vendored; scalars gmst/ra/dec come from the host)
//   SunZenithCorrectorBase.__call__ masking                 satpy/modifiers/geometry.py:48-71
//   _sunzen_corr_cos_ndarray                                satpy/modifiers/angles.py:555-584
//   atmospheric_path_length_correction (Li & Shibata)       satpy/utils.py:278-316
//
// dtype flow of the reference, reproduced on purpose: lon/lat are computed in float64 by the
// projection, ROUNDED to float32 when the data is float32 (angles.py:427-429); numpy then does
// deg2rad, sin(lat), cos(lat) in float32 and everything that meets pyorbital's float64 scalars
// (hour angle, products, sum) in float64; the correction factor is cast back to float32
// (angles.py:565,580) before multiplying the data.
#include "b200_common.cuh"
#include "b200_proj.cuh"
#include "b200_cal.cuh"
#include "b200_geos_fast.cuh"

#define B200_DEG2RAD_F 0.017453292f  // numpy's float32 deg2rad factor

struct SunzDev {
    b200_geo_t geo;
    GeosFast fast;  // geostationary areas: table-driven inverse projection
    double gmst, ra, sin_dec, cos_dec;
    double limit_rad, limit_cos, max_rad, max_cos, span, li_lim;
    int32_t has_max, lonlat_f32, method, alg;
    // angle-free cos(sza) for geostationary areas (alg = 1): rotation by gmst - ra + lon_0 and the band of cos(sza)
    // inside which the reference's rounded-angle flow is reproduced instead
    double cos_a, sin_a, alg_lo, alg_hi;
};

// ALG: the caller only needs the CORRECTION (not cos(sza) itself) to 1e-5 relative.  Then, on geostationary areas with
// float32 data, cos(sza) comes straight from the view vector: the ellipsoid normal is n = (vx, vy, vz / (1 - es)) and
// cos(sza) = n . s / |n| with s the sun direction rotated into the satellite's meridian frame -- one rsqrt, no angles.
// It differs from the reference's value (lon/lat rounded to float32, float32 deg2rad / sin / cos) by at most 4.7e-7
// absolute, i.e. < 2.4e-6 relative in 1 / cos(sza) above alg_hi = 0.2; below alg_lo the result is masked / constant
// whatever the last digits are.  In between (the 78.5 .. 95 degree band around the terminator, where the fall-off is
// steep) the reference's flow is reproduced digit by digit as before.
// FAST: geostationary area (table-driven inverse projection); separate instantiations keep the general projection code,
// and its register budget, out of the hot configuration.
template <bool ALG, bool FAST>
__device__ __forceinline__ double b200_cos_sza_pixel(const SunzDev &S, int64_t row, int64_t col) {
    if (ALG && FAST && S.alg) {
        double vx, vy, vz;
        if (!geos_fast_vector(S.fast, row, col, vx, vy, vz)) return __longlong_as_double(0x7ff8000000000000LL);
        const double u = S.fast.inv2 * vz;
        const double num = fma(u, S.sin_dec, S.cos_dec * fma(S.cos_a, vx, -S.sin_a * vy));
        const double cz = num * rsqrt(fma(u, u, fma(vx, vx, vy * vy)));
        if (cz > S.alg_hi || cz < S.alg_lo) return cz;
    }
    double lon, lat;
    bool ok = FAST ? geos_fast_lonlat(S.fast, row, col, lon, lat) : b200_geo_lonlat(S.geo, row, col, lon, lat);
    if (!ok || !(lon < 1e30) || !(lat < 1e30)) return __longlong_as_double(0x7ff8000000000000LL);
    double sl, cl, lon_r;
    if (S.lonlat_f32) {
        float lonf = (float)lon, latf = (float)lat;
        float lon_rf = __fmul_rn(lonf, B200_DEG2RAD_F);
        float lat_rf = __fmul_rn(latf, B200_DEG2RAD_F);
        sl = (double)sinf(lat_rf);
        cl = (double)cosf(lat_rf);
        lon_r = (double)lon_rf;
    } else {
        double lat_r = lat * B200_DEG2RAD;
        sl = sin(lat_r);
        cl = cos(lat_r);
        lon_r = lon * B200_DEG2RAD;
    }
    double h = __dsub_rn(__dadd_rn(S.gmst, lon_r), S.ra);
    return __dadd_rn(__dmul_rn(sl, S.sin_dec), __dmul_rn(__dmul_rn(cl, S.cos_dec), cos(h)));
}

__device__ __forceinline__ double b200_li_shibata(double cz) {
    return 24.35 / (2.0 * cz + sqrt(498.5225 * cz * cz + 1.0));
}

// correction factor for float64 data: everything stays float64 (the reference's .astype(data.dtype) is a no-op)
__device__ __forceinline__ double b200_sunz_corr_d(const SunzDev &S, double cz) {
    if (S.has_max && !(cz >= S.max_cos)) cz = __longlong_as_double(0x7ff8000000000000LL);
    if (isnan(cz)) return 0.0;
    const double base_lim = S.method == 0 ? 1.0 / S.limit_cos : S.li_lim;
    if (cz > S.limit_cos) return S.method == 0 ? 1.0 / cz : b200_li_shibata(cz);
    if (!S.has_max) return base_lim;
    double gf = (acos(cz) - S.limit_rad) / S.span;
    gf = 1.0 - log(gf + 1.0) / 0.6931471805599453;
    gf = gf < 0.0 ? 0.0 : gf;
    return S.method == 0 ? gf / S.limit_cos : gf * S.li_lim;
}

// correction factor from a float64 cos(sza) (computed-angles branch)
__device__ __forceinline__ float b200_sunz_corr(const SunzDev &S, double cz) {
    if (S.has_max && !(cz >= S.max_cos)) cz = __longlong_as_double(0x7ff8000000000000LL);
    if (isnan(cz)) return 0.0f;
    if (S.method == 0) {
        if (cz > S.limit_cos) return (float)(1.0 / cz);
        if (!S.has_max) return (float)(1.0 / S.limit_cos);
        double gf = (acos(cz) - S.limit_rad) / S.span;
        gf = 1.0 - log(gf + 1.0) / 0.6931471805599453;
        gf = gf < 0.0 ? 0.0 : gf;
        return (float)(gf / S.limit_cos);
    }
    if (cz > S.limit_cos) return (float)b200_li_shibata(cz);
    if (!S.has_max) return (float)S.li_lim;
    double gf = (acos(cz) - S.limit_rad) / S.span;
    gf = 1.0 - log(gf + 1.0) / 0.6931471805599453;
    gf = gf < 0.0 ? 0.0 : gf;
    return (float)(gf * S.li_lim);
}

template <bool FAST>
__global__ void __launch_bounds__(256)
cos_sza_kernel(double *__restrict__ out, const __grid_constant__ SunzDev S) {
    const int64_t W = S.geo.width;
    const int64_t n = S.geo.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        out[i] = b200_cos_sza_pixel<false, FAST>(S, S.geo.row0 + r, i - r * W);
    }
}

// One thread = one pixel, all bands (the factor is shared by every band of the area).
template <bool FAST>
__global__ void __launch_bounds__(256)
sunz_correct_kernel(const float *__restrict__ data, float *__restrict__ out, int nbands, int64_t band_stride,
                    const __grid_constant__ SunzDev S) {
    const int64_t W = S.geo.width;
    const int64_t n = S.geo.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        float corr = b200_sunz_corr(S, b200_cos_sza_pixel<true, FAST>(S, S.geo.row0 + r, i - r * W));
        for (int b = 0; b < nbands; ++b) out[b * band_stride + i] = __fmul_rn(data[b * band_stride + i], corr);
    }
}

// float64 data: lon/lat are not rounded (S.lonlat_f32 = 0), the product is float64
template <bool FAST>
__global__ void __launch_bounds__(256)
sunz_correct_f64_kernel(const double *__restrict__ data, double *__restrict__ out, int nbands, int64_t band_stride,
                        const __grid_constant__ SunzDev S) {
    const int64_t W = S.geo.width;
    const int64_t n = S.geo.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double corr = b200_sunz_corr_d(S, b200_cos_sza_pixel<false, FAST>(S, S.geo.row0 + r, i - r * W));
        for (int b = 0; b < nbands; ++b) out[b * band_stride + i] = __dmul_rn(data[b * band_stride + i], corr);
    }
}

__global__ void __launch_bounds__(256)
sunz_correct_sza_f64_kernel(const double *__restrict__ data, const double *__restrict__ sza, double *__restrict__ out,
                            int nbands, int64_t band_stride, int64_t n, const __grid_constant__ SunzDev S) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // the provided-SZA branch applies no max_sza mask (geometry.py:64-66): only the fall-off uses max_sza
        const double cz = cos(sza[i] * B200_DEG2RAD);
        double corr;
        if (isnan(cz)) corr = 0.0;
        else if (cz > S.limit_cos) corr = S.method == 0 ? 1.0 / cz : b200_li_shibata(cz);
        else {
            const double base = S.method == 0 ? 1.0 / S.limit_cos : S.li_lim;
            if (S.has_max) {
                double gf = (acos(cz) - S.limit_rad) / S.span;
                gf = 1.0 - log(gf + 1.0) / 0.6931471805599453;
                gf = gf < 0.0 ? 0.0 : gf;
                corr = gf * base;
            } else corr = base;
        }
        for (int b = 0; b < nbands; ++b) out[b * band_stride + i] = __dmul_rn(data[b * band_stride + i], corr);
    }
}

// optional_datasets branch: sza in degrees (float32) is given; everything that numpy does in
// float32 (deg2rad, cos, 1/cos, arccos) stays float32 (satpy/modifiers/geometry.py:64-66).
__global__ void __launch_bounds__(256)
sunz_correct_sza_kernel(const float *__restrict__ data, const float *__restrict__ sza, float *__restrict__ out,
                        int nbands, int64_t band_stride, int64_t n, const __grid_constant__ SunzDev S) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float czf = cosf(__fmul_rn(sza[i], B200_DEG2RAD_F));
        float corr;
        if (isnan(czf)) {
            corr = 0.0f;
        } else if ((double)czf > S.limit_cos) {
            corr = S.method == 0 ? __fdiv_rn(1.0f, czf) : (float)b200_li_shibata((double)czf);
        } else {
            double base = S.method == 0 ? 1.0 / S.limit_cos : S.li_lim;
            if (S.has_max) {
                double gf = ((double)acosf(czf) - S.limit_rad) / S.span;
                gf = 1.0 - log(gf + 1.0) / 0.6931471805599453;
                gf = gf < 0.0 ? 0.0 : gf;
                corr = (float)(gf * base);
            } else {
                corr = (float)base;
            }
        }
        for (int b = 0; b < nbands; ++b) out[b * band_stride + i] = __fmul_rn(data[b * band_stride + i], corr);
    }
}

struct FusedBands {
    const int16_t *counts[4];
    float *out[4];
    b200_abi_cal_t cal[4];
    int32_t nbands;
};

// Fused ABI reflective-band calibration + SunZenithCorrector: counts -> radiance -> % -> * corr.
template <bool FAST>
__global__ void __launch_bounds__(256)
abi_vis_sunz_kernel(const __grid_constant__ FusedBands B, const __grid_constant__ SunzDev S) {
    const int64_t W = S.geo.width;
    const int64_t n = S.geo.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        float corr = b200_sunz_corr(S, b200_cos_sza_pixel<true, FAST>(S, S.geo.row0 + r, i - r * W));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < B.nbands) {
                float refl = b200_abi_cal_vis((int)B.counts[b][i], B.cal[b]);  // mode 1 (checked by the entry point)
                st_stream_f1(B.out[b] + i, __fmul_rn(refl, corr));
            }
        }
    }
}

static int fill_sunz(SunzDev &S, const b200_geo_t *geo, const b200_sunz_t *sz, bool need_geo) {
    memset(&S.fast, 0, sizeof(S.fast));
    if (need_geo) {
        B200_CHECK_ARG(geo != nullptr, "geo is NULL");
        B200_CHECK_ARG(geo->width > 0 && geo->height > 0, "empty geometry");
        B200_CHECK_ARG(geo->row0 >= 0 && geo->nrows >= 0 && geo->row0 + geo->nrows <= geo->height,
                       "row window outside the grid");
        if (geo->is_swath) B200_CHECK_ARG(geo->lons && geo->lats, "swath without lon/lat arrays");
        S.geo = *geo;
    }
    B200_CHECK_ARG(sz != nullptr, "sunz params are NULL");
    B200_CHECK_ARG(sz->method == 0 || sz->method == 1, "unknown correction method");
    const double d2r = 0.017453292519943295;
    S.gmst = sz->gmst;
    S.ra = sz->ra;
    S.sin_dec = sin(sz->dec);
    S.cos_dec = cos(sz->dec);
    S.limit_rad = sz->limit_deg * d2r;
    S.limit_cos = cos(S.limit_rad);
    S.has_max = sz->has_max_sza;
    S.max_rad = sz->max_sza_deg * d2r;
    S.max_cos = cos(S.max_rad);
    S.span = 1.0;
    if (S.has_max) {
        B200_CHECK_ARG(sz->max_sza_deg != sz->limit_deg, "max_sza equals correction_limit");
        S.span = S.max_rad - S.limit_rad;
    }
    S.lonlat_f32 = sz->lonlat_f32;
    S.method = sz->method;
    S.li_lim = 24.35 / (2.0 * S.limit_cos + sqrt(498.5225 * S.limit_cos * S.limit_cos + 1.0));
    S.alg = 0;  // set by sunz_enable_alg once the geostationary tables exist
    S.cos_a = S.sin_a = S.alg_lo = S.alg_hi = 0.0;
    return B200_OK;
}

// angle-free cos(sza) where the correction tolerates it (see b200_cos_sza_pixel): geostationary area, float32 data
static void sunz_enable_alg(SunzDev &S) {
    if (!S.fast.enabled || !S.lonlat_f32) return;
    const double a = S.gmst - S.ra + S.fast.lam0;
    S.cos_a = cos(a);
    S.sin_a = sin(a);
    S.alg_hi = fmax(0.2, S.limit_cos + 1e-3);
    S.alg_lo = (S.has_max ? fmin(S.max_cos, S.limit_cos) : S.limit_cos) - 1e-3;
    S.alg = 1;
}

extern "C" int b200_cos_sza(const b200_geo_t *geo, const b200_sunz_t *sz, double *cos_sza, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, geo, sz, true);
    if (rc) return rc;
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(cos_sza != nullptr, "NULL output");
    rc = b200_geos_fast_prepare(geo, &S.fast, (cudaStream_t)stream);
    if (rc) return rc;
    if (S.fast.enabled) cos_sza_kernel<true><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(cos_sza, S);
    else cos_sza_kernel<false><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(cos_sza, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sunz_correct(const float *data, float *out, int32_t nbands, int64_t band_stride,
                                 const b200_geo_t *geo, const b200_sunz_t *sz, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, geo, sz, true);
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1, "nbands must be >= 1");
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    rc = b200_geos_fast_prepare(geo, &S.fast, (cudaStream_t)stream);
    if (rc) return rc;
    sunz_enable_alg(S);
    if (S.fast.enabled)
        sunz_correct_kernel<true><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, S);
    else
        sunz_correct_kernel<false><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sunz_correct_f64(const double *data, double *out, int32_t nbands, int64_t band_stride,
                                     const b200_geo_t *geo, const b200_sunz_t *sz, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, geo, sz, true);
    if (rc) return rc;
    S.lonlat_f32 = 0;  // float64 data: lon/lat keep their float64 values (angles.py:427-429 casts to the DATA dtype)
    B200_CHECK_ARG(nbands >= 1, "nbands must be >= 1");
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    rc = b200_geos_fast_prepare(geo, &S.fast, (cudaStream_t)stream);
    if (rc) return rc;
    if (S.fast.enabled)
        sunz_correct_f64_kernel<true><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, S);
    else
        sunz_correct_f64_kernel<false><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sunz_correct_sza_f64(const double *data, const double *sza_deg, double *out, int32_t nbands,
                                         int64_t band_stride, int64_t n, const b200_sunz_t *sz, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, nullptr, sz, false);
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1 && n >= 0, "bad shape");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data && sza_deg && out, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    sunz_correct_sza_f64_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(data, sza_deg, out, nbands,
                                                                                           band_stride, n, S);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_sunz_correct_sza(const float *data, const float *sza_deg, float *out, int32_t nbands,
                                     int64_t band_stride, int64_t n, const b200_sunz_t *sz, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, nullptr, sz, false);
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1 && n >= 0, "bad shape");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data && sza_deg && out, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    sunz_correct_sza_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        data, sza_deg, out, nbands, band_stride, n, S);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_abi_vis_sunz(const int16_t *const *counts, float *const *out, const b200_abi_cal_t *cal,
                                 int32_t nbands, const b200_geo_t *geo, const b200_sunz_t *sz, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, geo, sz, true);
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1 && nbands <= 4, "nbands must be 1..4");
    B200_CHECK_ARG(counts && out && cal, "NULL argument");
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    FusedBands B;
    B.nbands = nbands;
    for (int b = 0; b < 4; ++b) {
        B.counts[b] = b < nbands ? counts[b] : nullptr;
        B.out[b] = b < nbands ? out[b] : nullptr;
        B.cal[b] = cal[b < nbands ? b : 0];
        if (b < nbands) {
            B200_CHECK_ARG(counts[b] && out[b], "NULL band pointer");
            B200_CHECK_ARG(cal[b].mode == 1, "fused kernel is for reflectance calibration (mode 1)");
        }
    }
    rc = b200_geos_fast_prepare(geo, &S.fast, (cudaStream_t)stream);
    if (rc) return rc;
    sunz_enable_alg(S);
    if (S.fast.enabled) abi_vis_sunz_kernel<true><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(B, S);
    else abi_vis_sunz_kernel<false><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(B, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
