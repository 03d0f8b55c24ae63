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
#include "b200_reduce.cuh"

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

// =====================================================================================================================
// The hot configuration -- geostationary area, float32 data, correction only (ALG) -- in float32.
//
// Per pixel float32 only: the discriminant q = B_r (D_r - tx^2) of the ellipsoid intersection cancels towards the limb
// and is formed from float32 hi + lo pairs (relative error of q < 1e-6 everywhere, checked in
// scripts/sunz_f32_check.py) -- a float64 FMA plus the F64 -> F32 conversion saturated the XU pipe that the three
// MUFU operations need (ncu: sm__inst_executed_pipe_xu at its peak, 0.38 ms).  With s = sqrt(q) the view vector scaled to the ellipsoid is (vx, vy, vz) =
// ((1 + rg s), c ..., c ...) / (rg + s); the common factor drops out of cos(sza) = n . sun / |n|, so
//     X = 1 + rg s,  Y = Yr cy,  Z = Zr cz,  num = sd Z + cd_ca X - cd_sa Y,  n2 = X^2 + Y^2 + Z^2,
//     1 / cos(sza) = sqrt(n2) / num = n2 rsqrt(n2) rcp(num)                     (3 MUFU, ~20 FP32 instructions)
// Accuracy against the reference's rounded-angle flow, emulated in numpy float32 with +-2 ulp on every hardware
// approximation (scripts/sunz_f32_check.py; ABI 1 km and SEVIRI full disks, seven times of day): |d cos(sza)| <=
// 4.8e-7, 1 / cos(sza) within 2.3e-6 relative above alg_hi = 0.1 -- north_star's bar is 1e-5.  Pixels between alg_lo
// and alg_hi (the ~84.3 .. 95 degree band around the terminator, where the fall-off is steep) are flagged and redone
// digit by digit with the reference's flow (sunz_fast_fix, out of line).
//
// Layout: a warp owns 256 consecutive columns of one row: lane L works on columns 4L .. 4L+3 and 128 + 4L .. 128 + 4L+3,
// so that every warp-level access is one contiguous run (8-byte count loads: 256 B; 16-byte float loads / stores:
// 512 B) -- full 32-byte sectors, no half-used ones.  A CTA (4 warps) owns 1024 columns and walks down a band of
// rows with its column constants in registers; the row constants (32 B) are one broadcast load per row.
#define SUNZ_FAST_COLS_PER_CTA 1024
// ABI reflectance calibration with everything that does not depend on the pixel folded on the host.  The 16-bit count
// goes to float32 without the conversion unit: PRMT builds the float 2^23 + v from the packed halves and one FADD
// removes the bias (exact for all 16-bit values; signed counts are biased by 0x8000 first), then the reference's
// operation order (satpy/readers/core/abi.py:155-173, satpy/readers/abi_l1b.py:133-145,67-69).
struct SunzFastCal {
    uint32_t xor_mask;   // 0 (unsigned counts) or 0x80008000 (signed: flip the sign bits of both halves)
    float bias;          // 8388608 (+ 32768 for signed)
    float fill_f;        // _FillValue as float (NaN: none)
    float scale, offset, refl_factor;
    int32_t scaled;      // scale_factor != 1
    int32_t pad;
};

struct SunzFastArg {
    const void *in[4];   // MODE 0: float32 bands; MODE 1: int16 ABI counts per band
    float *out[4];
    SunzFastCal cal[4];
    b200_abi_cal_t cal_full[4];   // for the out-of-line fix-up path
    int32_t nbands, rows_per_cta;
    // per-scene float32 constants (kernel parameters are constant-bank operands: no registers)
    float rg, sd, cdca, cdsa, alg_hi, alg_lo, below;
    int32_t method;
};

__device__ __forceinline__ float sunz_cal_refl(uint32_t w, int half, const SunzFastCal &c) {
    const uint32_t m = half ? __byte_perm(w, 0x4b000000u, 0x7432) : __byte_perm(w, 0x4b000000u, 0x7410);
    float f = __fsub_rn(__uint_as_float(m), c.bias);
    f = (f == c.fill_f) ? __int_as_float(0x7fc00000) : f;
    const float rad = c.scaled ? __fadd_rn(__fmul_rn(f, c.scale), c.offset) : f;
    return __fmul_rn(__fmul_rn(rad, c.refl_factor), 100.0f);
}

__device__ __forceinline__ float sunz_rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float sunz_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sunz_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint2 sunz_ld_counts4(const int16_t *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 sunz_ld_float4(const float *p) {
    const int4 v = ld_stream_v4(p);
    return make_float4(__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z), __int_as_float(v.w));
}

// flagged pixels again, with the reference's rounded-angle flow (k: bit of `flags`, pixel col0 + 128 (k / 4) + k % 4)
template <int MODE>
static __device__ __noinline__ void sunz_fast_fix(const SunzFastArg &A, const SunzDev &S, int r, int col0, unsigned flags) {
    const int64_t W = S.geo.width;
    while (flags) {
        const int k = __ffs(flags) - 1;
        flags &= flags - 1;
        const int col = col0 + ((k >> 2) << 7) + (k & 3);
        const float corr = b200_sunz_corr(S, b200_cos_sza_pixel<false, true>(S, S.geo.row0 + r, col));
        const int64_t i = (int64_t)r * W + col;
        for (int b = 0; b < A.nbands; ++b) {
            float v;
            if (MODE == 1) v = b200_abi_cal_vis((int)static_cast<const int16_t *>(A.in[b])[i], A.cal_full[b]);
            else v = static_cast<const float *>(A.in[b])[i];
            A.out[b][i] = __fmul_rn(v, corr);
        }
    }
}

template <int MODE, bool SWEEP_X>
__global__ void __launch_bounds__(128, 8)
sunz_fast_kernel(const __grid_constant__ SunzFastArg A, const __grid_constant__ SunzDev S) {
    const int W = (int)S.geo.width;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col0 = blockIdx.x * SUNZ_FAST_COLS_PER_CTA + warp * 256 + lane * 4;
    if (col0 >= W) return;                      // W % 4 == 0 (checked by the caller): a group is all in or all out
    const bool second = col0 + 128 < W;
    float Th[8], Tl[8], cy[8], cz[8];  // tx^2 (hi, lo), tan x, hypot(1, tan x) of this thread's eight columns
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = min(col0 + ((k >> 2) << 7) + (k & 3), W - 1);
        const double2 ct = __ldg(S.fast.colt + c);
        const double T = ct.x * ct.x;
        Th[k] = (float)T;
        Tl[k] = (float)(T - (double)Th[k]);
        cy[k] = (float)ct.x;
        cz[k] = SWEEP_X ? 1.0f : (float)ct.y;
    }
    const float rg = A.rg, sd = A.sd, cdca = A.cdca, cdsa = A.cdsa, alg_hi = A.alg_hi, alg_lo = A.alg_lo, below = A.below;
    const int method = A.method;
    const int r_end = min((int)S.geo.nrows, (int)(blockIdx.y + 1) * A.rows_per_cta);
    const float4 *rowf = reinterpret_cast<const float4 *>(S.fast.rowf + S.geo.row0);
    for (int r = blockIdx.y * A.rows_per_cta; r < r_end; ++r) {
        const float4 f0 = __ldg(rowf + 2 * r), f1 = __ldg(rowf + 2 * r + 1);
        const float Dh = f0.x, Dl = f0.y, Bq = f0.z;
        const float Yr = f1.x, Zr = f1.y, c_lo = f1.z, c_hi = f1.w;
        const int64_t base = (int64_t)r * W + col0;
        // start the loads of the first band's data before the arithmetic (further bands are loaded where they are used)
        uint2 cnt[2];
        float4 dat[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (g == 1 && !second) break;
            if (MODE == 1) cnt[g] = sunz_ld_counts4(static_cast<const int16_t *>(A.in[0]) + base + g * 128);
            else dat[g] = sunz_ld_float4(static_cast<const float *>(A.in[0]) + base + g * 128);
        }
        unsigned flags = 0;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (g == 1 && !second) break;
            float corr[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float cg = (float)(col0 + g * 128);
            if (cg + 3.0f >= c_lo && cg <= c_hi) {   // else: the four pixels are off the disk, factor 0
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * g + j;
                    const float qf = Bq * __fadd_rn(__fsub_rn(Dh, Th[k]), __fsub_rn(Dl, Tl[k]));
                    const float s = sunz_sqrt_approx(fmaxf(qf, 0.0f));
                    const float X = fmaf(rg, s, 1.0f);
                    const float Y = Yr * cy[k];
                    const float Z = SWEEP_X ? Zr : Zr * cz[k];
                    const float num = fmaf(cdca, X, fmaf(-cdsa, Y, sd * Z));
                    const float n2 = fmaf(X, X, fmaf(Y, Y, Z * Z));
                    const float rs = sunz_rsqrt_approx(n2);
                    const float czf = num * rs;
                    float c;
                    if (method == 0) {
                        c = sunz_rcp_approx(czf);
                    } else {  // Li & Shibata: 24.35 / (2 cz + sqrt(498.5225 cz^2 + 1))
                        c = 24.35f * sunz_rcp_approx(fmaf(2.0f, czf, sunz_sqrt_approx(fmaf(498.5225f * czf, czf, 1.0f))));
                    }
                    const bool on = qf >= 0.0f;
                    const bool hi = czf > alg_hi, lo = czf < alg_lo;
                    c = hi ? c : below;
                    c = on ? c : 0.0f;   // off the disk: cos(sza) is NaN, the reference sets the factor to 0
                    if (on && !hi && !lo) flags |= 1u << k;
                    corr[j] = c;
                }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (b >= A.nbands) break;
                float v0, v1, v2, v3;
                if (MODE == 1) {
                    uint2 w = b == 0 ? cnt[g] : sunz_ld_counts4(static_cast<const int16_t *>(A.in[b]) + base + g * 128);
                    w.x ^= A.cal[b].xor_mask;
                    w.y ^= A.cal[b].xor_mask;
                    v0 = sunz_cal_refl(w.x, 0, A.cal[b]);
                    v1 = sunz_cal_refl(w.x, 1, A.cal[b]);
                    v2 = sunz_cal_refl(w.y, 0, A.cal[b]);
                    v3 = sunz_cal_refl(w.y, 1, A.cal[b]);
                } else {
                    const float4 d = b == 0 ? dat[g] : sunz_ld_float4(static_cast<const float *>(A.in[b]) + base + g * 128);
                    v0 = d.x, v1 = d.y, v2 = d.z, v3 = d.w;
                }
                st_stream_f4(A.out[b] + base + g * 128, __fmul_rn(v0, corr[0]), __fmul_rn(v1, corr[1]),
                             __fmul_rn(v2, corr[2]), __fmul_rn(v3, corr[3]));
            }
        }
        if (flags) sunz_fast_fix<MODE>(A, S, r, col0, flags);
    }
}

// true: the float32 kernel took the job
template <int MODE>
static bool sunz_fast_launch(const SunzDev &S, const void *const *in, float *const *out, const b200_abi_cal_t *cal,
                             int nbands, cudaStream_t s) {
    if (!S.fast.enabled || !S.alg || (S.geo.width & 3) || S.geo.width >= (1LL << 30) || S.geo.nrows >= (1LL << 30)) return false;
    for (int b = 0; b < nbands; ++b)
        if (((uintptr_t)in[b] & (MODE == 1 ? 7 : 15)) || ((uintptr_t)out[b] & 15)) return false;
    const int gx = (int)((S.geo.width + SUNZ_FAST_COLS_PER_CTA - 1) / SUNZ_FAST_COLS_PER_CTA);
    // short row bands: the terminator band costs ~20x a fast pixel and is spatially compact, so the CTAs that meet it
    // must be many and small or they become the tail of the launch
    int64_t rpc = 8;
    int64_t gy = (S.geo.nrows + rpc - 1) / rpc;
    while (gy > 65535) {
        rpc *= 2;
        gy = (S.geo.nrows + rpc - 1) / rpc;
    }
    for (int b0 = 0; b0 < nbands; b0 += 4) {
        SunzFastArg A;
        memset(&A, 0, sizeof(A));
        A.nbands = nbands - b0 < 4 ? nbands - b0 : 4;
        A.rows_per_cta = (int32_t)rpc;
        A.rg = (float)S.fast.rg;
        A.sd = (float)S.sin_dec;
        A.cdca = (float)(S.cos_dec * S.cos_a);
        A.cdsa = (float)(S.cos_dec * S.sin_a);
        A.alg_hi = (float)S.alg_hi;
        A.alg_lo = (float)S.alg_lo;
        // below alg_lo: masked by max_sza (0) or, without max_sza, the constant correction at the limit
        A.below = S.has_max ? 0.0f : (float)(S.method == 0 ? 1.0 / S.limit_cos : S.li_lim);
        A.method = S.method;
        for (int b = 0; b < A.nbands; ++b) {
            A.in[b] = in[b0 + b];
            A.out[b] = out[b0 + b];
            if (MODE == 1) {
                const b200_abi_cal_t &c = cal[b0 + b];
                A.cal_full[b] = c;
                SunzFastCal &f = A.cal[b];
                f.xor_mask = c.is_unsigned ? 0u : 0x80008000u;
                f.bias = c.is_unsigned ? 8388608.0f : 8388608.0f + 32768.0f;
                f.fill_f = c.has_fill ? (float)(c.is_unsigned ? (c.fill & 0xffff) : c.fill) : nanf("");
                f.scale = c.scale_factor;
                f.offset = c.add_offset;
                f.refl_factor = c.refl_factor;
                f.scaled = c.scale_factor != 1.0f;
            }
        }
        const dim3 grid(gx, (unsigned)gy);
        if (S.fast.sweep_x) sunz_fast_kernel<MODE, true><<<grid, 128, 0, s>>>(A, S);
        else sunz_fast_kernel<MODE, false><<<grid, 128, 0, s>>>(A, S);
    }
    return true;
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
    S.alg_hi = fmax(0.1, S.limit_cos + 1e-3);
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
    bool done = false;
    if (nbands <= 64 && (band_stride & 3) == 0) {
        const void *in_p[64];
        float *out_p[64];
        for (int b = 0; b < nbands; ++b) {
            in_p[b] = data + (int64_t)b * band_stride;
            out_p[b] = out + (int64_t)b * band_stride;
        }
        done = sunz_fast_launch<0>(S, in_p, out_p, nullptr, nbands, (cudaStream_t)stream);
    }
    if (done) {
    } else if (S.fast.enabled)
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

// SunZenithReducer without a solar-zenith dataset (satpy/modifiers/geometry.py:166-207): cos(sza) of the area in the
// reference's flow, masked beyond max_sza (geometry.py:62-63), sza = rad2deg(arccos(.)) (geometry.py:201) rounded to the
// data's float32, then the reduction factor -- one launch instead of a cos(sza) image and three element-wise passes.
template <bool FAST>
__global__ void __launch_bounds__(256)
sunz_reduce_area_kernel(const float *__restrict__ data, float *__restrict__ out, int nbands, int64_t band_stride,
                        float limit, float max_sza, float strength, const __grid_constant__ SunzDev S) {
    const int W = (int)S.geo.width;
    const int64_t nrows = S.geo.nrows;
    for (int64_t r = blockIdx.y; r < nrows; r += gridDim.y) {
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < W; c += gridDim.x * blockDim.x) {
            double cz = b200_cos_sza_pixel<false, FAST>(S, S.geo.row0 + r, c);
            if (S.has_max && !(cz >= S.max_cos)) cz = __longlong_as_double(0x7ff8000000000000LL);
            const float z = (float)(acos(cz) * 57.29577951308232);
            const float corr = b200_sunz_reduction(z, limit, max_sza, strength);
            const int64_t i = r * W + c;
            for (int b = 0; b < nbands; ++b) out[b * band_stride + i] = __fmul_rn(data[b * band_stride + i], corr);
        }
    }
}

extern "C" int b200_sunz_reduce_area(const float *data, float *out, int32_t nbands, int64_t band_stride,
                                     const b200_geo_t *geo, const b200_sunz_t *sz, float limit_deg, float max_sza_deg,
                                     float strength, void *stream) {
    SunzDev S;
    int rc = fill_sunz(S, geo, sz, true);
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1, "nbands must be >= 1");
    B200_CHECK_ARG(geo->width < (1LL << 31), "area too wide");
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || band_stride >= n, "band stride smaller than one image");
    rc = b200_geos_fast_prepare(geo, &S.fast, (cudaStream_t)stream);
    if (rc) return rc;
    const dim3 grid((unsigned)((geo->width + 255) / 256), (unsigned)(geo->nrows < 8192 ? geo->nrows : 8192));
    if (S.fast.enabled)
        sunz_reduce_area_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, limit_deg, max_sza_deg, strength, S);
    else
        sunz_reduce_area_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(data, out, nbands, band_stride, limit_deg, max_sza_deg, strength, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
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
    if (sunz_fast_launch<1>(S, reinterpret_cast<const void *const *>(counts), out, cal, nbands, (cudaStream_t)stream)) {
    } else if (S.fast.enabled) abi_vis_sunz_kernel<true><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(B, S);
    else abi_vis_sunz_kernel<false><<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(B, S);
    b200_geos_fast_release(&S.fast, (cudaStream_t)stream);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
