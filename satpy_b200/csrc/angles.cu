// Sensor and sun angles per pixel on sm_100a ("next" row of SURVEY.md section 8f).
// Replaces get_angles / _get_sensor_angles_ndarray / _get_sun_azimuth_ndarray / compute_relative_azimuth
// (satpy/modifiers/angles.py:336-402,443-531) including pyorbital's get_observer_look, observer_position and
// get_alt_az (un-vendored; restated from the library's published code, see oracle/solar.py).  Everything is float64
// like the reference (_get_valid_lonlats returns float64 lon/lat); lon/lat come from the projection in-kernel.
#include "b200_common.cuh"
#include "b200_proj.cuh"
#include "b200_geos_fast.cuh"

int b200_check_geo(const b200_geo_t *geo, const char *who);

#define WGS84_F (1.0 / 298.257223563)
#define WGS84_A_KM 6378.137

struct AnglesArg {
    b200_geo_t geo;
    GeosFast fast;
    double gmst, ra, sin_dec, cos_dec, tan_dec;
    double sx, sy, sz;  // satellite ECI position [km]
    double *sata, *satz, *suna, *sunz;
};

__global__ void __launch_bounds__(256) angles_kernel(const __grid_constant__ AnglesArg A) {
    const int64_t W = A.geo.width;
    const int64_t n = A.geo.nrows * W;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / W;
        double lon, lat;
        bool ok = A.fast.enabled ? geos_fast_lonlat(A.fast, A.geo.row0 + r, i - r * W, lon, lat)
                                 : b200_geo_lonlat(A.geo, A.geo.row0 + r, i - r * W, lon, lat);
        double sata = nan, satz = nan, suna = nan, sunz = nan;
        if (ok && lon < 1e30 && lat < 1e30) {
            const double lam = lon * B200_DEG2RAD, phi = lat * B200_DEG2RAD;
            double sl, cl;
            sincos(phi, &sl, &cl);
            if (A.sata || A.satz) {
                // observer_position(lon, lat, 0) and get_observer_look
                double theta = fmod(A.gmst + lam, 2.0 * B200_PI);
                if (theta < 0.0) theta += 2.0 * B200_PI;
                double st, ct;
                sincos(theta, &st, &ct);
                const double c = 1.0 / sqrt(1.0 + WGS84_F * (WGS84_F - 2.0) * sl * sl);
                const double sq = c * (1.0 - WGS84_F) * (1.0 - WGS84_F);
                const double achcp = (WGS84_A_KM * c) * cl;
                const double rx = A.sx - achcp * ct, ry = A.sy - achcp * st, rz = A.sz - (WGS84_A_KM * sq) * sl;
                const double top_s = sl * ct * rx + sl * st * ry - cl * rz;
                const double top_e = -st * rx + ct * ry;
                const double top_z = cl * ct * rx + cl * st * ry + sl * rz;
                double az = fmod(atan2(-top_e, top_s) + B200_PI, 2.0 * B200_PI);
                if (az < 0.0) az += 2.0 * B200_PI;
                const double rg = sqrt(rx * rx + ry * ry + rz * rz);
                const double el = asin(fmin(top_z / rg, 1.0));
                sata = az * B200_RAD2DEG;
                satz = 90.0 - el * B200_RAD2DEG;
            }
            if (A.suna || A.sunz) {
                const double h = (A.gmst + lam) - A.ra;
                double sh, ch;
                sincos(h, &sh, &ch);
                if (A.suna) {
                    double az = atan2(-sh, cl * A.tan_dec - sl * ch) * B200_RAD2DEG;
                    az = fmod(az, 360.0);
                    if (az < 0.0) az += 360.0;  // numpy's % takes the sign of the divisor
                    suna = az;
                }
                if (A.sunz) sunz = acos(sl * A.sin_dec + cl * A.cos_dec * ch) * B200_RAD2DEG;
            }
        }
        if (A.sata) A.sata[i] = sata;
        if (A.satz) A.satz[i] = satz;
        if (A.suna) A.suna[i] = suna;
        if (A.sunz) A.sunz[i] = sunz;
    }
}

extern "C" int b200_get_angles(const b200_geo_t *geo, const b200_sunz_t *sz, double sat_lon, double sat_lat,
                               double sat_alt_km, double *sata, double *satz, double *suna, double *sunz, void *stream) {
    int rc = b200_check_geo(geo, "b200_get_angles");
    if (rc) return rc;
    B200_CHECK_ARG(sz != nullptr, "time parameters are NULL");
    const int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(sata || satz || suna || sunz, "no output requested");
    cudaStream_t s = (cudaStream_t)stream;
    AnglesArg A;
    A.geo = *geo;
    A.gmst = sz->gmst;
    A.ra = sz->ra;
    A.sin_dec = sin(sz->dec);
    A.cos_dec = cos(sz->dec);
    A.tan_dec = tan(sz->dec);
    {   // observer_position(sat_lon, sat_lat, sat_alt) on the host
        const double lam = sat_lon * B200_DEG2RAD, phi = sat_lat * B200_DEG2RAD;
        double theta = fmod(sz->gmst + lam, 2.0 * B200_PI);
        if (theta < 0.0) theta += 2.0 * B200_PI;
        const double c = 1.0 / sqrt(1.0 + WGS84_F * (WGS84_F - 2.0) * sin(phi) * sin(phi));
        const double sq = c * (1.0 - WGS84_F) * (1.0 - WGS84_F);
        const double achcp = (WGS84_A_KM * c + sat_alt_km) * cos(phi);
        A.sx = achcp * cos(theta);
        A.sy = achcp * sin(theta);
        A.sz = (WGS84_A_KM * sq + sat_alt_km) * sin(phi);
    }
    A.sata = sata;
    A.satz = satz;
    A.suna = suna;
    A.sunz = sunz;
    rc = b200_geos_fast_prepare(geo, &A.fast, s);
    if (rc) return rc;
    angles_kernel<<<b200_grid_for(n, 256, 8), 256, 0, s>>>(A);
    b200_geos_fast_release(&A.fast, s);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
relative_azimuth_kernel(const T *__restrict__ sat_azi, const T *__restrict__ sun_azi, T *__restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T d = fabs(sun_azi[i] - sat_azi[i]);
        const T e = (T)360.0 - d;
        out[i] = (d != d || e != e) ? d + e : (d < e ? d : e);  // np.minimum propagates NaN
    }
}

extern "C" int b200_relative_azimuth(const void *sat_azi, const void *sun_azi, void *out, int64_t n, int32_t is_f32,
                                     void *stream) {
    B200_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(sat_azi && sun_azi && out, "NULL data pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (is_f32)
        relative_azimuth_kernel<float><<<b200_grid_for(n, 256, 8), 256, 0, s>>>((const float *)sat_azi, (const float *)sun_azi, (float *)out, n);
    else
        relative_azimuth_kernel<double><<<b200_grid_for(n, 256, 8), 256, 0, s>>>((const double *)sat_azi, (const double *)sun_azi, (double *)out, n);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
