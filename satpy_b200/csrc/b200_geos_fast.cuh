// Geostationary inverse projection, restructured for throughput (float64, same mathematics as PROJ's geos inverse in
// b200_proj.cuh, satpy/readers/core/utils.py:136-157):
//   * tan(scan angle) depends on the column only / the row only  -> two small tables per call, no tan per pixel;
//   * the chain  phi = atan(v_z cos(lam) / v_x);  phi = atan(tan(phi) / (1 - es))  collapses to ONE atan of
//     v_z / ((1 - es) hypot(v_x, v_y));
//   * the spherical cartesian point R (cos phi cos lon, cos phi sin lon, sin phi) follows from the same vector with
//     square roots and divisions only (no trigonometric function at all).
// Results differ from the step-by-step PROJ sequence in the last bits only (1e-16 relative).
#pragma once
#include "b200_common.cuh"
#include "b200_proj.cuh"

// Row constants of the float32 view-vector path (sunz.cu, sunz_fast_kernel).  With c = rg^2 - 1 the discriminant of the
// ellipsoid intersection is q = 1 - c (a - 1) = A_r - B_r tx^2 = B_r (D_r - tx^2), D_r = A_r / B_r; it cancels towards
// the limb (q -> 0), so D_r and tx^2 are carried as float32 hi + lo pairs: (Dh - Th) + (Dl - Tl) is accurate to a
// RELATIVE 1e-7 however small the difference (the hi parts subtract exactly once they are within a factor of two) --
// 4 FP32 instructions, no FP64 pipe, no conversion.  The (unnormalised) ellipsoid normal is
// (1 + rg sqrt(q), Yr cy, Zr cz) with the column constants cy = tan x_c, cz = 1 (sweep x) or hypot(1, tan x_c)
// (sweep y).  c_lo .. c_hi: columns outside this range are certainly off the disk in this row (conservative by a pixel).
struct GeosRowF {
    float Dh, Dl, Bq, pad;
    float Yr, Zr, c_lo, c_hi;
};

struct GeosFast {
    const double2 *colt;  // [width]  {tan(x_c), hypot(1, tan x_c)}
    const double4 *rowt;  // [height] {tan(y_r), hypot(1, tan y_r), (tan y_r)^2 / (1-es) + 1, unused}
    const GeosRowF *rowf; // [height]
    double rg, c_geos, inv2, lam0, sin_lam0, cos_lam0;
    int32_t sweep_x, enabled;
};

struct GeosTabArg {
    b200_area_t area;
};

static __global__ void __launch_bounds__(256)
geos_tables_kernel(const __grid_constant__ GeosTabArg A, double2 *__restrict__ colt, double4 *__restrict__ rowt,
                   GeosRowF *__restrict__ rowf) {
    const b200_proj_t &P = A.area.proj;
    const int64_t W = A.area.width, H = A.area.height;
    const double inv = 1.0 / (P.a * P.p[0]);  // projection metres -> scan angle
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W + H; i += stride) {
        if (i < W) {
            double x = (double)i * A.area.pixel_size_x + A.area.x_ul;
            double t = tan((x - P.x0) * inv);
            colt[i] = make_double2(t, hypot(1.0, t));
        } else {
            int64_t r = i - W;
            double y = (double)r * -A.area.pixel_size_y + A.area.y_ul;
            double t = tan((y - P.y0) * inv);
            rowt[r] = make_double4(t, hypot(1.0, t), t * t * P.p[5] + 1.0, 0.0);
            const double c = P.p[2], e_r = t * t * P.p[5];
            GeosRowF f;
            const double fa = 1.0 - c * e_r, fb = P.mode == 1 ? c * (1.0 + t * t) : c * (1.0 + e_r);
            const double fd = fa / fb;
            f.Dh = (float)fd;
            f.Dl = (float)(fd - (double)f.Dh);
            f.Bq = (float)fb;
            f.pad = 0.0f;
            if (fd > 0.0) {  // |tan x| <= sqrt(D): the columns this row of the disk can reach, widened by 1.5 pixels
                const double xr = atan(sqrt(fd)) / inv;
                f.c_lo = (float)floor((P.x0 - xr - A.area.x_ul) / A.area.pixel_size_x - 1.5);
                f.c_hi = (float)ceil((P.x0 + xr - A.area.x_ul) / A.area.pixel_size_x + 1.5);
            } else {
                f.c_lo = 1e9f;
                f.c_hi = -1e9f;
            }
            f.Yr = (float)(P.mode == 1 ? c * hypot(1.0, t) : c);
            f.Zr = (float)(c * P.p[5] * t);
            rowf[r] = f;
        }
    }
}

// tables live until the caller frees them on the same stream (cudaFreeAsync)
static inline int b200_geos_fast_prepare(const b200_geo_t *geo, GeosFast *F, cudaStream_t s) {
    memset(F, 0, sizeof(*F));
    if (geo->is_swath || geo->area.proj.kind != B200_PROJ_GEOS) return B200_OK;
    const b200_proj_t &P = geo->area.proj;
    b200_pool_keep_memory();
    // one allocation: [colt: width (padded to even) x 16 B][rowt: height x 32 B][rowf: height x 32 B]
    const size_t wpad = ((size_t)geo->area.width + 1) & ~(size_t)1, h = (size_t)geo->area.height;
    char *buf = nullptr;
    B200_CUDA(cudaMallocAsync((void **)&buf, wpad * sizeof(double2) + h * (sizeof(double4) + sizeof(GeosRowF)), s));
    double2 *colt = reinterpret_cast<double2 *>(buf);
    double4 *rowt = reinterpret_cast<double4 *>(buf + wpad * sizeof(double2));
    GeosRowF *rowf = reinterpret_cast<GeosRowF *>(buf + wpad * sizeof(double2) + h * sizeof(double4));
    GeosTabArg A;
    A.area = geo->area;
    geos_tables_kernel<<<b200_grid_for(geo->area.width + geo->area.height, 256, 8), 256, 0, s>>>(A, colt, rowt, rowf);
    F->colt = colt;
    F->rowt = rowt;
    F->rowf = rowf;
    F->rg = P.p[1];
    F->c_geos = P.p[2];
    F->inv2 = P.p[5];
    F->lam0 = P.lam0;
    F->sin_lam0 = sin(P.lam0);
    F->cos_lam0 = cos(P.lam0);
    F->sweep_x = P.mode == 1;
    F->enabled = 1;
    return B200_OK;
}

static inline void b200_geos_fast_release(GeosFast *F, cudaStream_t s) {
    if (F->colt) cudaFreeAsync((void *)F->colt, s);  // colt is the start of the one allocation
    F->colt = nullptr;
    F->rowt = nullptr;
    F->rowf = nullptr;
    F->enabled = 0;
}

// view vector scaled to the ellipsoid intersection: (vx, vy, vz) in units of the semi-major axis, earth-centred,
// x towards the satellite.  false: the line of sight misses the earth.
__device__ __forceinline__ bool geos_fast_vector(const GeosFast &F, int64_t row, int64_t col, double &vx, double &vy,
                                                 double &vz) {
    const double2 c = __ldg(F.colt + col);
    const double2 *rp = reinterpret_cast<const double2 *>(F.rowt + row);
    const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1);  // {ty, hz}, {ty^2 inv2 + 1, -}
    double a;
    if (F.sweep_x) {
        vz = r0.x;
        vy = c.x * r0.y;
        a = fma(vy, vy, r1.x);  // vy^2 + (vz / rp)^2 + 1
    } else {
        vy = c.x;
        vz = r0.x * c.y;
        a = c.y * c.y * r1.x;   // (1 + tx^2)(1 + ty^2 / (1 - es))
    }
    const double q = fma(-a, F.c_geos, F.rg * F.rg);  // det / 4
    if (!(q >= 0.0)) return false;
    const double k = (F.rg - sqrt(q)) / a;
    vx = F.rg - k;
    vy *= k;
    vz *= k;
    return true;
}

// lon/lat in degrees (+inf where off the disk), as b200_proj_inverse returns them
__device__ __forceinline__ bool geos_fast_lonlat(const GeosFast &F, int64_t row, int64_t col, double &lon, double &lat) {
    double vx, vy, vz;
    if (!geos_fast_vector(F, row, col, vx, vy, vz)) {
        lon = B200_INF;
        lat = B200_INF;
        return false;
    }
    const double lam = atan2(vy, vx);
    const double phi = atan(F.inv2 * vz / sqrt(fma(vx, vx, vy * vy)));
    lon = b200_adjlon(lam + F.lam0) * B200_RAD2DEG;
    lat = phi * B200_RAD2DEG;
    return true;
}

// spherical cartesian point of radius R from the geodetic lon/lat of the pixel, without trigonometric functions;
// also returns sin/cos of lon and lat (for callers that need the angles themselves)
__device__ __forceinline__ bool geos_fast_xyz(const GeosFast &F, int64_t row, int64_t col, double R, double &x, double &y,
                                              double &z, double &sin_lon, double &cos_lon, double &sin_lat,
                                              double &cos_lat) {
    double vx, vy, vz;
    if (!geos_fast_vector(F, row, col, vx, vy, vz)) return false;
    const double h2 = fma(vx, vx, vy * vy);
    const double inv_h = 1.0 / sqrt(h2);
    const double cl = vx * inv_h, sl = vy * inv_h;           // lon relative to the sub-satellite meridian
    cos_lon = cl * F.cos_lam0 - sl * F.sin_lam0;
    sin_lon = sl * F.cos_lam0 + cl * F.sin_lam0;
    const double t = F.inv2 * vz * inv_h;                    // tan(geodetic latitude)
    cos_lat = 1.0 / sqrt(fma(t, t, 1.0));
    sin_lat = t * cos_lat;
    x = R * cos_lat * cos_lon;
    y = R * cos_lat * sin_lon;
    z = R * sin_lat;
    return true;
}
