// Device-side map projections (float64): the on-GPU stand-in for pyproj/PROJ as
// reached through pyresample's AreaDefinition.get_lonlats
// (satpy/modifiers/angles.py:434-441, satpy/resample/kdtree.py:79-87).
// Formulas: Snyder 1987 + PROJ geos.cpp, the same algebra as oracle/projections.py
// (written independently; parity is checked to < 1e-9 deg in tests/test_gpu_geometry.py).
// Derived constants p[] are computed on the host by satpy_b200/geometry.py.
#pragma once
#include "b200_common.cuh"

#define B200_PI 3.14159265358979323846
#define B200_HALFPI 1.57079632679489661923
#define B200_DEG2RAD 0.017453292519943295
#define B200_RAD2DEG 57.29577951308232
#define B200_EPS10 1e-10
#define B200_INF __longlong_as_double(0x7ff0000000000000LL)

__device__ __forceinline__ double b200_adjlon(double lam) {
    if (fabs(lam) > B200_PI + 1e-12) {
        double w = lam + B200_PI;
        w = w - 2.0 * B200_PI * floor(w / (2.0 * B200_PI));
        lam = w - B200_PI;
    }
    return lam;
}

__device__ __forceinline__ double b200_tsfn(double phi, double sinphi, double e) {
    double con = e * sinphi;
    return tan(0.5 * (B200_HALFPI - phi)) / pow((1.0 - con) / (1.0 + con), 0.5 * e);
}

// Latitude from the conformal quantity ts (PROJ pj_phi2).  PROJ iterates phi <- pi/2 - 2 atan(ts ((1 - e sin phi) /
// (1 + e sin phi))^(e/2)) from the conformal latitude chi; each pass gains a factor ~e^2, six passes to 1e-13.  The
// same fixed-point iteration is run here to the same tolerance, but started from the series
// phi = chi + c2 sin 2chi + c4 sin 4chi + c6 sin 6chi + c8 sin 8chi (coefficients through e^8, error < 3e-12 rad), so two
// passes suffice; the converged value agrees with the plain iteration to the last bit or two (4e-16 rad).
__device__ __forceinline__ double b200_phi2(double ts, double e) {
    const double eccnth = 0.5 * e;
    const double chi = B200_HALFPI - 2.0 * atan(ts);
    const double es = e * e, es2 = es * es, es3 = es2 * es, es4 = es2 * es2;
    const double c2 = es / 2.0 + 5.0 * es2 / 24.0 + es3 / 12.0 + 13.0 * es4 / 360.0;
    const double c4 = 7.0 * es2 / 48.0 + 29.0 * es3 / 240.0 + 811.0 * es4 / 11520.0;
    const double c6 = 7.0 * es3 / 120.0 + 81.0 * es4 / 1120.0;
    const double c8 = 4279.0 * es4 / 161280.0;
    double s1, k1;
    sincos(chi, &s1, &k1);
    const double s2 = 2.0 * s1 * k1, k2 = 1.0 - 2.0 * s1 * s1;
    const double s4 = 2.0 * s2 * k2, k4 = 1.0 - 2.0 * s2 * s2;
    double phi = chi + c2 * s2 + c4 * s4 + c6 * (s2 * k4 + k2 * s4) + c8 * (2.0 * s4 * k4);
    for (int i = 0; i < 15; ++i) {
        double con = e * sin(phi);
        double dphi = B200_HALFPI - 2.0 * atan(ts * pow((1.0 - con) / (1.0 + con), eccnth)) - phi;
        phi += dphi;
        if (fabs(dphi) <= 1e-13) break;
    }
    return phi;
}

__device__ __forceinline__ double b200_qsfn(double sinphi, double e, double one_es) {
    if (e < 1e-7) return 2.0 * sinphi;
    double con = e * sinphi;
    return one_es * (sinphi / (1.0 - con * con) - (0.5 / e) * log((1.0 - con) / (1.0 + con)));
}

__device__ __forceinline__ double b200_authlat(double beta, const double *apa) {
    double t = beta + beta;
    return beta + apa[0] * sin(t) + apa[1] * sin(t + t) + apa[2] * sin(t + t + t);
}

// Inverse projection: projected metres -> lon/lat degrees.  Returns false (and +inf) when the
// point has no inverse (e.g. off the geostationary disk).
// NOT inlined on purpose: nvcc contracts a*b+c into FMAs differently depending on the surrounding code, and
// a last-ulp difference in lon/lat between two kernels would let them break exact distance ties differently.
// One out-of-line body per translation unit keeps every kernel's coordinates bit-identical.
static __device__ __noinline__ bool b200_proj_inverse(const b200_proj_t &P, double x, double y,
                                                  double &lon, double &lat) {
    if (P.kind == B200_PROJ_LONGLAT) {
        lon = x;
        lat = y;
        return true;
    }
    double xn = (x - P.x0) / P.a;
    double yn = (y - P.y0) / P.a;
    double lam, phi;
    switch (P.kind) {
    case B200_PROJ_EQC:
        lam = xn / P.p[0];
        phi = yn + P.phi0;
        break;
    case B200_PROJ_MERC:
        lam = xn / P.k0;
        if (P.es != 0.0)
            phi = b200_phi2(exp(-yn / P.k0), P.e);
        else
            phi = B200_HALFPI - 2.0 * atan(exp(-yn / P.k0));
        break;
    case B200_PROJ_GEOS: {
        const double radius_g_1 = P.p[0], radius_g = P.p[1], c_geos = P.p[2], radius_p = P.p[3],
                     radius_p_inv2 = P.p[5];
        double vx = -1.0, vy, vz;
        if (P.mode == 1) {
            vz = tan(yn / radius_g_1);
            vy = tan(xn / radius_g_1) * hypot(1.0, vz);
        } else {
            vy = tan(xn / radius_g_1);
            vz = tan(yn / radius_g_1) * hypot(1.0, vy);
        }
        double a = vz / radius_p;
        a = vy * vy + a * a + vx * vx;
        double b = 2.0 * radius_g * vx;
        double det = b * b - 4.0 * a * c_geos;
        if (!(det >= 0.0)) {
            lon = B200_INF;
            lat = B200_INF;
            return false;
        }
        double k = (-b - sqrt(det)) / (2.0 * a);
        vx = radius_g + k * vx;
        vy *= k;
        vz *= k;
        lam = atan2(vy, vx);
        phi = atan(vz * cos(lam) / vx);
        phi = atan(radius_p_inv2 * tan(phi));
        break;
    }
    case B200_PROJ_LCC: {
        const double n = P.p[0], c = P.p[1], rho0 = P.p[2];
        double xx = xn / P.k0;
        double yy = rho0 - yn / P.k0;
        double rho = hypot(xx, yy);
        if (n < 0.0) {
            rho = -rho;
            xx = -xx;
            yy = -yy;
        }
        if (P.es != 0.0)
            phi = b200_phi2(pow(rho / c, 1.0 / n), P.e);
        else
            phi = 2.0 * atan(pow(c / rho, 1.0 / n)) - B200_HALFPI;
        lam = atan2(xx, yy) / n;
        break;
    }
    case B200_PROJ_STERE: {
        const double akm1 = P.p[0];
        double rho = hypot(xn, yn);
        double yy = (P.mode == 0) ? -yn : yn;
        if (P.es != 0.0)
            phi = b200_phi2(rho / akm1, P.e);
        else
            phi = B200_HALFPI - 2.0 * atan(rho / akm1);
        if (P.mode == 1) phi = -phi;
        lam = (rho == 0.0) ? 0.0 : atan2(xn, yy);
        break;
    }
    case B200_PROJ_LAEA: {
        const double qp = P.p[0], rq = P.p[1], sinb1 = P.p[2], cosb1 = P.p[3], dd = P.p[4];
        if (P.es != 0.0) {
            if (P.mode >= 2) {
                double xx = xn / dd, yy = yn * dd;
                double rho = hypot(xx, yy);
                if (rho < B200_EPS10) {
                    lam = 0.0;
                    phi = P.phi0;
                } else {
                    double half = 0.5 * rho / rq;
                    if (half > 1.0) {
                        lon = B200_INF;
                        lat = B200_INF;
                        return false;
                    }
                    double sce = 2.0 * asin(half);
                    double cce = cos(sce);
                    sce = sin(sce);
                    double ab, y2;
                    if (P.mode == 3) {
                        ab = cce * sinb1 + yy * sce * cosb1 / rho;
                        y2 = rho * cosb1 * cce - yy * sinb1 * sce;
                    } else {
                        ab = yy * sce / rho;
                        y2 = rho * cce;
                    }
                    lam = atan2(xx * sce, y2);
                    phi = b200_authlat(asin(ab), &P.p[7]);
                }
            } else {
                double yy = (P.mode == 0) ? -yn : yn;
                double q = xn * xn + yy * yy;
                if (q == 0.0) {
                    lam = 0.0;
                    phi = P.phi0;
                } else {
                    double ab = 1.0 - q / qp;
                    if (P.mode == 1) ab = -ab;
                    lam = atan2(xn, yy);
                    phi = b200_authlat(asin(ab), &P.p[7]);
                }
            }
        } else {
            double rh = hypot(xn, yn);
            double half = 0.5 * rh;
            if (half > 1.0) {
                lon = B200_INF;
                lat = B200_INF;
                return false;
            }
            double z = 2.0 * asin(half);
            double sinz = sin(z), cosz = cos(z);
            double xx, yy;
            if (P.mode == 2) {
                phi = (rh <= B200_EPS10) ? 0.0 : asin(yn * sinz / rh);
                xx = xn * sinz;
                yy = cosz * rh;
                lam = (yy == 0.0) ? 0.0 : atan2(xx, yy);
            } else if (P.mode == 3) {
                phi = (rh <= B200_EPS10) ? P.phi0 : asin(cosz * sinb1 + yn * sinz * cosb1 / rh);
                xx = xn * sinz * cosb1;
                yy = (cosz - sin(phi) * sinb1) * rh;
                lam = (yy == 0.0) ? 0.0 : atan2(xx, yy);
            } else if (P.mode == 0) {
                phi = B200_HALFPI - z;
                lam = atan2(xn, -yn);
            } else {
                phi = z - B200_HALFPI;
                lam = atan2(xn, yn);
            }
        }
        break;
    }
    default:
        lon = B200_INF;
        lat = B200_INF;
        return false;
    }
    if (!(isfinite(lam) && isfinite(phi))) {
        lon = B200_INF;
        lat = B200_INF;
        return false;
    }
    lam = b200_adjlon(lam + P.lam0);
    lon = lam * B200_RAD2DEG;
    lat = phi * B200_RAD2DEG;
    return true;
}

// Forward projection: lon/lat degrees -> projected metres.  false (+inf) when not projectable.
static __device__ __noinline__ bool b200_proj_forward(const b200_proj_t &P, double lon, double lat,
                                                  double &x, double &y) {
    if (P.kind == B200_PROJ_LONGLAT) {
        x = lon;
        y = lat;
        return true;
    }
    if (!(fabs(lat) <= 90.0 + 1e-12) || !isfinite(lon)) {
        x = B200_INF;
        y = B200_INF;
        return false;
    }
    double lam = b200_adjlon(lon * B200_DEG2RAD - P.lam0);
    double phi = lat * B200_DEG2RAD;
    double xn, yn;
    switch (P.kind) {
    case B200_PROJ_EQC:
        xn = P.p[0] * lam;
        yn = phi - P.phi0;
        break;
    case B200_PROJ_MERC:
        xn = P.k0 * lam;
        if (P.es != 0.0)
            yn = -P.k0 * log(b200_tsfn(phi, sin(phi), P.e));
        else
            yn = P.k0 * log(tan(B200_PI / 4 + 0.5 * phi));
        break;
    case B200_PROJ_GEOS: {
        const double radius_g_1 = P.p[0], radius_g = P.p[1], radius_p = P.p[3], radius_p2 = P.p[4],
                     radius_p_inv2 = P.p[5];
        double pg = atan(radius_p2 * tan(phi));
        double r = radius_p / hypot(radius_p * cos(pg), sin(pg));
        double vx = r * cos(lam) * cos(pg);
        double vy = r * sin(lam) * cos(pg);
        double vz = r * sin(pg);
        if (((radius_g - vx) * vx - vy * vy - vz * vz * radius_p_inv2) < 0.0) {
            x = B200_INF;
            y = B200_INF;
            return false;
        }
        double tmp = radius_g - vx;
        if (P.mode == 1) {
            xn = radius_g_1 * atan(vy / hypot(vz, tmp));
            yn = radius_g_1 * atan(vz / tmp);
        } else {
            xn = radius_g_1 * atan(vy / tmp);
            yn = radius_g_1 * atan(vz / hypot(vy, tmp));
        }
        break;
    }
    case B200_PROJ_LCC: {
        const double n = P.p[0], c = P.p[1], rho0 = P.p[2];
        double rho;
        if (P.es != 0.0)
            rho = c * pow(b200_tsfn(phi, sin(phi), P.e), n);
        else
            rho = c * pow(tan(B200_PI / 4 + 0.5 * phi), -n);
        double l = lam * n;
        xn = P.k0 * (rho * sin(l));
        yn = P.k0 * (rho0 - rho * cos(l));
        break;
    }
    case B200_PROJ_STERE: {
        const double akm1 = P.p[0];
        double coslam = cos(lam), sinlam = sin(lam);
        if (P.mode == 1) {
            phi = -phi;
            coslam = -coslam;
        }
        double rho;
        if (P.es != 0.0)
            rho = akm1 * b200_tsfn(phi, sin(phi), P.e);
        else
            rho = akm1 * tan(B200_PI / 4 - 0.5 * phi);
        xn = rho * sinlam;
        yn = -rho * coslam;
        break;
    }
    case B200_PROJ_LAEA: {
        const double qp = P.p[0], sinb1 = P.p[2], cosb1 = P.p[3], xmf = P.p[5], ymf = P.p[6];
        double coslam = cos(lam), sinlam = sin(lam), sinphi = sin(phi);
        if (P.es != 0.0) {
            double q = b200_qsfn(sinphi, P.e, P.one_es);
            if (P.mode >= 2) {
                double sinb = q / qp;
                double cosb = sqrt(fmax(0.0, 1.0 - sinb * sinb));
                if (P.mode == 3) {
                    double b = 1.0 + sinb1 * sinb + cosb1 * cosb * coslam;
                    b = sqrt(2.0 / b);
                    xn = xmf * b * cosb * sinlam;
                    yn = ymf * b * (cosb1 * sinb - sinb1 * cosb * coslam);
                } else {
                    double b = sqrt(2.0 / (1.0 + cosb * coslam));
                    xn = xmf * b * cosb * sinlam;
                    yn = b * sinb * ymf;
                }
            } else if (P.mode == 0) {
                double b = sqrt(fmax(qp - q, 0.0));
                xn = b * sinlam;
                yn = -b * coslam;
            } else {
                double b = sqrt(fmax(qp + q, 0.0));
                xn = b * sinlam;
                yn = b * coslam;
            }
        } else {
            double cosphi = cos(phi);
            if (P.mode >= 2) {
                double kk = 1.0 + sinb1 * sinphi + cosb1 * cosphi * coslam;
                if (kk <= B200_EPS10) {
                    x = B200_INF;
                    y = B200_INF;
                    return false;
                }
                kk = sqrt(2.0 / kk);
                xn = kk * cosphi * sinlam;
                yn = kk * (cosb1 * sinphi - sinb1 * cosphi * coslam);
            } else if (P.mode == 0) {
                double yk = 2.0 * sin(B200_PI / 4 - 0.5 * phi);
                xn = yk * sinlam;
                yn = -yk * coslam;
            } else {
                double yk = 2.0 * cos(B200_PI / 4 - 0.5 * phi);
                xn = yk * sinlam;
                yn = yk * coslam;
            }
        }
        break;
    }
    default:
        x = B200_INF;
        y = B200_INF;
        return false;
    }
    if (!(isfinite(xn) && isfinite(yn))) {
        x = B200_INF;
        y = B200_INF;
        return false;
    }
    x = P.a * xn + P.x0;
    y = P.a * yn + P.y0;
    return true;
}

// lon/lat (deg, f64) of pixel (row, col) of a geometry; false when it has none.
__device__ __forceinline__ bool b200_geo_lonlat(const b200_geo_t &G, int64_t row, int64_t col,
                                                double &lon, double &lat) {
    if (G.is_swath) {
        int64_t i = row * G.width + col;
        if (G.lonlat_is_f32) {
            lon = (double)((const float *)G.lons)[i];
            lat = (double)((const float *)G.lats)[i];
        } else {
            lon = ((const double *)G.lons)[i];
            lat = ((const double *)G.lats)[i];
        }
        return true;
    }
    // pyresample AreaDefinition._get_proj_vectors: arange(n) * pixel_size + pixel_upper_left
    double x = (double)col * G.area.pixel_size_x + G.area.x_ul;
    double y = (double)row * -G.area.pixel_size_y + G.area.y_ul;
    return b200_proj_inverse(G.area.proj, x, y, lon, lat);
}

// pyresample kd_tree validity rule: -180 <= lon <= 180 and -90 <= lat <= 90 (NaN / inf fail).
__device__ __forceinline__ bool b200_lonlat_valid(double lon, double lat) {
    return (lon >= -180.0) && (lon <= 180.0) && (lat >= -90.0) && (lat <= 90.0);
}
