// Shared declarations of the nearest-neighbour search engines (nn.cu: bucket grid, nn_geos.cu: fixed grid).
#pragma once
#include "b200_common.cuh"
#include "b200_proj.cuh"
#include "b200_geos_fast.cuh"

#define B200_R_EARTH 6370997.0  // pyresample's sphere (kd_tree / geometry module constant)
#define B200_TWO_PI 6.283185307179586476925287

#define B200_NN_BUCKET 1
#define B200_NN_FIXED_GRID 2
#define B200_NN_FIXED_GRID_STRICT 3  /* request only: a fixed-grid structure with strict = 1 */
/* B200_NN_LIGHT (request flag, include/satpy_b200.h): light structure when the fixed-grid engine serves the source */

int b200_check_geo(const b200_geo_t *geo, const char *who);

// One search structure over the valid points of a source geometry (opaque to the caller).
struct b200_nn_grid {
    int mode;  // B200_NN_BUCKET or B200_NN_FIXED_GRID
    int strict;  // fixed grid: examine every point of the rigorous disc (no lattice-model pruning)
    int light;   // fixed grid, built for the fused path: float32 points only (no float64 records, no prefix table)
    int device;
    void *stream;  // the stream the structure was built on; its memory is stream-ordered on it
    int64_t n_src, n_valid;
    double radius;
    int32_t *prefix;  // [n_src] exclusive prefix sum of valid_input: raw -> compressed index
    const uint8_t *valid_input;  // the caller's mask buffer (needed again only for a lazily read n_valid)
    double4 *pts;     // bucket: [n_valid] records sorted by cell; fixed grid: [n_src] records by raw index
                      // record = {x, y, z, raw index bits}; fixed grid stores x = NaN for invalid pixels
    float4 *pts32;    // fixed grid only: float32 copy {x, y, z, 0} by raw index (distance pre-filter)
    int64_t nbytes;
    // ---- bucket mode
    int64_t ncells;
    int32_t nbands;
    double dphi;          // band height = cell size (radians): a few source points per cell, independent of the radius
    double phi0;          // lower edge of band 0; latitudes outside [phi0, phi0 + nbands * dphi) go to the edge bands
    int32_t *band_off;    // [nbands + 1] first cell of each band
    int32_t *cell_start;  // [ncells + 1] first record of each cell
    // ---- fixed-grid mode: the source area itself is the index
    b200_geo_t src;
    GeosFast fast;        // light structures: the projection tables stay with the grid (the exact search recomputes
                          // float64 points from them), released by b200_nn_grid_destroy
};

// Per-column / per-row quantities of a target whose lon depends on the column only and whose lat depends on the row
// only (longlat, eqc, merc): every transcendental of the query is evaluated W + H times instead of W * H times.
// Field order: what the fixed-grid kernel reads per pixel sits in aligned 16-byte groups (one 128-bit load each).
struct __align__(16) TgtCol {  // 48 B
    double sin_lam, cos_lam;       // exactly as b200_lonlat2xyz computes them                        [one 128-bit load]
    double lam;                    // radians
    float s_rel, c_rel;            // sin/cos(lam - lam0 of a geos source), float32 (fixed-grid mode); c_rel = NaN for
                                   // an invalid column
    int32_t valid, pad;
    double pad2;
};
struct __align__(16) TgtRow {  // 64 B
    double phi, sin_phi;
    double cos_phi;
    double dlam1;      // bucket mode, first phase: longitude half-width of the one-cell cap
    double dlam;       // longitude half-width of the search cap at this latitude (bucket mode); >= pi: whole circle
    float rc, rs;      // geos source: r*cos(pg), r*sin(pg) of the geocentric latitude (fixed-grid mode)
                       //                                              [dlam, rc, rs: one 128-bit load]
    int32_t valid, b_lo, b_hi;
    int32_t b_1;       // bucket mode, first phase (one-cell cap): b_lo1 | (b_hi1 - b_lo1) << 24
};

// lon depends on the column only and lat on the row only
static inline bool b200_geo_is_rectilinear(const b200_geo_t *g) {
    return !g->is_swath && (g->area.proj.kind == B200_PROJ_LONGLAT || g->area.proj.kind == B200_PROJ_EQC ||
                            g->area.proj.kind == B200_PROJ_MERC);
}

// lon/lat (degrees) -> spherical cartesian
__device__ __forceinline__ void b200_lonlat2xyz(double lon, double lat, double &x, double &y, double &z, double &lam,
                                                double &phi) {
    lam = lon * B200_DEG2RAD;
    phi = lat * B200_DEG2RAD;
    double sl, cl, sp, cp;
    sincos(lam, &sl, &cl);
    sincos(phi, &sp, &cp);
    x = B200_R_EARTH * cp * cl;
    y = B200_R_EARTH * cp * sl;
    z = B200_R_EARTH * sp;
}

// Rounding residuals of a float32 point, packed into one word: per coordinate rint((v - (float)v) * 2048) clamped to
// +-511 (10 bits, two's complement), x in bits 0-9, y in 10-19, z in 20-29.  v ~ (float)v + r * 2^-11 to 2^-11 m.
#define B200_RES_SCALE 2048.0
__device__ __forceinline__ int b200_pack_residuals(double x, double y, double z) {
    const int rx = max(-511, min(511, __double2int_rn((x - (double)(float)x) * B200_RES_SCALE)));
    const int ry = max(-511, min(511, __double2int_rn((y - (double)(float)y) * B200_RES_SCALE)));
    const int rz = max(-511, min(511, __double2int_rn((z - (double)(float)z) * B200_RES_SCALE)));
    return (rx & 0x3ff) | ((ry & 0x3ff) << 10) | ((rz & 0x3ff) << 20);
}
// the refined point: float32 value + residual, in float64
__device__ __forceinline__ void b200_refined_point(const float4 &p, double &x, double &y, double &z) {
    const int w = __float_as_int(p.w);
    x = (double)p.x + (double)((w << 22) >> 22) * (1.0 / B200_RES_SCALE);
    y = (double)p.y + (double)((w << 12) >> 22) * (1.0 / B200_RES_SCALE);
    z = (double)p.z + (double)((w << 2) >> 22) * (1.0 / B200_RES_SCALE);
}

struct Best {
    double d2;
    long long idx;
};

// squared chord distance with a FIXED operation order (both engines must rank candidates identically)
__device__ __forceinline__ double b200_nn_d2(double sx, double sy, double sz, double tx, double ty, double tz) {
    double dx = __dsub_rn(sx, tx), dy = __dsub_rn(sy, ty), dz = __dsub_rn(sz, tz);
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

// one 32-byte record {x, y, z, raw index}: two 128-bit loads
__device__ __forceinline__ void b200_nn_consider(const double4 *__restrict__ rec, double tx, double ty, double tz,
                                                 Best &best) {
    const double2 *p = reinterpret_cast<const double2 *>(rec);
    double2 a = __ldg(p);
    double2 b = __ldg(p + 1);
    double d2 = b200_nn_d2(a.x, a.y, b.x, tx, ty, tz);  // NaN (invalid pixel) fails both comparisons
    long long idx = __double_as_longlong(b.y);
    if (d2 < best.d2 || (d2 == best.d2 && idx < best.idx)) {
        best.d2 = d2;
        best.idx = idx;
    }
}

// Where a query kernel puts its result for one target pixel.
struct QueryOut {
    int32_t *index_array;   // compressed index or -1; may be NULL
    int32_t *gather_index;  // raw index or -1; may be NULL
    uint8_t *valid_output;  // may be NULL
    const float *data;      // fused gather: source bands (float32); may be NULL
    float *out;
    int32_t nbands;
    int64_t src_band_stride, dst_band_stride;
    float fill;
};

__device__ __forceinline__ void b200_nn_emit(const QueryOut &O, const int32_t *__restrict__ prefix, int64_t i,
                                             long long raw, bool valid) {
    if (O.valid_output) O.valid_output[i] = valid ? 1 : 0;
    if (O.gather_index) O.gather_index[i] = (int32_t)raw;
    if (O.index_array) O.index_array[i] = raw < 0 ? -1 : __ldg(prefix + raw);
    if (O.out) {
        for (int b = 0; b < O.nbands; ++b) {
            float v = raw < 0 ? O.fill : __ldg(O.data + b * O.src_band_stride + raw);
            st_stream_f1(O.out + b * O.dst_band_stride + i, v);
        }
    }
}

static inline double b200_cap_angle(double radius) {
    double h = radius / (2.0 * B200_R_EARTH);
    if (h >= 1.0) return B200_PI;
    return 2.0 * asin(h) * (1.0 + 1e-9) + 1e-12;
}

// implemented in nn.cu
int b200_nn_build_tables(const b200_geo_t *dst, double theta, double theta1, double phi0, double inv_dphi, int nbands,
                         int geos, double lam0, double rp, double rp2, TgtCol **cols, TgtRow **rows, cudaStream_t s);

// implemented in nn_geos.cu
int b200_nn_fixed_grid_supported(const b200_geo_t *src);
int b200_nn_fixed_grid_query(const b200_nn_grid *g, const b200_geo_t *dst, double radius, const QueryOut &O,
                             cudaStream_t s);
