// Fixed-grid nearest-neighbour search for geostationary area sources (ABI, AHI, SEVIRI, FCI ... full disks).
//
// Same contract as the bucket-grid engine in nn.cu (which stands in for pyresample's kd-tree,
// satpy/resample/kdtree.py:60-95): the exact float64 nearest valid source pixel within
// radius_of_influence, ties to the lowest source index.  Here the source grid itself is the index.
//
// A geos grid is a regular lattice of VIEW DIRECTIONS from the satellite: pixel (c, r) looks along scan
// angles (x_c, y_r) = pixel centre / h.  For any two points E1, E2 on the ellipsoid the straight-line
// distance is at least 2 h sin(alpha / 2), alpha being the angle between their view directions, because
// both are at least h (the sub-satellite range) away from the satellite.  The angle between two grid
// directions is at least 0.988 x their separation in the scan-angle plane (|scan angle| <= 0.152 rad for a
// full disk), and the distance pyresample measures -- between "spherical" points R(cos lat cos lon, ...)
// built from the geodetic lon/lat -- is at least 0.995 x the ellipsoidal chord.  Hence a source pixel at
// EUCLIDEAN distance g (in pixels) from the target's fractional grid position (fc, fr) is at least
//        0.975 * g * pixel_size   metres
// away (pixel_size = h * angular step, the projection-coordinate step).  The search therefore projects the
// target into the source grid (float32 is ample: only the lattice position is needed, the margin below
// absorbs 0.02 pixel of error), examines the 2x2 enclosing pixels with exact float64 distances (all eight
// 128-bit loads in flight together) and then only the lattice points inside the disc of radius
// best_distance / (0.975 * pixel_size) pixels around (fc, fr): typically 4 to 8 distance evaluations per
// target, no build step beyond one pass that stores the source points, no sort, no cell tables.
// Hidden (far-side) targets are rejected from the sign of the visibility term with a margin of 1.5 radii.
// Rectilinear targets are processed through a dynamic work queue so that all blocks sweep one compact front of
// target rows: the source records they share then stay in L2 although per-item cost varies 10x (on/off disk).
#include "nn_common.cuh"

#define GEOS_BOUND_FACTOR 0.975
#define GEOS_POS_ERR 0.02f

int b200_nn_fixed_grid_supported(const b200_geo_t *src) {
    if (src->is_swath || src->area.proj.kind != B200_PROJ_GEOS) return 0;
    const b200_proj_t &P = src->area.proj;
    // full-disk scan angles stay below ~0.152 rad; keep well inside the regime the bound was derived for
    double ang_x = fabs(src->area.x_ul) / (P.p[0] * P.a) + src->area.pixel_size_x * src->area.width / (P.p[0] * P.a);
    double ang_y = fabs(src->area.y_ul) / (P.p[0] * P.a) + src->area.pixel_size_y * src->area.height / (P.p[0] * P.a);
    if (!(ang_x < 0.5) || !(ang_y < 0.5)) return 0;
    if (!(P.p[0] > 3.0)) return 0;  // satellite height > 3 earth radii (geostationary: 5.6)
    return 1;
}

struct GeosDev {
    const double4 *pts;
    const int32_t *prefix;
    int32_t W, H;          // source window shape
    int32_t sweep_x, max_ring;
    double lam0;
    float rg, rp2, rp, inv_rp2;     // radius_g, 1-es, sqrt(1-es), 1/(1-es)
    float col_scale, col_off;       // fc = angle_x * col_scale + col_off
    float row_scale, row_off;       // fr = angle_y * row_scale + row_off  (row_scale < 0)
    float vis_min;                  // targets with vx below this are on the far side beyond the margin
    double inv_step2;               // 1 / (0.975 * min pixel size)^2
    double r2;
};

struct GeosQueryArg {
    b200_geo_t dst;
    GeosDev G;
    QueryOut O;
    const TgtCol *cols;
    const TgtRow *rows;
};

__device__ __forceinline__ void nn_consider_loaded(const double2 &a, const double2 &b, double tx, double ty, double tz,
                                                   Best &best) {
    double d2 = b200_nn_d2(a.x, a.y, b.x, tx, ty, tz);
    long long idx = __double_as_longlong(b.y);
    if (d2 < best.d2 || (d2 == best.d2 && idx < best.idx)) {
        best.d2 = d2;
        best.idx = idx;
    }
}

// Search around the fractional grid position (fc, fr) of a target with spherical coordinates (tx, ty, tz).
__device__ __forceinline__ long long nn_fixed_search(const GeosDev &G, float fc, float fr, double tx, double ty,
                                                     double tz) {
    Best best;
    best.d2 = G.r2;
    best.idx = -1;
    const float reach = (float)G.max_ring + 1.0f;
    if (!(fc > -reach) || !(fc < (float)G.W + reach) || !(fr > -reach) || !(fr < (float)G.H + reach)) return -1;
    const int c0 = (int)floorf(fc), r0 = (int)floorf(fr);
    {
        // ring 1 = the 2x2 enclosing pixels: all eight 128-bit loads are issued before the first use
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const double2 *p00 = reinterpret_cast<const double2 *>(G.pts + ((long long)r0 * G.W + c0));
        const double2 *p10 = p00 + 2 * (long long)G.W;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        double2 a00 = make_double2(nan, 0.0), a01 = a00, a10 = a00, a11 = a00, b00 = a00, b01 = a00, b10 = a00, b11 = a00;
        if (rk0 && ck0) { a00 = __ldg(p00); b00 = __ldg(p00 + 1); }
        if (rk0 && ck1) { a01 = __ldg(p00 + 2); b01 = __ldg(p00 + 3); }
        if (rk1 && ck0) { a10 = __ldg(p10); b10 = __ldg(p10 + 1); }
        if (rk1 && ck1) { a11 = __ldg(p10 + 2); b11 = __ldg(p10 + 3); }
        nn_consider_loaded(a00, b00, tx, ty, tz, best);
        nn_consider_loaded(a01, b01, tx, ty, tz, best);
        nn_consider_loaded(a10, b10, tx, ty, tz, best);
        nn_consider_loaded(a11, b11, tx, ty, tz, best);
    }
    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above).
    const float w = sqrtf((float)(best.d2 * G.inv_step2)) * 1.0001f + GEOS_POS_ERR;
    const float w2 = w * w;
    const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
    for (int r = rlo; r <= rhi; ++r) {
        const float dr = fr - (float)r;
        const float wc2 = w2 - dr * dr;
        if (!(wc2 >= 0.0f)) continue;
        const float wc = sqrtf(wc2);
        const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
        const bool block_row = (r == r0) || (r == r0 + 1);
        for (int c = clo; c <= chi; ++c) {
            if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
            b200_nn_consider(G.pts + ((long long)r * G.W + c), tx, ty, tz, best);
        }
    }
    return best.idx;
}

// target lon/lat -> fractional position in the source grid (PROJ geos forward, float32); false: far side
__device__ __forceinline__ bool nn_fixed_position(const GeosDev &G, float rc, float rs, float s_rel, float c_rel,
                                                  float &fc, float &fr) {
    float vx = rc * c_rel, vy = rc * s_rel, vz = rs;
    if (!(vx >= G.vis_min)) return false;
    float tmp = G.rg - vx;
    float ax, ay;
    if (G.sweep_x) {
        ax = atanf(vy / hypotf(vz, tmp));
        ay = atanf(vz / tmp);
    } else {
        ax = atanf(vy / tmp);
        ay = atanf(vz / hypotf(vy, tmp));
    }
    fc = fmaf(ax, G.col_scale, G.col_off);
    fr = fmaf(ay, G.row_scale, G.row_off);
    return true;
}

// general targets: lon/lat per pixel from the target projection (or swath arrays)
__global__ void __launch_bounds__(256) nn_fixed_query_kernel(const __grid_constant__ GeosQueryArg Q) {
    const int64_t W = Q.dst.width;
    const int64_t n = Q.dst.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(Q.dst, Q.dst.row0 + r, i - r * W, lon, lat);
        bool ok = b200_lonlat_valid(lon, lat);
        long long raw = -1;
        if (ok) {
            double tx, ty, tz, lam, phi;
            b200_lonlat2xyz(lon, lat, tx, ty, tz, lam, phi);
            float rel = (float)b200_adjlon(lam - Q.G.lam0);
            float s_rel, c_rel, sp, cp;
            sincosf(rel, &s_rel, &c_rel);
            // geocentric latitude: tan(pg) = (1 - es) tan(phi); (cos pg, sin pg) from (cos phi, (1-es) sin phi)
            sincosf((float)phi, &sp, &cp);
            float u = cp, v = Q.G.rp2 * sp;
            float inv = rsqrtf(u * u + v * v);
            float cpg = u * inv, spg = v * inv;
            float rr = Q.G.rp * rsqrtf(Q.G.rp2 * cpg * cpg + spg * spg);
            float fc, fr;
            if (nn_fixed_position(Q.G, rr * cpg, rr * spg, s_rel, c_rel, fc, fr))
                raw = nn_fixed_search(Q.G, fc, fr, tx, ty, tz);
        }
        b200_nn_emit(Q.O, Q.G.prefix, i, raw, ok);
    }
}

// rectilinear targets: one work item = 256 consecutive columns of one row, tables instead of trigonometry; items are
// handed out in order through an atomic counter (chunks of GEOS_CHUNK)
#define GEOS_CHUNK 4
__global__ void __launch_bounds__(256, 4)
nn_fixed_query_rect_kernel(const __grid_constant__ GeosQueryArg Q, unsigned long long *__restrict__ queue) {
    const int W = (int)Q.dst.width;
    const int nseg = (W + 255) >> 8;
    const long long items = (long long)Q.dst.nrows * nseg;
    __shared__ long long s_first;
    for (;;) {
        if (threadIdx.x == 0) s_first = (long long)atomicAdd(queue, (unsigned long long)GEOS_CHUNK);
        __syncthreads();
        const long long first = s_first;
        __syncthreads();
        if (first >= items) break;
        const long long last = min(first + GEOS_CHUNK, items);
        for (long long it = first; it < last; ++it) {
            int r = (int)(it / nseg);
            int col = (int)(it - (long long)r * nseg) * 256 + threadIdx.x;
            if (col >= W) continue;
            const TgtRow R = Q.rows[r];
            const TgtCol Cc = Q.cols[col];
            bool ok = R.valid && Cc.valid;
            long long raw = -1;
            float fc, fr;
            if (ok && nn_fixed_position(Q.G, R.rc, R.rs, Cc.s_rel, Cc.c_rel, fc, fr)) {
                double tx = B200_R_EARTH * R.cos_phi * Cc.cos_lam;
                double ty = B200_R_EARTH * R.cos_phi * Cc.sin_lam;
                double tz = B200_R_EARTH * R.sin_phi;
                raw = nn_fixed_search(Q.G, fc, fr, tx, ty, tz);
            }
            b200_nn_emit(Q.O, Q.G.prefix, (int64_t)r * W + col, raw, ok);
        }
    }
}

int b200_nn_fixed_grid_query(const b200_nn_grid *g, const b200_geo_t *dst, double radius, const QueryOut &O,
                             cudaStream_t s) {
    const b200_geo_t &src = g->src;
    const b200_proj_t &P = src.area.proj;
    GeosQueryArg Q;
    Q.dst = *dst;
    Q.O = O;
    Q.cols = nullptr;
    Q.rows = nullptr;
    GeosDev &G = Q.G;
    G.pts = g->pts;
    G.prefix = g->prefix;
    G.W = (int32_t)src.width;
    G.H = (int32_t)src.nrows;
    G.sweep_x = P.mode == 1;
    G.lam0 = P.lam0;
    G.rg = (float)P.p[1];
    G.rp2 = (float)P.p[4];
    G.rp = (float)P.p[3];
    G.inv_rp2 = (float)P.p[5];
    // projection coordinate X = a * radius_g_1 * angle (+ x0); column = (X - x_ul) / pixel_size_x
    const double hm = P.a * P.p[0];
    G.col_scale = (float)(hm / src.area.pixel_size_x);
    G.col_off = (float)((P.x0 - src.area.x_ul) / src.area.pixel_size_x);
    // row = (y_ul - Y) / pixel_size_y, relative to the first row of the source window
    G.row_scale = (float)(-hm / src.area.pixel_size_y);
    G.row_off = (float)((src.area.y_ul - P.y0) / src.area.pixel_size_y - (double)src.row0);
    const double ps_min = fmin(src.area.pixel_size_x, src.area.pixel_size_y);
    const double step = GEOS_BOUND_FACTOR * ps_min;
    G.inv_step2 = 1.0 / (step * step);
    G.r2 = radius * radius;
    double rings = ceil(radius / step + GEOS_POS_ERR);
    if (!(rings >= 1.0)) rings = 1.0;
    if (rings > 4096.0) {
        b200_set_error("b200_nn_query: radius_of_influence of %.0f m is %.0f source pixels; the fixed-grid search "
                       "supports up to 4096 (build the structure with the bucket method)", radius, rings);
        return B200_EUNSUPPORTED;
    }
    G.max_ring = (int32_t)rings;
    // visible iff vx >= 1 / radius_g (PROJ geos forward); keep hidden points up to 1.5 radii behind the limb
    G.vis_min = (float)(1.0 / P.p[1] - 1.5 * radius / P.a - 1e-5);

    if (b200_geo_is_rectilinear(dst) && dst->width < (1LL << 30)) {
        TgtCol *cols;
        TgtRow *rows;
        int rc = b200_nn_build_tables(dst, 0.0, 1.0, 1, 1, P.lam0, P.p[3], P.p[4], &cols, &rows, s);
        if (rc) return rc;
        Q.cols = cols;
        Q.rows = rows;
        unsigned long long *queue = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&queue, sizeof(unsigned long long), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s);
        if (e == cudaSuccess) {
            int64_t items = dst->nrows * ((dst->width + 255) / 256);
            nn_fixed_query_rect_kernel<<<b200_grid_for((items + GEOS_CHUNK - 1) / GEOS_CHUNK, 1, 4), 256, 0, s>>>(Q, queue);
            e = cudaGetLastError();
        }
        if (queue) cudaFreeAsync(queue, s);
        cudaFreeAsync(cols, s);
        cudaFreeAsync(rows, s);
        if (e != cudaSuccess) {
            b200_set_error("b200_nn_query: kernel launch failed: %s", cudaGetErrorString(e));
            return B200_ECUDA;
        }
        return B200_OK;
    }
    const int64_t n = dst->nrows * dst->width;
    nn_fixed_query_kernel<<<b200_grid_for(n, 256, 8), 256, 0, s>>>(Q);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
