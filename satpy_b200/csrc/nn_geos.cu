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
// absorbs 0.02 pixel of error), ranks the 2x2 enclosing pixels and then only the lattice points inside the disc of
// radius best_distance / (0.975 * pixel_size) pixels around (fc, fr): no build step beyond one pass that stores the
// source points, no sort, no cell tables.  Three layers decide a target (each only passes on what it cannot settle):
//   nn_fixed_search          float32 ranking over FIXED stencils (2x2 block, 4x4 window) under the bound above sharpened
//                            by the satellite distance of the target -- every nearest-neighbour configuration of
//                            BASELINE.json ends here for 99.7 % of its targets;
//   nn_fixed_refine          near-ties re-ranked in float64 from float32 points + their packed rounding residuals
//                            (0.5 mm points);
//   nn_fixed_search_exact    float64 points, ties to the lowest index;
// nn_fixed_search_general is the loop of the first release (any radius, grid borders, strict mode).
// Hidden (far-side) targets are rejected from the sign of the visibility term with a margin of 1.5 radii.
// Rectilinear targets are processed in warp-sized tiles handed out by a dynamic work queue so that all warps sweep one
// compact front of target rows: the source records they share then stay in L2 although the cost per tile varies 10x
// (on / off the disk).
#include "nn_common.cuh"
#include "b200_geos_fast.cuh"
#include "b200_bulk.cuh"
#include <stdlib.h>

#define GEOS_BOUND_FACTOR 0.975

// Path statistics for tuning (scratch builds with -DB200_DEBUG_COUNTERS only; never in the shipped library).
#ifdef B200_DEBUG_COUNTERS
__device__ unsigned long long g_geos_cnt[16];
#define GEOS_CNT(i)                                                                               \
    do {                                                                                          \
        unsigned m_ = __activemask();                                                             \
        if ((int)(threadIdx.x & 31) == __ffs(m_) - 1) atomicAdd(&g_geos_cnt[i], (unsigned long long)__popc(m_)); \
    } while (0)
extern "C" int b200_debug_counters(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_geos_cnt, sizeof(g_geos_cnt));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_geos_cnt, z, sizeof(z));
    }
    return 0;
}
#else
#define GEOS_CNT(i)
#endif
#define GEOS_POS_ERR 0.02f

int b200_nn_fixed_grid_supported(const b200_geo_t *src) {
    if (src->is_swath || src->area.proj.kind != B200_PROJ_GEOS) return 0;
    const b200_proj_t &P = src->area.proj;
    // full-disk scan angles stay below ~0.152 rad; keep well inside the regime the bound was derived for
    // largest scan angle of the grid (pixel centres, relative to the sub-satellite point): full disks stay below
    // 0.152 rad; nn_atan_scan's series and the 0.975 distance bound are good up to tan = 0.25
    const double hm = P.p[0] * P.a;
    const double x0 = src->area.x_ul - P.x0, x1 = x0 + src->area.pixel_size_x * (double)(src->area.width - 1);
    const double y0 = src->area.y_ul - P.y0, y1 = y0 - src->area.pixel_size_y * (double)(src->area.height - 1);
    const double ang_x = fmax(fabs(x0), fabs(x1)) / hm, ang_y = fmax(fabs(y0), fabs(y1)) / hm;
    if (!(ang_x < 0.24) || !(ang_y < 0.24)) return 0;
    if (!(P.p[0] > 3.0)) return 0;  // satellite height > 3 earth radii (geostationary: 5.6)
    return 1;
}

struct GeosDev {
    GeosFast fast;         // light structures (pts == NULL): projection tables, the exact path recomputes float64 points
    int64_t src_row0;      // first row of the source window in the area (for `fast`)
    const double4 *pts;
    const float4 *pts32;   // float32 copy of the points (pre-filter)
    const int32_t *prefix;
    int32_t W, H;          // source window shape
    int32_t sweep_x, max_ring;
    int32_t prune, pad_;            // 1: rule out disc points with the local lattice model (0: strict mode)
    double lam0;
    float rg, rp2, rp, inv_rp2;     // radius_g, 1-es, sqrt(1-es), 1/(1-es)
    float col_scale, col_off;       // fc = angle_x * col_scale + col_off
    float row_scale, row_off;       // fr = angle_y * row_scale + row_off  (row_scale < 0)
    float vis_min;                  // targets with vx below this are on the far side beyond the margin
    float inv_step;                 // 1 / (0.975 * min pixel size)
    float radius_f;                 // radius of influence
    float r2_slack;                 // (radius + GEOS_F32_SLACK)^2 * (1 + eps)
    float k_a, k_b, k_cap;          // satellite-distance factor of the tiered search: min(k_cap, k_a * rho^2 + k_b)
    int32_t Wm3, Hm3;               // max(0, W - 3), max(0, H - 3): blocks whose 4 x 4 window lies inside the grid
    uint32_t key_mask, pad2_;       // 0xfffffff0, as a kernel argument (see nn_key)
    double r2;
};

struct GeosQueryArg {
    b200_geo_t dst;
    GeosDev G;
    QueryOut O;
    const TgtCol *cols;
    const TgtRow *rows;
};

// float32 squared distance to a float32 copy of a source point: a PRE-FILTER only.  Both points carry at most
// 0.25 m of rounding per coordinate (|coordinate| < 2^23 m), so |d_f32 - d_exact| < 1 m; a candidate whose float32
// distance exceeds the best exact distance by more than GEOS_F32_SLACK metres can never win (or tie) and is skipped
// without touching its float64 record.  NaN (invalid pixel) fails the comparison and is skipped too.
#define GEOS_F32_SLACK 3.0f
__device__ __forceinline__ float nn_d2f(const float4 &p, float tx, float ty, float tz) {
    float dx = p.x - tx, dy = p.y - ty, dz = p.z - tz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Euclidean grid distance from the fractional position (fc, fr) to the nearest lattice point OUTSIDE its 2 x 2 enclosing
// block: 1 + the distance to the nearest edge of the block -- 1 for a target on an edge, 1.5 at the centre, which is
// where the nearest pixel is farthest away.  (The first release compared against the worst case 1: every target whose
// nearest pixel was more than 0.98 x 975 m away -- a third of config 5 -- took the disc search.)
__device__ __forceinline__ float nn_block_clearance(float fc, float fr, int c0, int r0) {
    const float fx = fc - (float)c0, fy = fr - (float)r0;
    return 1.0f + fminf(fminf(fx, 1.0f - fx), fminf(fy, 1.0f - fy));
}

// One candidate of the exact search.  Full structures read its float64 record; LIGHT structures (built for the fused
// path: float32 points only, 17 instead of 53 bytes written per source pixel) recompute the point from the projection
// tables with the function that would have produced the record (geos_fast_xyz), for the 0.3 % of the targets that get
// here.  An off-disk pixel has no point and loses, like the NaN of its record.
__device__ __forceinline__ void nn_exact_consider(const GeosDev &G, long long idx, double tx, double ty, double tz,
                                                  Best &best) {
    if (G.pts) {
        b200_nn_consider(G.pts + idx, tx, ty, tz, best);
        return;
    }
    const long long r = idx / G.W;
    double x, y, z, sl, cl, sp, cp;
    if (!geos_fast_xyz(G.fast, G.src_row0 + r, idx - r * G.W, B200_R_EARTH, x, y, z, sl, cl, sp, cp)) return;
    const double d2 = b200_nn_d2(x, y, z, tx, ty, tz);
    if (d2 < best.d2 || (d2 == best.d2 && idx < best.idx)) {
        best.d2 = d2;
        best.idx = idx;
    }
}

// EXACT search (float64 distances, float32 pre-filter).  Out of line: it only runs for the rare targets whose
// float32 ranking below is ambiguous (equidistant pixels, neighbour at the radius).
__device__ __noinline__ long long nn_fixed_search_exact(const GeosDev &G, float fc, float fr, double tx, double ty,
                                                        double tz) {
    Best best;
    best.d2 = G.r2;
    best.idx = -1;
    const float reach = (float)G.max_ring + 1.0f;
    if (!(fc > -reach) || !(fc < (float)G.W + reach) || !(fr > -reach) || !(fr < (float)G.H + reach)) return -1;
    const int c0 = (int)floorf(fc), r0 = (int)floorf(fr);
    const float txf = (float)tx, tyf = (float)ty, tzf = (float)tz;
    const float nanf_ = __int_as_float(0x7fc00000);
    const long long i00 = (long long)r0 * G.W + c0;
    float thr2 = G.r2_slack;  // (radius + slack)^2: nothing beyond it can be accepted
    {
        // the 2x2 enclosing pixels: four 128-bit loads in flight together
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const float4 *q00 = G.pts32 + i00, *q10 = q00 + G.W;
        float4 p00 = make_float4(nanf_, 0.f, 0.f, 0.f), p01 = p00, p10 = p00, p11 = p00;
        if (rk0 && ck0) p00 = __ldg(q00);
        if (rk0 && ck1) p01 = __ldg(q00 + 1);
        if (rk1 && ck0) p10 = __ldg(q10);
        if (rk1 && ck1) p11 = __ldg(q10 + 1);
        const float d00 = nn_d2f(p00, txf, tyf, tzf), d01 = nn_d2f(p01, txf, tyf, tzf);
        const float d10 = nn_d2f(p10, txf, tyf, tzf), d11 = nn_d2f(p11, txf, tyf, tzf);
        const float m = fminf(fminf(d00, d01), fminf(d10, d11));  // fminf ignores NaN
        if (m == m) {
            const float t = sqrtf(m) + GEOS_F32_SLACK;
            thr2 = fminf(thr2, t * t * 1.000001f);
        }
        // exact float64 evaluation of the few (usually one) that survive the pre-filter
        if (d00 <= thr2) nn_exact_consider(G, i00, tx, ty, tz, best);
        if (d01 <= thr2) nn_exact_consider(G, i00 + 1, tx, ty, tz, best);
        if (d10 <= thr2) nn_exact_consider(G, i00 + G.W, tx, ty, tz, best);
        if (d11 <= thr2) nn_exact_consider(G, i00 + G.W + 1, tx, ty, tz, best);
    }
    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above).
    const float bd = sqrtf((float)best.d2);
    const float w = bd * G.inv_step * 1.0001f + GEOS_POS_ERR;
    if (w < nn_block_clearance(fc, fr, c0, r0)) return best.idx;  // no pixel outside the 2x2 block is that close
    {
        const float t = bd * 1.000001f + GEOS_F32_SLACK;
        thr2 = fminf(G.r2_slack, t * t * 1.000001f);
    }
    const float w2 = w * w;
    const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
    for (int r = rlo; r <= rhi; ++r) {
        const float dr = fr - (float)r;
        const float wc2 = w2 - dr * dr;
        if (!(wc2 >= 0.0f)) continue;
        const float wc = sqrtf(wc2);
        const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
        const bool block_row = (r == r0) || (r == r0 + 1);
        const long long row_base = (long long)r * G.W;
        for (int c = clo; c <= chi; ++c) {
            if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
            const float d = nn_d2f(__ldg(G.pts32 + row_base + c), txf, tyf, tzf);
            if (d <= thr2) {
                const double before = best.d2;
                nn_exact_consider(G, row_base + c, tx, ty, tz, best);
                if (best.d2 < before) {
                    const float t = sqrtf((float)best.d2) * 1.000001f + GEOS_F32_SLACK;
                    thr2 = fminf(thr2, t * t * 1.000001f);
                }
            }
        }
    }
    return best.idx;
}

// MUFU.SQRT: 2^-22 relative (5e-4 m at 2 km), far inside the 0.01 m and 1e-4 relative slack of every test it feeds
__device__ __forceinline__ float nn_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// MUFU.RSQ / MUFU.RCP without the denormal scaling of rsqrtf / __fdividef: the arguments are view-vector lengths of
// 5 .. 7 earth radii (position only: 2^-22 relative is far below GEOS_POS_ERR)
__device__ __forceinline__ float nn_rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float nn_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// two smallest float32 squared distances seen so far, branch-free.  Invalid pixels are stored as (+inf, 0, 0) in the
// float32 copy, so their distance is +inf: it never becomes the minimum and leaves the runner-up at +inf.
__device__ __forceinline__ void nn_track(float d, int idx, float &m1, int &i1, float &m2) {
    i1 = d < m1 ? idx : i1;
    m2 = fminf(m2, fmaxf(d, m1));
    m1 = fminf(d, m1);
}

// Search around the fractional grid position (fc, fr) of a target with spherical coordinates (tx, ty, tz).
// Hot path: float32 only.  e_i = float32 distance to the float32 copy of pixel i satisfies |e_i - d_i| < 1 m
// (GEOS_F32_ERR).  If the runner-up is more than 2 m behind the leader, the leader is the exact float64 winner;
// if the leader is more than 1 m inside (outside) the radius it is accepted (rejected) exactly as float64 would.
// Anything closer than that is re-done by nn_fixed_search_exact.
#define GEOS_F32_ERR 1.0f
// Returns the raw source index, -1 (no neighbour within the radius) or GEOS_NEED_EXACT: the float32 ranking is
// ambiguous -- the caller runs nn_fixed_search_exact with the float64 coordinates.  (n_src < 2^31: checked at build.)
#define GEOS_NEED_EXACT (-2)
__device__ __noinline__ int nn_fixed_search_general(const GeosDev &G, float fc, float fr, float txf, float tyf,
                                                    float tzf) {
    const float reach = (float)G.max_ring + 1.0f;
    if (!(fc > -reach) || !(fc < (float)G.W + reach) || !(fr > -reach) || !(fr < (float)G.H + reach)) return -1;
    const int c0 = (int)floorf(fc), r0 = (int)floorf(fr);
    const float inf = __int_as_float(0x7f800000);
    float m1 = inf, m2 = inf;
    int i1 = -1;
    GEOS_CNT(0);
    float4 p00 = make_float4(inf, 0.f, 0.f, 0.f), p01 = p00, p10 = p00, p11 = p00;
    {
        // the 2x2 enclosing pixels: four 128-bit loads in flight together
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const int i00 = r0 * G.W + c0;
        const float4 *q00 = G.pts32 + i00, *q10 = q00 + G.W;   // n_src < 2^31 (checked at build)
        if (rk0 && ck0) p00 = __ldg(q00);
        if (rk0 && ck1) p01 = __ldg(q00 + 1);
        if (rk1 && ck0) p10 = __ldg(q10);
        if (rk1 && ck1) p11 = __ldg(q10 + 1);
        nn_track(nn_d2f(p00, txf, tyf, tzf), i00, m1, i1, m2);
        nn_track(nn_d2f(p01, txf, tyf, tzf), i00 + 1, m1, i1, m2);
        nn_track(nn_d2f(p10, txf, tyf, tzf), i00 + G.W, m1, i1, m2);
        nn_track(nn_d2f(p11, txf, tyf, tzf), i00 + G.W + 1, m1, i1, m2);
    }
    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above); the exact
    // best distance so far is at most sqrt(m1) + 1 m, and never needs to exceed the radius.
    const float bd = fminf(G.radius_f, nn_sqrt_approx(m1) + GEOS_F32_ERR);
    const float w = bd * G.inv_step * 1.0001f + GEOS_POS_ERR;
    if (w >= nn_block_clearance(fc, fr, c0, r0)) {  // else no pixel outside the 2x2 block can be closer
        GEOS_CNT(1);
        if (!(m1 < inf)) GEOS_CNT(8);
        const float w2 = w * w;
        const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
        // Where the lattice is locally a parallelogram lattice (it is, except next to the limb), most points of that
        // disc can be ruled out without loading them: with e_c, e_r the 3-D steps along a row / a column, the
        // distance to lattice point (i, j) is |h|^2 + Q(i - fx, j - fy), Q the quadratic form of the Gram matrix.
        // Only points with Q < (1.25 * best + 30 m)^2 are examined.  The model is used when all four block pixels are
        // valid and the fourth one lies within 2 % of a pixel of where the other three predict it (the measured
        // second-order term; over the few pixels of the disc it stays inside the 25 % + 30 m margin).
        bool model = G.prune && (p00.x < inf) && (p01.x < inf) && (p10.x < inf) && (p11.x < inf);
        float gcc = 0.f, gcr = 0.f, grr = 0.f, det = 0.f, ecx = 0.f, ecy = 0.f, ecz = 0.f, erx = 0.f, ery = 0.f, erz = 0.f;
        if (model) {
            ecx = p01.x - p00.x, ecy = p01.y - p00.y, ecz = p01.z - p00.z;
            erx = p10.x - p00.x, ery = p10.y - p00.y, erz = p10.z - p00.z;
            const float dx = p11.x - (p00.x + ecx + erx), dy = p11.y - (p00.y + ecy + ery), dz = p11.z - (p00.z + ecz + erz);
            gcc = ecx * ecx + ecy * ecy + ecz * ecz;
            grr = erx * erx + ery * ery + erz * erz;
            gcr = ecx * erx + ecy * ery + ecz * erz;
            det = gcc * grr - gcr * gcr;
            model = (dx * dx + dy * dy + dz * dz) <= 4e-4f * fminf(gcc, grr) && det > 1e-3f * gcc * grr;
        }
        if (model) {
            GEOS_CNT(2);
            const float d0x = txf - p00.x, d0y = tyf - p00.y, d0z = tzf - p00.z;
            const float bc = d0x * ecx + d0y * ecy + d0z * ecz, br = d0x * erx + d0y * ery + d0z * erz;
            const float inv_det = 1.0f / det;
            const float fx = (bc * grr - br * gcr) * inv_det, fy = (br * gcc - bc * gcr) * inv_det;  // model position
            const float B = 1.25f * bd + 30.0f, B2 = B * B;
            const float inv_gcc = 1.0f / gcc;
            // Most targets stop here: no lattice point OUTSIDE the 2x2 block has Q < B^2, so the loop below would not
            // examine anything.  Lower bounds of Q over the outside points: rows <= -1 and rows >= 2 through the
            // minimum over a whole row, v^2 det / gcc; columns <= -1 and >= 2 of rows 0 and 1 through the parabola
            // Q(., v) at the nearest admissible column (its vertex u* = -gcr v / gcc if that lies in the range).
            const float vd = 1.0f + fy, vu = 2.0f - fy, uL = -1.0f - fx, uR = 2.0f - fx;
            float lb = 0.0f;
            if (vd > 0.0f && vu > 0.0f && uL < 0.0f && uR > 0.0f) {
                const float vm = fminf(vd, vu);
                lb = vm * vm * det * inv_gcc;
                const float v0 = -fy, v1 = 1.0f - fy;
                const float s0 = -gcr * v0 * inv_gcc, s1 = -gcr * v1 * inv_gcc;
                const float g0 = grr * v0 * v0, g1 = grr * v1 * v1, h0 = 2.0f * gcr * v0, h1 = 2.0f * gcr * v1;
                float u;
                u = fminf(uL, s0), lb = fminf(lb, fmaf(fmaf(gcc, u, h0), u, g0));
                u = fmaxf(uR, s0), lb = fminf(lb, fmaf(fmaf(gcc, u, h0), u, g0));
                u = fminf(uL, s1), lb = fminf(lb, fmaf(fmaf(gcc, u, h1), u, g1));
                u = fmaxf(uR, s1), lb = fminf(lb, fmaf(fmaf(gcc, u, h1), u, g1));
            }
            int jlo = 0, jhi = -1;
            if (!(lb >= B2 * 1.0001f)) {
                GEOS_CNT(3);
                const float hb = B * nn_sqrt_approx(gcc * inv_det);
                jlo = max(rlo, r0 + (int)ceilf(fy - hb)), jhi = min(rhi, r0 + (int)floorf(fy + hb));
            }
            for (int r = jlo; r <= jhi; ++r) {
                const float b = (float)(r - r0) - fy;
                const float disc = gcr * b * gcr * b - gcc * (grr * b * b - B2);
                if (!(disc >= 0.0f)) continue;
                const float sq = nn_sqrt_approx(disc), mid = -gcr * b;
                int clo = c0 + (int)ceilf(fx + (mid - sq) * inv_gcc), chi = c0 + (int)floorf(fx + (mid + sq) * inv_gcc);
                // never outside the rigorous disc / the grid
                const float dr = fr - (float)r;
                const float wc2 = w2 - dr * dr;
                if (!(wc2 >= 0.0f)) continue;
                const float wc = nn_sqrt_approx(wc2);
                clo = max(clo, max(0, (int)ceilf(fc - wc)));
                chi = min(chi, min(G.W - 1, (int)floorf(fc + wc)));
                const bool block_row = (r == r0) || (r == r0 + 1);
                const int row_base = r * G.W;
                const float4 *row = G.pts32 + (long long)r * G.W;
                for (int c = clo; c <= chi; ++c) {
                    if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
                    GEOS_CNT(6);
                    nn_track(nn_d2f(__ldg(row + c), txf, tyf, tzf), row_base + c, m1, i1, m2);
                }
            }
        } else {
            GEOS_CNT(4);
            if ((p00.x < inf) && (p01.x < inf) && (p10.x < inf) && (p11.x < inf)) GEOS_CNT(9);
            for (int r = rlo; r <= rhi; ++r) {
                const float dr = fr - (float)r;
                const float wc2 = w2 - dr * dr;
                if (!(wc2 >= 0.0f)) continue;
                const float wc = nn_sqrt_approx(wc2);
                const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
                const bool block_row = (r == r0) || (r == r0 + 1);
                const int row_base = r * G.W;
                const float4 *row = G.pts32 + (long long)r * G.W;
                for (int c = clo; c <= chi; ++c) {
                    if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
                    GEOS_CNT(7);
                    nn_track(nn_d2f(__ldg(row + c), txf, tyf, tzf), row_base + c, m1, i1, m2);
                }
            }
        }
    }
    if (i1 < 0) {
        GEOS_CNT(10);
        return -1;  // no valid pixel anywhere near
    }
    const float s1 = nn_sqrt_approx(m1);
    const float t = s1 + 2.0f * GEOS_F32_ERR + 0.01f;
    const bool clear_winner = m2 > t * t * 1.000001f;
    const bool clear_radius = fabsf(s1 - G.radius_f) > GEOS_F32_ERR + 0.01f;
    if (clear_winner && clear_radius) {
        if (!(s1 < G.radius_f)) GEOS_CNT(11);
        return s1 < G.radius_f ? i1 : -1;
    }
    GEOS_CNT(5);
    return GEOS_NEED_EXACT;
}

// ---------------------------------------------------------------------------------------------------------------------
// TIERED search: the hot path for radii of a few source pixels (every nearest-neighbour configuration of BASELINE.json).
// Same float32 ranking and the same acceptance rules as nn_fixed_search_general, but the candidate sets are FIXED
// stencils, so there is no loop, no lattice model and no per-row square root:
//   tier 1   the 2 x 2 enclosing block.  Sufficient when every lattice point outside it is provably farther than the
//            best distance found: they are at Euclidean grid distance >= 1 + (distance to the nearest block edge).
//   tier 2   the other 12 points of the 4 x 4 window around the block; every lattice point outside the window is at
//            grid distance >= 2 + (distance to the nearest block edge).
//   else     nn_fixed_search_general (large radii, blocks at the border of the grid, strict mode).
// The metres-per-pixel bound of the header (0.975 x pixel size) assumes both points at least h from the satellite.  At
// the limb -- the only place where tier 1 does not suffice, because pixels are stretched radially there -- the points
// are up to 1.163 h away, and the chord between two directions grows with the smaller of the two ranges:
// d >= 2 rho_min sin(alpha / 2).  kf = min(1.10, 0.9999 + 0.36 (rho_T^2 / h^2 - 1)) <= rho / h for EVERY ellipsoid point
// seen within 3 pixels of the target's direction (pixel angles up to 250 urad; scratch check: the smallest
// rho_candidate / (h kf) over all view angles, azimuths, visible and just-hidden targets is 1.00009), so a grid distance
// g is at least 0.975 kf g pixel_size metres.  With it the 4 x 4 window always covers the disc of a miss at radius
// 2 x pixel size (w <= 2.07 / kf < 2), which a third of the far-limb targets are; without it a fifth of them fell
// through to the general search.
// Ranking keys: the float32 bits of the squared distance (monotone as unsigned integers) with the low four mantissa
// bits replaced by the candidate's slot in the window -- min / max of integers track winner, slot and runner-up in
// three instructions per candidate.  The four bits cost 2^-19 relative in d^2 (0.002 m at 2 km), inside the 0.01 m the
// acceptance rules add for it.  Invalid pixels are (+inf, 0, 0): key >= 0x7f800000.
#define GEOS_KEY_INF 0x7f800000u
// key = (distance bits & 0xfffffff0) | slot in ONE LOP3: the mask comes in as a kernel argument, because with two
// literal constants ptxas emits an AND and an OR (one immediate operand per instruction)
template <unsigned SLOT>
__device__ __forceinline__ unsigned nn_key(const GeosDev &G, const float4 &p, float tx, float ty, float tz) {
    unsigned k;
    asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(k) : "r"(__float_as_uint(nn_d2f(p, tx, ty, tz))), "n"(SLOT), "r"(G.key_mask));
    return k;
}
template <unsigned SLOT>
__device__ __forceinline__ void nn_key_track(const GeosDev &G, const float4 &p, float tx, float ty, float tz, unsigned &m1,
                                             unsigned &m2) {
    const unsigned k = nn_key<SLOT>(G, p, tx, ty, tz);
    m2 = min(m2, max(k, m1));
    m1 = min(m1, k);
}

// Returns the raw index, -1, GEOS_NEED_EXACT or GEOS_DEFER (not settled here: run nn_fixed_search_general).  It never
// calls an out-of-line routine itself, so that the caller's loop stays free of calls (their register conventions cost
// the hot loop spills); callers collect the deferred targets and finish them afterwards (nn_fixed_search_finish).
#define GEOS_DEFER (-3)
__device__ __forceinline__ int nn_fixed_search(const GeosDev &G, float fc, float fr, float kf, float txf, float tyf,
                                               float tzf) {
    const int c0 = __float2int_rd(fc), r0 = __float2int_rd(fr);   // NaN -> 0, out of range saturates: not interior
    if ((unsigned)(c0 - 1) >= (unsigned)G.Wm3 || (unsigned)(r0 - 1) >= (unsigned)G.Hm3) {
        // far outside the grid (what nn_fixed_search_general rejects first): a miss; the border zone is deferred
        const int reach = G.max_ring + 2;
        if ((unsigned)(c0 + reach) >= (unsigned)(G.W + 2 * reach) || (unsigned)(r0 + reach) >= (unsigned)(G.H + 2 * reach))
            return -1;
        return GEOS_DEFER;
    }
    GEOS_CNT(0);
    const int i00 = r0 * G.W + c0;
    const float4 *q0 = G.pts32 + i00, *q1 = q0 + G.W;
    unsigned m1, m2;
    {
        const float4 p00 = __ldg(q0), p01 = __ldg(q0 + 1), p10 = __ldg(q1), p11 = __ldg(q1 + 1);
        const unsigned ka = nn_key<5u>(G, p00, txf, tyf, tzf), kb = nn_key<6u>(G, p01, txf, tyf, tzf);
        m1 = min(ka, kb);
        m2 = max(ka, kb);
        nn_key_track<9u>(G, p10, txf, tyf, tzf, m1, m2);
        nn_key_track<10u>(G, p11, txf, tyf, tzf, m1, m2);
    }
#ifndef GEOS_NO_PREFETCH
    // the caller walks down the target rows of a tile: the next target's block is this one or the one below, whose
    // lower row is the only line not loaded yet -- ask for it now (interior block: row r0 + 2 exists)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(q1 + G.W));
#endif
    // best distance so far: at most sqrt(m1) + 1 m (+ the key's four bits), never more than the radius (NaN: no valid
    // pixel in the block -> fminf keeps the radius)
    float s1 = nn_sqrt_approx(__uint_as_float(m1));   // (recomputed after tier 2, which may lower m1)
    const float bd = fminf(G.radius_f, s1 + (GEOS_F32_ERR + 0.01f));
    const float wq = bd * G.inv_step * 1.0001f;
    const float fx = fc - (float)c0, fy = fr - (float)r0;
    const float edge = fminf(fminf(fx, 1.0f - fx), fminf(fy, 1.0f - fy));
    const float lim = (edge + (1.0f - GEOS_POS_ERR)) * kf;
    if (!(wq < lim)) {
        GEOS_CNT(1);
        if (!(wq < lim + kf)) return GEOS_DEFER;
        // Three batches of four loads (twelve float4 in flight at once would take 48 registers).  The row above the block
        // is at grid distance >= fy + 1 and the row below at >= 2 - fy: each is examined only if that is within reach --
        // the lanes of a warp are neighbours on one target row and mostly agree, so one of the two is usually skipped.
        const float4 *qm = q0 - G.W, *q2 = q1 + G.W;
        if (!(wq < (fy + (1.0f - GEOS_POS_ERR)) * kf)) {
            const float4 a0 = __ldg(qm - 1), a1 = __ldg(qm), a2 = __ldg(qm + 1), a3 = __ldg(qm + 2);
            nn_key_track<0u>(G, a0, txf, tyf, tzf, m1, m2);
            nn_key_track<1u>(G, a1, txf, tyf, tzf, m1, m2);
            nn_key_track<2u>(G, a2, txf, tyf, tzf, m1, m2);
            nn_key_track<3u>(G, a3, txf, tyf, tzf, m1, m2);
        }
        {
            const float4 b0 = __ldg(q0 - 1), b3 = __ldg(q0 + 2), c0_ = __ldg(q1 - 1), c3 = __ldg(q1 + 2);
            nn_key_track<4u>(G, b0, txf, tyf, tzf, m1, m2);
            nn_key_track<7u>(G, b3, txf, tyf, tzf, m1, m2);
            nn_key_track<8u>(G, c0_, txf, tyf, tzf, m1, m2);
            nn_key_track<11u>(G, c3, txf, tyf, tzf, m1, m2);
        }
        if (!(wq < ((2.0f - GEOS_POS_ERR) - fy) * kf)) {
            const float4 d0 = __ldg(q2 - 1), d1 = __ldg(q2), d2 = __ldg(q2 + 1), d3 = __ldg(q2 + 2);
            nn_key_track<12u>(G, d0, txf, tyf, tzf, m1, m2);
            nn_key_track<13u>(G, d1, txf, tyf, tzf, m1, m2);
            nn_key_track<14u>(G, d2, txf, tyf, tzf, m1, m2);
            nn_key_track<15u>(G, d3, txf, tyf, tzf, m1, m2);
        }
        s1 = nn_sqrt_approx(__uint_as_float(m1));
    }
    if (m1 >= GEOS_KEY_INF) {
        GEOS_CNT(10);
        return -1;  // no valid pixel within reach
    }
    const float t = s1 + (2.0f * GEOS_F32_ERR + 0.02f);
    const bool clear_winner = m2 > __float_as_uint(t * t * 1.000001f);
    const bool clear_radius = fabsf(s1 - G.radius_f) > GEOS_F32_ERR + 0.02f;
    if (clear_winner && clear_radius) {
        const int slot = (int)(m1 & 15u);
        return s1 < G.radius_f ? i00 + ((slot >> 2) - 1) * G.W + ((slot & 3) - 1) : -1;
    }
    GEOS_CNT(5);
    return GEOS_NEED_EXACT;
}

// REFINED ranking of a target whose float32 ranking was ambiguous (0.3 % of config 5: two pixels within 2 m of each
// other, or the nearest within 1 m of the radius).  The float32 points carry their rounding residuals (w, multiples of
// 2^-11 m; nn_source_points_kernel), so every candidate of the 4 x 4 window is known to 2^-11 m per coordinate:
// distances good to GEOS_REFINED_ERR = 1 mm, against 1 m for the float32 ranking.  The window is a superset of what the
// tiered search proved sufficient, so its exact nearest is the nearest.  Only what is still ambiguous at the millimetre
// -- systematic ties of symmetric geometries, one target in 10^5 -- goes on to nn_fixed_search_exact; before, every
// ambiguous target walked its disc again and recomputed or re-read float64 points (13 % of the kernel's instructions
// at 1.5 active lanes).  Only called for interior blocks (GEOS_NEED_EXACT from nn_fixed_search).
#define GEOS_REFINED_ERR 1.0e-3
__device__ __noinline__ int nn_fixed_refine(const GeosDev &G, float fc, float fr, double tx, double ty, double tz) {
    const int c0 = __float2int_rd(fc), r0 = __float2int_rd(fr);
    const float txf = (float)tx, tyf = (float)ty, tzf = (float)tz;
    const float4 *q = G.pts32 + ((r0 - 1) * G.W + c0 - 1);
    // float32 pass over the window: the smallest distance, then only what lies within 3 m of it is refined
    float m = __int_as_float(0x7f800000);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) m = fminf(m, nn_d2f(__ldg(q + j * G.W + i), txf, tyf, tzf));
    if (!(m < __int_as_float(0x7f800000))) return -1;
    const float lim = sqrtf(m) + (2.0f * GEOS_F32_ERR + 1.0f);
    const float lim2 = lim * lim * 1.000001f;
    double b1 = __longlong_as_double(0x7ff0000000000000LL), b2 = b1;   // two smallest refined squared distances
    int i1 = -1;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            const float4 p = __ldg(q + j * G.W + i);
            if (!(nn_d2f(p, txf, tyf, tzf) <= lim2)) continue;
            double x, y, z;
            b200_refined_point(p, x, y, z);
            const double d2 = b200_nn_d2(x, y, z, tx, ty, tz);
            if (d2 < b1) {
                b2 = b1;
                b1 = d2;
                i1 = (r0 - 1 + j) * G.W + c0 - 1 + i;
            } else if (d2 < b2) {
                b2 = d2;
            }
        }
    const double s1 = sqrt(b1), t = s1 + 2.5 * GEOS_REFINED_ERR;
    const double radius = sqrt(G.r2);
    if (!(b2 > t * t) || fabs(s1 - radius) < 1.25 * GEOS_REFINED_ERR) return GEOS_NEED_EXACT;
    return s1 < radius ? i1 : -1;
}

// what nn_fixed_search left open for one target
__device__ __forceinline__ long long nn_fixed_search_finish(const GeosDev &G, int found, float fc, float fr, double tx,
                                                            double ty, double tz) {
    if (found == GEOS_DEFER) {
        found = nn_fixed_search_general(G, fc, fr, (float)tx, (float)ty, (float)tz);
    } else if (found == GEOS_NEED_EXACT) {
        GEOS_CNT(12);
        found = nn_fixed_refine(G, fc, fr, tx, ty, tz);
    }
    if (found == GEOS_NEED_EXACT) GEOS_CNT(13);
    return found == GEOS_NEED_EXACT ? nn_fixed_search_exact(G, fc, fr, tx, ty, tz) : (long long)found;
}

// atan for scan angles: a full disk stays below tan = 0.16, where the odd series to x^11 is exact to float32
// (next term x^12 / 13 < 5e-9 relative at 0.25); anything larger takes the library routine.
__device__ __forceinline__ float nn_atan_scan(float x) {
    // |x| < 0.25 always: b200_nn_fixed_grid_supported admits sources whose scan angles stay below 0.24 rad, and every
    // point of the earth -- visible or just behind the limb -- is seen inside the disk's angular radius
    const float z = x * x;
    float p = fmaf(z, -1.0f / 11.0f, 1.0f / 9.0f);
    p = fmaf(z, p, -1.0f / 7.0f);
    p = fmaf(z, p, 0.2f);
    p = fmaf(z, p, -1.0f / 3.0f);
    return fmaf(x * z, p, x);
}

// target lon/lat -> fractional position in the source grid (PROJ geos forward, float32); false: far side.
// Fast reciprocal / rsqrt are fine here: only the lattice position is needed (error << GEOS_POS_ERR pixels).
// kf: the satellite-distance factor of the tiered search (see nn_fixed_search), from the squared range of the target
__device__ __forceinline__ bool nn_fixed_position(const GeosDev &G, float rc, float rs, float s_rel, float c_rel,
                                                  float &fc, float &fr, float &kf) {
    float vx = rc * c_rel, vy = rc * s_rel, vz = rs;
    if (!(vx >= G.vis_min)) return false;
    float tmp = G.rg - vx;
    float ax, ay;
    if (G.sweep_x) {
        const float q = fmaf(vz, vz, tmp * tmp);
        ax = nn_atan_scan(vy * nn_rsqrt_approx(q));
        ay = nn_atan_scan(vz * nn_rcp_approx(tmp));
        kf = fminf(G.k_cap, fmaf(fmaf(vy, vy, q), G.k_a, G.k_b));
    } else {
        const float q = fmaf(vy, vy, tmp * tmp);
        ax = nn_atan_scan(vy * nn_rcp_approx(tmp));
        ay = nn_atan_scan(vz * nn_rsqrt_approx(q));
        kf = fminf(G.k_cap, fmaf(fmaf(vz, vz, q), G.k_a, G.k_b));
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
            float fc, fr, kf;
            if (nn_fixed_position(Q.G, rr * cpg, rr * spg, s_rel, c_rel, fc, fr, kf)) {
                const int found = nn_fixed_search(Q.G, fc, fr, kf, (float)tx, (float)ty, (float)tz);
                raw = found >= -1 ? (long long)found : nn_fixed_search_finish(Q.G, found, fc, fr, tx, ty, tz);
            }
        }
        b200_nn_emit(Q.O, Q.G.prefix, i, raw, ok);
    }
}

// per group of GEOS_TILE_COLS columns of a rectilinear target: the largest cos(lon - lon_0) of its columns.  A group whose
// every column is on the far side for a given row (rc * cmax < vis_min) is filled without any per-pixel work.
#define GEOS_TILE_COLS 32   /* one warp */
#define GEOS_TILE_ROWS 16
__global__ void __launch_bounds__(256)
nn_seg_cmax_kernel(const TgtCol *__restrict__ cols, int W, int ngroups, float *__restrict__ seg_cmax) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    float m = -2.0f;
    for (int c = g * GEOS_TILE_COLS; c < min(W, (g + 1) * GEOS_TILE_COLS); ++c) m = fmaxf(m, cols[c].c_rel);   // NaN ignored
    seg_cmax[g] = m;
}

// One target of a rectilinear grid that the tiered search left open: everything recomputed from the tables (rare).
template <bool FUSED1>
__device__ __noinline__ void nn_rect_finish_item(const GeosQueryArg &Q, int r, int col, int found) {
    const TgtRow R = Q.rows[r];
    const TgtCol Cc = Q.cols[col];
    long long raw = -1;
    float fc, fr, kf;
    if (nn_fixed_position(Q.G, R.rc, R.rs, Cc.s_rel, Cc.c_rel, fc, fr, kf))
        raw = nn_fixed_search_finish(Q.G, found, fc, fr, R.dlam * Cc.cos_lam, R.dlam * Cc.sin_lam, R.dlam1);
    const int64_t i = (int64_t)r * Q.dst.width + col;
    if (FUSED1)
        st_stream_f1(Q.O.out + i, raw < 0 ? Q.O.fill : __ldg(Q.O.data + raw));
    else
        b200_nn_emit(Q.O, Q.G.prefix, i, raw, true);
}

// Rectilinear targets (lon from the column, lat from the row): tables instead of trigonometry, and one WARP per work
// unit -- a tile of GEOS_TILE_ROWS consecutive rows x 32 consecutive columns, walked row by row:
//   * a lane keeps its column for the whole tile: the column's table entry is read once and stays in registers;
//   * consecutive rows of a tile fall into the same or the next source rows, so the source points a row loads are the
//     ones the previous row left in L1 (the row-major order of the first release re-read them from L2 for every row:
//     L1 hit rate 47 %, 67 GB of L2 -> L1 traffic per launch);
//   * no block-level barrier: a warp draws its next tile from the atomic queue by itself, one tile ahead, so the
//     latency of the atomic is covered by the tile in hand (the per-CTA queue cost 13 % of the stall cycles at its two
//     barriers per chunk);
//   * tiles are handed out in row-band-major order, so all warps of the device sweep one compact front of 16 target
//     rows and the source rows they share stay in L2 although the cost per tile varies 10x (on / off the disk).
// A warp retires after GEOS_TILES_PER_WARP tiles, so SM slots keep freeing up for work on other streams (the exchange of
// the previous row block) while this kernel runs.
// FUSED1: the hot configuration -- fused gather of one float32 band, no index outputs.
#define GEOS_TILES_PER_WARP 64
#define GEOS_TILES_PER_DRAW 2   /* power of two, divides the tiles per warp */
#define GEOS_WARPS_PER_CTA 1   /* measured: 1 x 64 tiles 13.11 ms, 2 x 64 13.18, 4 x 64 13.32, 4 x 16 13.75 (scripts/query_knobs.sh) */
// 64 registers (32 warps / SM).  A 48-register build (40 warps / SM, 68 instead of 44 bytes of spills) measured the same:
// 12.95 against 12.99 ms with two warps per CTA.
template <bool FUSED1>
__global__ void __launch_bounds__(128, 8)   // launched with GEOS_WARPS_PER_CTA warps (up to 4 via the knob)
nn_fixed_query_rect_kernel(const __grid_constant__ GeosQueryArg Q, unsigned long long *__restrict__ queue,
                           const float *__restrict__ seg_cmax, unsigned ngroups, unsigned ntiles, int tiles_per_warp) {
    const int W = (int)Q.dst.width, H = (int)Q.dst.nrows;
    const unsigned lane = threadIdx.x & 31u;
    unsigned draw = 0u;   // first tile of the pair in hand (a draw = GEOS_TILES_PER_DRAW consecutive tiles, one atomic)
    if (lane == 0u) draw = (unsigned)atomicAdd(queue, (unsigned long long)GEOS_TILES_PER_DRAW);
    draw = __shfl_sync(0xffffffffu, draw, 0);
    unsigned nxt = 0xffffffffu;
#pragma unroll 1
    for (int turn = 0; turn < tiles_per_warp && draw < ntiles; ++turn) {
        const unsigned cur = draw + ((unsigned)turn & (GEOS_TILES_PER_DRAW - 1));
        if ((turn & (GEOS_TILES_PER_DRAW - 1)) == 0) {
            nxt = 0xffffffffu;   // the draw after this one: requested now, used GEOS_TILES_PER_DRAW tiles later
            if (lane == 0u && turn + GEOS_TILES_PER_DRAW < tiles_per_warp)
                nxt = (unsigned)atomicAdd(queue, (unsigned long long)GEOS_TILES_PER_DRAW);
        }
        if (cur >= ntiles) break;
        const unsigned band = cur / ngroups, grp = cur - band * ngroups;
        const int r_first = (int)band * GEOS_TILE_ROWS;
        const int count = min(GEOS_TILE_ROWS, H - r_first);
        const int col = (int)grp * GEOS_TILE_COLS + (int)lane;
        // Which rows of the tile can see the satellite at all?  Lane k looks at row k: the table reads of the whole tile
        // are in flight together.
        bool vis = false;
        if ((int)lane < count) {
            const TgtRow *Rk = Q.rows + r_first + lane;
            vis = Rk->valid && (Rk->rc * __ldg(seg_cmax + grp) >= Q.G.vis_min);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, vis);
        const bool in_w = col < W;
        const int64_t i_first = (int64_t)r_first * W + col;
        if (vmask == 0u) {
            // the whole tile is on the far side: every pixel is a miss
            if (in_w) {
                if (FUSED1) {
                    float *dst = Q.O.out + i_first;
                    for (int k = 0; k < count; ++k, dst += W) st_stream_f1(dst, Q.O.fill);
                } else {
                    const bool cv = Q.cols[col].valid != 0;
                    for (int k = 0; k < count; ++k)
                        b200_nn_emit(Q.O, Q.G.prefix, i_first + (int64_t)k * W, -1, cv && Q.rows[r_first + k].valid);
                }
            }
        } else if (in_w) {
            // the lane's column: {sin_lam, cos_lam} and {s_rel, c_rel}, one 128-bit and one 64-bit load per TILE
            const TgtCol *Cp = Q.cols + col;
            const double2 sc = __ldg(reinterpret_cast<const double2 *>(&Cp->sin_lam));
            const float2 rel = __ldg(reinterpret_cast<const float2 *>(&Cp->s_rel));
            // FUSED1 never reports validity: an invalid column carries c_rel = NaN (nn_target_tables_kernel), fails the
            // visibility test of nn_fixed_position and becomes a miss without its flag being read
            const bool ok = FUSED1 ? true : (Cp->valid != 0);
            // FUSED1: the gather of row k is issued at the end of its iteration and stored at the end of the NEXT one,
            // so the load's latency (a DRAM access for most pixels) is covered by the next row's search instead of
            // stalling the warp in front of the store.  Targets the tiered search does not settle (0.3 %: float32
            // ranking ambiguous, block at the border of the grid) only set a bit; they are finished after the loop,
            // which therefore contains no call.
            unsigned todo = 0u;
            float pend_v = 0.0f;
            int pend_k = -1;
            float *const out0 = FUSED1 ? Q.O.out + i_first : nullptr;
            const TgtRow *Rp = Q.rows + r_first;
            unsigned vm = vmask;
            for (int k = 0; k < count; ++k, ++Rp, vm >>= 1) {
                int found = -1;
                if (vm & 1u) {
                    // {dlam, rc, rs} of the row: one 128-bit load, the same address for every lane
                    const double2 rw = __ldg(reinterpret_cast<const double2 *>(&Rp->dlam));
                    const float rc = __int_as_float(__double2loint(rw.y)), rs = __int_as_float(__double2hiint(rw.y));
                    float fc, fr, kf;
                    if (ok && nn_fixed_position(Q.G, rc, rs, rel.x, rel.y, fc, fr, kf)) {
                        // R.dlam = R_earth cos(phi), R.dlam1 = R_earth sin(phi), R.b_lo = float32 bits of the latter
                        const double tx = rw.x * sc.y, ty = rw.x * sc.x;
                        found = nn_fixed_search(Q.G, fc, fr, kf, (float)tx, (float)ty, __int_as_float(__ldg(&Rp->b_lo)));
                    }
                    if (found < -1) {
                        todo |= (found == GEOS_DEFER ? 0x10000u : 1u) << k;
                        continue;
                    }
                }
                if (FUSED1) {
                    if (pend_k >= 0) st_stream_f1(out0 + pend_k * W, pend_v);
                    pend_k = k;
                    pend_v = found < 0 ? Q.O.fill : __ldg(Q.O.data + found);
                } else {
                    b200_nn_emit(Q.O, Q.G.prefix, i_first + (int64_t)k * W, (long long)found,
                                 ok && (vm & 1u ? true : Rp->valid != 0));
                }
            }
            if (FUSED1 && pend_k >= 0) st_stream_f1(out0 + pend_k * W, pend_v);
            while (todo) {
                const int bit = __ffs(todo) - 1;
                todo &= todo - 1u;
                nn_rect_finish_item<FUSED1>(Q, r_first + (bit & 15), col, bit >= 16 ? GEOS_DEFER : GEOS_NEED_EXACT);
            }
        }
        if ((turn & (GEOS_TILES_PER_DRAW - 1)) == GEOS_TILES_PER_DRAW - 1) draw = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

static int nn_fixed_grid_query_impl(const b200_nn_grid *g, const b200_geo_t *dst, double radius, const QueryOut &O,
                                    cudaStream_t s, const GeosFast &fast);

int b200_nn_fixed_grid_query(const b200_nn_grid *g, const b200_geo_t *dst, double radius, const QueryOut &O,
                             cudaStream_t s) {
    if (g->light && O.index_array) {
        b200_set_error("b200_nn_query: a light search structure (fused path) has no compressed index table");
        return B200_EUNSUPPORTED;
    }
    return nn_fixed_grid_query_impl(g, dst, radius, O, s, g->fast);   // (all zero for a full structure)
}

static int nn_fixed_grid_query_impl(const b200_nn_grid *g, const b200_geo_t *dst, double radius, const QueryOut &O,
                                    cudaStream_t s, const GeosFast &fast) {
    const b200_geo_t &src = g->src;
    const b200_proj_t &P = src.area.proj;
    GeosQueryArg Q;
    Q.dst = *dst;
    Q.O = O;
    Q.cols = nullptr;
    Q.rows = nullptr;
    GeosDev &G = Q.G;
    G.fast = fast;
    G.src_row0 = src.row0;
    G.pts = g->pts;
    G.pts32 = g->pts32;
    G.prune = g->strict ? 0 : 1;
    G.pad_ = 0;
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
    G.inv_step = (float)(1.0 / step);
    G.radius_f = (float)radius;
    G.r2_slack = (float)((radius + GEOS_F32_SLACK) * (radius + GEOS_F32_SLACK) * 1.000001);
    G.r2 = radius * radius;
    double rings = ceil(radius / step + GEOS_POS_ERR);
    if (!(rings >= 1.0)) rings = 1.0;
    if (rings > 4096.0) {
        b200_set_error("b200_nn_query: radius_of_influence of %.0f m is %.0f source pixels; the fixed-grid search "
                       "supports up to 4096 (build the structure with the bucket method)", radius, rings);
        return B200_EUNSUPPORTED;
    }
    G.max_ring = (int32_t)rings;
    // tiered search (nn_fixed_search): fixed stencils reach radii up to ~2.5 source pixels; strict mode never uses them.
    // The satellite-distance factor was verified for geostationary heights and pixel angles up to 250 urad
    // (kf = 1 otherwise: the plain bound).
    G.key_mask = 0xfffffff0u;
    G.pad2_ = 0;
    G.Wm3 = g->strict ? 0 : (int32_t)(src.width > 3 ? src.width - 3 : 0);
    G.Hm3 = g->strict ? 0 : (int32_t)(src.nrows > 3 ? src.nrows - 3 : 0);
    {
        const double ang = fmax(src.area.pixel_size_x, src.area.pixel_size_y) / hm;
        const bool kf_ok = P.p[0] > 5.3 && P.p[0] < 5.9 && ang <= 250e-6;
        const double c1 = kf_ok ? 0.36 : 0.0;
        // rho^2 / h^2 in units of the kernel's view vector (semi-major axes): (a / h)^2 = 1 / p0^2
        G.k_a = (float)(c1 / (P.p[0] * P.p[0]));
        G.k_b = (float)(kf_ok ? 0.9999 - c1 : 1.0);
        G.k_cap = (float)(kf_ok ? 1.10 : 1.0);
    }
    // visible iff vx >= 1 / radius_g (PROJ geos forward); keep hidden points up to 1.5 radii behind the limb
    G.vis_min = (float)(1.0 / P.p[1] - 1.5 * radius / P.a - 1e-5);

    if (b200_geo_is_rectilinear(dst) && dst->width < (1LL << 30)) {
        TgtCol *cols;
        TgtRow *rows;
        int rc = b200_nn_build_tables(dst, 0.0, 0.0, 0.0, 1.0, 1, 1, P.lam0, P.p[3], P.p[4], &cols, &rows, s);
        if (rc) return rc;
        Q.cols = cols;
        Q.rows = rows;
        unsigned long long *queue = nullptr;
        float *seg_cmax = nullptr;
        const int64_t ngroups = (dst->width + GEOS_TILE_COLS - 1) / GEOS_TILE_COLS;
        const int64_t ntiles = ngroups * ((dst->nrows + GEOS_TILE_ROWS - 1) / GEOS_TILE_ROWS);
        if (ntiles >= (1LL << 31) || (int64_t)GEOS_TILE_ROWS * dst->width >= (1LL << 31)) {
            cudaFreeAsync(cols, s);
            cudaFreeAsync(rows, s);
            b200_set_error("b200_nn_query: target too large for one call (%lld tiles): query it in row bands",
                           (long long)ntiles);
            return B200_EUNSUPPORTED;
        }
        cudaError_t e = cudaMallocAsync((void **)&queue, sizeof(unsigned long long), s);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&seg_cmax, (size_t)ngroups * sizeof(float), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s);
        if (e == cudaSuccess) {
            nn_seg_cmax_kernel<<<(int)((ngroups + 255) / 256), 256, 0, s>>>(cols, (int)dst->width, (int)ngroups, seg_cmax);
            // tuning knobs for measurements (scripts/query_knobs.sh); the defaults are what the library ships with
            static int warps_per_cta = 0, tiles_per_warp = 0;
            if (warps_per_cta == 0) {
                const char *e1 = getenv("B200_NN_WARPS_PER_CTA"), *e2 = getenv("B200_NN_TILES_PER_WARP");
                int w = e1 ? atoi(e1) : GEOS_WARPS_PER_CTA, t = e2 ? atoi(e2) : GEOS_TILES_PER_WARP;
                if (w != 1 && w != 2 && w != 4) w = GEOS_WARPS_PER_CTA;
                if (t < GEOS_TILES_PER_DRAW || t > 4096 || (t % GEOS_TILES_PER_DRAW)) t = GEOS_TILES_PER_WARP;
                tiles_per_warp = t;
                warps_per_cta = w;
            }
            const int64_t per_cta = (int64_t)tiles_per_warp * warps_per_cta;
            int grid = (int)((ntiles + per_cta - 1) / per_cta);
            if (grid < 1) grid = 1;
            const bool fused1 = O.out != nullptr && O.nbands == 1 && !O.index_array && !O.gather_index &&
                                !O.valid_output;
            const dim3 block(warps_per_cta * 32);
            if (fused1)
                nn_fixed_query_rect_kernel<true><<<grid, block, 0, s>>>(Q, queue, seg_cmax, (unsigned)ngroups, (unsigned)ntiles,
                                                                      tiles_per_warp);
            else
                nn_fixed_query_rect_kernel<false><<<grid, block, 0, s>>>(Q, queue, seg_cmax, (unsigned)ngroups, (unsigned)ntiles,
                                                                       tiles_per_warp);
            e = cudaGetLastError();
        }
        if (queue) cudaFreeAsync(queue, s);
        if (seg_cmax) cudaFreeAsync(seg_cmax, s);
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

// =====================================================================================================================
// Bilinear resampling: replaces pyresample's XArrayBilinearResampler.get_bil_info /
// get_sample_from_bil_info as driven by BilinearResampler.precompute / .compute (satpy/resample/kdtree.py:216-242,
// 281-293).  PARITY UNPINNED in the reference (its tests mock the library); the algorithm follows pyresample's published
// one (restated in oracle/bilinear.py):
//   the `neighbours` (32) nearest valid source pixels within the radius, nearest first; in TARGET-projection
//   coordinates the nearest one in each quadrant around the target pixel centre (strict inequalities); fractional
//   distances (t, s) from the general-quadrilateral quadratic with the "uprights parallel" and parallelogram fall-backs;
//   misses padded with source point 0 like _reduce_index_array does.
// The k-nearest part runs on the fixed-grid engine: lattice points inside a disc around the target's grid position are
// ranked by exact float64 distance; the disc grows until every quadrant's nearest pixel (or `neighbours` pixels) lie
// inside the radius the Euclidean bound of this file guarantees to be complete.
struct BilArg {
    b200_geo_t dst;
    GeosDev G;
    const double2 *in_xy;  // source pixel centres in target-projection coordinates (NaN = invalid), by raw index
    int32_t *index_array;  // [n][4] compressed indices (may be NULL)
    int32_t *gather_index; // [n][4] raw indices (may be NULL)
    double *t, *s;
    int32_t neighbours;
    int32_t first_valid;   // raw index of compressed point 0 (-1: no valid source point)
    double p0x, p0y;       // its target-projection coordinates
    float w_max;
};

struct SrcProjArg {
    b200_geo_t src;
    b200_proj_t dst_proj;
    GeosFast fast;
};

__global__ void __launch_bounds__(256)
bil_source_xy_kernel(const __grid_constant__ SrcProjArg A, const uint8_t *__restrict__ valid_input,
                     double2 *__restrict__ in_xy) {
    const int64_t W = A.src.width;
    const int64_t n = A.src.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double x = nan, y = nan;
        if (valid_input[i]) {
            int64_t r = i / W;
            double lon, lat;
            if (A.fast.enabled)
                geos_fast_lonlat(A.fast, A.src.row0 + r, i - r * W, lon, lat);
            else
                b200_geo_lonlat(A.src, A.src.row0 + r, i - r * W, lon, lat);
            if (!b200_proj_forward(A.dst_proj, lon, lat, x, y)) x = y = __longlong_as_double(0x7ff0000000000000LL);
        }
        in_xy[i] = make_double2(x, y);
    }
}

__device__ __forceinline__ bool bil_outside(double v) { return !(v >= 0.0 && v <= 1.0); }

// roots of a x^2 + b x + c in [0, 1]: x1 unless outside, then x2; NaN if neither (bilinear/_base.py _solve_quadratic)
__device__ __forceinline__ double bil_quadratic(double a, double b, double c) {
    double disc = b * b - 4.0 * a * c;
    double sq = sqrt(disc);
    double x1 = (-b + sq) / (2.0 * a), x2 = (-b - sq) / (2.0 * a);
    double x = bil_outside(x1) ? x2 : x1;
    return bil_outside(x) ? __longlong_as_double(0x7ff8000000000000LL) : x;
}

__device__ __forceinline__ double bil_other(double f, double y1, double y2, double y3, double y4, double out_y) {
    double y21 = y2 - y1, y43 = y4 - y3;
    double g = (out_y - y1 - y21 * f) / (y3 + y43 * f - y1 - y21 * f);
    return bil_outside(g) ? __longlong_as_double(0x7ff8000000000000LL) : g;
}

__device__ void bil_fractions(const double2 p1, const double2 p2, const double2 p3, const double2 p4, double ox, double oy,
                              double &t, double &s) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    {   // general quadrilateral
        double x21 = p2.x - p1.x, x31 = p3.x - p1.x, x42 = p4.x - p2.x;
        double y21 = p2.y - p1.y, y31 = p3.y - p1.y, y42 = p4.y - p2.y;
        double a = x31 * y42 - y31 * x42;
        double b = oy * (x42 - x31) - ox * (y42 - y31) + x31 * p2.y - y31 * p2.x + y42 * p1.x - x42 * p1.y;
        double c = oy * x21 - ox * y21 + p1.x * p2.y - p2.x * p1.y;
        t = bil_quadratic(a, b, c);
        s = bil_other(t, p1.y, p3.y, p2.y, p4.y, oy);
    }
    if (bil_outside(t)) {  // uprights parallel
        double x21 = p2.x - p1.x, x31 = p3.x - p1.x, x43 = p4.x - p3.x;
        double y21 = p2.y - p1.y, y31 = p3.y - p1.y, y43 = p4.y - p3.y;
        double a = x21 * y43 - y21 * x43;
        double b = oy * (x43 - x21) - ox * (y43 - y21) + x21 * p3.y - y21 * p3.x + y43 * p1.x - x43 * p1.y;
        double c = oy * x31 - ox * y31 + p1.x * p3.y - p3.x * p1.y;
        s = bil_quadratic(a, b, c);
        t = bil_other(s, p1.y, p2.y, p3.y, p4.y, oy);
    }
    if (bil_outside(t)) {  // parallelogram
        double x21 = p2.x - p1.x, x31 = p3.x - p1.x, y21 = p2.y - p1.y, y31 = p3.y - p1.y;
        t = (x21 * (oy - p1.y) - y21 * (ox - p1.x)) / (x21 * y31 - y21 * x31);
        if (bil_outside(t)) t = nan;
        s = (ox - p1.x + x31 * t) / x21;  // sign as published
        if (bil_outside(s)) s = nan;
    }
}

// quadrant of a source point relative to the target centre (d = out - in): 0 UL, 1 UR, 2 LL, 3 LR, -1 none
__device__ __forceinline__ int bil_quadrant(double dx, double dy) {
    if (dx > 0.0 && dy < 0.0) return 0;
    if (dx < 0.0 && dy < 0.0) return 1;
    if (dx > 0.0 && dy > 0.0) return 2;
    if (dx < 0.0 && dy > 0.0) return 3;
    return -1;
}

__global__ void __launch_bounds__(128) bil_info_kernel(const __grid_constant__ BilArg Q) {
    const int64_t W = Q.dst.width;
    const int64_t n = Q.dst.nrows * W;
    const GeosDev &G = Q.G;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / W, col = i - r * W;
        double lon, lat;
        b200_geo_lonlat(Q.dst, Q.dst.row0 + r, col, lon, lat);
        const double ox = (double)col * Q.dst.area.pixel_size_x + Q.dst.area.x_ul;
        const double oy = (double)(Q.dst.row0 + r) * -Q.dst.area.pixel_size_y + Q.dst.area.y_ul;
        double bd[4] = {inf, inf, inf, inf};   // nearest squared distance per quadrant
        long long bi[4] = {-1, -1, -1, -1};
        double nd = inf;                       // nearest overall
        long long ni = -1;
        int n_found = 0;
        bool complete = false;                 // every pixel within the radius has been examined
        float fc = 0.f, fr = 0.f;
        bool have_pos = false;
        double tx = 0, ty = 0, tz = 0;
        if (b200_lonlat_valid(lon, lat)) {
            double lam, phi;
            b200_lonlat2xyz(lon, lat, tx, ty, tz, lam, phi);
            float rel = (float)b200_adjlon(lam - G.lam0);
            float s_rel, c_rel, sp, cp;
            sincosf(rel, &s_rel, &c_rel);
            sincosf((float)phi, &sp, &cp);
            float u = cp, v = G.rp2 * sp;
            float inv = rsqrtf(u * u + v * v);
            float cpg = u * inv, spg = v * inv;
            float rr = G.rp * rsqrtf(G.rp2 * cpg * cpg + spg * spg);
            float kf_;
            have_pos = nn_fixed_position(G, rr * cpg, rr * spg, s_rel, c_rel, fc, fr, kf_);
        }
        const float reach = Q.w_max + 1.0f;
        if (have_pos && fc > -reach && fc < (float)G.W + reach && fr > -reach && fr < (float)G.H + reach) {
            float w = 2.0f;  // ~13 lattice points: enough for the four quadrant pixels of a well-sampled target
            int n_examined = 0;
            for (;;) {
                if (w > Q.w_max) w = Q.w_max;
                const double bound = (double)(w - GEOS_POS_ERR) / (double)G.inv_step;  // metres: complete up to here
                const double bound2 = bound * bound;
                for (int q = 0; q < 4; ++q) {
                    bd[q] = inf;
                    bi[q] = -1;
                }
                nd = inf;
                ni = -1;
                n_found = 0;
                n_examined = 0;
                int n_in_bound = 0;
                const float w2 = w * w;
                const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
                for (int rr_ = rlo; rr_ <= rhi; ++rr_) {
                    const float dr = fr - (float)rr_;
                    const float wc2 = w2 - dr * dr;
                    if (!(wc2 >= 0.0f)) continue;
                    const float wc = sqrtf(wc2);
                    const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
                    for (int c = clo; c <= chi; ++c) {
                        const long long idx = (long long)rr_ * G.W + c;
                        const double2 *p = reinterpret_cast<const double2 *>(G.pts + idx);
                        const double2 a = __ldg(p), b = __ldg(p + 1);
                        const double d2 = b200_nn_d2(a.x, a.y, b.x, tx, ty, tz);
                        ++n_examined;
                        if (!(d2 < G.r2)) continue;  // invalid pixel (NaN) or outside the radius
                        ++n_found;
                        if (d2 <= bound2) ++n_in_bound;
                        if (d2 < nd || (d2 == nd && idx < ni)) {
                            nd = d2;
                            ni = idx;
                        }
                        const double2 xy = __ldg(Q.in_xy + idx);
                        const int q = bil_quadrant(ox - xy.x, oy - xy.y);
                        if (q >= 0 && (d2 < bd[q] || (d2 == bd[q] && idx < bi[q]))) {
                            bd[q] = d2;
                            bi[q] = idx;
                        }
                    }
                }
                complete = bound2 >= G.r2 || w >= Q.w_max;
                const bool all4 = bd[0] <= bound2 && bd[1] <= bound2 && bd[2] <= bound2 && bd[3] <= bound2;
                if (complete || all4 || n_in_bound >= Q.neighbours) break;
                w *= 1.6f;
            }
            // rank of each quadrant's pixel among all neighbours (nearest first); beyond `neighbours` it is not seen
            int rank[4] = {0, 0, 0, 0};
            if (n_examined >= Q.neighbours) {  // with fewer points examined no rank can reach `neighbours`
                const float w2 = w * w;
                const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
                for (int rr_ = rlo; rr_ <= rhi; ++rr_) {
                    const float dr = fr - (float)rr_;
                    const float wc2 = w2 - dr * dr;
                    if (!(wc2 >= 0.0f)) continue;
                    const float wc = sqrtf(wc2);
                    const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
                    for (int c = clo; c <= chi; ++c) {
                        const long long idx = (long long)rr_ * G.W + c;
                        const double2 *p = reinterpret_cast<const double2 *>(G.pts + idx);
                        const double2 a = __ldg(p), b = __ldg(p + 1);
                        const double d2 = b200_nn_d2(a.x, a.y, b.x, tx, ty, tz);
                        if (!(d2 < G.r2)) continue;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (bi[q] >= 0 && (d2 < bd[q] || (d2 == bd[q] && idx < bi[q]))) ++rank[q];
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (bi[q] >= 0 && rank[q] >= Q.neighbours) bi[q] = -1;
        }
        // corners; fewer than `neighbours` hits are padded with source point 0 (pyresample _reduce_index_array), which
        // becomes the corner of any quadrant it happens to lie in when no real neighbour does
        double2 pt[4];
        long long out_idx[4];
        const bool padded = n_found < Q.neighbours && Q.first_valid >= 0;
        const int q0 = padded ? bil_quadrant(ox - Q.p0x, oy - Q.p0y) : -1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (bi[q] >= 0) {
                pt[q] = __ldg(Q.in_xy + bi[q]);
                out_idx[q] = bi[q];
            } else if (q == q0) {
                pt[q] = make_double2(Q.p0x, Q.p0y);
                out_idx[q] = Q.first_valid;
            } else {
                pt[q] = make_double2(nan, nan);
                out_idx[q] = ni >= 0 ? ni : Q.first_valid;  // the reference keeps neighbour 0's index here
            }
        }
        double t, s;
        bil_fractions(pt[0], pt[1], pt[2], pt[3], ox, oy, t, s);
        Q.t[i] = t;
        Q.s[i] = s;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (Q.gather_index) Q.gather_index[4 * i + q] = (int32_t)out_idx[q];
            if (Q.index_array) Q.index_array[4 * i + q] = out_idx[q] < 0 ? 0 : __ldg(G.prefix + out_idx[q]);
        }
    }
}

__global__ void __launch_bounds__(256) bil_first_valid_kernel(const uint8_t *__restrict__ valid, int64_t n, int *out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (valid[i]) {
            atomicMin(out, (int)i);
            return;  // later pixels of this thread have larger indices
        }
}

extern "C" int b200_bil_info(const b200_nn_grid_t *grid, const uint8_t *valid_input, const b200_geo_t *dst,
                             double radius, int32_t neighbours, int32_t *index_array, int32_t *gather_index, double *t,
                             double *s, void *stream) {
    B200_CHECK_ARG(grid != nullptr, "grid is NULL");
    if (grid->mode != B200_NN_FIXED_GRID) {
        b200_set_error("b200_bil_info: bilinear resampling needs the fixed-grid search (a geostationary area source)");
        return B200_EUNSUPPORTED;
    }
    int rc = b200_check_geo(dst, "b200_bil_info");
    if (rc) return rc;
    if (dst->is_swath) {
        b200_set_error("b200_bil_info: the target must be an area (its projection is the interpolation plane)");
        return B200_EUNSUPPORTED;
    }
    B200_CHECK_ARG(radius > 0.0 && isfinite(radius), "radius_of_influence must be positive and finite");
    B200_CHECK_ARG(neighbours >= 4 && neighbours <= 1024, "neighbours must be in 4..1024");
    B200_CHECK_ARG(valid_input != nullptr && t != nullptr && s != nullptr, "NULL argument");
    const int64_t n = dst->nrows * dst->width;
    if (n == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const b200_geo_t &src = grid->src;
    const b200_proj_t &P = src.area.proj;
    BilArg Q;
    Q.dst = *dst;
    if (grid->light) {
        b200_set_error("b200_bil_info: needs a full search structure (float64 records), not a light one");
        return B200_EUNSUPPORTED;
    }
    GeosDev &G = Q.G;
    memset(&G.fast, 0, sizeof(G.fast));
    G.src_row0 = src.row0;
    G.pts = grid->pts;
    G.pts32 = grid->pts32;
    G.prune = 0;
    G.pad_ = 0;
    G.prefix = grid->prefix;
    G.W = (int32_t)src.width;
    G.H = (int32_t)src.nrows;
    G.sweep_x = P.mode == 1;
    G.lam0 = P.lam0;
    G.rg = (float)P.p[1];
    G.rp2 = (float)P.p[4];
    G.rp = (float)P.p[3];
    G.inv_rp2 = (float)P.p[5];
    const double hm = P.a * P.p[0];
    G.col_scale = (float)(hm / src.area.pixel_size_x);
    G.col_off = (float)((P.x0 - src.area.x_ul) / src.area.pixel_size_x);
    G.row_scale = (float)(-hm / src.area.pixel_size_y);
    G.row_off = (float)((src.area.y_ul - P.y0) / src.area.pixel_size_y - (double)src.row0);
    const double step = GEOS_BOUND_FACTOR * fmin(src.area.pixel_size_x, src.area.pixel_size_y);
    G.inv_step = (float)(1.0 / step);
    G.radius_f = (float)radius;
    G.r2_slack = (float)((radius + GEOS_F32_SLACK) * (radius + GEOS_F32_SLACK) * 1.000001);
    G.r2 = radius * radius;
    double rings = ceil(radius / step + GEOS_POS_ERR) + 1.0;
    if (rings > 4096.0) rings = 4096.0;
    G.max_ring = (int32_t)rings;
    G.key_mask = 0xfffffff0u;
    G.pad2_ = 0;
    G.Wm3 = G.Hm3 = 0;   // (the tiered nearest search is not used here)
    G.k_a = 0.0f;
    G.k_b = G.k_cap = 1.0f;
    G.vis_min = (float)(1.0 / P.p[1] - 1.5 * radius / P.a - 1e-5);
    Q.w_max = (float)rings;
    Q.neighbours = neighbours;
    Q.index_array = index_array;
    Q.gather_index = gather_index;
    Q.t = t;
    Q.s = s;

    double2 *in_xy = nullptr;
    int *d_first = nullptr;
    int h_first = 0x7fffffff;
    double2 h_p0 = make_double2(0.0, 0.0);
    B200_CUDA(cudaMallocAsync((void **)&in_xy, (size_t)grid->n_src * sizeof(double2), st));
    cudaError_t e = cudaMallocAsync((void **)&d_first, sizeof(int), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_first, &h_first, sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        SrcProjArg A;
        A.src = src;
        A.dst_proj = dst->area.proj;
        b200_geos_fast_prepare(&src, &A.fast, st);
        bil_source_xy_kernel<<<b200_grid_for(grid->n_src, 256, 8), 256, 0, st>>>(A, valid_input, in_xy);
        b200_geos_fast_release(&A.fast, st);
        bil_first_valid_kernel<<<b200_grid_for(grid->n_src, 256, 8), 256, 0, st>>>(valid_input, grid->n_src, d_first);
        e = cudaMemcpyAsync(&h_first, d_first, sizeof(int), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && h_first != 0x7fffffff)
        e = cudaMemcpy(&h_p0, in_xy + h_first, sizeof(double2), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
        Q.in_xy = in_xy;
        Q.first_valid = h_first == 0x7fffffff ? -1 : h_first;
        Q.p0x = h_p0.x;
        Q.p0y = h_p0.y;
        bil_info_kernel<<<b200_grid_for(n, 128, 8), 128, 0, st>>>(Q);
        e = cudaGetLastError();
    }
    if (in_xy) cudaFreeAsync(in_xy, st);
    if (d_first) cudaFreeAsync(d_first, st);
    if (e != cudaSuccess) {
        b200_set_error("b200_bil_info: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
    }
    return B200_OK;
}

// out = p1 (1-s)(1-t) + p2 s (1-t) + p3 (1-s) t + p4 s t in float64 (numpy promotes float32 data with the float64
// weights), NaN -> fill, cast to float32.  28 B of index/weight traffic + 4 B written per output pixel.
__global__ void __launch_bounds__(256)
bil_sample_f32_kernel(const float *__restrict__ data, float *__restrict__ out, const int32_t *__restrict__ gidx,
                      const double *__restrict__ t, const double *__restrict__ s, int64_t n_out, int nbands,
                      int64_t src_band_stride, int64_t dst_band_stride, float fill) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        const int4 ix = ld_stream_v4(gidx + 4 * i);
        const double tt = t[i], ss = s[i];
        const double w1 = (1.0 - ss) * (1.0 - tt), w2 = ss * (1.0 - tt), w3 = (1.0 - ss) * tt, w4 = ss * tt;
        for (int b = 0; b < nbands; ++b) {
            const float *src = data + b * src_band_stride;
            float v = fill;
            if (ix.x >= 0 && ix.y >= 0 && ix.z >= 0 && ix.w >= 0) {
                // ((p1 * (1-s)) * (1-t)) ... in the reference's order of operations
                double r1 = (double)__ldg(src + ix.x) * (1.0 - ss) * (1.0 - tt);
                double r2 = (double)__ldg(src + ix.y) * ss * (1.0 - tt);
                double r3 = (double)__ldg(src + ix.z) * (1.0 - ss) * tt;
                double r4 = (double)__ldg(src + ix.w) * ss * tt;
                double res = ((r1 + r2) + r3) + r4;
                v = isnan(res) ? fill : (float)res;
            }
            (void)w1; (void)w2; (void)w3; (void)w4;
            st_stream_f1(out + b * dst_band_stride + i, v);
        }
    }
}

// Bulk-staged variant (one band): the three input streams (4 corner indices, t, s: 32 B per output) come in through
// shared memory with cp.async.bulk + mbarrier, three tiles in flight per CTA, the result goes out the same way; the
// threads only issue the four scattered source loads.  Same structure as nn_gather_bulk_kernel (nn.cu).
#define BIL_TILE 1024
#define BIL_STAGES 3
struct __align__(128) BilStage {
    int32_t idx[4 * BIL_TILE];
    double t[BIL_TILE];
    double s[BIL_TILE];
    float out[BIL_TILE];
};

__global__ void __launch_bounds__(256)
bil_sample_bulk_kernel(const float *__restrict__ data, float *__restrict__ out, const int32_t *__restrict__ gidx,
                       const double *__restrict__ tt_, const double *__restrict__ ss_, int64_t n_tiles, float fill) {
    extern __shared__ __align__(128) unsigned char bil_smem[];
    BilStage *stage = reinterpret_cast<BilStage *>(bil_smem);
    __shared__ uint64_t full[BIL_STAGES];
    const int t = threadIdx.x;
    if (t == 0) {
        for (int k = 0; k < BIL_STAGES; ++k) mbar_init(&full[k], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    const uint32_t in_bytes = BIL_TILE * (16 + 8 + 8);
    auto load_tile = [&](int k, int64_t tile) {
        mbar_expect_tx(&full[k], in_bytes);
        bulk_g2s(stage[k].idx, gidx + tile * 4 * BIL_TILE, BIL_TILE * 16, &full[k]);
        bulk_g2s(stage[k].t, tt_ + tile * BIL_TILE, BIL_TILE * 8, &full[k]);
        bulk_g2s(stage[k].s, ss_ + tile * BIL_TILE, BIL_TILE * 8, &full[k]);
    };
    if (t == 0)
        for (int k = 0; k < BIL_STAGES; ++k)
            if (first + (int64_t)k * step < n_tiles) load_tile(k, first + (int64_t)k * step);
    int k = 0;
    uint32_t phase = 0;
    for (int64_t tile = first; tile < n_tiles; tile += step) {
        mbar_wait(&full[k], phase);
        if (t == 0) bulk_wait_read<BIL_STAGES - 1>();
        __syncthreads();
#pragma unroll
        for (int j = 0; j < BIL_TILE / 256; ++j) {
            const int o = j * 256 + t;
            const int4 ix = *reinterpret_cast<const int4 *>(stage[k].idx + 4 * o);
            const double tt = stage[k].t[o], ss = stage[k].s[o];
            float v = fill;
            if (ix.x >= 0 && ix.y >= 0 && ix.z >= 0 && ix.w >= 0) {
                // ((p1 * (1-s)) * (1-t)) ... in the reference's order of operations (as bil_sample_f32_kernel)
                double r1 = (double)__ldg(data + ix.x) * (1.0 - ss) * (1.0 - tt);
                double r2 = (double)__ldg(data + ix.y) * ss * (1.0 - tt);
                double r3 = (double)__ldg(data + ix.z) * (1.0 - ss) * tt;
                double r4 = (double)__ldg(data + ix.w) * ss * tt;
                double res = ((r1 + r2) + r3) + r4;
                v = isnan(res) ? fill : (float)res;
            }
            stage[k].out[o] = v;
        }
        bulk_fence_smem_writes();
        __syncthreads();
        if (t == 0) {
            bulk_s2g(out + tile * BIL_TILE, stage[k].out, BIL_TILE * 4);
            bulk_commit();
            const int64_t next = tile + (int64_t)BIL_STAGES * step;
            if (next < n_tiles) load_tile(k, next);
        }
        if (++k == BIL_STAGES) {
            k = 0;
            phase ^= 1;
        }
    }
    if (t == 0) bulk_wait_all();
}

extern "C" int b200_bil_sample_f32(const float *data, float *out, const int32_t *gather_index, const double *t,
                                   const double *s, int64_t n_out, int32_t nbands, int64_t src_band_stride,
                                   int64_t dst_band_stride, float fill, void *stream) {
    B200_CHECK_ARG(n_out >= 0 && nbands >= 1, "bad shape");
    if (n_out == 0) return B200_OK;
    B200_CHECK_ARG(data && out && gather_index && t && s, "NULL pointer");
    B200_CHECK_ARG(((uintptr_t)gather_index & 15) == 0, "gather_index must be 16-byte aligned");
    B200_CHECK_ARG(nbands == 1 || dst_band_stride >= n_out, "band stride smaller than one image");
    int64_t done = 0;
    static int variant = -1;   // B200_BIL_VARIANT: 0 direct, 1 bulk-staged (default)
    if (variant < 0) {
        const char *e = getenv("B200_BIL_VARIANT");
        variant = e ? (atoi(e) != 0) : 1;
    }
    const bool aligned = ((((uintptr_t)out) | ((uintptr_t)t) | ((uintptr_t)s)) & 15) == 0;
    if (variant == 1 && nbands == 1 && aligned && n_out >= BIL_TILE) {
        const int64_t n_tiles = n_out / BIL_TILE;
        const size_t smem = sizeof(BilStage) * BIL_STAGES + 128;
        static bool attr_set = false;
        if (!attr_set) {
            B200_CUDA(cudaFuncSetAttribute(bil_sample_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        int64_t grid = (int64_t)B200_NUM_SMS * 2;
        if (grid > n_tiles) grid = n_tiles;
        bil_sample_bulk_kernel<<<(int)grid, 256, smem, (cudaStream_t)stream>>>(data, out, gather_index, t, s, n_tiles, fill);
        B200_LAUNCH_CHECK();
        done = n_tiles * BIL_TILE;
        if (done == n_out) return B200_OK;
    }
    bil_sample_f32_kernel<<<b200_grid_for(n_out - done, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        data, out + done, gather_index + 4 * done, t + done, s + done, n_out - done, nbands, src_band_stride, dst_band_stride, fill);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
