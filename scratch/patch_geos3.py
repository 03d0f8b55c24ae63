p='satpy_b200/csrc/nn_geos.cu'
s=open(p).read()
a=s.index("__device__ __forceinline__ void nn_consider_loaded(")
b=s.index("// target lon/lat -> fractional position in the source grid")
new=r'''// float32 squared distance to a float32 copy of a source point: a PRE-FILTER only.  Both points carry at most
// 0.25 m of rounding per coordinate (|coordinate| < 2^23 m), so |d_f32 - d_exact| < 1 m; a candidate whose float32
// distance exceeds the best exact distance by more than GEOS_F32_SLACK metres can never win (or tie) and is skipped
// without touching its float64 record.  NaN (invalid pixel) fails the comparison and is skipped too.
#define GEOS_F32_SLACK 3.0f
__device__ __forceinline__ float nn_d2f(const float4 &p, float tx, float ty, float tz) {
    float dx = p.x - tx, dy = p.y - ty, dz = p.z - tz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
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
        if (d00 <= thr2) b200_nn_consider(G.pts + i00, tx, ty, tz, best);
        if (d01 <= thr2) b200_nn_consider(G.pts + i00 + 1, tx, ty, tz, best);
        if (d10 <= thr2) b200_nn_consider(G.pts + i00 + G.W, tx, ty, tz, best);
        if (d11 <= thr2) b200_nn_consider(G.pts + i00 + G.W + 1, tx, ty, tz, best);
    }
    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above).
    const float bd = sqrtf((float)best.d2);
    const float w = bd * G.inv_step * 1.0001f + GEOS_POS_ERR;
    if (w < 1.0f) return best.idx;  // every pixel outside the 2x2 block is at least one pixel away
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
                b200_nn_consider(G.pts + row_base + c, tx, ty, tz, best);
                if (best.d2 < before) {
                    const float t = sqrtf((float)best.d2) * 1.000001f + GEOS_F32_SLACK;
                    thr2 = fminf(thr2, t * t * 1.000001f);
                }
            }
        }
    }
    return best.idx;
}

'''
s=s[:a]+new+s[b:]
s=s.replace("    const double4 *pts;\n    const int32_t *prefix;\n    int32_t W, H;","    const double4 *pts;\n    const float4 *pts32;   // float32 copy of the points (pre-filter)\n    const int32_t *prefix;\n    int32_t W, H;")
s=s.replace("    double inv_step2;               // 1 / (0.975 * min pixel size)^2\n","    float inv_step;                 // 1 / (0.975 * min pixel size)\n    float r2_slack;                 // (radius + GEOS_F32_SLACK)^2 * (1 + eps)\n")
s=s.replace("    G.inv_step2 = 1.0 / (step * step);\n","    G.inv_step = (float)(1.0 / step);\n    G.r2_slack = (float)((radius + GEOS_F32_SLACK) * (radius + GEOS_F32_SLACK) * 1.000001);\n")
s=s.replace("    G.pts = g->pts;\n","    G.pts = g->pts;\n    G.pts32 = g->pts32;\n")
open(p,'w').write(s)

p='satpy_b200/csrc/nn_common.cuh'
s=open(p).read()
s=s.replace("    int64_t nbytes;\n    // ---- bucket mode","    float4 *pts32;    // fixed grid only: float32 copy {x, y, z, 0} by raw index (distance pre-filter)\n    int64_t nbytes;\n    // ---- bucket mode")
open(p,'w').write(s)

p='satpy_b200/csrc/nn.cu'
s=open(p).read()
s=s.replace('''                        uint8_t *__restrict__ valid_input, double4 *__restrict__ rec, int32_t *__restrict__ cell_of,
                        const int32_t *__restrict__ band_off, double inv_dphi, int nbands) {''','''                        uint8_t *__restrict__ valid_input, double4 *__restrict__ rec, float4 *__restrict__ rec32,
                        int32_t *__restrict__ cell_of, const int32_t *__restrict__ band_off, double inv_dphi,
                        int nbands) {''')
s=s.replace('''        p[1] = make_double2(z, __longlong_as_double((long long)i));
        if (cell_of) cell_of[i]''','''        p[1] = make_double2(z, __longlong_as_double((long long)i));
        if (rec32) rec32[i] = make_float4((float)x, (float)y, (float)z, 0.0f);
        if (cell_of) cell_of[i]''')
s=s.replace('''    NN_TRY(cudaMallocAsync((void **)&rec, (size_t)n_src * sizeof(double4), s));
    nn_source_points_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(S, mask, valid_input, rec, cell_of,
                                                                        g->band_off, inv_dphi, nbands);''','''    NN_TRY(cudaMallocAsync((void **)&rec, (size_t)n_src * sizeof(double4), s));
    if (!bucket) NN_TRY(cudaMallocAsync((void **)&g->pts32, (size_t)n_src * sizeof(float4), s));
    nn_source_points_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(S, mask, valid_input, rec, g->pts32, cell_of,
                                                                        g->band_off, inv_dphi, nbands);''')
s=s.replace("        g->nbytes = n_src * (int64_t)(sizeof(double4) + sizeof(int32_t));","        g->nbytes = n_src * (int64_t)(sizeof(double4) + sizeof(float4) + sizeof(int32_t));")
s=s.replace("    if (grid->pts) cudaFreeAsync(grid->pts, s);\n","    if (grid->pts) cudaFreeAsync(grid->pts, s);\n    if (grid->pts32) cudaFreeAsync(grid->pts32, s);\n")
open(p,'w').write(s)
