p='satpy_b200/csrc/nn_geos.cu'
s=open(p).read()
# rename current exact search
s=s.replace('''// Search around the fractional grid position (fc, fr) of a target with spherical coordinates (tx, ty, tz).
__device__ __forceinline__ long long nn_fixed_search(const GeosDev &G, float fc, float fr, double tx, double ty,
                                                     double tz) {''','''// EXACT search (float64 distances, float32 pre-filter).  Out of line: it only runs for the rare targets whose
// float32 ranking below is ambiguous (equidistant pixels, neighbour at the radius).
__device__ __noinline__ long long nn_fixed_search_exact(const GeosDev &G, float fc, float fr, double tx, double ty,
                                                        double tz) {''')
b=s.index("// target lon/lat -> fractional position in the source grid")
new=r'''// two smallest float32 squared distances seen so far (NaN = invalid pixel is ignored by both comparisons)
__device__ __forceinline__ void nn_track(float d, int idx, float &m1, int &i1, float &m2) {
    if (d < m1) {
        m2 = m1;
        m1 = d;
        i1 = idx;
    } else if (d < m2) {
        m2 = d;
    }
}

// Search around the fractional grid position (fc, fr) of a target with spherical coordinates (tx, ty, tz).
// Hot path: float32 only.  e_i = float32 distance to the float32 copy of pixel i satisfies |e_i - d_i| < 1 m
// (GEOS_F32_ERR).  If the runner-up is more than 2 m behind the leader, the leader is the exact float64 winner;
// if the leader is more than 1 m inside (outside) the radius it is accepted (rejected) exactly as float64 would.
// Anything closer than that is re-done by nn_fixed_search_exact.
#define GEOS_F32_ERR 1.0f
__device__ __forceinline__ long long nn_fixed_search(const GeosDev &G, float fc, float fr, double tx, double ty,
                                                     double tz) {
    const float reach = (float)G.max_ring + 1.0f;
    if (!(fc > -reach) || !(fc < (float)G.W + reach) || !(fr > -reach) || !(fr < (float)G.H + reach)) return -1;
    const int c0 = (int)floorf(fc), r0 = (int)floorf(fr);
    const float txf = (float)tx, tyf = (float)ty, tzf = (float)tz;
    const float inf = __int_as_float(0x7f800000), nanf_ = __int_as_float(0x7fc00000);
    float m1 = inf, m2 = inf;
    int i1 = -1;
    {
        // the 2x2 enclosing pixels: four 128-bit loads in flight together
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const int i00 = r0 * G.W + c0;
        const float4 *q00 = G.pts32 + (long long)r0 * G.W + c0, *q10 = q00 + G.W;
        float4 p00 = make_float4(nanf_, 0.f, 0.f, 0.f), p01 = p00, p10 = p00, p11 = p00;
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
    const float bd = fminf(G.radius_f, sqrtf(m1) + GEOS_F32_ERR);
    const float w = bd * G.inv_step * 1.0001f + GEOS_POS_ERR;
    if (w >= 1.0f) {  // else every pixel outside the 2x2 block is at least one pixel away
        const float w2 = w * w;
        const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
        for (int r = rlo; r <= rhi; ++r) {
            const float dr = fr - (float)r;
            const float wc2 = w2 - dr * dr;
            if (!(wc2 >= 0.0f)) continue;
            const float wc = sqrtf(wc2);
            const int clo = max(0, (int)ceilf(fc - wc)), chi = min(G.W - 1, (int)floorf(fc + wc));
            const bool block_row = (r == r0) || (r == r0 + 1);
            const int row_base = r * G.W;
            const float4 *row = G.pts32 + (long long)r * G.W;
            for (int c = clo; c <= chi; ++c) {
                if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
                nn_track(nn_d2f(__ldg(row + c), txf, tyf, tzf), row_base + c, m1, i1, m2);
            }
        }
    }
    if (i1 < 0) return -1;  // no valid pixel anywhere near
    const float s1 = sqrtf(m1);
    const float t = s1 + 2.0f * GEOS_F32_ERR + 0.01f;
    const bool clear_winner = m2 > t * t * 1.000001f;
    const bool clear_radius = fabsf(s1 - G.radius_f) > GEOS_F32_ERR + 0.01f;
    if (clear_winner && clear_radius) return s1 < G.radius_f ? (long long)i1 : -1;
    return nn_fixed_search_exact(G, fc, fr, tx, ty, tz);
}

'''
s=s[:b]+new+s[b:]
s=s.replace("    float inv_step;                 // 1 / (0.975 * min pixel size)\n","    float inv_step;                 // 1 / (0.975 * min pixel size)\n    float radius_f;                 // radius of influence\n")
s=s.replace("    G.inv_step = (float)(1.0 / step);\n","    G.inv_step = (float)(1.0 / step);\n    G.radius_f = (float)radius;\n")
open(p,'w').write(s)
