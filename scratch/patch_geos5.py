p='satpy_b200/csrc/nn_geos.cu'
s=open(p).read()
old=s[s.index("    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above); the exact\n    // best distance so far is at most sqrt(m1) + 1 m"):s.index("    if (i1 < 0) return -1;  // no valid pixel anywhere near")]
new=r'''    // Any closer pixel lies inside the disc of radius w pixels around (fc, fr) (Euclidean bound above); the exact
    // best distance so far is at most sqrt(m1) + 1 m, and never needs to exceed the radius.
    const float bd = fminf(G.radius_f, sqrtf(m1) + GEOS_F32_ERR);
    const float w = bd * G.inv_step * 1.0001f + GEOS_POS_ERR;
    if (w >= 1.0f) {  // else every pixel outside the 2x2 block is at least one pixel away
        const float w2 = w * w;
        const int rlo = max(0, (int)ceilf(fr - w)), rhi = min(G.H - 1, (int)floorf(fr + w));
        // Where the lattice is locally a parallelogram lattice (it is, except next to the limb), most points of that
        // disc can be ruled out without loading them: with e_c, e_r the 3-D steps along a row / a column, the
        // distance to lattice point (i, j) is |h|^2 + Q(i - fx, j - fy), Q the quadratic form of the Gram matrix.
        // Only points with Q < (1.5 * best + 50 m)^2 are examined.  The model is used when all four block pixels are
        // valid and the fourth one lies within 2 % of a pixel of where the other three predict it (the measured
        // second-order term; over the few pixels of the disc it stays far inside the 50 % + 50 m margin).
        bool model = G.prune && (p00.x == p00.x) && (p01.x == p01.x) && (p10.x == p10.x) && (p11.x == p11.x);
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
            const float d0x = txf - p00.x, d0y = tyf - p00.y, d0z = tzf - p00.z;
            const float bc = d0x * ecx + d0y * ecy + d0z * ecz, br = d0x * erx + d0y * ery + d0z * erz;
            const float inv_det = 1.0f / det;
            const float fx = (bc * grr - br * gcr) * inv_det, fy = (br * gcc - bc * gcr) * inv_det;  // model position
            const float B = 1.5f * bd + 50.0f, B2 = B * B;
            const float hb = B * sqrtf(gcc * inv_det);
            const int jlo = max(rlo, r0 + (int)ceilf(fy - hb)), jhi = min(rhi, r0 + (int)floorf(fy + hb));
            const float inv_gcc = 1.0f / gcc;
            for (int r = jlo; r <= jhi; ++r) {
                const float b = (float)(r - r0) - fy;
                const float disc = gcr * b * gcr * b - gcc * (grr * b * b - B2);
                if (!(disc >= 0.0f)) continue;
                const float sq = sqrtf(disc), mid = -gcr * b;
                int clo = c0 + (int)ceilf(fx + (mid - sq) * inv_gcc), chi = c0 + (int)floorf(fx + (mid + sq) * inv_gcc);
                // never outside the rigorous disc / the grid
                const float dr = fr - (float)r;
                const float wc2 = w2 - dr * dr;
                if (!(wc2 >= 0.0f)) continue;
                const float wc = sqrtf(wc2);
                clo = max(clo, max(0, (int)ceilf(fc - wc)));
                chi = min(chi, min(G.W - 1, (int)floorf(fc + wc)));
                const bool block_row = (r == r0) || (r == r0 + 1);
                const int row_base = r * G.W;
                const float4 *row = G.pts32 + (long long)r * G.W;
                for (int c = clo; c <= chi; ++c) {
                    if (block_row && (c == c0 || c == c0 + 1)) continue;  // already examined
                    nn_track(nn_d2f(__ldg(row + c), txf, tyf, tzf), row_base + c, m1, i1, m2);
                }
            }
        } else {
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
    }
'''
s=s.replace(old,new)
# p00.. must be visible outside the ring-1 block: restructure declarations
s=s.replace('''    float m1 = inf, m2 = inf;
    int i1 = -1;
    {
        // the 2x2 enclosing pixels: four 128-bit loads in flight together
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const int i00 = r0 * G.W + c0;
        const float4 *q00 = G.pts32 + (long long)r0 * G.W + c0, *q10 = q00 + G.W;
        float4 p00 = make_float4(nanf_, 0.f, 0.f, 0.f), p01 = p00, p10 = p00, p11 = p00;''','''    float m1 = inf, m2 = inf;
    int i1 = -1;
    float4 p00 = make_float4(nanf_, 0.f, 0.f, 0.f), p01 = p00, p10 = p00, p11 = p00;
    {
        // the 2x2 enclosing pixels: four 128-bit loads in flight together
        const bool rk0 = (unsigned)r0 < (unsigned)G.H, rk1 = (unsigned)(r0 + 1) < (unsigned)G.H;
        const bool ck0 = (unsigned)c0 < (unsigned)G.W, ck1 = (unsigned)(c0 + 1) < (unsigned)G.W;
        const int i00 = r0 * G.W + c0;
        const float4 *q00 = G.pts32 + (long long)r0 * G.W + c0, *q10 = q00 + G.W;''')
s=s.replace("    int32_t sweep_x, max_ring;\n","    int32_t sweep_x, max_ring;\n    int32_t prune, pad_;            // 1: rule out disc points with the local lattice model (0: strict mode)\n")
s=s.replace("    G.pts = g->pts;\n    G.pts32 = g->pts32;\n","    G.pts = g->pts;\n    G.pts32 = g->pts32;\n    G.prune = g->strict ? 0 : 1;\n    G.pad_ = 0;\n",1)
# bilinear setup also sets fields; add prune there
s=s.replace("    G.pts = grid->pts;\n    G.pts32 = grid->pts32;\n","    G.pts = grid->pts;\n    G.pts32 = grid->pts32;\n    G.prune = 0;\n    G.pad_ = 0;\n")
# hidden path: one warp per segment does the fill
s=s.replace('''                if (FUSED1) {
                    const int n = min(256, W - seg * 256);
                    float *dst = Q.O.out + (int64_t)r * W + seg * 256;
                    const bool aligned = (((uintptr_t)dst) & 15) == 0;
                    if (aligned && (int)threadIdx.x * 4 + 3 < n) {
                        st_stream_f4(dst + threadIdx.x * 4, Q.O.fill, Q.O.fill, Q.O.fill, Q.O.fill);
                    } else if (aligned && (int)threadIdx.x * 4 < n) {
                        for (int j = threadIdx.x * 4; j < n; ++j) st_stream_f1(dst + j, Q.O.fill);
                    } else if (!aligned && col < W) {
                        st_stream_f1(Q.O.out + i, Q.O.fill);
                    }
                } else if (col < W) {''','''                if (FUSED1) {
                    if (threadIdx.x < 64) {  // two warps write the 1 KB of the segment, the others move on
                        const int n = min(256, W - seg * 256);
                        float *dst = Q.O.out + (int64_t)r * W + seg * 256;
                        // rows are 8-byte aligned at worst (even widths): peel to a 16-byte boundary
                        const int head = (int)(((16 - (((uintptr_t)dst) & 15)) & 15) >> 2);
                        const int j = head + (int)threadIdx.x * 4;
                        if (j + 3 < n) st_stream_f4(dst + j, Q.O.fill, Q.O.fill, Q.O.fill, Q.O.fill);
                        else for (int k2 = j; k2 < n; ++k2) st_stream_f1(dst + k2, Q.O.fill);
                        if ((int)threadIdx.x < head && (int)threadIdx.x < n) st_stream_f1(dst + threadIdx.x, Q.O.fill);
                    }
                } else if (col < W) {''')
open(p,'w').write(s)

p='satpy_b200/csrc/nn_common.cuh'
s=open(p).read()
s=s.replace("    int mode;  // B200_NN_BUCKET or B200_NN_FIXED_GRID\n","    int mode;  // B200_NN_BUCKET or B200_NN_FIXED_GRID\n    int strict;  // fixed grid: examine every point of the rigorous disc (no lattice-model pruning)\n")
s=s.replace("#define B200_NN_FIXED_GRID 2\n","#define B200_NN_FIXED_GRID 2\n#define B200_NN_FIXED_GRID_STRICT 3  /* request only: a fixed-grid structure with strict = 1 */\n")
open(p,'w').write(s)

p='satpy_b200/csrc/nn.cu'
s=open(p).read()
s=s.replace('''    B200_CHECK_ARG(method >= 0 && method <= 2, "method must be 0 (auto), 1 (bucket grid) or 2 (fixed grid)");''','''    B200_CHECK_ARG(method >= 0 && method <= 3, "method must be 0 (auto), 1 (bucket), 2 (fixed grid) or 3 (fixed grid, strict)");''')
s=s.replace('''    if (method == B200_NN_FIXED_GRID && !fixed_ok) {''','''    const int strict = method == B200_NN_FIXED_GRID_STRICT;
    if (strict) method = B200_NN_FIXED_GRID;
    if (method == B200_NN_FIXED_GRID && !fixed_ok) {''')
s=s.replace('''    g->n_src = n_src;
    g->radius = radius;
    g->src = *src;''','''    g->strict = strict;
    g->n_src = n_src;
    g->radius = radius;
    g->src = *src;''')
open(p,'w').write(s)

p='include/satpy_b200.h'
s=open(p).read()
s=s.replace('''/* Same, choosing the search engine: 0 = automatic, 1 = bucket grid (any source), 2 = fixed grid
 * (geostationary area sources only: the source lattice itself is the index, see csrc/nn_geos.cu;
 * B200_EUNSUPPORTED for other sources).  Both engines return identical indices. */''','''/* Same, choosing the search engine: 0 = automatic, 1 = bucket grid (any source), 2 = fixed grid
 * (geostationary area sources only: the source lattice itself is the index, see csrc/nn_geos.cu;
 * B200_EUNSUPPORTED for other sources), 3 = fixed grid in strict mode (every lattice point of the
 * rigorous search disc is examined; 2 rules most of them out with a local lattice model first).
 * All engines return identical indices. */''')
open(p,'w').write(s)
p='satpy_b200/resample/nearest.py'
s=open(p).read()
s=s.replace('SEARCH_METHODS = {"auto": 0, "bucket": 1, "fixed_grid": 2}','SEARCH_METHODS = {"auto": 0, "bucket": 1, "fixed_grid": 2, "fixed_grid_strict": 3}')
s=s.replace('"method": {1: "bucket", 2: "fixed_grid"}[used.value]}','"method": {1: "bucket", 2: "fixed_grid"}[used.value] + ("_strict" if method == "fixed_grid_strict" else "")}')
open(p,'w').write(s)
