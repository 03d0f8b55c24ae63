p='satpy_b200/csrc/nn_geos.cu'
s=open(p).read()

old=s[s.index("// target lon/lat -> fractional position in the source grid"):s.index("// general targets: lon/lat per pixel from the target projection")]
new=r'''// target lon/lat -> fractional position in the source grid (PROJ geos forward, float32); false: far side.
// Fast reciprocal / rsqrt are fine here: only the lattice position is needed (error << GEOS_POS_ERR pixels).
__device__ __forceinline__ bool nn_fixed_position(const GeosDev &G, float rc, float rs, float s_rel, float c_rel,
                                                  float &fc, float &fr) {
    float vx = rc * c_rel, vy = rc * s_rel, vz = rs;
    if (!(vx >= G.vis_min)) return false;
    float tmp = G.rg - vx;
    float ax, ay;
    if (G.sweep_x) {
        ax = atanf(vy * rsqrtf(fmaf(vz, vz, tmp * tmp)));
        ay = atanf(__fdividef(vz, tmp));
    } else {
        ax = atanf(__fdividef(vy, tmp));
        ay = atanf(vz * rsqrtf(fmaf(vy, vy, tmp * tmp)));
    }
    fc = fmaf(ax, G.col_scale, G.col_off);
    fr = fmaf(ay, G.row_scale, G.row_off);
    return true;
}

'''
s=s.replace(old,new)

s=s.replace('''    const float w = sqrtf((float)(best.d2 * G.inv_step2)) * 1.0001f + GEOS_POS_ERR;
    const float w2 = w * w;''','''    const float w = sqrtf((float)(best.d2 * G.inv_step2)) * 1.0001f + GEOS_POS_ERR;
    if (w < 1.0f) return best.idx;  // every pixel outside the 2x2 block is at least one pixel away
    const float w2 = w * w;''')

old=s[s.index("// rectilinear targets: one work item = 256 consecutive columns of one row"):s.index("int b200_nn_fixed_grid_query(")]
new=r'''// per 256-column segment of a rectilinear target: the largest cos(lon - lon_0) of its columns.  A segment whose every
// column is on the far side for a given row (rc * cmax < vis_min) is filled without any per-pixel work.
__global__ void __launch_bounds__(256)
nn_seg_cmax_kernel(const TgtCol *__restrict__ cols, int W, int nseg, float *__restrict__ seg_cmax) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nseg) return;
    float m = -2.0f;
    for (int c = warp * 256 + lane; c < min(W, warp * 256 + 256); c += 32) m = fmaxf(m, cols[c].c_rel);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) seg_cmax[warp] = m;
}

// rectilinear targets: one work item = 256 consecutive columns of one row, tables instead of trigonometry; items are
// handed out in order through an atomic counter (chunks of GEOS_CHUNK consecutive segments).
// FUSED1: the hot configuration -- fused gather of one float32 band, no index outputs.
#define GEOS_CHUNK 8
template <bool FUSED1>
__global__ void __launch_bounds__(256, 4)
nn_fixed_query_rect_kernel(const __grid_constant__ GeosQueryArg Q, unsigned long long *__restrict__ queue,
                           const float *__restrict__ seg_cmax) {
    const int W = (int)Q.dst.width;
    const int nseg = (W + 255) >> 8;
    const long long items = (long long)Q.dst.nrows * nseg;
    __shared__ int s_row, s_seg, s_count;
    for (;;) {
        if (threadIdx.x == 0) {
            long long first = (long long)atomicAdd(queue, (unsigned long long)GEOS_CHUNK);
            long long left = items - first;
            s_count = left <= 0 ? 0 : (left < GEOS_CHUNK ? (int)left : GEOS_CHUNK);
            if (left > 0) {
                long long r = first / nseg;
                s_row = (int)r;
                s_seg = (int)(first - r * nseg);
            }
        }
        __syncthreads();
        const int count = s_count;
        int r = s_row, seg = s_seg;
        __syncthreads();
        if (count == 0) break;
        for (int k = 0; k < count; ++k, ++seg) {
            if (seg == nseg) {
                seg = 0;
                ++r;
            }
            const TgtRow R = Q.rows[r];
            const int col = seg * 256 + (int)threadIdx.x;
            const int64_t i = (int64_t)r * W + col;
            if (!R.valid || !(R.rc * __ldg(seg_cmax + seg) >= Q.G.vis_min)) {
                // nothing of this segment can see the satellite: every pixel is a miss
                if (FUSED1) {
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
                } else if (col < W) {
                    b200_nn_emit(Q.O, Q.G.prefix, i, -1, R.valid && Q.cols[col].valid);
                }
                continue;
            }
            if (col >= W) continue;
            const TgtCol Cc = Q.cols[col];
            const bool ok = Cc.valid;
            long long raw = -1;
            float fc, fr;
            if (ok && nn_fixed_position(Q.G, R.rc, R.rs, Cc.s_rel, Cc.c_rel, fc, fr)) {
                double tx = B200_R_EARTH * R.cos_phi * Cc.cos_lam;
                double ty = B200_R_EARTH * R.cos_phi * Cc.sin_lam;
                double tz = B200_R_EARTH * R.sin_phi;
                raw = nn_fixed_search(Q.G, fc, fr, tx, ty, tz);
            }
            if (FUSED1)
                st_stream_f1(Q.O.out + i, raw < 0 ? Q.O.fill : __ldg(Q.O.data + raw));
            else
                b200_nn_emit(Q.O, Q.G.prefix, i, raw, ok);
        }
    }
}

'''
s=s.replace(old,new)
old=s[s.index("        unsigned long long *queue = nullptr;"):s.index("        cudaFreeAsync(cols, s);")]
new=r'''        unsigned long long *queue = nullptr;
        float *seg_cmax = nullptr;
        const int nseg = (int)((dst->width + 255) / 256);
        cudaError_t e = cudaMallocAsync((void **)&queue, sizeof(unsigned long long), s);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&seg_cmax, (size_t)nseg * sizeof(float), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s);
        if (e == cudaSuccess) {
            nn_seg_cmax_kernel<<<(nseg * 32 + 255) / 256, 256, 0, s>>>(cols, (int)dst->width, nseg, seg_cmax);
            int64_t items = dst->nrows * (int64_t)nseg;
            int grid = b200_grid_for((items + GEOS_CHUNK - 1) / GEOS_CHUNK, 1, 4);
            const bool fused1 = O.out != nullptr && O.nbands == 1 && !O.index_array && !O.gather_index &&
                                !O.valid_output;
            if (fused1)
                nn_fixed_query_rect_kernel<true><<<grid, 256, 0, s>>>(Q, queue, seg_cmax);
            else
                nn_fixed_query_rect_kernel<false><<<grid, 256, 0, s>>>(Q, queue, seg_cmax);
            e = cudaGetLastError();
        }
        if (queue) cudaFreeAsync(queue, s);
        if (seg_cmax) cudaFreeAsync(seg_cmax, s);
'''
s=s.replace(old,new)
open(p,'w').write(s)
