p='satpy_b200/csrc/nn_geos.cu'
s=open(p).read()
a=s.index("#define GEOS_CHUNK 8")
b=s.index("int b200_nn_fixed_grid_query(")
new=r'''#define GEOS_CHUNK 16
template <bool FUSED1>
__global__ void __launch_bounds__(256, 4)
nn_fixed_query_rect_kernel(const __grid_constant__ GeosQueryArg Q, unsigned long long *__restrict__ queue,
                           const float *__restrict__ seg_cmax) {
    const int W = (int)Q.dst.width;
    const int nseg = (W + 255) >> 8;
    const long long items = (long long)Q.dst.nrows * nseg;
    const int lane = threadIdx.x & 31;
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
        // Which items of the chunk can see the satellite at all?  Lane k looks at item k, so the (row, segment) table
        // reads of the whole chunk are in flight together instead of one dependent read per item.
        bool vis = false;
        if (lane < count) {
            int rk = r, sk = seg + lane;
            while (sk >= nseg) {
                sk -= nseg;
                ++rk;
            }
            const TgtRow *Rk = Q.rows + rk;
            vis = Rk->valid && (Rk->rc * __ldg(seg_cmax + sk) >= Q.G.vis_min);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, vis);
        int64_t row_base = (int64_t)r * W;
        for (int k = 0; k < count; ++k, ++seg) {
            if (seg == nseg) {
                seg = 0;
                ++r;
                row_base += W;
            }
            const int col = seg * 256 + (int)threadIdx.x;
            const int64_t i = row_base + col;
            if (!((vmask >> k) & 1u)) {
                // nothing of this segment can see the satellite: every pixel is a miss
                if (FUSED1) {
                    if (threadIdx.x < 64) {  // two warps write the 1 KB of the segment, the others move on
                        const int n = min(256, W - seg * 256);
                        float *dst = Q.O.out + row_base + seg * 256;
                        // rows are 8-byte aligned at worst (even widths): peel to a 16-byte boundary
                        const int head = (int)(((16 - (((uintptr_t)dst) & 15)) & 15) >> 2);
                        const int j = head + (int)threadIdx.x * 4;
                        if (j + 3 < n) st_stream_f4(dst + j, Q.O.fill, Q.O.fill, Q.O.fill, Q.O.fill);
                        else for (int k2 = j; k2 < n; ++k2) st_stream_f1(dst + k2, Q.O.fill);
                        if ((int)threadIdx.x < head && (int)threadIdx.x < n) st_stream_f1(dst + threadIdx.x, Q.O.fill);
                    }
                } else if (col < W) {
                    b200_nn_emit(Q.O, Q.G.prefix, i, -1, Q.rows[r].valid && Q.cols[col].valid);
                }
                continue;
            }
            if (col >= W) continue;
            const TgtRow R = Q.rows[r];
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
s=s[:a]+new+s[b:]
# tighter model margin
s=s.replace("const float B = 1.5f * bd + 50.0f, B2 = B * B;","const float B = 1.25f * bd + 30.0f, B2 = B * B;")
s=s.replace("// Only points with Q < (1.5 * best + 50 m)^2 are examined.","// Only points with Q < (1.25 * best + 30 m)^2 are examined.")
s=s.replace("second-order term; over the few pixels of the disc it stays far inside the 50 % + 50 m margin).","second-order term; over the few pixels of the disc it stays inside the 25 % + 30 m margin).")
open(p,'w').write(s)
p='DESIGN.md'
s=open(p).read()
s=s.replace("places farther than 1.5 × best + 50 m are ruled out","places farther than 1.25 × best + 30 m are ruled out")
open(p,'w').write(s)
