// ABI L1b reader calibration on sm_100a.
//
// Replaces the per-pixel arithmetic of
//   NC_ABI_BASE._adjust_data     satpy/readers/core/abi.py:138-174
//   NC_ABI_L1B._vis_calibrate    satpy/readers/abi_l1b.py:133-145  (+ "*100" :67-69)
//   NC_ABI_L1B._ir_calibrate     satpy/readers/abi_l1b.py:157-173
// in one pass: 2 B read + 4 B written per pixel (the reference makes 3-5 passes).
//
// The float32 operation order of the reference is kept (explicit round-to-nearest
// mul/add, no FMA contraction) so radiance and reflectance are bit-identical to
// numpy; brightness temperature differs only by logf's last ulp.
#include <stdlib.h>
#include "b200_common.cuh"
#include "b200_cal.cuh"
#include "b200_bulk.cuh"

#ifndef B200_CAL_DEFAULT_VARIANT
#define B200_CAL_DEFAULT_VARIANT 1
#endif

struct AbiCalDev {
    b200_abi_cal_t c;
};

// 16 pixels per thread per iteration: two independent 128-bit loads in flight, four 128-bit stores.
// BT: the brightness-temperature instantiation (75 registers); radiance / reflectance keep their 38.
template <bool BT>
__global__ void __launch_bounds__(256, BT ? 3 : 6)
abi_calibrate_vec_kernel(const int16_t *__restrict__ counts, float *__restrict__ out, int64_t n_vec16,
                         const __grid_constant__ AbiCalDev P) {
    const b200_abi_cal_t &c = P.c;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec16; i += stride) {
        const int4 *src = reinterpret_cast<const int4 *>(counts) + 2 * i;
        int4 a = ld_stream_v4(src);
        int4 b = ld_stream_v4(src + 1);
        int w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float r[16];
#if !B200_BT_IEEE
        if (BT) {
            // brightness temperature: the hardware-unit chain for all 16 pixels, then the rare unsafe ones again
            unsigned slow_mask = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int raw = (k & 1) ? (int)(short)((unsigned)w[k >> 1] >> 16) : (int)(short)(w[k >> 1] & 0xffff);
                bool slow;
                r[k] = b200_bt_fast(b200_abi_rad(raw, c), c, slow);
                slow_mask |= (slow ? 1u : 0u) << k;
            }
            if (slow_mask) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (!((slow_mask >> k) & 1u)) continue;
                    const int raw = (k & 1) ? (int)(short)((unsigned)w[k >> 1] >> 16) : (int)(short)(w[k >> 1] & 0xffff);
                    r[k] = b200_bt_ieee(b200_abi_rad(raw, c), c);
                }
            }
        } else
#endif
        {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (BT) {  // only with B200_BT_IEEE
                    r[2 * k] = b200_abi_cal_one((int)(short)(w[k] & 0xffff), c);
                    r[2 * k + 1] = b200_abi_cal_one((int)(short)((unsigned)w[k] >> 16), c);
                } else {
                    r[2 * k] = b200_abi_cal_vis((int)(short)(w[k] & 0xffff), c);
                    r[2 * k + 1] = b200_abi_cal_vis((int)(short)((unsigned)w[k] >> 16), c);
                }
            }
        }
        float *dst = out + 16 * i;
#pragma unroll
        for (int k = 0; k < 4; ++k) st_stream_f4(dst + 4 * k, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 1: lane-interleaved accesses.  A warp-iteration covers 256 consecutive pixels: lane L loads the four counts
// 4L .. 4L+3 (one 8-byte load: the warp-level instruction reads 256 contiguous bytes, whole 32-byte sectors) and stores
// their four results with one 128-bit store (512 contiguous bytes per warp-level instruction).  The first kernel gave
// every thread 32 contiguous input bytes and 64 contiguous output bytes, so each warp-level LDG.128 / STG.128 touched
// half / a quarter of every sector it addressed and L2 merged the pieces (68 % of the HBM copy rate).
template <bool BT>
__device__ __forceinline__ void abi_cal4(const uint2 w, const b200_abi_cal_t &c, float r[4]) {
    const int raw[4] = {(int)(short)(w.x & 0xffff), (int)(short)(w.x >> 16), (int)(short)(w.y & 0xffff),
                        (int)(short)(w.y >> 16)};
#if !B200_BT_IEEE
    if (BT) {
        unsigned slow_mask = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool slow;
            r[k] = b200_bt_fast(b200_abi_rad(raw[k], c), c, slow);
            slow_mask |= (slow ? 1u : 0u) << k;
        }
        if (slow_mask) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((slow_mask >> k) & 1u) r[k] = b200_bt_ieee(b200_abi_rad(raw[k], c), c);
        }
        return;
    }
#endif
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = BT ? b200_abi_cal_one(raw[k], c) : b200_abi_cal_vis(raw[k], c);
}

__device__ __forceinline__ uint2 ld_stream_v2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

#define CAL_UNROLL 4   // independent 8-byte loads in flight per thread
template <bool BT>
__global__ void __launch_bounds__(256, BT ? 3 : 6)
abi_calibrate_lane_kernel(const int16_t *__restrict__ counts, float *__restrict__ out, int64_t n_vec4,
                          const __grid_constant__ AbiCalDev P) {
    const b200_abi_cal_t &c = P.c;
    // thread t of the grid handles quads t, t + T, t + 2T ... (T = grid threads): consecutive lanes, consecutive quads
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + (CAL_UNROLL - 1) * T < n_vec4; q += CAL_UNROLL * T) {
        uint2 w[CAL_UNROLL];
#pragma unroll
        for (int u = 0; u < CAL_UNROLL; ++u) w[u] = ld_stream_v2(counts + 4 * (q + u * T));
#pragma unroll
        for (int u = 0; u < CAL_UNROLL; ++u) {
            float r[4];
            abi_cal4<BT>(w[u], c, r);
            st_stream_f4(out + 4 * (q + u * T), r[0], r[1], r[2], r[3]);
        }
    }
    for (; q < n_vec4; q += T) {
        float r[4];
        abi_cal4<BT>(ld_stream_v2(counts + 4 * q), c, r);
        st_stream_f4(out + 4 * q, r[0], r[1], r[2], r[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 2: bulk-asynchronous staging (TMA's 1-D form).  One thread of the CTA issues cp.async.bulk global -> shared
// for a tile of CAL_TILE counts (8 KB) and the tile's arrival is signalled on an mbarrier (complete_tx); all threads
// convert from shared memory into a shared output tile (16 KB), which one thread sends back with cp.async.bulk
// shared -> global (bulk_group).  CAL_STAGES tiles are in flight per CTA; no thread issues a global load or store.
// SASS: UBLKCP (both directions), SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK (the mbarrier).
#define CAL_TILE 4096
#define CAL_STAGES 3

struct __align__(128) CalStage {
    float out[CAL_TILE];
    int16_t in[CAL_TILE];
};

template <bool BT>
__global__ void __launch_bounds__(256)
abi_calibrate_bulk_kernel(const int16_t *__restrict__ counts, float *__restrict__ out, int64_t n_tiles,
                          const __grid_constant__ AbiCalDev P) {
    extern __shared__ __align__(128) unsigned char cal_smem[];
    CalStage *stage = reinterpret_cast<CalStage *>(cal_smem);
    __shared__ uint64_t full[CAL_STAGES];
    const b200_abi_cal_t &c = P.c;
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < CAL_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    // prologue: the first CAL_STAGES tiles of this CTA
    if (t == 0) {
        for (int s = 0; s < CAL_STAGES; ++s) {
            const int64_t tile = first + (int64_t)s * step;
            if (tile < n_tiles) {
                mbar_expect_tx(&full[s], CAL_TILE * 2);
                bulk_g2s(stage[s].in, counts + tile * CAL_TILE, CAL_TILE * 2, &full[s]);
            }
        }
    }
    int s = 0;
    uint32_t phase = 0;
    for (int64_t tile = first; tile < n_tiles; tile += step) {
        mbar_wait(&full[s], phase);
        // the bulk store that last read stage[s].out (CAL_STAGES tiles ago) must have finished reading it
        if (t == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(CAL_STAGES - 1) : "memory");
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CAL_TILE / 1024; ++j) {
            const uint2 w = *reinterpret_cast<const uint2 *>(stage[s].in + j * 1024 + 4 * t);
            float r[4];
            abi_cal4<BT>(w, c, r);
            *reinterpret_cast<float4 *>(stage[s].out + j * 1024 + 4 * t) = make_float4(r[0], r[1], r[2], r[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk copy
        __syncthreads();
        if (t == 0) {
            bulk_s2g(out + tile * CAL_TILE, stage[s].out, CAL_TILE * 4);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            const int64_t next = tile + (int64_t)CAL_STAGES * step;
            if (next < n_tiles) {   // stage[s].in has been consumed by every thread (barrier above): refill it
                mbar_expect_tx(&full[s], CAL_TILE * 2);
                bulk_g2s(stage[s].in, counts + next * CAL_TILE, CAL_TILE * 2, &full[s]);
            }
        }
        if (++s == CAL_STAGES) {
            s = 0;
            phase ^= 1;
        }
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
}

__global__ void __launch_bounds__(256)
abi_calibrate_scalar_kernel(const int16_t *__restrict__ counts, float *__restrict__ out, int64_t start,
                            int64_t n, const __grid_constant__ AbiCalDev P) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = b200_abi_cal_one((int)counts[i], P.c);
}

static int g_cal_variant = -1;   // -1: not read yet; 0 per-thread blocks (first release), 1 lane-interleaved, 2 bulk-async

extern "C" int b200_abi_calibrate_variant(const int16_t *counts, float *out, int64_t n, const b200_abi_cal_t *cal,
                                          int32_t variant, void *stream);

extern "C" int b200_abi_calibrate(const int16_t *counts, float *out, int64_t n, const b200_abi_cal_t *cal,
                                  void *stream) {
    if (g_cal_variant < 0) {
        const char *e = getenv("B200_CAL_VARIANT");
        g_cal_variant = e ? atoi(e) : B200_CAL_DEFAULT_VARIANT;
        if (g_cal_variant < 0 || g_cal_variant > 2) g_cal_variant = B200_CAL_DEFAULT_VARIANT;
    }
    return b200_abi_calibrate_variant(counts, out, n, cal, g_cal_variant, stream);
}

extern "C" int b200_abi_calibrate_variant(const int16_t *counts, float *out, int64_t n, const b200_abi_cal_t *cal,
                                          int32_t variant, void *stream) {
    B200_CHECK_ARG(cal != nullptr, "cal is NULL");
    B200_CHECK_ARG(n >= 0, "negative size");
    B200_CHECK_ARG(cal->mode >= 0 && cal->mode <= 2, "mode must be 0 (radiance), 1 (reflectance) or 2 (BT)");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(counts != nullptr && out != nullptr, "NULL data pointer");
    cudaStream_t s = (cudaStream_t)stream;
    AbiCalDev P;
    P.c = *cal;
    B200_CHECK_ARG(variant >= 0 && variant <= 2, "variant must be 0, 1 or 2");
    bool aligned = (((uintptr_t)counts & 15) == 0) && (((uintptr_t)out & 15) == 0);
    int64_t done = 0;
    if (aligned && variant == 2 && n >= CAL_TILE) {
        const int64_t n_tiles = n / CAL_TILE;
        const size_t smem = sizeof(CalStage) * CAL_STAGES + 128;
        static bool attr_set[2] = {false, false};
        const int bt = cal->mode == 2;
        if (!attr_set[bt]) {
            if (bt) B200_CUDA(cudaFuncSetAttribute(abi_calibrate_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else B200_CUDA(cudaFuncSetAttribute(abi_calibrate_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[bt] = true;
        }
        int64_t grid = (int64_t)B200_NUM_SMS * 3;
        if (grid > n_tiles) grid = n_tiles;
        if (bt) abi_calibrate_bulk_kernel<true><<<(int)grid, 256, smem, s>>>(counts, out, n_tiles, P);
        else abi_calibrate_bulk_kernel<false><<<(int)grid, 256, smem, s>>>(counts, out, n_tiles, P);
        B200_LAUNCH_CHECK();
        done = n_tiles * CAL_TILE;
    } else if (aligned && variant == 1 && n >= 4) {
        const int64_t n_vec4 = n / 4;
        int grid = b200_grid_for((n_vec4 + CAL_UNROLL - 1) / CAL_UNROLL, 256, 8);
        if (cal->mode == 2) abi_calibrate_lane_kernel<true><<<grid, 256, 0, s>>>(counts, out, n_vec4, P);
        else abi_calibrate_lane_kernel<false><<<grid, 256, 0, s>>>(counts, out, n_vec4, P);
        B200_LAUNCH_CHECK();
        done = n_vec4 * 4;
    } else if (aligned && n >= 16) {
        const int64_t n_vec = n / 16;
        int grid = b200_grid_for(n_vec, 256, 8);
        if (cal->mode == 2) abi_calibrate_vec_kernel<true><<<grid, 256, 0, s>>>(counts, out, n_vec, P);
        else abi_calibrate_vec_kernel<false><<<grid, 256, 0, s>>>(counts, out, n_vec, P);
        B200_LAUNCH_CHECK();
        done = n_vec * 16;
    }
    if (done < n) {
        int grid = b200_grid_for(n - done, 256, 8);
        abi_calibrate_scalar_kernel<<<grid, 256, 0, s>>>(counts, out, done, n, P);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}
