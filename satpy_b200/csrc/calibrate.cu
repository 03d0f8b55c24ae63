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
#include "b200_common.cuh"
#include "b200_cal.cuh"

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

__global__ void __launch_bounds__(256)
abi_calibrate_scalar_kernel(const int16_t *__restrict__ counts, float *__restrict__ out, int64_t start,
                            int64_t n, const __grid_constant__ AbiCalDev P) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = b200_abi_cal_one((int)counts[i], P.c);
}

extern "C" int b200_abi_calibrate(const int16_t *counts, float *out, int64_t n, const b200_abi_cal_t *cal,
                                  void *stream) {
    B200_CHECK_ARG(cal != nullptr, "cal is NULL");
    B200_CHECK_ARG(n >= 0, "negative size");
    B200_CHECK_ARG(cal->mode >= 0 && cal->mode <= 2, "mode must be 0 (radiance), 1 (reflectance) or 2 (BT)");
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(counts != nullptr && out != nullptr, "NULL data pointer");
    cudaStream_t s = (cudaStream_t)stream;
    AbiCalDev P;
    P.c = *cal;
    bool aligned = (((uintptr_t)counts & 15) == 0) && (((uintptr_t)out & 15) == 0);
    int64_t n_vec = aligned ? n / 16 : 0;
    if (n_vec > 0) {
        int grid = b200_grid_for(n_vec, 256, 8);
        if (cal->mode == 2) abi_calibrate_vec_kernel<true><<<grid, 256, 0, s>>>(counts, out, n_vec, P);
        else abi_calibrate_vec_kernel<false><<<grid, 256, 0, s>>>(counts, out, n_vec, P);
        B200_LAUNCH_CHECK();
    }
    int64_t done = n_vec * 16;
    if (done < n) {
        int grid = b200_grid_for(n - done, 256, 8);
        abi_calibrate_scalar_kernel<<<grid, 256, 0, s>>>(counts, out, done, n, P);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}
