// Per-pixel ABI calibration arithmetic shared by calibrate.cu and the fused kernel in sunz.cu.
// Follows satpy/readers/core/abi.py:138-174 and satpy/readers/abi_l1b.py:133-173 in float32,
// with the reference's operation order (explicit round-to-nearest, no FMA contraction).
#pragma once
#include "b200_common.cuh"
#ifndef B200_BT_IEEE
#define B200_BT_IEEE 0  /* 1: IEEE divisions + logf in the brightness-temperature branch */
#endif

// brightness temperature with correctly rounded divisions and the libm logarithm: ~100 instructions per pixel
static __device__ __noinline__ float b200_bt_ieee(float rad, const b200_abi_cal_t &c) {
    float t = __fadd_rn(__fdiv_rn(c.fk1, rad), 1.0f);
    t = __fdiv_rn(c.fk2, logf(t));
    return __fdiv_rn(__fsub_rn(t, c.bc1), c.bc2);
}

// counts -> radiance (with the negative-radiance clip of the brightness-temperature branch)
__device__ __forceinline__ float b200_abi_rad(int raw16, const b200_abi_cal_t &c) {
    // raw16: the stored 16-bit pattern, sign-extended (int16 view)
    int v = c.is_unsigned ? (raw16 & 0xffff) : raw16;
    int fill = c.is_unsigned ? (c.fill & 0xffff) : c.fill;
    float f = (float)v;
    if (c.has_fill && v == fill) f = __int_as_float(0x7fc00000);
    // data * float32(scale) + float32(offset): two roundings, as numpy does it
    float rad = (c.scale_factor != 1.0f) ? __fadd_rn(__fmul_rn(f, c.scale_factor), c.add_offset) : f;
    if (c.mode == 2 && c.clip_negative_radiances) rad = (rad < c.min_rad) ? c.min_rad : rad;  // NaN stays NaN like np.clip
    return rad;
}

// Planck inversion with the hardware reciprocal / log2 units (2-3 ulp each): within 3.5e-7 relative of the correctly
// rounded chain on a full ABI C13 image (north_star's bar is 1e-5), and HBM bound instead of issue bound.  numpy's own
// float32 log is not correctly rounded either, so neither variant is bit-identical to the reference.
// Below t = 2 the short cuts are not safe (__logf is only 2^-21 ABSOLUTE there, and for radiance near -fk1 the sum
// fk1/rad + 1 cancels and magnifies the division's 2 ulp): such pixels -- radiance above fk1 or below -fk1, no physical
// scene, but the reference's test vector has them -- are flagged `slow` and redone with b200_bt_ieee by the caller.
__device__ __forceinline__ float b200_bt_fast(float rad, const b200_abi_cal_t &c, bool &slow) {
    float t = __fadd_rn(__fdividef(c.fk1, rad), 1.0f);
    // fill values (NaN) and radiances in (-fk1, 0) give the logarithm of a negative number: NaN like numpy
    slow = (t >= -1e-3f) && (t < 2.0f);
    t = __fdividef(c.fk2, __logf(t));  // NaN in, or t < 0: NaN out
    return __fdividef(__fsub_rn(t, c.bc1), c.bc2);
}

// radiance (mode 0) or reflectance in % (mode 1) only
__device__ __forceinline__ float b200_abi_cal_vis(int raw16, const b200_abi_cal_t &c) {
    float rad = b200_abi_rad(raw16, c);
    return c.mode == 0 ? rad : __fmul_rn(__fmul_rn(rad, c.refl_factor), 100.0f);
}

__device__ __forceinline__ float b200_abi_cal_one(int raw16, const b200_abi_cal_t &c) {
    float rad = b200_abi_rad(raw16, c);
    if (c.mode == 0) return rad;
    if (c.mode == 1) return __fmul_rn(__fmul_rn(rad, c.refl_factor), 100.0f);
#if B200_BT_IEEE
    return b200_bt_ieee(rad, c);
#else
    bool slow;
    float bt = b200_bt_fast(rad, c, slow);
    return slow ? b200_bt_ieee(rad, c) : bt;
#endif
}
