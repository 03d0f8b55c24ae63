// Per-pixel ABI calibration arithmetic shared by calibrate.cu and the fused kernel in sunz.cu.
// Follows satpy/readers/core/abi.py:138-174 and satpy/readers/abi_l1b.py:133-173 in float32,
// with the reference's operation order (explicit round-to-nearest, no FMA contraction).
#pragma once
#include "b200_common.cuh"

__device__ __forceinline__ float b200_abi_cal_one(int raw16, const b200_abi_cal_t &c) {
    // raw16: the stored 16-bit pattern, sign-extended (int16 view)
    int v = c.is_unsigned ? (raw16 & 0xffff) : raw16;
    int fill = c.is_unsigned ? (c.fill & 0xffff) : c.fill;
    float f = (float)v;
    if (c.has_fill && v == fill) f = __int_as_float(0x7fc00000);
    // data * float32(scale) + float32(offset): two roundings, as numpy does it
    float rad = (c.scale_factor != 1.0f) ? __fadd_rn(__fmul_rn(f, c.scale_factor), c.add_offset) : f;
    if (c.mode == 0) return rad;
    if (c.mode == 1) return __fmul_rn(__fmul_rn(rad, c.refl_factor), 100.0f);
    if (c.clip_negative_radiances) rad = (rad < c.min_rad) ? c.min_rad : rad;  // NaN stays NaN like np.clip
    float t = __fadd_rn(__fdiv_rn(c.fk1, rad), 1.0f);
    t = __fdiv_rn(c.fk2, logf(t));
    return __fdiv_rn(__fsub_rn(t, c.bc1), c.bc2);
}

