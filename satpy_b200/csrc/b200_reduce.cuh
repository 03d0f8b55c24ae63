// SunZenithReducer's per-pixel factor (_sunzen_reduction_ndarray, satpy/modifiers/angles.py:594-624) for float32 data and a
// float32 solar zenith angle in degrees; shared by the kernel that takes the angle as a dataset (calibrate_more.cu) and
// the one that derives it from the area (sunz.cu).
#pragma once
#include "b200_common.cuh"

__device__ __forceinline__ float b200_sunz_reduction(float z, float limit, float max_sza, float strength) {
    const float nanf_ = __int_as_float(0x7fc00000);
    float rf = __fdiv_rn(__fsub_rn(z, limit), __fsub_rn(max_sza, limit));
    rf = fminf(fmaxf(rf, 0.0f), 1.0f);
    if (isnan(z)) rf = nanf_;  // clip keeps NaN
    rf = __fsub_rn(1.0f, log2f(__fadd_rn(rf, 1.0f)));
    const float p = powf(rf, strength), q = powf(__fsub_rn(1.0f, rf), strength);
    rf = __fdiv_rn(p, __fadd_rn(p, q));
    float corr = z < limit ? 1.0f : rf;
    if (isnan(z)) corr = 0.0f;
    return corr;
}
