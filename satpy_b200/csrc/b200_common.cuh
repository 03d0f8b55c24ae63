// Shared helpers for the sm_100a kernels of libsatpy_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/satpy_b200.h"

#define B200_NUM_SMS 148

void b200_set_error(const char *fmt, ...);
void b200_pool_keep_memory();

#define B200_CHECK_ARG(cond, msg)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            b200_set_error("%s: %s", __func__, msg);      \
            return B200_EINVAL;                           \
        }                                                 \
    } while (0)

#define B200_CUDA(call)                                                             \
    do {                                                                            \
        cudaError_t err__ = (call);                                                 \
        if (err__ != cudaSuccess) {                                                 \
            b200_set_error("%s: %s failed: %s", __func__, #call,                    \
                           cudaGetErrorString(err__));                              \
            return err__ == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;   \
        }                                                                           \
    } while (0)

#define B200_LAUNCH_CHECK()                                                         \
    do {                                                                            \
        cudaError_t err__ = cudaGetLastError();                                     \
        if (err__ != cudaSuccess) {                                                 \
            b200_set_error("%s: kernel launch failed: %s", __func__,                \
                           cudaGetErrorString(err__));                              \
            return B200_ECUDA;                                                      \
        }                                                                           \
    } while (0)

// Streaming (read-once / write-once) 128-bit global accesses: keep them out of L1.
__device__ __forceinline__ int4 ld_stream_v4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void *p, const int4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f4(float *p, float a, float b, float c, float d) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f1(float *p, float a) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

static inline int b200_grid_for(int64_t work_items, int per_block, int ctas_per_sm) {
    // persistent-style grid: a multiple of the SM count, capped by the work available
    int64_t blocks = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)B200_NUM_SMS * ctas_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
