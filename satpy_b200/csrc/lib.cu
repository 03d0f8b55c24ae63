// Library-level entry points of libsatpy_b200.so: version, thread-local error text,
// device probing and the lon/lat helper (AreaDefinition.get_lonlats stand-in,
// satpy/modifiers/angles.py:438).
#include <stdarg.h>
#include "b200_common.cuh"
#include "b200_proj.cuh"

static thread_local char g_err[512] = "";

void b200_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Keep freed stream-ordered allocations in the default memory pool instead of returning them to the driver at every
// synchronisation: temporaries (projection tables, search structures) are re-allocated for every scene.
void b200_pool_keep_memory() {
    static bool done[64] = {false};  // per device: a process may drive more than one GPU
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[dev] = true;
}

extern "C" const char *b200_version(void) { return "satpy_b200 0.1.0 (sm_100a)"; }

extern "C" const char *b200_last_error(void) { return g_err; }

extern "C" int b200_device_count(int *count) {
    B200_CHECK_ARG(count != nullptr, "count is NULL");
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        b200_set_error("b200_device_count: %s", cudaGetErrorString(err));
        return B200_ECUDA;
    }
    *count = n;
    return B200_OK;
}

int b200_check_geo(const b200_geo_t *geo, const char *who) {
    b200_pool_keep_memory();  // every entry point that allocates temporaries validates a geometry first
    if (geo == nullptr) {
        b200_set_error("%s: geometry is NULL", who);
        return B200_EINVAL;
    }
    if (geo->width <= 0 || geo->height <= 0) {
        b200_set_error("%s: empty geometry (%lld x %lld)", who, (long long)geo->height, (long long)geo->width);
        return B200_EINVAL;
    }
    if (geo->row0 < 0 || geo->nrows < 0 || geo->row0 + geo->nrows > geo->height) {
        b200_set_error("%s: row window [%lld, +%lld) outside the %lld-row grid", who, (long long)geo->row0,
                       (long long)geo->nrows, (long long)geo->height);
        return B200_EINVAL;
    }
    if (geo->is_swath) {
        if (geo->lons == nullptr || geo->lats == nullptr) {
            b200_set_error("%s: swath geometry without lon/lat arrays", who);
            return B200_EINVAL;
        }
    } else {
        int k = geo->area.proj.kind;
        if (k < B200_PROJ_LONGLAT || k > B200_PROJ_STERE) {
            b200_set_error("%s: unknown projection kind %d", who, k);
            return B200_EINVAL;
        }
        if (!(geo->area.pixel_size_x > 0.0) || !(geo->area.pixel_size_y > 0.0)) {
            b200_set_error("%s: pixel sizes must be positive", who);
            return B200_EINVAL;
        }
    }
    return B200_OK;
}

struct GeoArg {
    b200_geo_t g;
};

__global__ void __launch_bounds__(256)
area_lonlats_kernel(double *__restrict__ lons, double *__restrict__ lats, const __grid_constant__ GeoArg A) {
    const int64_t W = A.g.width;
    const int64_t n = A.g.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(A.g, A.g.row0 + r, i - r * W, lon, lat);
        lons[i] = lon;
        lats[i] = lat;
    }
}

extern "C" int b200_area_lonlats(const b200_geo_t *geo, double *lons, double *lats, void *stream) {
    int rc = b200_check_geo(geo, "b200_area_lonlats");
    if (rc) return rc;
    int64_t n = geo->nrows * geo->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(lons != nullptr && lats != nullptr, "NULL output");
    GeoArg A;
    A.g = *geo;
    area_lonlats_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(lons, lats, A);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Assembling a row-tiled image: the part of a row block the source can never reach is known from
// the geometry, so every rank writes that fill value itself and only the column span that can hold data crosses NVLink.
__global__ void __launch_bounds__(256)
fill_f32_2d_kernel(float *__restrict__ dst, int64_t pitch, int64_t width, int64_t height, float value) {
    // one warp-iteration = 128 consecutive floats of one row (512 B per warp-level store where 16-byte aligned)
    const int64_t chunks_per_row = (width + 127) >> 7;
    const int64_t n = chunks_per_row * height;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = warp; it < n; it += nwarps) {
        const int64_t r = it / chunks_per_row, c0 = (it - r * chunks_per_row) << 7;
        float *p = dst + r * pitch + c0 + lane * 4;
        const int64_t left = width - (c0 + lane * 4);
        if (left >= 4 && (((uintptr_t)p) & 15) == 0) st_stream_f4(p, value, value, value, value);
        else for (int k = 0; k < 4 && k < left; ++k) st_stream_f1(p + k, value);
    }
}

extern "C" int b200_fill_f32_2d(float *dst, int64_t pitch, int64_t width, int64_t height, float value, void *stream) {
    B200_CHECK_ARG(width >= 0 && height >= 0 && pitch >= width, "bad shape");
    if (width == 0 || height == 0) return B200_OK;
    B200_CHECK_ARG(dst != nullptr, "NULL pointer");
    const int64_t warps = ((width + 127) >> 7) * height;
    fill_f32_2d_kernel<<<b200_grid_for(warps * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(dst, pitch, width, height, value);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_copy_2d(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                            int64_t width_bytes, int64_t height, void *stream) {
    B200_CHECK_ARG(width_bytes >= 0 && height >= 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
                   "bad shape");
    if (width_bytes == 0 || height == 0) return B200_OK;
    B200_CHECK_ARG(dst != nullptr && src != nullptr, "NULL pointer");
    // copy engines; dst may be a peer GPU's memory mapped into this process (unified addressing)
    B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                                (size_t)height, cudaMemcpyDefault, (cudaStream_t)stream));
    return B200_OK;
}

// up to 64 rectangles of one image in ONE launch (the fill value outside the reachable spans of all the row blocks a
// rank does not own: 56 rectangles per step at 8 ranks x 4 blocks -- one launch instead of 56)
struct FillRects {
    int64_t row0[64], nrows[64], col0[64], ncols[64];
    int64_t first_chunk[65];   // prefix sum of (chunks per row) * rows
    int32_t n;
};

__global__ void __launch_bounds__(256)
fill_f32_rects_kernel(float *__restrict__ img, int64_t pitch, const __grid_constant__ FillRects R, float value) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = R.first_chunk[R.n];
    int k = 0;
    for (int64_t it = warp; it < total; it += nwarps) {
        while (it >= R.first_chunk[k + 1]) ++k;    // `it` only grows: the rectangle index never goes back
        const int64_t local = it - R.first_chunk[k];
        const int64_t cpr = (R.ncols[k] + 127) >> 7;
        const int64_t r = local / cpr, c0 = (local - r * cpr) << 7;
        float *p = img + (R.row0[k] + r) * pitch + R.col0[k] + c0 + lane * 4;
        const int64_t left = R.ncols[k] - (c0 + lane * 4);
        if (left >= 4 && (((uintptr_t)p) & 15) == 0) st_stream_f4(p, value, value, value, value);
        else for (int j = 0; j < 4 && j < left; ++j) st_stream_f1(p + j, value);
    }
}

extern "C" int b200_fill_f32_rects(float *img, int64_t pitch, const int64_t *rects /* n x {row0, nrows, col0, ncols} */,
                                   int32_t n, float value, void *stream) {
    B200_CHECK_ARG(n >= 0 && (n == 0 || rects != nullptr), "bad rectangle list");
    B200_CHECK_ARG(n == 0 || img != nullptr, "NULL pointer");
    for (int32_t first = 0; first < n; first += 64) {
        FillRects R;
        memset(&R, 0, sizeof(R));
        const int32_t m = n - first < 64 ? n - first : 64;
        R.n = m;
        for (int32_t k = 0; k < m; ++k) {
            const int64_t *q = rects + 4 * (int64_t)(first + k);
            B200_CHECK_ARG(q[0] >= 0 && q[1] >= 0 && q[2] >= 0 && q[3] >= 0 && q[2] + q[3] <= pitch, "rectangle outside the image");
            R.row0[k] = q[0], R.nrows[k] = q[1], R.col0[k] = q[2], R.ncols[k] = q[3];
            R.first_chunk[k + 1] = R.first_chunk[k] + ((q[3] + 127) >> 7) * q[1];
        }
        const int64_t warps = R.first_chunk[m];
        if (warps == 0) continue;
        fill_f32_rects_kernel<<<b200_grid_for(warps * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(img, pitch, R, value);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// HOST helper (no device work): write `value` into a rectangle of a page-locked float32 image with non-temporal
// stores.  Results that go back to the host only carry the reachable column spans over PCIe (satpy_b200/_xfer.py); the
// rest of the image is fill value, written by host threads meanwhile.  Ordinary stores read every cache line before
// overwriting it (read-for-ownership), which halves the useful memory bandwidth; streaming stores do not.
#include <emmintrin.h>
extern "C" int b200_host_fill_f32(float *dst, int64_t pitch, int64_t width, int64_t height, float value) {
    B200_CHECK_ARG(width >= 0 && height >= 0 && pitch >= width, "bad shape");
    if (width == 0 || height == 0) return B200_OK;
    B200_CHECK_ARG(dst != nullptr, "NULL pointer");
    const __m128 v = _mm_set1_ps(value);
    for (int64_t r = 0; r < height; ++r) {
        float *p = dst + r * pitch;
        int64_t c = 0;
        while (c < width && (((uintptr_t)(p + c)) & 15)) p[c++] = value;     // head up to a 16-byte boundary
        for (; c + 16 <= width; c += 16) {
            _mm_stream_ps(p + c, v);
            _mm_stream_ps(p + c + 4, v);
            _mm_stream_ps(p + c + 8, v);
            _mm_stream_ps(p + c + 12, v);
        }
        for (; c < width; ++c) p[c] = value;
    }
    _mm_sfence();
    return B200_OK;
}
