// Nearest-neighbour resampling on sm_100a: bucket-grid search + gather.
//
// Replaces
//   pyresample.kd_tree.XArrayResamplerNN.get_neighbour_info   (pykdtree build + query; reached from
//       KDTreeResampler.precompute, satpy/resample/kdtree.py:60-95)
//   XArrayResamplerNN.get_sample_from_neighbour_info          (KDTreeResampler.compute,
//       satpy/resample/kdtree.py:184-190)
// with the same semantics: spherical cartesian coordinates with
// R = 6370997 m in float64, only VALID source points take part, a neighbour counts when its chord
// distance is < radius_of_influence, misses are -1, and index_array indexes the compressed
// valid-input vector.
//
// The kd-tree is replaced by a uniform bucket grid on the sphere: latitude bands of height >= the
// search cap angle, each band cut into ~square longitude cells; cells of a band are contiguous, and
// the valid source points are counting-sorted by cell into one array of 32-byte records
// {x, y, z, raw index}.  A query turns the spherical cap of the search radius into (bands) x (one or
// two contiguous cell ranges per band), i.e. a handful of contiguous record ranges that it scans
// with 128-bit loads.  Ties (exactly equal float64 distance) go to the lowest source index, so the
// result does not depend on the (atomic, unordered) placement inside a cell.
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include "b200_common.cuh"
#include "b200_proj.cuh"

int b200_check_geo(const b200_geo_t *geo, const char *who);

#define B200_R_EARTH 6370997.0
#define B200_TWO_PI 6.283185307179586476925287
#define B200_NN_MAX_CELLS (160LL * 1000 * 1000)

struct b200_nn_grid {
    int64_t n_src, n_valid, ncells;
    int64_t src_width, src_height;
    int32_t nbands;
    double radius, dphi;
    int32_t *band_off;    // [nbands + 1] first cell of each band
    int32_t *cell_start;  // [ncells + 1] first record of each cell
    double4 *pts;         // [n_valid] {x, y, z, raw source index (bit pattern)} sorted by cell
    int32_t *prefix;      // [n_src] exclusive prefix sum of valid_input: raw -> compressed index
    int64_t nbytes;
    int device;
};

struct GridDev {
    const int32_t *band_off;
    const int32_t *cell_start;
    const double4 *pts;
    const int32_t *prefix;
    double inv_dphi;
    int32_t nbands;
};

struct SrcArg {
    b200_geo_t g;
};

// lon/lat (degrees) -> spherical cartesian, pyresample's R (kd_tree / geometry module constant)
__device__ __forceinline__ void b200_lonlat2xyz(double lon, double lat, double &x, double &y, double &z,
                                                double &lam, double &phi) {
    lam = lon * B200_DEG2RAD;
    phi = lat * B200_DEG2RAD;
    double sl, cl, sp, cp;
    sincos(lam, &sl, &cl);
    sincos(phi, &sp, &cp);
    x = B200_R_EARTH * cp * cl;
    y = B200_R_EARTH * cp * sl;
    z = B200_R_EARTH * sp;
}

__device__ __forceinline__ int b200_band_of(double phi, double inv_dphi, int nbands) {
    int b = (int)floor((phi + B200_HALFPI) * inv_dphi);
    return min(nbands - 1, max(0, b));
}

__device__ __forceinline__ int b200_cell_of(const int32_t *__restrict__ band_off, double lam, double phi,
                                            double inv_dphi, int nbands) {
    int b = b200_band_of(phi, inv_dphi, nbands);
    int off = __ldg(band_off + b);
    int nb = __ldg(band_off + b + 1) - off;
    int c = (int)floor((lam + B200_PI) * ((double)nb / B200_TWO_PI));
    c = min(nb - 1, max(0, c));
    return off + c;
}

// ---------------------------------------------------------------------------------------- build
// pass 1: validity + per-cell histogram
__global__ void __launch_bounds__(256)
nn_count_kernel(const __grid_constant__ SrcArg S, const uint8_t *__restrict__ mask, uint8_t *__restrict__ valid_input,
                int32_t *__restrict__ counts, const int32_t *__restrict__ band_off, double inv_dphi, int nbands) {
    const int64_t W = S.g.width;
    const int64_t n = S.g.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(S.g, S.g.row0 + r, i - r * W, lon, lat);
        bool ok = b200_lonlat_valid(lon, lat) && !(mask != nullptr && mask[i]);
        valid_input[i] = ok ? 1 : 0;
        if (ok) {
            int cell = b200_cell_of(band_off, lon * B200_DEG2RAD, lat * B200_DEG2RAD, inv_dphi, nbands);
            atomicAdd(counts + cell, 1);
        }
    }
}

// pass 2: scatter the records into their cells
__global__ void __launch_bounds__(256)
nn_scatter_kernel(const __grid_constant__ SrcArg S, const uint8_t *__restrict__ valid_input,
                  int32_t *__restrict__ cursor, double4 *__restrict__ pts, const int32_t *__restrict__ band_off,
                  double inv_dphi, int nbands) {
    const int64_t W = S.g.width;
    const int64_t n = S.g.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!valid_input[i]) continue;
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(S.g, S.g.row0 + r, i - r * W, lon, lat);
        double x, y, z, lam, phi;
        b200_lonlat2xyz(lon, lat, x, y, z, lam, phi);
        int cell = b200_cell_of(band_off, lam, phi, inv_dphi, nbands);
        int slot = atomicAdd(cursor + cell, 1);
        pts[slot] = make_double4(x, y, z, __longlong_as_double((long long)i));
    }
}

struct U8ToI32 {
    __host__ __device__ __forceinline__ int32_t operator()(const uint8_t &v) const { return (int32_t)v; }
};

static double cap_angle(double radius) {
    double h = radius / (2.0 * B200_R_EARTH);
    if (h >= 1.0) return B200_PI;
    return 2.0 * asin(h) * (1.0 + 1e-9) + 1e-12;
}

extern "C" int b200_nn_grid_destroy(b200_nn_grid_t *grid) {
    if (grid == nullptr) return B200_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(grid->device);
    cudaFree(grid->band_off);
    cudaFree(grid->cell_start);
    cudaFree(grid->pts);
    cudaFree(grid->prefix);
    cudaSetDevice(prev);
    delete grid;
    return B200_OK;
}

extern "C" int b200_nn_grid_nbytes(const b200_nn_grid_t *grid, int64_t *nbytes) {
    B200_CHECK_ARG(grid != nullptr && nbytes != nullptr, "NULL argument");
    *nbytes = grid->nbytes;
    return B200_OK;
}

#define NN_CUDA_OR_FREE(call)                                                                       \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess) {                                                                 \
            b200_set_error("b200_nn_grid_create: %s failed: %s", #call, cudaGetErrorString(err__)); \
            cudaFree(counts);                                                                       \
            cudaFree(tmp);                                                                          \
            b200_nn_grid_destroy(g);                                                                \
            return err__ == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;                   \
        }                                                                                           \
    } while (0)

extern "C" int b200_nn_grid_create(const b200_geo_t *src, const uint8_t *mask, double radius, uint8_t *valid_input,
                                   int64_t *n_valid, b200_nn_grid_t **grid, void *stream) {
    int rc = b200_check_geo(src, "b200_nn_grid_create");
    if (rc) return rc;
    B200_CHECK_ARG(grid != nullptr && n_valid != nullptr, "NULL output argument");
    B200_CHECK_ARG(valid_input != nullptr, "valid_input is NULL");
    B200_CHECK_ARG(radius > 0.0 && isfinite(radius), "radius_of_influence must be positive and finite");
    // the grid covers the source row window [row0, row0 + nrows); indices are relative to its first pixel
    const int64_t n_src = src->nrows * src->width;
    B200_CHECK_ARG(n_src > 0, "empty source row window");
    B200_CHECK_ARG(n_src < (1LL << 31), "more than 2^31 source pixels");
    cudaStream_t s = (cudaStream_t)stream;

    // ---- host: band layout.  Band height >= cap angle; widened if the table would get too large.
    double theta = cap_angle(radius);
    double dphi = theta;
    double min_dphi = sqrt(4.0 * B200_PI / (double)B200_NN_MAX_CELLS);
    if (dphi < min_dphi) dphi = min_dphi;
    if (dphi > B200_PI) dphi = B200_PI;
    int nbands = (int)ceil(B200_PI / dphi);
    if (nbands < 1) nbands = 1;
    int32_t *h_off = new (std::nothrow) int32_t[nbands + 1];
    if (!h_off) {
        b200_set_error("b200_nn_grid_create: out of host memory");
        return B200_ENOMEM;
    }
    int64_t ncells = 0;
    for (int b = 0; b < nbands; ++b) {
        double lo = -B200_HALFPI + b * dphi, hi = lo + dphi;
        if (hi > B200_HALFPI) hi = B200_HALFPI;
        double cmax = (lo <= 0.0 && hi >= 0.0) ? 1.0 : fmax(cos(lo), cos(hi));
        int64_t nb = (int64_t)floor(B200_TWO_PI * cmax / dphi);
        if (nb < 1) nb = 1;
        h_off[b] = (int32_t)ncells;
        ncells += nb;
    }
    h_off[nbands] = (int32_t)ncells;

    b200_nn_grid *g = new (std::nothrow) b200_nn_grid();
    int32_t *counts = nullptr;
    void *tmp = nullptr;
    if (!g) {
        delete[] h_off;
        b200_set_error("b200_nn_grid_create: out of host memory");
        return B200_ENOMEM;
    }
    g->n_src = n_src;
    g->ncells = ncells;
    g->nbands = nbands;
    g->radius = radius;
    g->dphi = dphi;
    g->src_width = src->width;
    g->src_height = src->nrows;
    g->band_off = nullptr;
    g->cell_start = nullptr;
    g->pts = nullptr;
    g->prefix = nullptr;
    cudaGetDevice(&g->device);
    const double inv_dphi = 1.0 / dphi;

    cudaError_t e0 = cudaMalloc(&g->band_off, (size_t)(nbands + 1) * sizeof(int32_t));
    if (e0 == cudaSuccess)
        e0 = cudaMemcpyAsync(g->band_off, h_off, (size_t)(nbands + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(s);  // h_off is pageable: finish before freeing it
    delete[] h_off;
    NN_CUDA_OR_FREE(e0);
    NN_CUDA_OR_FREE(cudaMalloc(&g->cell_start, (size_t)(ncells + 1) * sizeof(int32_t)));
    NN_CUDA_OR_FREE(cudaMalloc(&g->prefix, (size_t)n_src * sizeof(int32_t)));
    NN_CUDA_OR_FREE(cudaMalloc(&counts, (size_t)(ncells + 1) * sizeof(int32_t)));
    NN_CUDA_OR_FREE(cudaMemsetAsync(counts, 0, (size_t)(ncells + 1) * sizeof(int32_t), s));

    SrcArg S;
    S.g = *src;
    nn_count_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(S, mask, valid_input, counts, g->band_off,
                                                                inv_dphi, nbands);
    NN_CUDA_OR_FREE(cudaGetLastError());

    // exclusive scans: cell histogram -> cell_start, valid mask -> compressed index
    size_t tmp_a = 0, tmp_b = 0;
    thrust::transform_iterator<U8ToI32, const uint8_t *> vit(valid_input, U8ToI32());
    NN_CUDA_OR_FREE(cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, counts, g->cell_start, (int)(ncells + 1), s));
    NN_CUDA_OR_FREE(cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, vit, g->prefix, (int)n_src, s));
    size_t tmp_bytes = tmp_a > tmp_b ? tmp_a : tmp_b;
    NN_CUDA_OR_FREE(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    NN_CUDA_OR_FREE(cub::DeviceScan::ExclusiveSum(tmp, tmp_a, counts, g->cell_start, (int)(ncells + 1), s));
    NN_CUDA_OR_FREE(cub::DeviceScan::ExclusiveSum(tmp, tmp_b, vit, g->prefix, (int)n_src, s));
    int32_t total = 0;
    NN_CUDA_OR_FREE(cudaMemcpyAsync(&total, g->cell_start + ncells, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NN_CUDA_OR_FREE(cudaStreamSynchronize(s));
    g->n_valid = total;

    // counts becomes the per-cell write cursor
    NN_CUDA_OR_FREE(cudaMemcpyAsync(counts, g->cell_start, (size_t)ncells * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    NN_CUDA_OR_FREE(cudaMalloc(&g->pts, (size_t)(total > 0 ? total : 1) * sizeof(double4)));
    if (total > 0) {
        nn_scatter_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(S, valid_input, counts, g->pts, g->band_off,
                                                                      inv_dphi, nbands);
        NN_CUDA_OR_FREE(cudaGetLastError());
    }
    NN_CUDA_OR_FREE(cudaStreamSynchronize(s));
    cudaFree(counts);
    cudaFree(tmp);
    g->nbytes = (int64_t)(nbands + 1) * 4 + (ncells + 1) * 4 + n_src * 4 + (int64_t)(total > 0 ? total : 1) * 32;
    *n_valid = total;
    *grid = g;
    return B200_OK;
}

// ---------------------------------------------------------------------------------------- query
struct Best {
    double d2;
    long long idx;
};

__device__ __forceinline__ void nn_scan_range(const double4 *__restrict__ pts, int s, int e, double tx, double ty,
                                              double tz, Best &best) {
    for (int j = s; j < e; ++j) {
        const double2 *p = reinterpret_cast<const double2 *>(pts + j);
        double2 a = __ldg(p);
        double2 b = __ldg(p + 1);
        double dx = a.x - tx, dy = a.y - ty, dz = b.x - tz;
        double d2 = dx * dx + dy * dy + dz * dz;
        long long idx = __double_as_longlong(b.y);
        if (d2 < best.d2 || (d2 == best.d2 && idx < best.idx)) {
            best.d2 = d2;
            best.idx = idx;
        }
    }
}

// nearest valid source point of one target lon/lat (degrees); -1 when none within the radius
__device__ __forceinline__ long long nn_query_point(const GridDev &G, double lon, double lat, double theta,
                                                    double sin_theta, double r2) {
    double tx, ty, tz, lam, phi;
    b200_lonlat2xyz(lon, lat, tx, ty, tz, lam, phi);
    Best best;
    best.d2 = r2;  // strict "<": a neighbour exactly at the radius is not accepted
    best.idx = -1;
    int b_lo = b200_band_of(phi - theta, G.inv_dphi, G.nbands);
    int b_hi = b200_band_of(phi + theta, G.inv_dphi, G.nbands);
    bool full = (fabs(phi) + theta >= B200_HALFPI - 1e-12);
    double dlam = B200_PI;
    if (!full) {
        double q = sin_theta / cos(phi);
        dlam = (q >= 1.0) ? B200_PI : asin(q) * (1.0 + 1e-12) + 1e-12;
        if (q >= 1.0) full = true;
    }
    for (int b = b_lo; b <= b_hi; ++b) {
        int off = __ldg(G.band_off + b);
        int nb = __ldg(G.band_off + b + 1) - off;
        double scale = (double)nb / B200_TWO_PI;
        if (full || 2.0 * dlam * scale + 2.0 >= (double)nb) {
            nn_scan_range(G.pts, __ldg(G.cell_start + off), __ldg(G.cell_start + off + nb), tx, ty, tz, best);
            continue;
        }
        int c0 = (int)floor((lam - dlam + B200_PI) * scale);
        int c1 = (int)floor((lam + dlam + B200_PI) * scale);
        if (c0 < 0) {
            nn_scan_range(G.pts, __ldg(G.cell_start + off + c0 + nb), __ldg(G.cell_start + off + nb), tx, ty, tz, best);
            c0 = 0;
        }
        if (c1 >= nb) {
            nn_scan_range(G.pts, __ldg(G.cell_start + off), __ldg(G.cell_start + off + c1 - nb + 1), tx, ty, tz, best);
            c1 = nb - 1;
        }
        nn_scan_range(G.pts, __ldg(G.cell_start + off + c0), __ldg(G.cell_start + off + c1 + 1), tx, ty, tz, best);
    }
    return best.idx;
}

struct QueryArg {
    b200_geo_t dst;
    GridDev G;
    double theta, sin_theta, r2;
};

__global__ void __launch_bounds__(256)
nn_query_kernel(const __grid_constant__ QueryArg Q, int32_t *__restrict__ index_array,
                int32_t *__restrict__ gather_index, uint8_t *__restrict__ valid_output) {
    const int64_t W = Q.dst.width;
    const int64_t n = Q.dst.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(Q.dst, Q.dst.row0 + r, i - r * W, lon, lat);
        bool ok = b200_lonlat_valid(lon, lat);
        long long raw = ok ? nn_query_point(Q.G, lon, lat, Q.theta, Q.sin_theta, Q.r2) : -1;
        if (valid_output) valid_output[i] = ok ? 1 : 0;
        if (gather_index) gather_index[i] = (int32_t)raw;
        if (index_array) index_array[i] = raw < 0 ? -1 : __ldg(Q.G.prefix + raw);
    }
}

__global__ void __launch_bounds__(256)
nn_query_gather_f32_kernel(const __grid_constant__ QueryArg Q, const float *__restrict__ data,
                           float *__restrict__ out, int nbands, int64_t src_band_stride, int64_t dst_band_stride,
                           float fill) {
    const int64_t W = Q.dst.width;
    const int64_t n = Q.dst.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(Q.dst, Q.dst.row0 + r, i - r * W, lon, lat);
        long long raw = b200_lonlat_valid(lon, lat) ? nn_query_point(Q.G, lon, lat, Q.theta, Q.sin_theta, Q.r2) : -1;
        for (int b = 0; b < nbands; ++b) {
            float v = raw < 0 ? fill : __ldg(data + b * src_band_stride + raw);
            st_stream_f1(out + b * dst_band_stride + i, v);
        }
    }
}

static int fill_query(QueryArg &Q, const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius, const char *who) {
    if (grid == nullptr) {
        b200_set_error("%s: grid is NULL", who);
        return B200_EINVAL;
    }
    int rc = b200_check_geo(dst, who);
    if (rc) return rc;
    if (!(radius > 0.0) || !isfinite(radius)) {
        b200_set_error("%s: radius_of_influence must be positive and finite", who);
        return B200_EINVAL;
    }
    Q.dst = *dst;
    Q.G.band_off = grid->band_off;
    Q.G.cell_start = grid->cell_start;
    Q.G.pts = grid->pts;
    Q.G.prefix = grid->prefix;
    Q.G.inv_dphi = 1.0 / grid->dphi;
    Q.G.nbands = grid->nbands;
    Q.theta = cap_angle(radius);
    Q.sin_theta = Q.theta >= B200_HALFPI ? 1.0 : sin(Q.theta);
    Q.r2 = radius * radius;
    return B200_OK;
}

extern "C" int b200_nn_query(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius, int32_t *index_array,
                             int32_t *gather_index, uint8_t *valid_output, void *stream) {
    QueryArg Q;
    int rc = fill_query(Q, grid, dst, radius, "b200_nn_query");
    if (rc) return rc;
    int64_t n = dst->nrows * dst->width;
    if (n == 0) return B200_OK;
    nn_query_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(Q, index_array, gather_index,
                                                                               valid_output);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_nn_query_gather_f32(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius,
                                        const float *data, float *out, int32_t nbands, int64_t src_band_stride,
                                        int64_t dst_band_stride, float fill, void *stream) {
    QueryArg Q;
    int rc = fill_query(Q, grid, dst, radius, "b200_nn_query_gather_f32");
    if (rc) return rc;
    B200_CHECK_ARG(nbands >= 1, "nbands must be >= 1");
    int64_t n = dst->nrows * dst->width;
    if (n == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
    B200_CHECK_ARG(nbands == 1 || (src_band_stride >= grid->n_src && dst_band_stride >= n), "band stride too small");
    nn_query_gather_f32_kernel<<<b200_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        Q, data, out, nbands, src_band_stride, dst_band_stride, fill);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// --------------------------------------------------------------------------------------- gather
// get_sample_from_neighbour_info: out = data.ravel()[valid_input][index]; fill where index == -1.
// Works on RAW source indices so the boolean compress of the reference never materialises.
template <typename T>
__global__ void __launch_bounds__(256)
nn_gather_kernel(const T *__restrict__ data, T *__restrict__ out, const int32_t *__restrict__ gidx, int64_t n_out,
                 int nbands, int64_t src_band_stride, int64_t dst_band_stride, T fill) {
    // four consecutive outputs per thread: one 128-bit index load, four independent source loads in
    // flight, and (for 4-byte elements) one 128-bit store
    const int64_t n4 = n_out >> 2;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        int4 ix = ld_stream_v4(gidx + 4 * q);
        for (int b = 0; b < nbands; ++b) {
            const T *src = data + b * src_band_stride;
            T v0 = ix.x < 0 ? fill : __ldg(src + ix.x);
            T v1 = ix.y < 0 ? fill : __ldg(src + ix.y);
            T v2 = ix.z < 0 ? fill : __ldg(src + ix.z);
            T v3 = ix.w < 0 ? fill : __ldg(src + ix.w);
            T *dst = out + b * dst_band_stride + 4 * q;
            if (sizeof(T) == 4 && ((dst_band_stride & 3) == 0 || nbands == 1)) {
                int4 w;
                w.x = *reinterpret_cast<int *>(&v0);
                w.y = *reinterpret_cast<int *>(&v1);
                w.z = *reinterpret_cast<int *>(&v2);
                w.w = *reinterpret_cast<int *>(&v3);
                st_stream_v4(dst, w);
            } else {
                dst[0] = v0;
                dst[1] = v1;
                dst[2] = v2;
                dst[3] = v3;
            }
        }
    }
    // tail (< 4 outputs)
    if (blockIdx.x == 0 && threadIdx.x < (n_out & 3)) {
        int64_t i = (n4 << 2) + threadIdx.x;
        int32_t ix = gidx[i];
        for (int b = 0; b < nbands; ++b)
            out[b * dst_band_stride + i] = ix < 0 ? fill : data[b * src_band_stride + ix];
    }
}

template <typename T>
static int launch_gather(const void *data, void *out, const int32_t *gidx, int64_t n_out, int nbands,
                         int64_t src_band_stride, int64_t dst_band_stride, const void *fill, cudaStream_t s) {
    T f;
    memcpy(&f, fill, sizeof(T));
    int64_t work = (n_out + 3) / 4;
    nn_gather_kernel<T><<<b200_grid_for(work, 256, 8), 256, 0, s>>>((const T *)data, (T *)out, gidx, n_out, nbands,
                                                                   src_band_stride, dst_band_stride, f);
    return B200_OK;
}

extern "C" int b200_nn_gather(const void *data, void *out, const int32_t *gather_index, int64_t n_out, int32_t nbands,
                              int64_t src_band_stride, int64_t dst_band_stride, int32_t elem_size, const void *fill,
                              void *stream) {
    B200_CHECK_ARG(n_out >= 0 && nbands >= 1, "bad shape");
    B200_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8,
                   "elem_size must be 1, 2, 4 or 8");
    B200_CHECK_ARG(fill != nullptr, "fill is NULL");
    if (n_out == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr && gather_index != nullptr, "NULL data pointer");
    B200_CHECK_ARG(((uintptr_t)gather_index & 15) == 0, "gather_index must be 16-byte aligned");
    B200_CHECK_ARG(elem_size != 4 || ((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    B200_CHECK_ARG(nbands == 1 || dst_band_stride >= n_out, "band stride smaller than one image");
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_size) {
    case 1: launch_gather<uint8_t>(data, out, gather_index, n_out, nbands, src_band_stride, dst_band_stride, fill, s); break;
    case 2: launch_gather<uint16_t>(data, out, gather_index, n_out, nbands, src_band_stride, dst_band_stride, fill, s); break;
    case 4: launch_gather<uint32_t>(data, out, gather_index, n_out, nbands, src_band_stride, dst_band_stride, fill, s); break;
    default: launch_gather<uint64_t>(data, out, gather_index, n_out, nbands, src_band_stride, dst_band_stride, fill, s); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ------------------------------------------------------------- cached index arrays (compressed -> raw)
__global__ void __launch_bounds__(256)
nn_raw_of_kernel(const uint8_t *__restrict__ valid_input, const int32_t *__restrict__ prefix, int64_t n_src,
                 int32_t *__restrict__ raw_of) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += stride)
        if (valid_input[i]) raw_of[prefix[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
nn_map_index_kernel(const int32_t *__restrict__ index_array, const int32_t *__restrict__ raw_of, int64_t n_valid,
                    int32_t *__restrict__ gather_index, int64_t n_out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        int32_t c = index_array[i];
        gather_index[i] = (c < 0 || c >= n_valid) ? -1 : raw_of[c];
    }
}

extern "C" int b200_nn_compressed_to_raw(const uint8_t *valid_input, int64_t n_src, const int32_t *index_array,
                                         int32_t *gather_index, int64_t n_out, void *stream) {
    B200_CHECK_ARG(n_src >= 0 && n_out >= 0, "negative size");
    B200_CHECK_ARG(n_src < (1LL << 31), "more than 2^31 source pixels");
    if (n_out == 0) return B200_OK;
    B200_CHECK_ARG(index_array != nullptr && gather_index != nullptr, "NULL index pointer");
    B200_CHECK_ARG(n_src == 0 || valid_input != nullptr, "valid_input is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *prefix = nullptr, *raw_of = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = B200_OK;
    cudaError_t err = cudaSuccess;
    int64_t n_alloc = n_src > 0 ? n_src : 1;
    thrust::transform_iterator<U8ToI32, const uint8_t *> vit(valid_input, U8ToI32());
    if ((err = cudaMalloc(&prefix, (size_t)(n_alloc + 1) * 4)) != cudaSuccess) goto fail;
    if ((err = cudaMalloc(&raw_of, (size_t)n_alloc * 4)) != cudaSuccess) goto fail;
    // 0xff.. = -1: compressed indices >= n_valid (the kd-tree miss convention) map to "no neighbour"
    if ((err = cudaMemsetAsync(raw_of, 0xff, (size_t)n_alloc * 4, s)) != cudaSuccess) goto fail;
    if (n_src > 0) {
        if ((err = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, vit, prefix, (int)n_src, s)) != cudaSuccess) goto fail;
        if ((err = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) goto fail;
        if ((err = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, vit, prefix, (int)n_src, s)) != cudaSuccess) goto fail;
        nn_raw_of_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(valid_input, prefix, n_src, raw_of);
    }
    // n_src bounds the compressed index; slots in [n_valid, n_src) still hold -1
    nn_map_index_kernel<<<b200_grid_for(n_out, 256, 8), 256, 0, s>>>(index_array, raw_of, n_src, gather_index, n_out);
    if ((err = cudaGetLastError()) != cudaSuccess) goto fail;
    if ((err = cudaStreamSynchronize(s)) != cudaSuccess) goto fail;
    goto done;
fail:
    b200_set_error("b200_nn_compressed_to_raw: %s", cudaGetErrorString(err));
    rc = err == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
done:
    cudaFree(prefix);
    cudaFree(raw_of);
    cudaFree(tmp);
    return rc;
}
