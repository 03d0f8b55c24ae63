// Nearest-neighbour resampling on sm_100a: search-structure life cycle, the bucket-grid engine, the gather.
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
// Two exact engines stand in for the kd-tree; both break exact float64 distance ties towards the lowest
// source index, so they return identical indices:
//   * bucket grid (this file; any source, swath or area): latitude bands over the source's latitude range,
//     each band cut into ~square longitude cells sized to hold a few source points (from a sampled estimate
//     of the point density); cells of a band are contiguous, and the valid source points are counting-sorted
//     by cell into one array of 32-byte records {x, y, z, raw index}.  A query turns a spherical cap into
//     (bands) x (one or two contiguous cell ranges per band), i.e. a handful of contiguous record ranges
//     that it scans with 128-bit loads -- first the cap of ONE cell size, which already holds the answer
//     wherever the source is dense, and only then the cap of the full radius of influence.
//   * fixed grid (nn_geos.cu; geostationary area sources): the source grid itself is the index.
// Targets whose lon depends on the column only and lat on the row only (longlat / eqc / merc) get
// per-column and per-row tables so that no transcendental is evaluated per pixel.
#include <climits>
#include <stdlib.h>
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include "nn_common.cuh"
#include "b200_bulk.cuh"
#include "b200_geos_fast.cuh"

#define B200_NN_MAX_CELLS (160LL * 1000 * 1000)
#define B200_NN_CELL_POINTS 4.0  /* source points per cell the layout aims for */

struct GridDev {
    const int32_t *band_off;
    const int32_t *cell_start;
    const double4 *pts;
    const int32_t *prefix;
    double inv_dphi, phi0;
    int32_t nbands;
};

struct SrcArg {
    b200_geo_t g;
};

// latitudes below / above the banded range belong to the first / last band (monotone, so a cap [phi - t, phi + t]
// always maps to a band range that contains every point inside it)
__device__ __forceinline__ int b200_band_of(double phi, double phi0, double inv_dphi, int nbands) {
    double b = floor((phi - phi0) * inv_dphi);
    return (int)fmin((double)(nbands - 1), fmax(0.0, b));
}

__device__ __forceinline__ int b200_cell_of(const int32_t *__restrict__ band_off, double lam, double phi, double phi0,
                                            double inv_dphi, int nbands) {
    int b = b200_band_of(phi, phi0, inv_dphi, nbands);
    int off = __ldg(band_off + b);
    int nb = __ldg(band_off + b + 1) - off;
    int c = (int)floor((lam + B200_PI) * ((double)nb / B200_TWO_PI));
    c = min(nb - 1, max(0, c));
    return off + c;
}

// ---------------------------------------------------------------------------------------- build
// ONE pass over the source for both engines: validity, the record {x, y, z, raw index} by raw index (x = NaN marks an
// invalid pixel) and, for the bucket engine, the cell of every valid point.
// FAST: geostationary source (table-driven view vectors); a separate instantiation so that the register budget of the
// general projection code does not lower the occupancy of the hot configuration.
template <bool FAST>
__global__ void __launch_bounds__(256)
nn_source_points_kernel(const __grid_constant__ SrcArg S, const uint8_t *__restrict__ mask,
                        uint8_t *__restrict__ valid_input, double4 *__restrict__ rec, float4 *__restrict__ rec32,
                        int32_t *__restrict__ cell_of, const int32_t *__restrict__ band_off, double phi0, double inv_dphi,
                        int nbands, const __grid_constant__ GeosFast F) {
    // 32-bit pixel numbers (n_src < 2^31 is checked when the structure is created): the row comes from a 32-bit
    // division instead of the 64-bit one (~50 instructions of the 236 this kernel spent per pixel)
    const unsigned W = (unsigned)S.g.width;
    const unsigned n = (unsigned)(S.g.nrows * S.g.width);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned r = i / W;
        double x = __longlong_as_double(0x7ff8000000000000LL), y = 0.0, z = 0.0, lam = 0.0, phi = 0.0;
        bool ok;
        if (FAST) {
            // geostationary source: every pixel on the disk has a valid lon/lat; the point comes from the view
            // vector without trigonometric functions (the angles are only needed to pick a bucket cell)
            double sl, cl, sp, cp;
            ok = geos_fast_xyz(F, S.g.row0 + r, i - r * W, B200_R_EARTH, x, y, z, sl, cl, sp, cp) &&
                 !(mask != nullptr && mask[i]);
            if (!ok) x = __longlong_as_double(0x7ff8000000000000LL);
            if (ok && cell_of) {
                lam = atan2(sl, cl);
                phi = atan2(sp, cp);
            }
        } else {
            double lon, lat;
            b200_geo_lonlat(S.g, S.g.row0 + r, i - r * W, lon, lat);
            ok = b200_lonlat_valid(lon, lat) && !(mask != nullptr && mask[i]);
            if (ok) b200_lonlat2xyz(lon, lat, x, y, z, lam, phi);
        }
        valid_input[i] = ok ? 1 : 0;
        if (rec) {
            double2 *p = reinterpret_cast<double2 *>(rec + i);
            p[0] = make_double2(x, y);
            p[1] = make_double2(z, __longlong_as_double((long long)i));
        }
        // float32 copy; w carries what the rounding dropped, per coordinate, as signed 10-bit multiples of 2^-11 m
        // (|residual| <= 0.25 m for |coordinate| < 2^23 m): nn_geos.cu re-ranks near-ties with these 0.5 mm points
        // instead of going back to the float64 records (nn_fixed_refine)
        if (rec32) rec32[i] = ok ? make_float4((float)x, (float)y, (float)z, __int_as_float(b200_pack_residuals(x, y, z)))
                                 : make_float4(__int_as_float(0x7f800000), 0.0f, 0.0f, 0.0f);
        if (cell_of) cell_of[i] = ok ? b200_cell_of(band_off, lam, phi, phi0, inv_dphi, nbands) : -1;
    }
}

__global__ void __launch_bounds__(256)
nn_cell_count_kernel(const int32_t *__restrict__ cell_of, int64_t n, int32_t *__restrict__ counts) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int c = cell_of[i];
        if (c >= 0) atomicAdd(counts + c, 1);
    }
}

// counting sort, second half: move the records into their cells
__global__ void __launch_bounds__(256)
nn_cell_scatter_kernel(const int32_t *__restrict__ cell_of, const double4 *__restrict__ rec, int64_t n,
                       int32_t *__restrict__ cursor, double4 *__restrict__ pts) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int c = cell_of[i];
        if (c < 0) continue;
        int slot = atomicAdd(cursor + c, 1);
        const double2 *q = reinterpret_cast<const double2 *>(rec + i);
        double2 *p = reinterpret_cast<double2 *>(pts + slot);
        p[0] = q[0];
        p[1] = q[1];
    }
}

// Density and latitude range of the source from a regular sample of its pixels (at most NN_STATS_SIDE^2): the mean
// area per pixel (|step to the right| x |step down|) sizes the cells, the latitude range places the bands.  Both only
// tune the layout -- a point outside the sampled range lands in an edge band and is still found.
#define NN_STATS_SIDE 512
struct SrcStats {
    double area_sum;
    unsigned long long pairs;
    long long phi_min_k, phi_max_k;  // fixed point, 1e-12 rad
};

__device__ __forceinline__ bool nn_stats_point(const SrcArg &S, const uint8_t *__restrict__ mask, int64_t r, int64_t c,
                                               double &x, double &y, double &z, double &phi) {
    if (r >= S.g.nrows || c >= S.g.width) return false;
    if (mask != nullptr && mask[r * S.g.width + c]) return false;
    double lon, lat, lam;
    b200_geo_lonlat(S.g, S.g.row0 + r, c, lon, lat);
    if (!b200_lonlat_valid(lon, lat)) return false;
    b200_lonlat2xyz(lon, lat, x, y, z, lam, phi);
    return true;
}

__global__ void __launch_bounds__(256)
nn_source_stats_kernel(const __grid_constant__ SrcArg S, const uint8_t *__restrict__ mask, int rs, int cs,
                       SrcStats *__restrict__ st) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double area = 0.0;
    unsigned long long pairs = 0;
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;
    if (t < rs * cs) {
        const int64_t r = (int64_t)(t / cs) * S.g.nrows / rs, c = (int64_t)(t % cs) * S.g.width / cs;
        double x, y, z, phi, x1, y1, z1, x2, y2, z2, p_;
        if (nn_stats_point(S, mask, r, c, x, y, z, phi)) {
            kmin = kmax = llrint(phi * 1e12);
            if (nn_stats_point(S, mask, r, c + 1, x1, y1, z1, p_) && nn_stats_point(S, mask, r + 1, c, x2, y2, z2, p_)) {
                area = sqrt(b200_nn_d2(x1, y1, z1, x, y, z) * b200_nn_d2(x2, y2, z2, x, y, z));
                pairs = 1;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        area += __shfl_xor_sync(0xffffffffu, area, o);
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
        atomicMin(&st->phi_min_k, kmin);
        atomicMax(&st->phi_max_k, kmax);
        if (pairs) {
            atomicAdd(&st->area_sum, area);
            atomicAdd(&st->pairs, pairs);
        }
    }
}

struct U8ToI32 {
    __host__ __device__ __forceinline__ int32_t operator()(const uint8_t &v) const { return (int32_t)v; }
};

// Device memory of a search structure comes from the stream-ordered pool (b200_pool_keep_memory, lib.cu): after the
// first build, rebuilding for the next scene allocates without a device synchronisation.  All uses of a structure must
// be on the stream it was created on (the caller's current stream).
extern "C" int b200_nn_grid_destroy(b200_nn_grid_t *grid) {
    if (grid == nullptr) return B200_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(grid->device);
    cudaStream_t s = (cudaStream_t)grid->stream;
    if (grid->band_off) cudaFreeAsync(grid->band_off, s);
    if (grid->cell_start) cudaFreeAsync(grid->cell_start, s);
    if (grid->pts) cudaFreeAsync(grid->pts, s);
    if (grid->pts32) cudaFreeAsync(grid->pts32, s);
    if (grid->prefix) cudaFreeAsync(grid->prefix, s);
    b200_geos_fast_release(&grid->fast, s);
    cudaSetDevice(prev);
    delete grid;
    return B200_OK;
}

extern "C" int b200_nn_grid_nbytes(const b200_nn_grid_t *grid, int64_t *nbytes) {
    B200_CHECK_ARG(grid != nullptr && nbytes != nullptr, "NULL argument");
    *nbytes = grid->nbytes;
    return B200_OK;
}

#define NN_TRY(call)                                                                        \
    do {                                                                                    \
        cudaError_t err__ = (call);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            b200_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(err__)); \
            rc = err__ == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;             \
            goto fail;                                                                      \
        }                                                                                   \
    } while (0)

// exclusive prefix sum of the valid mask (raw -> compressed index) and the number of valid points
// lazy: do not read the count back (no host synchronisation); b200_nn_grid_n_valid does it on demand
static int nn_prefix(b200_nn_grid *g, const uint8_t *valid_input, cudaStream_t s, bool lazy) {
    int rc = B200_OK;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int32_t last_prefix = 0;
    uint8_t last_valid = 0;
    thrust::transform_iterator<U8ToI32, const uint8_t *> vit(valid_input, U8ToI32());
    NN_TRY(cudaMallocAsync((void **)&g->prefix, (size_t)g->n_src * sizeof(int32_t), s));
    NN_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, vit, g->prefix, (int)g->n_src, s));
    NN_TRY(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
    NN_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, vit, g->prefix, (int)g->n_src, s));
    g->valid_input = valid_input;
    if (lazy) {
        g->n_valid = -1;
        goto fail;  // (common clean-up)
    }
    NN_TRY(cudaMemcpyAsync(&last_prefix, g->prefix + g->n_src - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NN_TRY(cudaMemcpyAsync(&last_valid, valid_input + g->n_src - 1, 1, cudaMemcpyDeviceToHost, s));
    NN_TRY(cudaStreamSynchronize(s));
    g->n_valid = (int64_t)last_prefix + (last_valid ? 1 : 0);
fail:
    if (tmp) cudaFreeAsync(tmp, s);
    return rc;
}

static int nn_build(b200_nn_grid *g, const b200_geo_t *src, const uint8_t *mask, uint8_t *valid_input,
                    cudaStream_t s, bool lazy_count) {
    int rc = B200_OK;
    const bool bucket = g->mode == B200_NN_BUCKET;
    const int64_t n_src = g->n_src;
    int32_t *counts = nullptr, *cell_of = nullptr, *h_off = nullptr;
    double4 *rec = nullptr;
    SrcStats *d_st = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int64_t ncells = 0;
    int nbands = 1;
    double inv_dphi = 1.0, phi0 = -B200_HALFPI;
    SrcArg S;
    S.g = *src;
    GeosFast fast;
    memset(&fast, 0, sizeof(fast));
    b200_pool_keep_memory();
    if (bucket) {
        // ---- cell size from the sampled point density, bands over the sampled latitude range
        SrcStats h_st;
        h_st.area_sum = 0.0;
        h_st.pairs = 0;
        h_st.phi_min_k = LLONG_MAX;
        h_st.phi_max_k = LLONG_MIN;
        const int rs = (int)(src->nrows < NN_STATS_SIDE ? src->nrows : NN_STATS_SIDE);
        const int cs = (int)(src->width < NN_STATS_SIDE ? src->width : NN_STATS_SIDE);
        NN_TRY(cudaMallocAsync((void **)&d_st, sizeof(SrcStats), s));
        NN_TRY(cudaMemcpyAsync(d_st, &h_st, sizeof(SrcStats), cudaMemcpyHostToDevice, s));
        nn_source_stats_kernel<<<(rs * cs + 255) / 256, 256, 0, s>>>(S, mask, rs, cs, d_st);
        NN_TRY(cudaGetLastError());
        NN_TRY(cudaMemcpyAsync(&h_st, d_st, sizeof(SrcStats), cudaMemcpyDeviceToHost, s));
        NN_TRY(cudaStreamSynchronize(s));
        cudaFreeAsync(d_st, s);
        d_st = nullptr;
        double dphi = b200_cap_angle(g->radius);  // no usable sample: one cell per search cap
        if (h_st.pairs >= 16 && h_st.area_sum > 0.0)
            dphi = sqrt(B200_NN_CELL_POINTS * h_st.area_sum / (double)h_st.pairs) / B200_R_EARTH;
        if (!(dphi > 1e-9)) dphi = 1e-9;
        if (dphi > B200_PI) dphi = B200_PI;
        double lo = -B200_HALFPI, hi = B200_HALFPI;
        if (h_st.phi_min_k <= h_st.phi_max_k) {
            lo = fmax(lo, (double)h_st.phi_min_k * 1e-12 - 2.0 * dphi);
            hi = fmin(hi, (double)h_st.phi_max_k * 1e-12 + 2.0 * dphi);
        }
        // cells are ~dphi x dphi: keep the table below B200_NN_MAX_CELLS
        const double zone = B200_TWO_PI * fmax(sin(hi) - sin(lo), 1e-12);
        if (zone / (dphi * dphi) > (double)B200_NN_MAX_CELLS) dphi = sqrt(zone / (double)B200_NN_MAX_CELLS);
        nbands = (int)ceil((hi - lo) / dphi);
        if (nbands < 1) nbands = 1;
        inv_dphi = 1.0 / dphi;
        phi0 = lo;
        h_off = new (std::nothrow) int32_t[nbands + 1];
        if (!h_off) {
            b200_set_error("b200_nn_grid_create: out of host memory");
            return B200_ENOMEM;
        }
        for (int b = 0; b < nbands; ++b) {
            double blo = lo + b * dphi, bhi = fmin(blo + dphi, B200_HALFPI);
            double cmax = (blo <= 0.0 && bhi >= 0.0) ? 1.0 : fmax(cos(blo), cos(bhi));
            int64_t nb = (int64_t)floor(B200_TWO_PI * cmax / dphi);
            if (nb < 1) nb = 1;
            h_off[b] = (int32_t)ncells;
            ncells += nb;
        }
        h_off[nbands] = (int32_t)ncells;
        g->ncells = ncells;
        g->nbands = nbands;
        g->dphi = dphi;
        g->phi0 = phi0;
        NN_TRY(cudaMallocAsync((void **)&g->band_off, (size_t)(nbands + 1) * sizeof(int32_t), s));
        NN_TRY(cudaMemcpyAsync(g->band_off, h_off, (size_t)(nbands + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        NN_TRY(cudaStreamSynchronize(s));  // h_off is pageable: finish before freeing it
        NN_TRY(cudaMallocAsync((void **)&cell_of, (size_t)n_src * sizeof(int32_t), s));
    }
    rc = b200_geos_fast_prepare(src, &fast, s);
    if (rc) goto fail;
    if (!fast.enabled) g->light = 0;   // (a light structure recomputes exact points from these tables)
    if (!g->light) NN_TRY(cudaMallocAsync((void **)&rec, (size_t)n_src * sizeof(double4), s));
    if (!bucket) NN_TRY(cudaMallocAsync((void **)&g->pts32, (size_t)n_src * sizeof(float4), s));
    if (fast.enabled)
        nn_source_points_kernel<true><<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(
            S, mask, valid_input, rec, g->pts32, cell_of, g->band_off, phi0, inv_dphi, nbands, fast);
    else
        nn_source_points_kernel<false><<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(
            S, mask, valid_input, rec, g->pts32, cell_of, g->band_off, phi0, inv_dphi, nbands, fast);
    NN_TRY(cudaGetLastError());
    if (g->light && fast.enabled) {   // the tables stay with the structure
        g->fast = fast;
        memset(&fast, 0, sizeof(fast));
    }
    b200_geos_fast_release(&fast, s);
    if (g->light) {   // fused path: no compressed indices are ever asked for -- no prefix table, no count
        g->valid_input = valid_input;
        g->n_valid = -1;
        g->nbytes = n_src * (int64_t)sizeof(float4);
        goto fail;  // (common clean-up)
    }
    rc = nn_prefix(g, valid_input, s, lazy_count && !bucket);
    if (rc) goto fail;
    if (!bucket) {
        g->pts = rec;  // the raw-indexed records ARE the search structure
        rec = nullptr;
        g->nbytes = n_src * (int64_t)(sizeof(double4) + sizeof(float4) + sizeof(int32_t));
        goto fail;  // (common clean-up)
    }
    NN_TRY(cudaMallocAsync((void **)&g->cell_start, (size_t)(ncells + 1) * sizeof(int32_t), s));
    NN_TRY(cudaMallocAsync((void **)&counts, (size_t)(ncells + 1) * sizeof(int32_t), s));
    NN_TRY(cudaMemsetAsync(counts, 0, (size_t)(ncells + 1) * sizeof(int32_t), s));
    nn_cell_count_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(cell_of, n_src, counts);
    NN_TRY(cudaGetLastError());
    NN_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, g->cell_start, (int)(ncells + 1), s));
    NN_TRY(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
    NN_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, g->cell_start, (int)(ncells + 1), s));
    // counts becomes the per-cell write cursor
    NN_TRY(cudaMemcpyAsync(counts, g->cell_start, (size_t)ncells * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    NN_TRY(cudaMallocAsync((void **)&g->pts, (size_t)(g->n_valid > 0 ? g->n_valid : 1) * sizeof(double4), s));
    if (g->n_valid > 0) {
        nn_cell_scatter_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(cell_of, rec, n_src, counts, g->pts);
        NN_TRY(cudaGetLastError());
    }
    g->nbytes = (int64_t)(nbands + 1) * 4 + (ncells + 1) * 4 + n_src * 4 + (g->n_valid > 0 ? g->n_valid : 1) * 32;
fail:
    delete[] h_off;
    b200_geos_fast_release(&fast, s);
    if (counts) cudaFreeAsync(counts, s);
    if (cell_of) cudaFreeAsync(cell_of, s);
    if (rec) cudaFreeAsync(rec, s);
    if (d_st) cudaFreeAsync(d_st, s);
    if (tmp) cudaFreeAsync(tmp, s);
    return rc;
}

extern "C" int b200_nn_grid_create_ex(const b200_geo_t *src, const uint8_t *mask, double radius, uint8_t *valid_input,
                                      int64_t *n_valid, b200_nn_grid_t **grid, int32_t method, void *stream) {
    int rc = b200_check_geo(src, "b200_nn_grid_create");
    if (rc) return rc;
    B200_CHECK_ARG(grid != nullptr, "NULL output argument");
    B200_CHECK_ARG(valid_input != nullptr, "valid_input is NULL");
    B200_CHECK_ARG(radius > 0.0 && isfinite(radius), "radius_of_influence must be positive and finite");
    const bool want_light = (method & B200_NN_LIGHT) != 0;
    method &= ~B200_NN_LIGHT;
    B200_CHECK_ARG(method >= 0 && method <= 3, "method must be 0 (auto), 1 (bucket), 2 (fixed grid) or 3 (fixed grid, strict)");
    // the structure covers the source row window [row0, row0 + nrows); indices are relative to its first pixel
    const int64_t n_src = src->nrows * src->width;
    B200_CHECK_ARG(n_src > 0, "empty source row window");
    B200_CHECK_ARG(n_src < (1LL << 31), "more than 2^31 source pixels");
    const bool fixed_ok = b200_nn_fixed_grid_supported(src) != 0;
    const int strict = method == B200_NN_FIXED_GRID_STRICT;
    if (strict) method = B200_NN_FIXED_GRID;
    if (method == B200_NN_FIXED_GRID && !fixed_ok) {
        b200_set_error("b200_nn_grid_create: the fixed-grid search needs a geostationary area source");
        return B200_EUNSUPPORTED;
    }
    b200_nn_grid *g = new (std::nothrow) b200_nn_grid();
    if (!g) {
        b200_set_error("b200_nn_grid_create: out of host memory");
        return B200_ENOMEM;
    }
    memset(g, 0, sizeof(*g));
    g->mode = (method == 0) ? (fixed_ok ? B200_NN_FIXED_GRID : B200_NN_BUCKET) : method;
    g->strict = strict;
    // light: only the fixed-grid engine, only without a mask (an off-disk test cannot reproduce a mask), only when the
    // caller does not want the count now
    g->light = want_light && g->mode == B200_NN_FIXED_GRID && mask == nullptr && n_valid == nullptr;
    g->n_src = n_src;
    g->radius = radius;
    g->src = *src;
    cudaGetDevice(&g->device);
    cudaStream_t s = (cudaStream_t)stream;
    g->stream = stream;
    // n_valid == NULL: the caller does not need the count now.  The fixed-grid engine then builds without any host
    // synchronisation (b200_nn_grid_n_valid reads the count back later); the bucket engine sizes its tables from it.
    rc = nn_build(g, src, mask, valid_input, s, n_valid == nullptr);
    if (rc) {
        b200_nn_grid_destroy(g);
        return rc;
    }
    if (n_valid) *n_valid = g->n_valid;
    *grid = g;
    return B200_OK;
}

extern "C" int b200_nn_grid_n_valid(b200_nn_grid_t *grid, int64_t *n_valid) {
    B200_CHECK_ARG(grid != nullptr && n_valid != nullptr, "NULL argument");
    if (grid->n_valid < 0 && grid->light) {
        b200_set_error("b200_nn_grid_n_valid: a light search structure keeps no count");
        return B200_EUNSUPPORTED;
    }
    if (grid->n_valid < 0) {  // built lazily: read the last prefix entry + the last mask byte back now
        int32_t last_prefix = 0;
        uint8_t last_valid = 0;
        cudaStream_t s = (cudaStream_t)grid->stream;
        B200_CUDA(cudaMemcpyAsync(&last_prefix, grid->prefix + grid->n_src - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(&last_valid, grid->valid_input + grid->n_src - 1, 1, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        grid->n_valid = (int64_t)last_prefix + (last_valid ? 1 : 0);
    }
    *n_valid = grid->n_valid;
    return B200_OK;
}

extern "C" int b200_nn_grid_create(const b200_geo_t *src, const uint8_t *mask, double radius, uint8_t *valid_input,
                                   int64_t *n_valid, b200_nn_grid_t **grid, void *stream) {
    return b200_nn_grid_create_ex(src, mask, radius, valid_input, n_valid, grid, 0, stream);
}

extern "C" int b200_nn_grid_method(const b200_nn_grid_t *grid, int32_t *method) {
    B200_CHECK_ARG(grid != nullptr && method != nullptr, "NULL argument");
    *method = grid->mode;
    return B200_OK;
}

// ---------------------------------------------------------------------------------------- query
__device__ __forceinline__ void nn_scan_range(const double4 *__restrict__ pts, int s, int e, double tx, double ty,
                                              double tz, Best &best) {
    for (int j = s; j < e; ++j) b200_nn_consider(pts + j, tx, ty, tz, best);
}

// longitude half-width of the spherical cap of angular radius theta centred at latitude phi (>= pi: all longitudes)
__device__ __forceinline__ double nn_cap_dlam(double phi, double theta, double sin_theta) {
    if (fabs(phi) + theta >= B200_HALFPI - 1e-12) return 4.0;
    double q = sin_theta / cos(phi);
    if (q >= 1.0) return 4.0;
    return asin(q) * (1.0 + 1e-12) + 1e-12;
}

// scan every record of the cells that meet the cap (bands b_lo..b_hi, longitudes lam +- dlam)
__device__ __forceinline__ void nn_bucket_scan(const GridDev &G, double lam, double tx, double ty, double tz, int b_lo,
                                               int b_hi, double dlam, Best &best) {
    const bool full = dlam >= B200_PI;
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
}

// Search caps of one target: the full radius of influence and, when cells are much smaller than that, a first cap of
// one cell size (radius r1).
struct Caps {
    int b_lo, b_hi, b_lo1, b_hi1;
    double dlam, dlam1;
};

// nearest valid source point of one target given its spherical coordinates; -1 when none within the radius.
// The first cap covers every point closer than r1: a hit closer than r1 in it is the overall nearest (everything
// not scanned is at least r1 away), which settles the targets inside a dense source after a few records.
__device__ __forceinline__ long long nn_bucket_point(const GridDev &G, double lam, double tx, double ty, double tz,
                                                     const Caps &C, double r2, double r1_2) {
    Best best;
    best.d2 = r2;  // strict "<": a neighbour exactly at the radius is not accepted
    best.idx = -1;
    if (r1_2 > 0.0) {
        nn_bucket_scan(G, lam, tx, ty, tz, C.b_lo1, C.b_hi1, C.dlam1, best);
        if (best.d2 < r1_2) return best.idx;
    }
    nn_bucket_scan(G, lam, tx, ty, tz, C.b_lo, C.b_hi, C.dlam, best);
    return best.idx;
}

struct QueryArg {
    b200_geo_t dst;
    GridDev G;
    QueryOut O;
    double theta, sin_theta, r2;
    double theta1, sin_theta1, r1_2;  // first cap; r1_2 = 0: single phase
    const TgtCol *cols;  // rectilinear targets only
    const TgtRow *rows;
};

// general targets: everything per pixel
__global__ void __launch_bounds__(256) nn_bucket_query_kernel(const __grid_constant__ QueryArg Q) {
    const int64_t W = Q.dst.width;
    const int64_t n = Q.dst.nrows * W;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t r = i / W;
        double lon, lat;
        b200_geo_lonlat(Q.dst, Q.dst.row0 + r, i - r * W, lon, lat);
        bool ok = b200_lonlat_valid(lon, lat);
        long long raw = -1;
        if (ok) {
            double tx, ty, tz, lam, phi;
            b200_lonlat2xyz(lon, lat, tx, ty, tz, lam, phi);
            Caps C;
            C.b_lo = b200_band_of(phi - Q.theta, Q.G.phi0, Q.G.inv_dphi, Q.G.nbands);
            C.b_hi = b200_band_of(phi + Q.theta, Q.G.phi0, Q.G.inv_dphi, Q.G.nbands);
            C.dlam = nn_cap_dlam(phi, Q.theta, Q.sin_theta);
            C.b_lo1 = C.b_hi1 = 0;
            C.dlam1 = 0.0;
            if (Q.r1_2 > 0.0) {
                C.b_lo1 = b200_band_of(phi - Q.theta1, Q.G.phi0, Q.G.inv_dphi, Q.G.nbands);
                C.b_hi1 = b200_band_of(phi + Q.theta1, Q.G.phi0, Q.G.inv_dphi, Q.G.nbands);
                C.dlam1 = nn_cap_dlam(phi, Q.theta1, Q.sin_theta1);
            }
            raw = nn_bucket_point(Q.G, lam, tx, ty, tz, C, Q.r2, Q.r1_2);
        }
        b200_nn_emit(Q.O, Q.G.prefix, i, raw, ok);
    }
}

// rectilinear targets: one block iteration = 256 consecutive columns of one row, tables instead of trigonometry
__global__ void __launch_bounds__(256) nn_bucket_query_rect_kernel(const __grid_constant__ QueryArg Q) {
    const int W = (int)Q.dst.width;
    const int nseg = (W + 255) >> 8;
    const int64_t items = (int64_t)Q.dst.nrows * nseg;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        int r = (int)(it / nseg);
        int col = (int)(it - (int64_t)r * nseg) * 256 + threadIdx.x;
        if (col >= W) continue;
        const TgtRow R = Q.rows[r];
        const TgtCol Cc = Q.cols[col];
        bool ok = R.valid && Cc.valid;
        long long raw = -1;
        if (ok) {
            double tx = B200_R_EARTH * R.cos_phi * Cc.cos_lam;
            double ty = B200_R_EARTH * R.cos_phi * Cc.sin_lam;
            double tz = B200_R_EARTH * R.sin_phi;
            Caps C;
            C.b_lo = R.b_lo;
            C.b_hi = R.b_hi;
            C.dlam = R.dlam;
            C.b_lo1 = R.b_1 & 0xffffff;
            C.b_hi1 = C.b_lo1 + (int)((unsigned)R.b_1 >> 24);
            C.dlam1 = R.dlam1;
            raw = nn_bucket_point(Q.G, Cc.lam, tx, ty, tz, C, Q.r2, Q.r1_2);
        }
        b200_nn_emit(Q.O, Q.G.prefix, (int64_t)r * W + col, raw, ok);
    }
}

// Tables of a rectilinear target (lon from the column, lat from the row), bit-identical to the per-pixel path.
struct TableArg {
    b200_geo_t dst;
    double theta, sin_theta, inv_dphi;  // bucket mode
    double theta1, sin_theta1, phi0;    // ... first cap (theta1 = 0: none), lower edge of band 0
    int32_t nbands, geos;
    double lam0;                        // fixed-grid mode: geos source constants
    double rp, rp2;
};

__global__ void __launch_bounds__(256)
nn_target_tables_kernel(const __grid_constant__ TableArg A, TgtCol *__restrict__ cols, TgtRow *__restrict__ rows) {
    const int64_t W = A.dst.width, H = A.dst.nrows;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < W + H; i += stride) {
        double lon, lat;
        if (i < W) {
            b200_geo_lonlat(A.dst, A.dst.row0, i, lon, lat);
            TgtCol c;
            c.lam = lon * B200_DEG2RAD;
            sincos(c.lam, &c.sin_lam, &c.cos_lam);
            c.valid = (lon >= -180.0 && lon <= 180.0) ? 1 : 0;
            double rel = b200_adjlon(c.lam - A.lam0);
            c.s_rel = (float)sin(rel);
            c.c_rel = (float)cos(rel);
            // an invalid column fails every visibility test (rc * c_rel >= vis_min) without its flag being read
            if (!c.valid) c.c_rel = __int_as_float(0x7fc00000);
            c.pad = 0;
            c.pad2 = 0.0;
            cols[i] = c;
        } else {
            int64_t r = i - W;
            b200_geo_lonlat(A.dst, A.dst.row0 + r, 0, lon, lat);
            TgtRow t;
            t.phi = lat * B200_DEG2RAD;
            sincos(t.phi, &t.sin_phi, &t.cos_phi);
            t.valid = (lat >= -90.0 && lat <= 90.0) ? 1 : 0;
            t.dlam = t.dlam1 = 0.0;
            t.b_lo = t.b_hi = t.b_1 = 0;
            t.rc = t.rs = 0.0f;
            if (t.valid) {
                if (A.geos) {
                    // geocentric latitude and ellipsoid radius of PROJ's geos forward, per row
                    double pg = atan(A.rp2 * tan(t.phi));
                    double rr = A.rp / hypot(A.rp * cos(pg), sin(pg));
                    t.rc = (float)(rr * cos(pg));
                    t.rs = (float)(rr * sin(pg));
                    // fixed-grid mode has no use for the search-cap fields: they carry the row's part of the target's
                    // spherical point instead (R cos(phi), R sin(phi) in float64, the latter also rounded to float32),
                    // so that the query kernel does one float64 multiplication per coordinate, not two
                    t.dlam = B200_R_EARTH * t.cos_phi;
                    t.dlam1 = B200_R_EARTH * t.sin_phi;
                    t.b_lo = __float_as_int((float)t.dlam1);
                } else {
                    t.dlam = nn_cap_dlam(t.phi, A.theta, A.sin_theta);
                    t.b_lo = b200_band_of(t.phi - A.theta, A.phi0, A.inv_dphi, A.nbands);
                    t.b_hi = b200_band_of(t.phi + A.theta, A.phi0, A.inv_dphi, A.nbands);
                    if (A.theta1 > 0.0) {
                        const int lo1 = b200_band_of(t.phi - A.theta1, A.phi0, A.inv_dphi, A.nbands);
                        const int hi1 = b200_band_of(t.phi + A.theta1, A.phi0, A.inv_dphi, A.nbands);
                        t.b_1 = lo1 | ((hi1 - lo1) << 24);
                        t.dlam1 = nn_cap_dlam(t.phi, A.theta1, A.sin_theta1);
                    }
                }
            }
            rows[r] = t;
        }
    }
}

int b200_nn_build_tables(const b200_geo_t *dst, double theta, double theta1, double phi0, double inv_dphi, int nbands,
                         int geos, double lam0, double rp, double rp2, TgtCol **cols, TgtRow **rows, cudaStream_t s) {
    *cols = nullptr;
    *rows = nullptr;
    B200_CUDA(cudaMallocAsync((void **)cols, (size_t)dst->width * sizeof(TgtCol), s));
    cudaError_t e = cudaMallocAsync((void **)rows, (size_t)(dst->nrows > 0 ? dst->nrows : 1) * sizeof(TgtRow), s);
    if (e != cudaSuccess) {
        cudaFreeAsync(*cols, s);
        b200_set_error("b200_nn_query: table allocation failed: %s", cudaGetErrorString(e));
        return B200_ENOMEM;
    }
    TableArg A;
    A.dst = *dst;
    A.theta = theta;
    A.sin_theta = theta >= B200_HALFPI ? 1.0 : sin(theta);
    A.theta1 = theta1;
    A.sin_theta1 = theta1 >= B200_HALFPI ? 1.0 : sin(theta1);
    A.phi0 = phi0;
    A.inv_dphi = inv_dphi;
    A.nbands = nbands;
    A.geos = geos;
    A.lam0 = lam0;
    A.rp = rp;
    A.rp2 = rp2;
    nn_target_tables_kernel<<<b200_grid_for(dst->width + dst->nrows, 256, 8), 256, 0, s>>>(A, *cols, *rows);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

static int nn_bucket_query(const b200_nn_grid *grid, const b200_geo_t *dst, double radius, const QueryOut &O,
                           cudaStream_t s) {
    QueryArg Q;
    Q.dst = *dst;
    Q.G.band_off = grid->band_off;
    Q.G.cell_start = grid->cell_start;
    Q.G.pts = grid->pts;
    Q.G.prefix = grid->prefix;
    Q.G.inv_dphi = 1.0 / grid->dphi;
    Q.G.phi0 = grid->phi0;
    Q.G.nbands = grid->nbands;
    Q.O = O;
    Q.theta = b200_cap_angle(radius);
    Q.sin_theta = Q.theta >= B200_HALFPI ? 1.0 : sin(Q.theta);
    Q.r2 = radius * radius;
    // first cap: one cell size, when that is well below the radius of influence (and the band span fits the tables)
    const double r1 = grid->dphi * B200_R_EARTH;
    Q.theta1 = Q.sin_theta1 = Q.r1_2 = 0.0;
    if (r1 < 0.6 * radius) {
        Q.theta1 = b200_cap_angle(r1);
        Q.sin_theta1 = sin(Q.theta1);
        Q.r1_2 = r1 * r1;
    }
    Q.cols = nullptr;
    Q.rows = nullptr;
    const int64_t n = dst->nrows * dst->width;
    if (b200_geo_is_rectilinear(dst) && dst->width < (1LL << 30)) {
        TgtCol *cols;
        TgtRow *rows;
        int rc = b200_nn_build_tables(dst, Q.theta, Q.theta1, Q.G.phi0, Q.G.inv_dphi, Q.G.nbands, 0, 0.0, 0.0, 0.0, &cols,
                                      &rows, s);
        if (rc) return rc;
        Q.cols = cols;
        Q.rows = rows;
        int64_t items = dst->nrows * ((dst->width + 255) / 256);
        nn_bucket_query_rect_kernel<<<b200_grid_for(items, 1, 16), 256, 0, s>>>(Q);
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(cols, s);
        cudaFreeAsync(rows, s);
        if (e != cudaSuccess) {
            b200_set_error("b200_nn_query: kernel launch failed: %s", cudaGetErrorString(e));
            return B200_ECUDA;
        }
        return B200_OK;
    }
    nn_bucket_query_kernel<<<b200_grid_for(n, 256, 8), 256, 0, s>>>(Q);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

static int nn_query_dispatch(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius, const QueryOut &O,
                             void *stream, const char *who) {
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
    if (dst->nrows * dst->width == 0) return B200_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (grid->mode == B200_NN_FIXED_GRID) return b200_nn_fixed_grid_query(grid, dst, radius, O, s);
    return nn_bucket_query(grid, dst, radius, O, s);
}

extern "C" int b200_nn_query(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius, int32_t *index_array,
                             int32_t *gather_index, uint8_t *valid_output, void *stream) {
    QueryOut O;
    memset(&O, 0, sizeof(O));
    O.index_array = index_array;
    O.gather_index = gather_index;
    O.valid_output = valid_output;
    return nn_query_dispatch(grid, dst, radius, O, stream, "b200_nn_query");
}

extern "C" int b200_nn_query_gather_f32(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius,
                                        const float *data, float *out, int32_t nbands, int64_t src_band_stride,
                                        int64_t dst_band_stride, float fill, void *stream) {
    B200_CHECK_ARG(nbands >= 1, "nbands must be >= 1");
    B200_CHECK_ARG(dst != nullptr, "geometry is NULL");
    int64_t n = dst->nrows * dst->width;
    if (n > 0) {
        B200_CHECK_ARG(data != nullptr && out != nullptr, "NULL data pointer");
        B200_CHECK_ARG(grid == nullptr || nbands == 1 || (src_band_stride >= grid->n_src && dst_band_stride >= n),
                       "band stride too small");
    }
    QueryOut O;
    memset(&O, 0, sizeof(O));
    O.data = data;
    O.out = out;
    O.nbands = nbands;
    O.src_band_stride = src_band_stride;
    O.dst_band_stride = dst_band_stride;
    O.fill = fill;
    return nn_query_dispatch(grid, dst, radius, O, stream, "b200_nn_query_gather_f32");
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
    const bool out_aligned = (((uintptr_t)out) & 15) == 0;  // else (e.g. a row band of an odd-width image): scalar stores
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
            if (sizeof(T) == 4 && out_aligned && ((dst_band_stride & 3) == 0 || nbands == 1)) {
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
static void launch_gather(const void *data, void *out, const int32_t *gidx, int64_t n_out, int nbands,
                          int64_t src_band_stride, int64_t dst_band_stride, const void *fill, cudaStream_t s) {
    T f;
    memcpy(&f, fill, sizeof(T));
    int64_t work = (n_out + 3) / 4;
    nn_gather_kernel<T><<<b200_grid_for(work, 256, 8), 256, 0, s>>>((const T *)data, (T *)out, gidx, n_out, nbands,
                                                                   src_band_stride, dst_band_stride, f);
}

// Bulk-staged variant (4-byte elements, any number of bands): the index stream comes in and the result goes out through
// shared memory with cp.async.bulk (b200_bulk.cuh), GATHER_STAGES tiles in flight per CTA; threads only issue the
// scattered source loads.  A/B against the direct kernel in scripts/variant_bench.py.
#define GATHER_TILE 2048
#define GATHER_STAGES 3
struct __align__(128) GatherStage {
    int32_t idx[GATHER_TILE];
    uint32_t out[GATHER_TILE];
};

// MULTI = false: one band, compiled without the band loop (the single-band kernel lost 9 % when the loop was added)
template <bool MULTI>
__global__ void __launch_bounds__(256)
nn_gather_bulk_kernel(const uint32_t *__restrict__ data, uint32_t *__restrict__ out, const int32_t *__restrict__ gidx,
                      int64_t n_tiles, uint32_t fill, int nbands_, int64_t src_band_stride, int64_t dst_band_stride) {
    const int nbands = MULTI ? nbands_ : 1;
    extern __shared__ __align__(128) unsigned char gather_smem[];
    GatherStage *stage = reinterpret_cast<GatherStage *>(gather_smem);
    __shared__ uint64_t full[GATHER_STAGES];
    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < GATHER_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    if (t == 0) {
        for (int s = 0; s < GATHER_STAGES; ++s) {
            const int64_t tile = first + (int64_t)s * step;
            if (tile < n_tiles) {
                mbar_expect_tx(&full[s], GATHER_TILE * 4);
                bulk_g2s(stage[s].idx, gidx + tile * GATHER_TILE, GATHER_TILE * 4, &full[s]);
            }
        }
    }
    int s = 0;
    uint32_t phase = 0;
    for (int64_t tile = first; tile < n_tiles; tile += step) {
        mbar_wait(&full[s], phase);
#pragma unroll 1
        for (int b = 0; b < nbands; ++b) {
            // the bulk store that last read stage[s].out must have finished reading it: the one GATHER_STAGES tiles ago
            // for the first band, the previous band's for the others (one output tile per stage serves every band)
            if (t == 0) {
                if (b == 0) bulk_wait_read<GATHER_STAGES - 1>();
                else bulk_wait_read<0>();
            }
            __syncthreads();
            const uint32_t *src = MULTI ? data + (int64_t)b * src_band_stride : data;
#pragma unroll
            for (int j = 0; j < GATHER_TILE / 1024; ++j) {
                const int4 ix = *reinterpret_cast<const int4 *>(stage[s].idx + j * 1024 + 4 * t);
                uint4 v;
                v.x = ix.x < 0 ? fill : __ldg(src + ix.x);
                v.y = ix.y < 0 ? fill : __ldg(src + ix.y);
                v.z = ix.z < 0 ? fill : __ldg(src + ix.z);
                v.w = ix.w < 0 ? fill : __ldg(src + ix.w);
                *reinterpret_cast<uint4 *>(stage[s].out + j * 1024 + 4 * t) = v;
            }
            bulk_fence_smem_writes();
            __syncthreads();
            if (t == 0) {
                bulk_s2g((MULTI ? out + (int64_t)b * dst_band_stride : out) + tile * GATHER_TILE, stage[s].out, GATHER_TILE * 4);
                bulk_commit();
            }
        }
        if (t == 0) {
            const int64_t next = tile + (int64_t)GATHER_STAGES * step;
            if (next < n_tiles) {   // every thread has read stage[s].idx for the last band (barrier above)
                mbar_expect_tx(&full[s], GATHER_TILE * 4);
                bulk_g2s(stage[s].idx, gidx + next * GATHER_TILE, GATHER_TILE * 4, &full[s]);
            }
        }
        if (++s == GATHER_STAGES) {
            s = 0;
            phase ^= 1;
        }
    }
    if (t == 0) bulk_wait_all();
}

static int g_gather_variant = -1;
#ifndef B200_GATHER_DEFAULT_VARIANT
#define B200_GATHER_DEFAULT_VARIANT 1
#endif

extern "C" int b200_nn_gather_variant(const void *data, void *out, const int32_t *gather_index, int64_t n_out,
                                      int32_t nbands, int64_t src_band_stride, int64_t dst_band_stride,
                                      int32_t elem_size, const void *fill_value, int32_t variant, void *stream);

extern "C" int b200_nn_gather(const void *data, void *out, const int32_t *gather_index, int64_t n_out, int32_t nbands,
                              int64_t src_band_stride, int64_t dst_band_stride, int32_t elem_size, const void *fill,
                              void *stream) {
    if (g_gather_variant < 0) {
        const char *e = getenv("B200_GATHER_VARIANT");
        g_gather_variant = e ? atoi(e) : B200_GATHER_DEFAULT_VARIANT;
        if (g_gather_variant < 0 || g_gather_variant > 1) g_gather_variant = B200_GATHER_DEFAULT_VARIANT;
    }
    return b200_nn_gather_variant(data, out, gather_index, n_out, nbands, src_band_stride, dst_band_stride, elem_size,
                                  fill, g_gather_variant, stream);
}

extern "C" int b200_nn_gather_variant(const void *data, void *out, const int32_t *gather_index, int64_t n_out,
                                      int32_t nbands, int64_t src_band_stride, int64_t dst_band_stride,
                                      int32_t elem_size, const void *fill, int32_t variant, void *stream) {
    B200_CHECK_ARG(n_out >= 0 && nbands >= 1, "bad shape");
    B200_CHECK_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8,
                   "elem_size must be 1, 2, 4 or 8");
    B200_CHECK_ARG(variant == 0 || variant == 1, "variant must be 0 (direct) or 1 (bulk-staged)");
    B200_CHECK_ARG(fill != nullptr, "fill is NULL");
    if (n_out == 0) return B200_OK;
    B200_CHECK_ARG(data != nullptr && out != nullptr && gather_index != nullptr, "NULL data pointer");
    B200_CHECK_ARG(((uintptr_t)gather_index & 15) == 0, "gather_index must be 16-byte aligned");
    B200_CHECK_ARG(((uintptr_t)out & (uintptr_t)(elem_size - 1)) == 0 && ((uintptr_t)data & (uintptr_t)(elem_size - 1)) == 0,
                   "data and out must be aligned to the element size");
    B200_CHECK_ARG(nbands == 1 || dst_band_stride >= n_out, "band stride smaller than one image");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t done = 0;
    if (variant == 1 && elem_size == 4 && (nbands == 1 || (dst_band_stride & 3) == 0) && (((uintptr_t)out) & 15) == 0 &&
        n_out >= GATHER_TILE) {
        // whole tiles through the bulk-staged kernel, the rest (< one tile) through the direct one
        const int64_t n_tiles = n_out / GATHER_TILE;
        const size_t smem = sizeof(GatherStage) * GATHER_STAGES + 128;
        static bool attr_set = false;
        if (!attr_set) {
            B200_CUDA(cudaFuncSetAttribute(nn_gather_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CUDA(cudaFuncSetAttribute(nn_gather_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        int64_t grid = (int64_t)B200_NUM_SMS * 4;
        if (grid > n_tiles) grid = n_tiles;
        uint32_t f;
        memcpy(&f, fill, 4);
        if (nbands == 1)
            nn_gather_bulk_kernel<false><<<(int)grid, 256, smem, s>>>(static_cast<const uint32_t *>(data), static_cast<uint32_t *>(out),
                                                                    gather_index, n_tiles, f, 1, 0, 0);
        else
            nn_gather_bulk_kernel<true><<<(int)grid, 256, smem, s>>>(static_cast<const uint32_t *>(data), static_cast<uint32_t *>(out),
                                                                   gather_index, n_tiles, f, nbands, src_band_stride, dst_band_stride);
        B200_LAUNCH_CHECK();
        done = n_tiles * GATHER_TILE;
        if (done == n_out) return B200_OK;
    }
    const char *o = static_cast<char *>(out) + done * elem_size;
    const int32_t *gi = gather_index + done;
    const int64_t n = n_out - done;
    switch (elem_size) {
    case 1: launch_gather<uint8_t>(data, (void *)o, gi, n, nbands, src_band_stride, dst_band_stride, fill, s); break;
    case 2: launch_gather<uint16_t>(data, (void *)o, gi, n, nbands, src_band_stride, dst_band_stride, fill, s); break;
    case 4: launch_gather<uint32_t>(data, (void *)o, gi, n, nbands, src_band_stride, dst_band_stride, fill, s); break;
    default: launch_gather<uint64_t>(data, (void *)o, gi, n, nbands, src_band_stride, dst_band_stride, fill, s); break;
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
    int64_t n_alloc = n_src > 0 ? n_src : 1;
    thrust::transform_iterator<U8ToI32, const uint8_t *> vit(valid_input, U8ToI32());
    NN_TRY(cudaMalloc(&prefix, (size_t)(n_alloc + 1) * 4));
    NN_TRY(cudaMalloc(&raw_of, (size_t)n_alloc * 4));
    // 0xff.. = -1: compressed indices >= n_valid (the kd-tree miss convention) map to "no neighbour"
    NN_TRY(cudaMemsetAsync(raw_of, 0xff, (size_t)n_alloc * 4, s));
    if (n_src > 0) {
        NN_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, vit, prefix, (int)n_src, s));
        NN_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
        NN_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, vit, prefix, (int)n_src, s));
        nn_raw_of_kernel<<<b200_grid_for(n_src, 256, 8), 256, 0, s>>>(valid_input, prefix, n_src, raw_of);
    }
    // n_src bounds the compressed index; slots in [n_valid, n_src) still hold -1
    nn_map_index_kernel<<<b200_grid_for(n_out, 256, 8), 256, 0, s>>>(index_array, raw_of, n_src, gather_index, n_out);
    NN_TRY(cudaGetLastError());
    NN_TRY(cudaStreamSynchronize(s));
fail:
    cudaFree(prefix);
    cudaFree(raw_of);
    cudaFree(tmp);
    return rc;
}
