/*
 * satpy_b200.h -- C ABI of libsatpy_b200.so (hand-written sm_100a CUDA kernels).
 *
 * This is the drop-in boundary for satpy's per-pixel hot path.  satpy itself is
 * pure Python and has no FFI for this path: the arithmetic lives inline in its
 * Python file handlers / modifiers and in the pyresample / pykdtree / pyorbital
 * wheels.  Each entry point below names the reference interface whose work it
 * replaces (file:line under /root/reference/satpy unless stated), so a
 * maintainer can bind it with ctypes.
 *
 * Conventions
 *   - every pointer that is not marked "host" is a DEVICE pointer into memory
 *     owned by the caller (e.g. torch.Tensor.data_ptr());
 *   - sizes are int64_t, images are row-major (row 0 = northernmost line);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success, a B200_E* code otherwise, never
 *     aborts and never prints; b200_last_error() gives the thread-local text;
 *   - kernels are asynchronous on `stream`; nothing here synchronises except
 *     where stated (grid build returns counts through a device->host copy).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns B200_ECUDA.
 */
#ifndef SATPY_B200_H
#define SATPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL 1   /* bad argument (maps to ValueError)          */
#define B200_ECUDA 2    /* CUDA runtime error (maps to RuntimeError)  */
#define B200_ENOMEM 3   /* allocation failed (maps to MemoryError)    */
#define B200_EUNSUPPORTED 4 /* valid request this build cannot serve (NotImplementedError) */

/* ------------------------------------------------------------------ geometry
 * Projection + grid description, filled on the host by satpy_b200/geometry.py.
 * Stands in for pyresample.geometry.AreaDefinition (+ pyproj) as used through
 * area.get_lonlats() at modifiers/angles.py:434-441 and resample/kdtree.py:79-87.
 */
enum b200_proj_kind {
    B200_PROJ_LONGLAT = 0,
    B200_PROJ_EQC = 1,
    B200_PROJ_MERC = 2,
    B200_PROJ_GEOS = 3,
    B200_PROJ_LAEA = 4,
    B200_PROJ_LCC = 5,
    B200_PROJ_STERE = 6
};

typedef struct b200_proj {
    int32_t kind;   /* enum b200_proj_kind */
    int32_t mode;   /* geos: 1 = sweep x; laea: 0 N,1 S,2 EQ,3 OB; stere: 0 north,1 south */
    double a, es, e, one_es;
    double lam0, phi0, x0, y0, k0;
    double p[12];   /* kind-specific derived constants, see satpy_b200/geometry.py */
} b200_proj_t;

typedef struct b200_area {
    b200_proj_t proj;
    int64_t width, height;
    double pixel_size_x, pixel_size_y; /* both positive */
    double x_ul, y_ul;                 /* projection coords of the CENTRE of pixel (0,0) */
} b200_area_t;

/* A source/target geometry is either an area (lon/lat computed on device) or a
 * swath (explicit lon/lat arrays, float64 or float32, degrees). */
typedef struct b200_geo {
    int32_t is_swath;      /* 0: use `area`; 1: use lons/lats                      */
    int32_t lonlat_is_f32; /* swath arrays are float32 (else float64)              */
    b200_area_t area;      /* used when !is_swath                                  */
    const void *lons;      /* device, height*width                                 */
    const void *lats;      /* device, height*width                                 */
    int64_t width, height; /* grid shape (duplicates area.width/height for areas)  */
    int64_t row0, nrows;   /* the row window [row0, row0+nrows) this call covers   */
} b200_geo_t;

/* ------------------------------------------------------------------ library */
const char *b200_version(void);
const char *b200_last_error(void);
int b200_device_count(int *count /* host */);

/* lon/lat of pixel centres, float64 degrees, +inf where the projection has no
 * answer.  Replaces AreaDefinition.get_lonlats (modifiers/angles.py:438). */
int b200_area_lonlats(const b200_geo_t *geo /* host */, double *lons, double *lats, void *stream);

/* --------------------------------------------------------------- calibration
 * ABI L1b: readers/core/abi.py:138-174 (_adjust_data) + readers/abi_l1b.py:133-173.
 * mode 0 radiance, 1 reflectance [%], 2 brightness temperature [K].
 * Radiance and reflectance are bit-identical to numpy's float32 arithmetic; the Planck inversion of
 * mode 2 uses the hardware reciprocal / log2 units (within 4e-7 relative of the correctly rounded
 * chain; numpy's own float32 log is not correctly rounded either) and redoes pixels whose
 * logarithm argument is below 2 with IEEE divisions and logf.
 */
typedef struct b200_abi_cal {
    int32_t mode;
    int32_t is_unsigned;  /* _Unsigned == "true": reinterpret int16 as uint16      */
    int32_t has_fill;
    int32_t fill;         /* _FillValue as stored (int16 bit pattern, sign-extended) */
    int32_t clip_negative_radiances;
    float scale_factor, add_offset;
    float refl_factor;    /* float32(pi * esd^2 / esun)                             */
    float min_rad;        /* float32(ceil(-offset/scale)*scale+offset), abi_l1b.py:147-155 */
    float fk1, fk2, bc1, bc2;
} b200_abi_cal_t;

int b200_abi_calibrate(const int16_t *counts, float *out, int64_t n,
                       const b200_abi_cal_t *cal /* host */, void *stream);
/* The same with the kernel variant chosen by the caller (benchmarks, tests): 0 = one 32-byte block per thread (first
 * release), 1 = lane-interleaved 8-byte loads / 16-byte stores (every warp-level access one contiguous run),
 * 2 = bulk-asynchronous staging through shared memory (cp.async.bulk + mbarrier in both directions: no thread issues
 * a global load or store).  b200_abi_calibrate uses B200_CAL_VARIANT from the environment, else the measured winner.
 * All variants give bit-identical results. */
int b200_abi_calibrate_variant(const int16_t *counts, float *out, int64_t n, const b200_abi_cal_t *cal,
                               int32_t variant, void *stream);

/* AHI HSD (readers/ahi_hsd.py:607-611,675-750): uint16 counts -> radiance / reflectance [%] / BT [K].
 * The Planck constants are doubles; bt_in_f64 selects numpy's dtype flow when they are numpy float64
 * scalars (a real file header: float64 arithmetic) or Python floats (float32 arithmetic). */
typedef struct b200_ahi_cal {
    int32_t mode;          /* 0 radiance, 1 reflectance, 2 brightness temperature            */
    int32_t has_invalid;   /* mask count_outside_scan / count_error (-> NaN)                 */
    int32_t count_outside_scan, count_error;
    int32_t bt_in_f64, reserved;
    float gain, offset;    /* float32(dn_gain), float32(dn_offset)                           */
    float coeff, pad;      /* float32(coeff_rad2albedo_conversion)                           */
    double a;              /* h c / (k cwl)                                                  */
    double num;            /* 2 h c^2                                                        */
    double cwl5;           /* cwl^5, cwl = central_wave_length * 1e-6                        */
    double c0, c1, c2;     /* rad2tb polynomial                                              */
} b200_ahi_cal_t;

int b200_ahi_calibrate(const uint16_t *counts, float *out, int64_t n, const b200_ahi_cal_t *cal /* host */,
                       void *stream);

/* SEVIRI (readers/core/seviri.py:574-633,660-686): counts (float32 or uint16) -> radiance / reflectance
 * [%] / BT [K].  All constants already rounded to float32 by the host (numpy weak-scalar rule). */
typedef struct b200_seviri_cal {
    int32_t mode;          /* 0 radiance, 1 reflectance, 2 BT from effective, 3 BT from spectral radiance */
    int32_t reserved;
    float gain, offset;
    float pi_f, solar_irradiance, sun_earth_dist2; /* float32(pi), F, float32(d*d)            */
    float c1, nu3, c2nu;   /* float32(C1), float32(nu^3), float32(C2 * nu)                   */
    float alpha, beta;     /* effective radiance: (T - beta) / alpha                          */
    float fit_a, fit_b, fit_c; /* spectral radiance: a T^2 + b T + c (BTFIT)                  */
} b200_seviri_cal_t;

int b200_seviri_calibrate(const void *counts, int32_t counts_are_f32, float *out, int64_t n,
                          const b200_seviri_cal_t *cal /* host */, void *stream);

/* VIIRS SDR (readers/core/viirs_atms_sdr.py:197-214,328-361): fill masking (ints >= fill_min_int, floats
 * <= fill_max_float -> NaN) then data * factor + offset with one pair per granule of rows.
 * factors: host, n_granules x 2 float32; rows_per_granule: host, n_granules.  Synchronises `stream`. */
int b200_viirs_scale(const void *data, int32_t data_is_f32, float *out, int64_t rows, int64_t cols,
                     const float *factors /* host */, const int32_t *rows_per_granule /* host */,
                     int32_t n_granules, int32_t fill_min_int, float fill_max_float, void *stream);

/* ------------------------------------------------------------- sun zenith
 * get_cos_sza (modifiers/angles.py:418-469, pyorbital cos_zen) and
 * _sunzen_corr_cos_ndarray (modifiers/angles.py:555-584) as driven by
 * SunZenithCorrectorBase.__call__ (modifiers/geometry.py:48-71).
 */
typedef struct b200_sunz {
    double gmst, ra, dec;     /* radians, from start_time (host, satpy_b200/modifiers/angles.py) */
    double limit_deg;         /* correction_limit, default 88                        */
    double max_sza_deg;       /* default 95                                          */
    int32_t has_max_sza;      /* 0 => max_sza is None                                */
    int32_t lonlat_f32;       /* 1 => lon/lat rounded to float32 first (float data)  */
    int32_t method;           /* 0 SunZenithCorrector, 1 EffectiveSolarPathLength (utils.py:278-316) */
    int32_t reserved;
} b200_sunz_t;

/* cos(sza) per pixel (float64, NaN off-disk); rows [geo->row0, +nrows). */
int b200_cos_sza(const b200_geo_t *geo, const b200_sunz_t *sz, double *cos_sza, void *stream);

/* out = data * corr(cos_sza(lon,lat)); lon/lat derived in-kernel from geo.
 * nbands images of geo->nrows*width pixels, band stride = band_stride elements.
 * On geostationary areas cos(sza) comes straight from the view vector (no angles) where the
 * correction is insensitive to its last digits (cos(sza) > 0.2 or beyond max_sza: within 2.4e-6
 * relative of the reference); the band around the terminator reproduces the reference's float32
 * rounding of lon/lat digit by digit, as b200_cos_sza always does. */
int b200_sunz_correct(const float *data, float *out, int32_t nbands, int64_t band_stride,
                      const b200_geo_t *geo, const b200_sunz_t *sz, void *stream);

/* Same, with the solar zenith angle given in degrees (optional_datasets branch,
 * modifiers/geometry.py:64-66); sza is float32, one image shared by all bands. */
int b200_sunz_correct_sza(const float *data, const float *sza_deg, float *out, int32_t nbands,
                          int64_t band_stride, int64_t n, const b200_sunz_t *sz, void *stream);

/* float64 data (the reference's tests run both dtypes, tests/test_modifiers.py:116-176): lon/lat are then
 * NOT rounded to float32 (angles.py:427-429 casts them to the data dtype) and the product is float64. */
int b200_sunz_correct_f64(const double *data, double *out, int32_t nbands, int64_t band_stride,
                          const b200_geo_t *geo, const b200_sunz_t *sz, void *stream);
int b200_sunz_correct_sza_f64(const double *data, const double *sza_deg, double *out, int32_t nbands,
                              int64_t band_stride, int64_t n, const b200_sunz_t *sz, void *stream);

/* get_angles (modifiers/angles.py:377-402,443-531 + pyorbital get_observer_look / get_alt_az): sensor azimuth
 * and zenith, solar azimuth (0..360) and zenith, degrees, float64, NaN where the area has no lon/lat.  Any output
 * may be NULL.  sz supplies gmst / ra / dec of start_time; the satellite position is the one get_satpos returns
 * (longitude, latitude in degrees, altitude in km). */
int b200_get_angles(const b200_geo_t *geo, const b200_sunz_t *sz, double sat_lon, double sat_lat, double sat_alt_km,
                    double *sata, double *satz, double *suna, double *sunz, void *stream);
/* compute_relative_azimuth (modifiers/angles.py:336-374): min(|sun - sat|, 360 - |sun - sat|). */
int b200_relative_azimuth(const void *sat_azi, const void *sun_azi, void *out, int64_t n, int32_t is_f32,
                          void *stream);

/* SunZenithReducer (modifiers/geometry.py:166-207 -> angles.py:594-624): out = data * reduction(sza). */
int b200_sunz_reduce(const float *data, const float *sunz_deg, float *out, int32_t nbands, int64_t band_stride,
                     int64_t n, float limit_deg, float max_sza_deg, float strength, void *stream);
/* SunZenithReducer when no solar-zenith dataset is given (geometry.py:166-207): the angle is derived from the area in
 * the same launch -- cos(sza) in the reference's flow (b200_cos_sza), masked beyond sz->max_sza_deg (geometry.py:62-63),
 * sza = rad2deg(arccos(.)) rounded to float32 (geometry.py:201), then b200_sunz_reduce's factor. */
int b200_sunz_reduce_area(const float *data, float *out, int32_t nbands, int64_t band_stride, const b200_geo_t *geo,
                          const b200_sunz_t *sz, float limit_deg, float max_sza_deg, float strength, void *stream);

/* Fused reader calibration + SunZenithCorrector for up to 4 ABI reflective bands
 * sharing one area: counts -> radiance -> reflectance[%] -> * corr.  6 B/pixel/band. */
int b200_abi_vis_sunz(const int16_t *const *counts /* host array of device ptrs */,
                      float *const *out /* host array of device ptrs */,
                      const b200_abi_cal_t *cal /* host, nbands entries */, int32_t nbands,
                      const b200_geo_t *geo, const b200_sunz_t *sz, void *stream);

/* ---------------------------------------------------------------- nearest
 * Replaces pyresample.kd_tree.XArrayResamplerNN.get_neighbour_info (kd-tree
 * build + query; called from resample/kdtree.py:60-95) with a uniform
 * grid-bucket search in spherical ECEF (R = 6370997 m), and
 * get_sample_from_neighbour_info (resample/kdtree.py:184-190) with a gather.
 */
typedef struct b200_nn_grid b200_nn_grid_t; /* opaque, owned by the library */

/* Build the search structure over the valid source points (bucket grid: cells of a few points
 * each, sized from a sample of the source, independent of `radius`; fixed grid: the source
 * lattice itself).
 *   src            source geometry; the grid covers its row window [row0, row0+nrows) (the
 *                  whole grid normally; a band of rows when the target is row-tiled across GPUs,
 *                  cf. Scene._reduce_data, scene.py:930-954).  Source pixel indices and the
 *                  mask / valid_input arrays are relative to the first pixel of that window.
 *   mask           optional uint8 (1 = exclude), window pixels; may be NULL
 *   radius         radius_of_influence [m]
 *   valid_input    out, uint8 src pixels (NN_COORDINATES "valid_input_index")
 *   n_valid        out (host): number of valid source points
 * Synchronises `stream` once (sizes come back to the host). */
int b200_nn_grid_create(const b200_geo_t *src, const uint8_t *mask, double radius,
                        uint8_t *valid_input, int64_t *n_valid /* host */,
                        b200_nn_grid_t **grid /* host out */, void *stream);
/* Same, choosing the search engine: 0 = automatic, 1 = bucket grid (any source), 2 = fixed grid
 * (geostationary area sources only: the source lattice itself is the index, see csrc/nn_geos.cu;
 * B200_EUNSUPPORTED for other sources), 3 = fixed grid in strict mode (every lattice point of the
 * rigorous search disc is examined by the general loop; 2 settles radii of a few source pixels with
 * fixed 2x2 / 4x4 stencils and a satellite-distance-corrected bound first).
 * All engines return identical indices.
 * `n_valid` may be NULL: the fixed-grid engine then builds without synchronising the stream (the
 * count is read back by b200_nn_grid_n_valid when somebody asks); the bucket engine always needs it.
 * B200_NN_LIGHT may be OR-ed into `method`: when the fixed-grid engine serves the source, no mask
 * is given and n_valid is NULL, the structure keeps float32 points only (16 instead of 52 bytes per
 * source pixel; the rare exact float64 decisions recompute their candidates from the projection).
 * Such a structure serves b200_nn_query_gather_f32 and b200_nn_query with gather_index /
 * valid_output; b200_nn_query(index_array), b200_nn_grid_n_valid and b200_bil_info return
 * B200_EUNSUPPORTED for it.  The flag is ignored where the conditions do not hold. */
#define B200_NN_LIGHT 0x100
int b200_nn_grid_create_ex(const b200_geo_t *src, const uint8_t *mask, double radius,
                           uint8_t *valid_input, int64_t *n_valid /* host, may be NULL */,
                           b200_nn_grid_t **grid /* host out */, int32_t method, void *stream);
/* engine actually used by a grid (1 or 2) */
int b200_nn_grid_method(const b200_nn_grid_t *grid, int32_t *method /* host */);
/* number of valid source points of a grid (synchronises the grid's stream if it was built with
 * n_valid == NULL).  `valid_input` of the create call must still be alive then.  Replaces the
 * implicit `valid_input_index.sum()` of pyresample's get_neighbour_info (kdtree.py:60-95). */
int b200_nn_grid_n_valid(b200_nn_grid_t *grid, int64_t *n_valid /* host */);
int b200_nn_grid_destroy(b200_nn_grid_t *grid);
/* bytes of device memory held by the grid */
int b200_nn_grid_nbytes(const b200_nn_grid_t *grid, int64_t *nbytes /* host */);

/* Query k=1 for the target rows [dst->row0, +nrows).
 *   index_array    out int32, compressed valid-input index, -1 = no neighbour
 *                  (NN_COORDINATES "index_array", z2 = 1); may be NULL
 *   gather_index   out int32, RAW flat source pixel index, -1 = none; may be NULL
 *   valid_output   out uint8 ("valid_output_index"); may be NULL */
int b200_nn_query(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius,
                  int32_t *index_array, int32_t *gather_index, uint8_t *valid_output,
                  void *stream);

/* out[b][i] = gather_index[i] < 0 ? fill : data[b][gather_index[i]]
 * elem_size in {1,2,4,8}; fill points at one host element of that size. */
int b200_nn_gather(const void *data, void *out, const int32_t *gather_index, int64_t n_out,
                   int32_t nbands, int64_t src_band_stride, int64_t dst_band_stride,
                   int32_t elem_size, const void *fill /* host */, void *stream);
/* The same with the kernel variant chosen by the caller: 0 = direct (128-bit index loads and stores from the threads),
 * 1 = index stream in and result out through shared memory with cp.async.bulk + mbarrier (4-byte elements, one band;
 * anything else takes variant 0).  b200_nn_gather uses B200_GATHER_VARIANT from the environment, else the measured
 * winner.  Both give identical results. */
int b200_nn_gather_variant(const void *data, void *out, const int32_t *gather_index, int64_t n_out, int32_t nbands,
                           int64_t src_band_stride, int64_t dst_band_stride, int32_t elem_size,
                           const void *fill_value, int32_t variant, void *stream);

/* Fused query + gather (no index materialised): float32 data only. */
int b200_nn_query_gather_f32(const b200_nn_grid_t *grid, const b200_geo_t *dst, double radius,
                             const float *data, float *out, int32_t nbands,
                             int64_t src_band_stride, int64_t dst_band_stride, float fill,
                             void *stream);

/* compressed index -> raw flat index (and back) helpers for cached index arrays */
int b200_nn_compressed_to_raw(const uint8_t *valid_input, int64_t n_src, const int32_t *index_array,
                              int32_t *gather_index, int64_t n_out, void *stream);

/* ------------------------------------------------------- native resampler, SimulatedGreen
 * NativeResampler (resample/native.py:41-159): repeat every pixel fy x fx times (any dtype of elem_size
 * bytes, nbands leading images) or np.nanmean over fy x fx blocks (2-D float32).
 * SimulatedGreen (composites/abi.py:56-61): c01 * f1 + c02 * f2 + c03 * f3 in float32. */
int b200_native_replicate(const void *data, void *out, int64_t nbands, int64_t rows, int64_t cols, int32_t fy,
                          int32_t fx, int32_t elem_size, void *stream);
int b200_native_aggregate_f32(const float *data, float *out, int64_t out_rows, int64_t out_cols, int32_t fy,
                              int32_t fx, void *stream);
int b200_simulated_green(const float *c01, const float *c02, const float *c03, float *out, int64_t n, float f1,
                         float f2, float f3, void *stream);

/* RatioSharpenedRGB / SelfSharpenedRGB (composites/resolution.py:32-233) with GenericCompositor's common channel
 * mask (composites/core.py:489-490): out (3 x H x W) from three bands of one shape.
 *   high        the high-resolution band (NULL: stack + mask only, "high_resolution_band: None")
 *   hi          colour (0 red, 1 green, 2 blue) `high` stands for; out[hi] = high
 *   neutral     colour left untouched (-1: none)
 *   self_mean   1: ratio = high / four_element_average(high) (SelfSharpenedRGB, _mean4 :168-192, blocks anchored at
 *               the parities row_off / col_off of the area's crop offset); 0: ratio = high / band[hi]
 *   ratio       1 where not finite or negative, clipped to 1.5 (:157-165) */
int b200_ratio_sharpen_f32(const float *red, const float *green, const float *blue, const float *high, float *out,
                           int64_t H, int64_t W, int32_t hi, int32_t neutral, int32_t self_mean, int32_t row_off,
                           int32_t col_off, int32_t common_mask, void *stream);
int b200_ratio_sharpen_f64(const double *red, const double *green, const double *blue, const double *high, double *out,
                           int64_t H, int64_t W, int32_t hi, int32_t neutral, int32_t self_mean, int32_t row_off,
                           int32_t col_off, int32_t common_mask, void *stream);

/* ---------------------------------------------------------------- assembling a row-tiled image
 * Multi-GPU decomposition of the target (the reference's analogue is dask chunking of the
 * target, satpy/resample/base.py:32, which never replicates fill values either): the column span of a row block
 * that the source can reach follows from the geometry (satpy_b200/parallel.py::block_column_span), so each rank
 * writes the fill value outside it locally (b200_fill_f32_2d) and only the span crosses NVLink (b200_copy_2d:
 * cudaMemcpy2DAsync on the copy engines; `dst` may be a peer GPU's image mapped into this process).
 *   pitch / *_pitch_bytes   distance between consecutive rows (elements / bytes) */
int b200_fill_f32_2d(float *dst, int64_t pitch, int64_t width, int64_t height, float value, void *stream);
/* the same for `n` rectangles {row0, nrows, col0, ncols} (host array of 4 n int64) of one image, 64 per launch */
int b200_fill_f32_rects(float *img, int64_t pitch, const int64_t *rects, int32_t n, float value, void *stream);
/* HOST helper: `value` into a width x height rectangle of a (page-locked) float32 host image with non-temporal stores;
 * runs on the calling thread (callers split the rows over their own threads).  No device work. */
int b200_host_fill_f32(float *dst, int64_t pitch, int64_t width, int64_t height, float value);
int b200_copy_2d(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes,
                 int64_t width_bytes, int64_t height, void *stream);

/* ---------------------------------------------------------------- bilinear
 * Replaces pyresample XArrayBilinearResampler.get_bil_info / get_sample_from_bil_info as driven by
 * BilinearResampler.precompute / .compute (resample/kdtree.py:216-242, 281-293).  PARITY UNPINNED in the
 * reference (mocked in its tests); follows the library's published algorithm (oracle/bilinear.py).
 *   grid          a FIXED-GRID search structure (geostationary area source); B200_EUNSUPPORTED otherwise
 *   valid_input   the uint8 array b200_nn_grid_create filled for that structure
 *   dst           target AREA (its projection is the interpolation plane), rows [row0, +nrows)
 *   neighbours    32 in satpy (kdtree.py:231)
 *   index_array   out int32 [n][4] compressed indices of the UL, UR, LL, LR corner; may be NULL
 *   gather_index  out int32 [n][4] raw source indices (-1: no valid source at all); may be NULL
 *   t, s          out float64 [n] vertical / horizontal fractional distances, NaN = no interpolation possible
 * Synchronises `stream` once. */
int b200_bil_info(const b200_nn_grid_t *grid, const uint8_t *valid_input, const b200_geo_t *dst, double radius,
                  int32_t neighbours, int32_t *index_array, int32_t *gather_index, double *t, double *s,
                  void *stream);

/* out = p1 (1-s)(1-t) + p2 s (1-t) + p3 (1-s) t + p4 s t, NaN -> fill; float32 data, float64 arithmetic. */
int b200_bil_sample_f32(const float *data, float *out, const int32_t *gather_index, const double *t,
                        const double *s, int64_t n_out, int32_t nbands, int64_t src_band_stride,
                        int64_t dst_band_stride, float fill, void *stream);

/* --------------------------------------------------------------------- EWA
 * Replaces pyresample.ewa ll2cr / fornav (Cython + C++), exposed by satpy as resampler "ewa"
 * (resample/ewa.py:3-11).  PARITY UNPINNED in the reference (no test exercises it); follows the
 * library's published algorithm (oracle/ewa.py).  Defaults are pyresample's.
 */
typedef struct b200_ewa_params {
    int32_t weight_count;        /* 10000 */
    int32_t maximum_weight_mode; /* 0 */
    float weight_min;            /* 0.01 */
    float weight_distance_max;   /* 1.0  */
    float weight_delta_max;      /* 10.0 */
    float weight_sum_min;        /* -1.0: use weight_min */
} b200_ewa_params_t;

/* Swath lon/lat -> fractional (col, row) of the target AREA (col 0.0 / row 0.0 = centre of the top-left
 * cell), NaN where not projectable.  cols/rows have the dtype of the swath's lon/lat arrays (float32 or
 * float64; float64 for area sources).  points_in_grid (host, may be NULL): points with
 * -1 <= col <= width+1 and -1 <= row <= height+1; passing it synchronises `stream`. */
int b200_ll2cr(const b200_geo_t *swath, const b200_geo_t *dst, void *cols, void *rows,
               int64_t *points_in_grid /* host */, void *stream);

/* Add the weighted samples of float32 `data` (nbands images of swath_rows x swath_cols) to the
 * accumulation grids grid_weights / grid_accums ([nbands][grid_rows * grid_cols] float32, zeroed by the
 * caller before the first call; several calls -- e.g. one per scan chunk or per GPU -- accumulate).
 * rows_per_scan <= 0 means the whole swath is one scan.  img_fill: input value to skip (NaN always is).
 * maximum_weight_mode 1 keeps, per cell, the sample with the largest weight: all samples must come through ONE
 * call.  To split the samples over several calls / ranks: mode 2 (only the maximum of the weights, into
 * grid_weights), take the element-wise maximum of the weight grids, then mode 3 with that grid (writes the samples
 * whose weight equals the maximum into grid_accums, NaN for an invalid sample) and combine the value grids -- see
 * satpy_b200/resample/ewa.py::B200EWAResampler.compute(group=...). */
int b200_fornav_accum(const void *cols, const void *rows, int32_t cr_is_f32, int64_t swath_rows,
                      int64_t swath_cols, int32_t rows_per_scan, const float *data, int32_t nbands,
                      int64_t band_stride, float img_fill, int64_t grid_rows, int64_t grid_cols,
                      float *grid_weights, float *grid_accums, const b200_ewa_params_t *params, void *stream);

/* out = sum(w v) / sum(w) where sum(w) >= weight_sum_min, else fill.  valid_counts (host, nbands entries,
 * may be NULL): number of non-fill cells per band; passing it synchronises `stream`. */
int b200_fornav_finalize(const float *grid_weights, const float *grid_accums, float *out, int32_t nbands,
                         int64_t grid_size, const b200_ewa_params_t *params, float fill,
                         int64_t *valid_counts /* host */, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SATPY_B200_H */
