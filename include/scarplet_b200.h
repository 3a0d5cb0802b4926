/* scarplet_b200 — C ABI of the B200-native template-matching hot path.
 *
 * The reference (stgl/scarplet 0.1.4) is pure Python and has no FFI; its boundary for
 * this path is the Python API of scarplet/core.py and the WindowedTemplate plugin
 * duck type.  Each entry point below replaces the reference interface cited beside
 * it; scarplet_b200/core.py (the host-side mirror) binds them with ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a non-zero status and leaves a
 * message retrievable with sb_last_error().  A plan is bound to one CUDA device and
 * is used from one host thread at a time; work is stream-ordered on the plan's
 * stream.  `*_host` pointers are host memory (copied inside the call), `*_dev`
 * pointers are device memory on the plan's device.  Inputs are borrowed and never
 * modified; outputs are caller-allocated.
 */
#ifndef SCARPLET_B200_H
#define SCARPLET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_plan sb_plan;

/* One search orientation: the direction of the curvature
 * (dem.py:68-107 `_calculate_directional_laplacian(alpha)`).  The host passes the
 * float64 values NumPy computes so the device reproduces them bit for bit. */
typedef struct sb_angle {
    double cos_a;   /* np.cos(angle)        */
    double sin_a;   /* np.sin(angle)        */
    double cos2_a;  /* np.cos(angle) ** 2   */
    double sin2_a;  /* np.sin(angle) ** 2   */
} sb_angle;

enum { SB_KIND_SCARP = 0, SB_KIND_RICKER = 1,
       SB_KIND_RASTER = 2 /* values supplied by the caller: sb_match_template_raster */ };
enum { SB_ERRMASK_NONE = 0, SB_ERRMASK_XR_LE0 = 1, SB_ERRMASK_XR_GE0 = 2 };

/* One (scale, age, angle) template: what `Template(scale, age, angle, nx, ny, de)`
 * (core.py:345) plus `.template()`, `.get_window_limits()`, `.get_err_mask()`
 * (core.py:346, 369-375) determine.  Everything that must be float64-exact is
 * precomputed by the host. */
typedef struct sb_template {
    double cos_t, sin_t;  /* cos/sin of the template's alpha = -angle (WindowedTemplate.py:151) */
    double c, d;          /* window half-widths (WindowedTemplate.py:63)                     */
    double k0, k1;        /* scarp: 2*kt**1.5*sqrt(pi), 4*kt;  ricker: pi*f, 0               */
    double sign;          /* -1 for RightFacingUpperBreakScarp (WindowedTemplate.py:254)    */
    double tscale;        /* power of two ~ 1/rms(t): balances t against M in the packed FFT  */
    int32_t kind;         /* SB_KIND_*                                                       */
    int32_t errmode;      /* SB_ERRMASK_* (WindowedTemplate.py:257-267, 294-304)             */
    int32_t sy_lo, sy_hi, sx_lo, sx_hi; /* conservative support box, offsets from (ny//2, nx//2) */
    int32_t i_lo, i_hi, j_lo, j_hi;     /* rows/cols NOT masked by get_window_limits (inclusive) */
    int32_t angle_id;     /* index into the sb_angle array                                  */
    int32_t idx;          /* flat result index: tie priority and key into age_of/angle_of   */
    int32_t state;        /* best state the template folds into (0 .. option "states" - 1): one per
                             template scale in a multi-scale search (CHANGELOG.md:20-24)    */
    int32_t reserved;     /* 0 */
} sb_template;

/* flags for sb_plan_create */
#define SB_PLAN_DEFAULT 0u

const char* sb_last_error(void);
/* library / build information: "cuda sm_100a <hash of the sources it was compiled from>" for the product build */
const char* sb_build_info(void);

/* Plan = raster geometry + device workspace.  dx2/dy2 are `dx ** 2`, `dy ** 2` as the
 * host evaluates them (dem.py:95,99).  device < 0 keeps the current device.
 * stream is a cudaStream_t (0 = the plan creates its own). */
int sb_plan_create(sb_plan** plan, int ny, int nx, double dx, double dx2, double dy2,
                   int device, void* stream, unsigned flags);
int sb_plan_destroy(sb_plan* plan);

/* knobs: key in {"workspace_mb", "max_fft", "force_pad", "mixed_tiles", "profile",
 * "precision" (32 = complex64 pipeline, default; 64 = complex128 pipeline),
 * "states" (number of independent best states, default 1; resets them),
 * and the kernel-selection switches, all on by default (off = the slower variant, same results
 * to rounding): "fast" (pipelined complex64 kernels), "conv_persist" (persistent column kernel:
 * minimum templates per orientation run, 0 = never), "conv_r64" (radix-64 column kernel at
 * Py = 4096), "lincomb" (curvature spectra as combinations of nine plane spectra),
 * "fit_substreams" (0 = automatic, 1..8 = sub-streams of the row kernel on small rasters)};
 * returns 0 if known */
int sb_plan_set_option(sb_plan* plan, const char* key, long value);

/* Spatial sharding of one raster over several GPUs (BASELINE config 5; the reference's
 * analogue is file-level tiling, dem.py:249-278).  The plan keeps the geometry of the WHOLE raster
 * (template centring, edge masks and the circular wrap-around of core.py:359 are evaluated in
 * full-raster coordinates) but computes only raster rows row_lo .. row_hi - 1 and holds only the
 * DEM rows row_lo - halo .. row_hi + halo - 1 (periodic in ny).  Call before sb_set_dem_*; the DEM
 * handed to sb_set_dem_* is then that slab (sb_plan_dem_rows: first raster row, row count).
 * halo >= template support along y + 2.  The best state covers the plan's rows only. */
int sb_plan_set_slab(sb_plan* plan, int row_lo, int row_hi, int halo);
int sb_plan_dem_rows(const sb_plan* plan, int* row0, int* drows);
/* curvature statistics of the plan's own rows (sum of dxx^2 + dyy^2, pixel count): a non-finite
 * sum marks a NaN in the DEM.  Slabs of one raster add theirs up and hand the totals back, so that
 * every rank packs with the same scale and knows about a NaN held by another (dem.py:105). */
int sb_plan_curv_stats(const sb_plan* plan, double* sumsq, double* count);
int sb_plan_set_curv_stats(sb_plan* plan, double sumsq, double count);
/* the cudaStream_t the plan's work is ordered on */
void* sb_plan_stream(const sb_plan* plan);
/* device memory held by the plan (bytes) */
long sb_plan_device_bytes(const sb_plan* plan);
/* sum over the FFT tiles of the last sweep of Py * Px (tiling overhead = this / (rows * nx)) */
double sb_plan_last_fft_area(const sb_plan* plan);
/* per-kernel device time (CUDA events on the plan's stream) accumulated while the option
 * "profile" is 1: ms[6], launches[6] in the order k_curv_rows, k_curv_cols, k_tmpl_rows,
 * k_tmpl_sums, k_conv_cols, k_fit_rows.  reset != 0 clears the counters afterwards. */
int sb_plan_profile(sb_plan* plan, double* ms6, long* launches6, int reset);
/* number of kernels launched by this plan so far */
long sb_plan_launch_count(const sb_plan* plan);
/* FFT domain and tile grid chosen by the last sweep: out[0..5] = largest Py, largest Px, tiles_y, tiles_x,
 * angle batch, template batch (tiles of different lengths are mixed along an axis when that wastes less) */
int sb_plan_last_geometry(const sb_plan* plan, int* out6);

/* DEM upload (DEMGrid._griddata, float64 row-major ny x nx) and the centred axis
 * vectors x[nx], y[ny] of WindowedTemplate.py:50-53. */
int sb_set_dem_host(sb_plan* plan, const double* dem_host);
int sb_set_dem_dev(sb_plan* plan, const double* dem_dev);
int sb_set_axes_host(sb_plan* plan, const double* x_host, const double* y_host);

/* DEMGrid._calculate_directional_laplacian (dem.py:68-107): float64 out[ny*nx]. */
int sb_directional_laplacian(sb_plan* plan, const sb_angle* angle, double* out, int out_is_device);

/* WindowedTemplate.template() (WindowedTemplate.py:159-183, 497-520): float64 out[ny*nx]. */
int sb_render_template(sb_plan* plan, const sb_template* tmpl, double* out, int out_is_device);

/* core.match_template (core.py:297-377): amp[ny*nx], snr[ny*nx] float64 for one template. */
int sb_match_template(sb_plan* plan, const sb_angle* angle, const sb_template* tmpl,
                      double* amp, double* snr, int out_is_device);

/* core.match_template for a template the library has no generator for -- the plugin surface
 * of core.py:345-348: any class with `.template()`.  The caller renders the template on the
 * host and passes its values on the box of rows sy_lo..sy_hi x columns sx_lo..sx_hi (offsets
 * from (ny//2, nx//2), row-major float64) that holds every non-zero; outside the box the
 * template is zero.  amp / snr are the raw planes of core.py:360-367: `get_err_mask` /
 * `get_window_limits` (core.py:369-375) are the caller's to apply.  tscale: a power of two
 * near 1 / rms(template) (0 = 1). */
int sb_match_template_raster(sb_plan* plan, const sb_angle* angle, const double* box_host,
                             int sy_lo, int sy_hi, int sx_lo, int sx_hi, double tscale,
                             double* amp, double* snr, int out_is_device);

/* Best-fit state (running result of core.compare over a sweep). */
int sb_best_reset(sb_plan* plan);

/* The fan-out of core.calculate_best_fit_parameters / match (core.py:139-195, 266-294)
 * and the fold of core.compare (core.py:198-243) for a list of templates; accumulates
 * into the plan's best state (call sb_best_reset first for a fresh search). */
int sb_sweep(sb_plan* plan, const sb_angle* angles, int n_angles,
             const sb_template* tmpls, int n_tmpls);

/* The same for a SHARE of a larger search (one rank's orientations): whole_search5 =
 * {sy_lo, sy_hi, sx_lo, sx_hi of the union of all support boxes, templates per orientation} of
 * the undivided search.  FFT domains, tiles and kernel variants are then chosen exactly as the
 * undivided search chooses them, which makes the shares' results bit-identical to its. */
int sb_sweep_ex(sb_plan* plan, const sb_angle* angles, int n_angles, const sb_template* tmpls, int n_tmpls,
                const int32_t* whole_search5);

/* Decode the best state to the reference's stack order [amp, age, angle, snr]
 * (core.py:190-193): out4 is float64[4*ny*nx]; age_of/angle_of are host float64
 * tables indexed by sb_template.idx (n_idx entries). */
int sb_finalize(sb_plan* plan, const double* age_of_host, const double* angle_of_host, int n_idx,
                double* out4, int out_is_device);

/* The same for best state `state` and raster rows row_lo .. row_hi - 1 only:
 * out4 is float64[4 * (row_hi - row_lo) * nx]. */
int sb_finalize_ex(sb_plan* plan, int state, int row_lo, int row_hi, const double* age_of_host,
                   const double* angle_of_host, int n_idx, double* out4, int out_is_device);

/* Raw best state for a cross-GPU merge: device pointers into the plan, one value per pixel of
 * the plan's rows. */
int sb_best_state(sb_plan* plan, float** snr_dev, float** amp_dev, int32_t** idx_dev);
int sb_best_state_ex(sb_plan* plan, int state, float** snr_dev, float** amp_dev, int32_t** idx_dev);

/* Fold n_cands candidate best states for raster rows row_lo .. row_hi - 1 (device buffers laid out
 * [candidate][row][nx], e.g. what an all-to-all of the ranks' best states delivers) into best
 * state `state`: highest SNR, then lowest flat index; a NaN SNR sticks (core.py:230-240). */
int sb_best_merge(sb_plan* plan, int state, int row_lo, int row_hi, int n_cands, const float* snr_c,
                  const float* amp_c, const int32_t* idx_c);

/* Cross-GPU merge of best states (the parent-side reduce over Pool results,
 * core.py:185, when the search is sharded over ranks).  All buffers are device memory
 * of ny*nx elements owned by the caller (e.g. torch tensors handed to NCCL):
 *   1. sb_best_pack    keys = (SNR bits << 32) | (0xFFFFFFFF - idx)  -> all-reduce MAX (int64)
 *   2. sb_best_select  amp_out = amp where this rank owns the winning key, else 0 -> all-reduce SUM
 *   3. sb_best_unpack  overwrite the plan's best state with the merged result */
int sb_best_pack(sb_plan* plan, unsigned long long* keys_dev);
int sb_best_select(sb_plan* plan, const unsigned long long* gkeys_dev, float* amp_out_dev);
int sb_best_unpack(sb_plan* plan, const unsigned long long* gkeys_dev, const float* amp_dev);

/* core.compare (core.py:198-243) with the reference's exact strict-compare semantics on
 * float64 host planes: folds (amp, age, angle, snr) into best[4][n].  age/angle may be
 * NULL, in which case the scalars age_s / angle_s are used. */
int sb_compare_host(sb_plan* plan, double* best4_host, const double* amp, const double* age,
                    const double* angle, const double* snr, double age_s, double angle_s);

/* DEMGrid._estimate_curvature_noiselevel (dem.py:152-179): sums, over the pixels whose Gaussian
 * window (sigma, truncate as scipy.ndimage.gaussian_filter) holds no NaN, of the high-passed second
 * differences h = d - lowpass(d) and their products -- out10 = [count, hxx, hxy, hyy, hxx^2, hxy^2,
 * hyy^2, hxx hxy, hxx hyy, hxy hyy].  The directional curvature is linear in (dxx, dxy, dyy)
 * (dem.py:103-104), so mean and standard deviation for every direction follow from these. */
int sb_curvature_noise_moments(sb_plan* plan, double sigma, double truncate, double* out10_host);

/* One pass of nodata filling in place of rasterio.fill.fillnodata (dem.py:406-408): NaN cells of
 * the host raster (ny x nx float64) become the inverse-distance-weighted mean of the nearest valid
 * cell along the eight row / column / diagonal rays within max_search_distance cells;
 * *remaining = cells still NaN. */
int sb_fill_nodata(sb_plan* plan, double* dem_host_inout, double max_search_distance, long* remaining);

/* unit-test hook: batched complex64 FFT of length n (power of two, 128..8192) over rows;
 * inverse: bit 0 = inverse transform, bit 1 = use the radix-64 core (n = 4096 only) */
int sb_debug_fft(sb_plan* plan, int n, int rows, const float* in_host, float* out_host, int inverse);

/* developer hook: device time (ms per launch, CUDA events) of the batched FFT kernel alone on
 * resident data; radix64 != 0 selects the 64 x 64 core (n = 4096) */
int sb_debug_fft_bench(sb_plan* plan, int n, int rows, int reps, int radix64, float* ms_per_launch);

/* synchronise the plan's stream */
int sb_sync(sb_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* SCARPLET_B200_H */
