// Host side of the C ABI declared in include/scarplet_b200.h: plan, workspace,
// tile / batch scheduling and kernel launches.  No torch types, no CPU compute path.
#include "../../include/scarplet_b200.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "sb_kernels.cuh"
#include "sb_fast.cuh"

static_assert(sizeof(sb_template) == sizeof(sb::Tmpl), "sb_template / sb::Tmpl layout");
static_assert(sizeof(sb_angle) == sizeof(sb::Angle), "sb_angle / sb::Angle layout");
static_assert(offsetof(sb_template, kind) == offsetof(sb::Tmpl, kind), "layout");
static_assert(offsetof(sb_template, idx) == offsetof(sb::Tmpl, idx), "layout");
static_assert(offsetof(sb_template, i_lo) == offsetof(sb::Tmpl, i_lo), "layout");
static_assert(offsetof(sb_template, tscale) == offsetof(sb::Tmpl, tscale), "layout");

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#define SB_TRY(expr)                                                                      \
    do {                                                                                  \
        int _e = (expr);                                                                  \
        if (_e != 0)                                                                      \
            return fail(std::string(#expr) + ": " + sb_rt_error_string(_e) + " (" +       \
                        std::to_string(_e) + ")");                                        \
    } while (0)

#define SB_OK(expr)                     \
    do {                                \
        int _s = (expr);                \
        if (_s != 0) return _s;         \
    } while (0)

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
int next_pow2(long n) {
    long p = 1;
    while (p < n) p <<= 1;
    return (int)p;
}
constexpr int kMinFft = 128;
constexpr int kMaxFftSupported = 8192;

struct AxisPlan {
    int P = 0;          // FFT length
    int tiles = 1;
    int tile_out = 0;   // output pixels per tile (last tile may be shorter)
    int lo = 0, hi = 0; // support offsets
    bool periodic = false;
};

}  // namespace

struct sb_plan {
    int ny = 0, nx = 0;
    double dx = 1, dx2 = 1, dy2 = 1;
    int device = 0;
    sb_stream_t stream = 0;
    bool own_stream = false;
    double* d_dem = nullptr;
    bool own_dem = false;
    double* d_diffs = nullptr;      // dxx, dxy, dyy planes of the current DEM (dem.py:88-99)
    float4* d_diffs32 = nullptr;    // the same, float32, interleaved per pixel (complex64 pipeline)
    double* d_x = nullptr;
    double* d_y = nullptr;
    float* d_bsnr = nullptr;
    float* d_bamp = nullptr;
    int* d_bidx = nullptr;
    std::map<long, void*> tw;      // twiddle tables keyed by 2 * n + (float64 ? 1 : 0)
    int precision = 32;            // 32: complex64 pipeline, 64: complex128 pipeline
    Buf cr, fct, trt, part, gbuf, sums, fit, tmpls, angles, tables, raw, tbox;
    int fast = 1;                  // 1: pipelined complex64 kernels (sb_fast.cuh), 0: simple kernels
    // persistent column kernel: 1 always, 0 never, -1 (default) when a search angle carries at least
    // four templates -- with fewer, half of its thread groups idle and the per-angle staging of
    // the spectrum columns is not amortised (C2: 94 ms persistent, 69 ms per-template)
    int conv_persist = std::getenv("SB_CONV_P") ? std::atoi(std::getenv("SB_CONV_P")) : -1;
    int fit_threads = std::getenv("SB_FIT_THREADS") ? std::atoi(std::getenv("SB_FIT_THREADS")) : 0;
    long launches = 0;
    double c2_scale = 1.0;
    bool dem_nonfinite = false;    // the DEM holds a NaN / Inf: every FFT domain is poisoned like the reference's
    int profile = 0;
    std::vector<sb_event_t> ev_pool;
    struct EvPair { int kind; sb_event_t a, b; };
    std::vector<EvPair> ev_live;
    size_t ev_used = 0;
    double prof_ms[6] = {0, 0, 0, 0, 0, 0};
    long prof_n[6] = {0, 0, 0, 0, 0, 0};
    long workspace_mb = 0;         // 0: half of the free device memory, at most 48 GB
    int max_fft = kMaxFftSupported;
    int force_pad = 0;
    int last_geom[6] = {0, 0, 0, 0, 0, 0};
};

namespace {

int ensure(Buf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) sb_rt_free(b.p);
    b.p = nullptr;
    b.cap = 0;
    SB_TRY(sb_rt_malloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

void release(Buf& b) {
    if (b.p) sb_rt_free(b.p);
    b.p = nullptr;
    b.cap = 0;
}

template <typename R>
int twiddles(sb_plan* pl, int n, const typename Vec<R>::v2** out) {
    typedef typename Vec<R>::v2 C2;
    const long key = 2L * n + (sizeof(R) == 8 ? 1 : 0);
    auto it = pl->tw.find(key);
    if (it != pl->tw.end()) {
        *out = (const C2*)it->second;
        return 0;
    }
    int count = 0;
    {   // same stage walk as sbfft::fill_twiddles
        int ns = 1;
        for (int s = 0; n > ns; ++s) {
            int rest = n / ns, r = rest >= 16 ? 16 : rest;
            if (s > 0) count += (r - 1) * ns;
            ns *= r;
        }
    }
    std::vector<C2> host(std::max(count, 1));
    sbfft::fill_twiddles<R>(n, host.data());
    void* d = nullptr;
    SB_TRY(sb_rt_malloc(&d, host.size() * sizeof(C2)));
    SB_TRY(sb_rt_h2d(d, host.data(), host.size() * sizeof(C2), pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    pl->tw[key] = d;
    *out = (const C2*)d;
    return 0;
}

template <typename F>
int dispatch_n(int n, F&& f) {
#ifdef SB_DEV_N   // developer builds (scratch/devbuild.sh): one FFT length, a tenth of the compile time
    if (n == SB_DEV_N) return f(std::integral_constant<int, SB_DEV_N>{});
    return fail("developer build: only FFT length " + std::to_string(SB_DEV_N));
#else
    switch (n) {
        case 128: return f(std::integral_constant<int, 128>{});
        case 256: return f(std::integral_constant<int, 256>{});
        case 512: return f(std::integral_constant<int, 512>{});
        case 1024: return f(std::integral_constant<int, 1024>{});
        case 2048: return f(std::integral_constant<int, 2048>{});
        case 4096: return f(std::integral_constant<int, 4096>{});
        case 8192: return f(std::integral_constant<int, 8192>{});
        default: return fail("unsupported FFT length " + std::to_string(n));
    }
#endif
}

template <int N, typename R>
struct Shape {
    static constexpr int T = N / sbfft::E;
    static constexpr int threads = T > 256 ? T : 256;
    static constexpr int GP = threads / T;
    static constexpr size_t smem = (size_t)GP * sbfft::padded_len(N) * sizeof(typename Vec<R>::v2);
    // k_conv_cols adds a park buffer of N elements per group
    static constexpr size_t smem_conv = (size_t)GP * (sbfft::padded_len(N) + N) * sizeof(typename Vec<R>::v2);
    // pipelined kernels (sb_fast.cuh): two exchange buffers per group; k_fit_rows_f adds the
    // batch's scalars, the active-template list (+ its length) and the flags
    static constexpr size_t smem_conv_f = (size_t)GP * 2 * sbfft::padded_len(N) * sizeof(float2);
    static constexpr size_t smem_fit_f = smem_conv_f + sb::kFitMaxBatch * sizeof(sb::FitT) +
                                         (2 * sb::kFitMaxBatch + 2) * sizeof(int) +
                                         (size_t)sbfft::ctw_count<float>(N) * sizeof(float2);
};

#ifndef SB_EMU
template <typename K>
int allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        SB_TRY((int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}
#define SB_ALLOW_SMEM(kern, bytes) SB_OK(allow_smem(kern, bytes))
#else
#define SB_ALLOW_SMEM(kern, bytes)
#endif

int check_launch(sb_plan* pl, const char* what) {
    pl->launches++;
    int e = sb_rt_last_error();
    if (e != 0) return fail(std::string(what) + ": " + sb_rt_error_string(e));
    return 0;
}

int div_up(long a, long b) { return (int)((a + b - 1) / b); }

enum { K_CURV_ROWS = 0, K_CURV_COLS, K_TMPL_ROWS, K_TMPL_SUMS, K_CONV_COLS, K_FIT_ROWS };

sb_event_t take_event(sb_plan* pl) {
    if (pl->ev_used == pl->ev_pool.size()) {
        sb_event_t e;
        sb_rt_event_create(&e);
        pl->ev_pool.push_back(e);
    }
    return pl->ev_pool[pl->ev_used++];
}

struct ProfScope {   // brackets one launch with events when profiling is on
    sb_plan* pl;
    int kind;
    sb_event_t a{}, b{};
    ProfScope(sb_plan* p, int k) : pl(p), kind(k) {
        if (pl->profile) { a = take_event(pl); sb_rt_event_record(a, pl->stream); }
    }
    ~ProfScope() {
        if (pl->profile) {
            b = take_event(pl);
            sb_rt_event_record(b, pl->stream);
            pl->ev_live.push_back({kind, a, b});
        }
    }
};

void drain_profile(sb_plan* pl) {
    if (pl->ev_live.empty()) return;
    sb_rt_sync(pl->stream);
    for (auto& e : pl->ev_live) {
        pl->prof_ms[e.kind] += sb_rt_event_ms(e.a, e.b);
        pl->prof_n[e.kind]++;
    }
    pl->ev_live.clear();
    pl->ev_used = 0;
}

// curvature RMS -> power-of-two factor that brings curv**2 to the magnitude of curv
int update_curv_scale(sb_plan* pl) {
    {   // the angle-independent second differences, once per DEM
        const long n = (long)pl->ny * pl->nx;
        if (!pl->d_diffs) SB_TRY(sb_rt_malloc((void**)&pl->d_diffs, (size_t)n * 3 * sizeof(double)));
        if (!pl->d_diffs32) SB_TRY(sb_rt_malloc((void**)&pl->d_diffs32, (size_t)n * sizeof(float4)));
        SB_LAUNCH(sb::k_second_differences, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, pl->ny, pl->nx,
                  (const double*)pl->d_dem, pl->dx, pl->dx2, pl->dy2, pl->d_diffs, pl->d_diffs32);
        SB_OK(check_launch(pl, "k_second_differences"));
    }
    const int blocks = std::max(1, std::min(1024, div_up((long)pl->ny * pl->nx, 256)));
    SB_OK(ensure(pl->raw, (size_t)blocks * sizeof(double)));
    SB_LAUNCH(sb::k_curv_sumsq, dim3(blocks), dim3(256), 256 * sizeof(double), pl->stream, pl->ny, pl->nx,
              (const double*)pl->d_dem, pl->dx, pl->dx2, pl->dy2, (double*)pl->raw.p);
    SB_OK(check_launch(pl, "k_curv_sumsq"));
    std::vector<double> part(blocks);
    SB_TRY(sb_rt_d2h(part.data(), pl->raw.p, (size_t)blocks * sizeof(double), pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    double sum = 0.0;
    for (double v : part) sum += v;
    const double sigma = std::sqrt(sum / (2.0 * (double)pl->ny * (double)pl->nx));
    pl->c2_scale = 1.0;
    if (std::isfinite(sigma) && sigma > 1e-150 && sigma < 1e150)
        pl->c2_scale = std::exp2(std::round(-std::log2(sigma)));
    // a NaN anywhere in the DEM reaches every output pixel through the reference's full-raster
    // fft2 (dem.py:105, core.py:353-363); tiles that do not hold the cell have to be told
    pl->dem_nonfinite = !std::isfinite(sum);
    return 0;
}

// choose FFT length / tiling along one axis
int plan_axis(const sb_plan* pl, int n, int lo, int hi, AxisPlan* ax) {
    lo = std::min(lo, 0);
    hi = std::max(hi, 0);
    ax->lo = lo;
    ax->hi = hi;
    const int ext = hi - lo + 1;
    // complex128 at 8192 would need 270 KB of shared memory per column in k_conv_cols
    const int max_fft = std::min(pl->max_fft, pl->precision == 64 ? 4096 : kMaxFftSupported);
    if (is_pow2(n) && n >= kMinFft && n <= max_fft && !pl->force_pad) {
        ax->P = n;
        ax->tiles = 1;
        ax->tile_out = n;
        ax->periodic = true;
        return 0;
    }
    ax->periodic = false;
    long best_cost = -1;
    for (int P = kMinFft; P <= max_fft; P <<= 1) {
        if (P < ext + 1 || hi >= P / 2 || -lo > P / 2) continue;
        const int cap = P - ext;
        if (cap < 1) continue;
        const int tiles = div_up(n, cap);
        const long cost = (long)tiles * P;
        // equal cost: 4096, the length the pipelined kernels are tuned for, beats 2048
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && P == 4096)) {
            best_cost = cost;
            ax->P = P;
            ax->tiles = tiles;
            ax->tile_out = div_up(n, tiles);
        }
    }
    if (best_cost < 0)
        return fail("template support (" + std::to_string(ext) + " px) does not fit max_fft=" +
                    std::to_string(max_fft));
    return 0;
}

void fill_axis(const AxisPlan& ax, int n, int tile, int* o, int* out_n, int* split, int* dl,
               int* need_lo, int* need_hi) {
    *dl = -(n & 1);
    *o = tile * ax.tile_out;
    *out_n = std::min(ax.tile_out, n - *o);
    if (ax.periodic) {
        *split = ax.P;
        *need_lo = 0;
        *need_hi = ax.P - 1;
    } else {
        *split = *out_n - *dl - ax.lo + 1;
        *need_lo = -*dl - ax.hi;
        *need_hi = *out_n - 1 - *dl - ax.lo;
    }
}

struct SweepOut {
    double* raw_amp = nullptr;
    double* raw_snr = nullptr;
};

template <typename R>
int run_sweep_t(sb_plan* pl, const sb_angle* angles, int n_angles, const sb_template* tmpls_in,
                int n_tmpls, SweepOut so) {
    typedef typename Vec<R>::v2 C2;
    typedef typename Vec<R>::v4 C4;
    if (!pl->d_dem) return fail("no DEM set (sb_set_dem_host / sb_set_dem_dev)");
    if (pl->ny < 3 || pl->nx < 3) return fail("raster too small for a template search");
    if (!pl->d_x || !pl->d_y) return fail("no axis vectors set (sb_set_axes_host)");
    if (n_tmpls <= 0 || n_angles <= 0) return 0;

    // order templates by search angle so that each curvature spectrum is built once
    std::vector<sb_template> tm(tmpls_in, tmpls_in + n_tmpls);
    std::stable_sort(tm.begin(), tm.end(),
                     [](const sb_template& a, const sb_template& b) { return a.angle_id < b.angle_id; });
    int lo_y = 0, hi_y = 0, lo_x = 0, hi_x = 0, syp = 1;
    for (auto& t : tm) {
        if (t.kind == SB_KIND_RASTER && (n_tmpls != 1 || !pl->tbox.p))
            return fail("raster templates go through sb_match_template_raster, one at a time");
        if (t.angle_id < 0 || t.angle_id >= n_angles) return fail("template angle_id out of range");
        if (t.sy_hi < t.sy_lo || t.sx_hi < t.sx_lo) return fail("empty template support box");
        if (t.sy_lo < -(pl->ny / 2) || t.sy_hi > pl->ny - 1 - pl->ny / 2 || t.sx_lo < -(pl->nx / 2) ||
            t.sx_hi > pl->nx - 1 - pl->nx / 2)
            return fail("template support box exceeds the raster");
        lo_y = std::min(lo_y, t.sy_lo);
        hi_y = std::max(hi_y, t.sy_hi);
        lo_x = std::min(lo_x, t.sx_lo);
        hi_x = std::max(hi_x, t.sx_hi);
        syp = std::max(syp, t.sy_hi - t.sy_lo + 1);
    }
    syp += syp & 1;      // the row kernels store row pairs
    AxisPlan ay, ax;
    SB_OK(plan_axis(pl, pl->ny, lo_y, hi_y, &ay));
    SB_OK(plan_axis(pl, pl->nx, lo_x, hi_x, &ax));
    const int Py = ay.P, Px = ax.P;
    const int KX = Px / 2 + 1;
    const int kpitch = Px / 2 + 8;

    const C2 *twy = nullptr, *twx = nullptr;
    SB_OK(twiddles<R>(pl, Py, &twy));
    SB_OK(twiddles<R>(pl, Px, &twx));

    // batch sizes from the workspace budget
    const int need_rows_max = ay.periodic ? Py : std::min(Py, ay.tile_out + (hi_y - lo_y) + 2);
    const size_t per_angle = (size_t)(need_rows_max + 1) * KX * sizeof(C4) + (size_t)2 * KX * Py * sizeof(C2);
    const size_t per_tmpl = (size_t)KX * syp * sizeof(C4) + (size_t)Py * kpitch * sizeof(C4) +
                            (size_t)syp * sizeof(double2) + sizeof(sb::TSum);
    size_t budget = (size_t)pl->workspace_mb << 20;
    if (pl->workspace_mb <= 0) {
        // auto: the workspace already held counts as free
        size_t free_b = 0, total_b = 0, held = 0;
        for (const Buf* b : {&pl->cr, &pl->fct, &pl->trt, &pl->part, &pl->gbuf}) held += b->cap;
        if (sb_rt_mem_info(&free_b, &total_b) != 0) free_b = (size_t)16 << 30;
        budget = std::min<size_t>((size_t)48 << 30, std::max<size_t>((size_t)1 << 30, (free_b + held) / 2));
    }
    // templates per angle (max) decides the split of the budget
    std::vector<int> first(n_angles + 1, 0);
    for (auto& t : tm) first[t.angle_id + 1]++;
    int max_per_angle = 1;
    for (int a = 0; a < n_angles; ++a) {
        max_per_angle = std::max(max_per_angle, first[a + 1]);
        first[a + 1] += first[a];
    }
    int Bt = (int)std::max<size_t>(1, std::min<size_t>(64, (budget * 6 / 10) / per_tmpl));
    int Ba = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_angles, (budget * 4 / 10) / per_angle));
    Ba = std::min(Ba, 64);
    if (max_per_angle == 1) Bt = std::min(Bt, Ba), Ba = std::min(Ba, Bt);
    Bt = std::min(Bt, n_tmpls);
    // whole angles per batch when they fit: templates of one angle share the curvature spectra
    if (max_per_angle > 1 && Bt >= max_per_angle) Bt = (Bt / max_per_angle) * max_per_angle;
    else if (Bt > 1) Bt &= ~1;       // k_fit_rows_f walks the batch two templates at a time

    const int rpitch_max = need_rows_max + (need_rows_max & 1);
    SB_OK(ensure(pl->cr, (size_t)Ba * KX * rpitch_max * sizeof(C4)));
    SB_OK(ensure(pl->fct, (size_t)Ba * 2 * KX * Py * sizeof(C2)));
    SB_OK(ensure(pl->trt, (size_t)Bt * KX * syp * sizeof(C4)));
    SB_OK(ensure(pl->part, (size_t)Bt * syp * sizeof(double2)));
    SB_OK(ensure(pl->gbuf, (size_t)Bt * Py * kpitch * sizeof(C4)));
    SB_OK(ensure(pl->sums, (size_t)Bt * sizeof(sb::TSum)));
    SB_OK(ensure(pl->fit, (size_t)Bt * sizeof(sb::FitT)));
    SB_OK(ensure(pl->tmpls, (size_t)n_tmpls * sizeof(sb::Tmpl)));
    SB_OK(ensure(pl->angles, (size_t)n_angles * sizeof(sb::Angle)));
    SB_TRY(sb_rt_h2d(pl->tmpls.p, tm.data(), (size_t)n_tmpls * sizeof(sb::Tmpl), pl->stream));
    SB_TRY(sb_rt_h2d(pl->angles.p, angles, (size_t)n_angles * sizeof(sb::Angle), pl->stream));
    // the host vectors must outlive the async copies
    SB_TRY(sb_rt_sync(pl->stream));

    pl->last_geom[0] = Py; pl->last_geom[1] = Px; pl->last_geom[2] = ay.tiles; pl->last_geom[3] = ax.tiles;
    pl->last_geom[4] = Ba; pl->last_geom[5] = Bt;

    const sb::Tmpl* d_tm = (const sb::Tmpl*)pl->tmpls.p;
    const sb::Angle* d_an = (const sb::Angle*)pl->angles.p;
    sb::FitOut fo;
    fo.best_snr = pl->d_bsnr; fo.best_amp = pl->d_bamp; fo.best_idx = pl->d_bidx;
    fo.raw_amp = so.raw_amp; fo.raw_snr = so.raw_snr;

    for (int ty = 0; ty < ay.tiles; ++ty)
        for (int tx = 0; tx < ax.tiles; ++tx) {
            sb::Geom g;
            g.ny = pl->ny; g.nx = pl->nx; g.Py = Py; g.Px = Px;
            fill_axis(ay, pl->ny, ty, &g.oy, &g.out_ny, &g.split_y, &g.dly, &g.need_y_lo, &g.need_y_hi);
            fill_axis(ax, pl->nx, tx, &g.ox, &g.out_nx, &g.split_x, &g.dlx, &g.need_x_lo, &g.need_x_hi);
            g.kpitch = kpitch; g.syp = syp;
            g.dx = pl->dx; g.dx2 = pl->dx2; g.dy2 = pl->dy2;
            g.norm = 1.0 / ((double)Px * (double)Py);
            g.c2_scale = pl->c2_scale;
            g.dbg = std::getenv("SB_DBG") ? std::atoi(std::getenv("SB_DBG")) : 0;
            g.poison = pl->dem_nonfinite ? 1 : 0;
            const int need_rows = g.need_y_hi - g.need_y_lo + 1;
            g.rpitch = need_rows + (need_rows & 1);

            for (int a0 = 0; a0 < n_angles; a0 += Ba) {
                const int a1 = std::min(n_angles, a0 + Ba);
                const int p0 = first[a0], p1 = first[a1];
                if (p1 == p0) continue;
                if (std::is_same<R, float>::value && pl->fast) {
                    SB_OK(dispatch_n(Px, [&](auto nn) {
                        constexpr int N = decltype(nn)::value;
                        using S = Shape<N, float>;
                        auto kern = sb::k_curv_rows_f<N>;
                        SB_ALLOW_SMEM(kern, S::smem_conv_f);
                        ProfScope prof(pl, K_CURV_ROWS);
                        SB_LAUNCH(kern, dim3(a1 - a0, div_up(div_up(need_rows, 2), S::GP)), dim3(S::threads),
                                  S::smem_conv_f, pl->stream, g, (const float4*)pl->d_diffs32, d_an, a0,
                                  (float4*)pl->cr.p, (const float2*)twx);
                        return check_launch(pl, "k_curv_rows_f");
                    }));
                } else
                SB_OK(dispatch_n(Px, [&](auto nn) {
                    constexpr int N = decltype(nn)::value;
                    using S = Shape<N, R>;
                    auto kern = sb::k_curv_rows<N, R>;
                    SB_ALLOW_SMEM(kern, S::smem);
                    ProfScope prof(pl, K_CURV_ROWS);
                    SB_LAUNCH(kern, dim3(a1 - a0, div_up(div_up(need_rows, 2), S::GP)), dim3(S::threads),
                              S::smem, pl->stream, g, (const double*)pl->d_diffs, d_an, a0, (C4*)pl->cr.p, twx);
                    return check_launch(pl, "k_curv_rows");
                }));
                SB_OK(dispatch_n(Py, [&](auto nn) {
                    constexpr int N = decltype(nn)::value;
                    using S = Shape<N, R>;
                    auto kern = sb::k_curv_cols<N, R>;
                    SB_ALLOW_SMEM(kern, S::smem);
                    ProfScope prof(pl, K_CURV_COLS);
                    SB_LAUNCH(kern, dim3(div_up(KX, S::GP), a1 - a0), dim3(S::threads), S::smem,
                              pl->stream, g, (const C4*)pl->cr.p, (C2*)pl->fct.p, twy);
                    return check_launch(pl, "k_curv_cols");
                }));
                for (int pb = p0; pb < p1; pb += Bt) {
                    const int cnt = std::min(Bt, p1 - pb);
                    SB_OK(dispatch_n(Px, [&](auto nn) {
                        constexpr int N = decltype(nn)::value;
                        using S = Shape<N, R>;
                        ProfScope prof(pl, K_TMPL_ROWS);
                        // narrow templates: two samples per thread (needs a radix-16 first stage)
                        if constexpr (N >= 256) {
                            if (hi_x <= S::T - 1 && lo_x >= -S::T) {
                                auto kern = sb::k_tmpl_rows<N, R, true>;
                                SB_ALLOW_SMEM(kern, S::smem);
                                SB_LAUNCH(kern, dim3(div_up(syp / 2, S::GP), cnt), dim3(S::threads), S::smem,
                                          pl->stream, g, d_tm, pb, (const double*)pl->d_x, (const double*)pl->d_y,
                                          (C4*)pl->trt.p, (double2*)pl->part.p, twx, (const double*)pl->tbox.p);
                                return check_launch(pl, "k_tmpl_rows");
                            }
                        }
                        auto kern = sb::k_tmpl_rows<N, R, false>;
                        SB_ALLOW_SMEM(kern, S::smem);
                        SB_LAUNCH(kern, dim3(div_up(syp / 2, S::GP), cnt), dim3(S::threads), S::smem,
                                  pl->stream, g, d_tm, pb, (const double*)pl->d_x, (const double*)pl->d_y,
                                  (C4*)pl->trt.p, (double2*)pl->part.p, twx, (const double*)pl->tbox.p);
                        return check_launch(pl, "k_tmpl_rows");
                    }));
                    {
                        ProfScope prof(pl, K_TMPL_SUMS);
                        SB_LAUNCH(sb::k_tmpl_sums, dim3(cnt), dim3(32), 32 * sizeof(double2), pl->stream, g, d_tm, pb, cnt,
                                  (const double2*)pl->part.p, (sb::TSum*)pl->sums.p, (sb::FitT*)pl->fit.p);
                        SB_OK(check_launch(pl, "k_tmpl_sums"));
                    }
                    bool err_masks = false;
                    for (int i = pb; i < pb + cnt; ++i) err_masks |= tm[i].errmode != 0;
                    const bool fast = std::is_same<R, float>::value && pl->fast;
                    const bool fast_fit = fast && !so.raw_amp && !err_masks && cnt <= sb::kFitMaxBatch;
                    if constexpr (std::is_same<R, float>::value) {
                        if (fast) {
                            SB_OK(dispatch_n(Py, [&](auto nn) {
                                constexpr int N = decltype(nn)::value;
                                using S = Shape<N, float>;
                                // every template column has at most two non-zero inputs per thread
                                const bool sparse = hi_y <= S::T - 1 && lo_y >= -S::T;
                                ProfScope prof(pl, K_CONV_COLS);
                                if constexpr (N >= 1024 && N <= 4096) {
                                    const bool persist = pl->conv_persist > 0 || (pl->conv_persist < 0 && max_per_angle >= 4);
                                    if (persist && cnt <= sb::kConvPMaxBatch) {
                                        constexpr size_t smem_p =
                                            (size_t)(sb::kConvPThreads / S::T) * 2 * sbfft::padded_len(N) * sizeof(float2) +
                                            (size_t)2 * N * sizeof(float2) + sb::kConvPMaxBatch * 4 * sizeof(int) +
                                            (size_t)sbfft::ctw_count<float>(N) * sizeof(float2);
                                        if (sparse) {
                                            auto kern = sb::k_conv_cols_p<N, true>;
                                            SB_ALLOW_SMEM(kern, smem_p);
                                            SB_LAUNCH(kern, dim3(KX), dim3(sb::kConvPThreads), smem_p, pl->stream, g, d_tm,
                                                      pb, cnt, a0, (const float4*)pl->trt.p, (const float2*)pl->fct.p,
                                                      (float4*)pl->gbuf.p, (const float2*)twy);
                                        } else {
                                            auto kern = sb::k_conv_cols_p<N, false>;
                                            SB_ALLOW_SMEM(kern, smem_p);
                                            SB_LAUNCH(kern, dim3(KX), dim3(sb::kConvPThreads), smem_p, pl->stream, g, d_tm,
                                                      pb, cnt, a0, (const float4*)pl->trt.p, (const float2*)pl->fct.p,
                                                      (float4*)pl->gbuf.p, (const float2*)twy);
                                        }
                                        return check_launch(pl, "k_conv_cols_p");
                                    }
                                }
                                const dim3 grid(cnt, div_up(KX, S::GP));
                                if (sparse) {
                                    auto kern = sb::k_conv_cols_f<N, true>;
                                    SB_ALLOW_SMEM(kern, S::smem_conv_f);
                                    SB_LAUNCH(kern, grid, dim3(S::threads), S::smem_conv_f, pl->stream, g, d_tm, pb, a0,
                                              (const float4*)pl->trt.p, (const float2*)pl->fct.p, (float4*)pl->gbuf.p,
                                              (const float2*)twy);
                                } else {
                                    auto kern = sb::k_conv_cols_f<N, false>;
                                    SB_ALLOW_SMEM(kern, S::smem_conv_f);
                                    SB_LAUNCH(kern, grid, dim3(S::threads), S::smem_conv_f, pl->stream, g, d_tm, pb, a0,
                                              (const float4*)pl->trt.p, (const float2*)pl->fct.p, (float4*)pl->gbuf.p,
                                              (const float2*)twy);
                                }
                                return check_launch(pl, "k_conv_cols_f");
                            }));
                        }
                        if (fast_fit) {
                            SB_OK(dispatch_n(Px, [&](auto nn) {
                                constexpr int N = decltype(nn)::value;
                                using S = Shape<N, float>;
                                ProfScope prof(pl, K_FIT_ROWS);
                                if (pl->fit_threads == 0) {
                                    auto kern = sb::k_fit_rows_g<N>;
                                    SB_ALLOW_SMEM(kern, S::smem_fit_f);
                                    SB_LAUNCH(kern, dim3(div_up(Py / 2, S::GP)), dim3(S::threads), S::smem_fit_f,
                                              pl->stream, g, cnt, (const sb::FitT*)pl->fit.p,
                                              (const float4*)pl->gbuf.p, pl->d_bsnr, pl->d_bamp, pl->d_bidx,
                                              (const float2*)twx);
                                    return check_launch(pl, "k_fit_rows_g");
                                }
                                if (S::T <= 256 && pl->fit_threads == 512) {
                                    constexpr int threads = S::T > 512 ? S::T : 512;
                                    constexpr int GP = threads / S::T;
                                    constexpr size_t smem = S::smem_fit_f - S::smem_conv_f +
                                                            (size_t)GP * 2 * sbfft::padded_len(N) * sizeof(float2);
                                    auto kern = sb::k_fit_rows_f<N, 512>;
                                    SB_ALLOW_SMEM(kern, smem);
                                    SB_LAUNCH(kern, dim3(div_up(g.out_ny, GP)), dim3(threads), smem, pl->stream, g, cnt,
                                              (const sb::FitT*)pl->fit.p, (const float4*)pl->gbuf.p, pl->d_bsnr,
                                              pl->d_bamp, pl->d_bidx, (const float2*)twx);
                                } else {
                                    auto kern = sb::k_fit_rows_f<N, 256>;
                                    SB_ALLOW_SMEM(kern, S::smem_fit_f);
                                    SB_LAUNCH(kern, dim3(div_up(g.out_ny, S::GP)), dim3(S::threads), S::smem_fit_f,
                                              pl->stream, g, cnt, (const sb::FitT*)pl->fit.p,
                                              (const float4*)pl->gbuf.p, pl->d_bsnr, pl->d_bamp, pl->d_bidx,
                                              (const float2*)twx);
                                }
                                return check_launch(pl, "k_fit_rows_f");
                            }));
                            if (g.poison) {
                                SB_LAUNCH(sb::k_poison_windows, dim3(div_up((long)g.out_ny * g.out_nx, 256)), dim3(256), 0,
                                          pl->stream, g, cnt, (const sb::FitT*)pl->fit.p, pl->d_bsnr);
                                SB_OK(check_launch(pl, "k_poison_windows"));
                            }
                        }
                    }
                    if (!fast) {
                    SB_OK(dispatch_n(Py, [&](auto nn) {
                        constexpr int N = decltype(nn)::value;
                        using S = Shape<N, R>;
                        auto kern = sb::k_conv_cols<N, R>;
                        SB_ALLOW_SMEM(kern, S::smem_conv);
                        ProfScope prof(pl, K_CONV_COLS);
                        SB_LAUNCH(kern, dim3(div_up(KX, S::GP), cnt), dim3(S::threads), S::smem_conv,
                                  pl->stream, g, d_tm, pb, a0, (const C4*)pl->trt.p, (const C2*)pl->fct.p,
                                  (C4*)pl->gbuf.p, twy);
                        return check_launch(pl, "k_conv_cols");
                    }));
                    }
                    if (!fast_fit) {
                    SB_OK(dispatch_n(Px, [&](auto nn) {
                        constexpr int N = decltype(nn)::value;
                        using S = Shape<N, R>;
                        auto kern = sb::k_fit_rows<N, R>;
                        SB_ALLOW_SMEM(kern, S::smem);
                        ProfScope prof(pl, K_FIT_ROWS);
                        SB_LAUNCH(kern, dim3(div_up(g.out_ny, S::GP)), dim3(S::threads), S::smem,
                                  pl->stream, g, d_tm, pb, cnt, (const sb::TSum*)pl->sums.p,
                                  (const C4*)pl->gbuf.p, (const double*)pl->d_x, (const double*)pl->d_y, fo, twx);
                        return check_launch(pl, "k_fit_rows");
                    }));
                    }
                }
            }
        }
    drain_profile(pl);
    return 0;
}

int run_sweep(sb_plan* pl, const sb_angle* angles, int n_angles, const sb_template* tmpls_in, int n_tmpls,
              SweepOut so) {
    if (pl->precision == 64) return run_sweep_t<double>(pl, angles, n_angles, tmpls_in, n_tmpls, so);
    return run_sweep_t<float>(pl, angles, n_angles, tmpls_in, n_tmpls, so);
}

int copy_out(sb_plan* pl, double* dst, const double* src_dev, size_t count, int out_is_device) {
    if (out_is_device)
        SB_TRY(sb_rt_d2d(dst, src_dev, count * sizeof(double), pl->stream));
    else
        SB_TRY(sb_rt_d2h(dst, src_dev, count * sizeof(double), pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

}  // namespace

extern "C" {

const char* sb_last_error(void) { return g_err.c_str(); }

const char* sb_build_info(void) {
#ifdef SB_EMU
    return "cpu-emulator (test infrastructure)";
#else
    return "cuda sm_100a";
#endif
}

int sb_plan_create(sb_plan** plan, int ny, int nx, double dx, double dx2, double dy2, int device,
                   void* stream, unsigned flags) {
    (void)flags;
    if (!plan || ny < 1 || nx < 1) return fail("sb_plan_create: bad arguments");
    sb_plan* pl = new sb_plan();
    pl->ny = ny; pl->nx = nx; pl->dx = dx; pl->dx2 = dx2; pl->dy2 = dy2;
#ifndef SB_EMU
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        delete pl;
        return fail("no CUDA device available: scarplet_b200 has no CPU path");
    }
    if (device >= 0) {
        if (cudaSetDevice(device) != cudaSuccess) { delete pl; return fail("cudaSetDevice failed"); }
    }
    cudaGetDevice(&pl->device);
    if (stream) {
        pl->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete pl;
            return fail("cudaStreamCreate failed");
        }
        pl->own_stream = true;
    }
#else
    (void)device; (void)stream;
#endif
    const size_t n = (size_t)ny * nx;
    int e = 0;
    e |= sb_rt_malloc((void**)&pl->d_bsnr, n * sizeof(float));
    e |= sb_rt_malloc((void**)&pl->d_bamp, n * sizeof(float));
    e |= sb_rt_malloc((void**)&pl->d_bidx, n * sizeof(int));
    if (e) { sb_plan_destroy(pl); return fail("sb_plan_create: out of device memory"); }
    *plan = pl;
    return sb_best_reset(pl);
}

int sb_plan_destroy(sb_plan* pl) {
    if (!pl) return 0;
#ifndef SB_EMU
    cudaSetDevice(pl->device);
#endif
    sb_rt_sync(pl->stream);
    if (pl->own_dem && pl->d_dem) sb_rt_free(pl->d_dem);
    if (pl->d_diffs) sb_rt_free(pl->d_diffs);
    if (pl->d_diffs32) sb_rt_free(pl->d_diffs32);
    if (pl->d_x) sb_rt_free(pl->d_x);
    if (pl->d_y) sb_rt_free(pl->d_y);
    if (pl->d_bsnr) sb_rt_free(pl->d_bsnr);
    if (pl->d_bamp) sb_rt_free(pl->d_bamp);
    if (pl->d_bidx) sb_rt_free(pl->d_bidx);
    for (auto& kv : pl->tw) sb_rt_free(kv.second);
    for (auto e : pl->ev_pool) sb_rt_event_destroy(e);
    for (Buf* b : {&pl->cr, &pl->fct, &pl->trt, &pl->part, &pl->gbuf, &pl->sums, &pl->fit, &pl->tmpls, &pl->angles,
                   &pl->tables, &pl->raw, &pl->tbox})
        release(*b);
#ifndef SB_EMU
    if (pl->own_stream) cudaStreamDestroy(pl->stream);
#endif
    delete pl;
    return 0;
}

int sb_plan_set_option(sb_plan* pl, const char* key, long value) {
    if (!pl || !key) return fail("sb_plan_set_option: null");
    std::string k(key);
    if (k == "workspace_mb") { pl->workspace_mb = value <= 0 ? 0 : std::max(64L, value); return 0; }
    if (k == "max_fft") {
        if (!is_pow2((int)value) || value < kMinFft || value > kMaxFftSupported)
            return fail("max_fft must be a power of two in [128, 8192]");
        pl->max_fft = (int)value;
        return 0;
    }
    if (k == "force_pad") { pl->force_pad = value != 0; return 0; }
    if (k == "profile") { pl->profile = value != 0; return 0; }
    if (k == "fast") { pl->fast = value != 0; return 0; }
    if (k == "precision") {
        if (value != 32 && value != 64) return fail("precision must be 32 or 64");
        pl->precision = (int)value;
        return 0;
    }
    return fail("unknown option " + k);
}

long sb_plan_launch_count(const sb_plan* pl) { return pl ? pl->launches : 0; }

int sb_plan_profile(sb_plan* pl, double* ms6, long* launches6, int reset) {
    if (!pl) return fail("null plan");
    drain_profile(pl);
    for (int i = 0; i < 6; ++i) {
        if (ms6) ms6[i] = pl->prof_ms[i];
        if (launches6) launches6[i] = pl->prof_n[i];
        if (reset) { pl->prof_ms[i] = 0; pl->prof_n[i] = 0; }
    }
    return 0;
}

int sb_plan_last_geometry(const sb_plan* pl, int* out6) {
    if (!pl || !out6) return fail("null");
    for (int i = 0; i < 6; ++i) out6[i] = pl->last_geom[i];
    return 0;
}

int sb_set_dem_host(sb_plan* pl, const double* dem_host) {
    if (!pl || !dem_host) return fail("sb_set_dem_host: null");
    const size_t bytes = (size_t)pl->ny * pl->nx * sizeof(double);
    if (!pl->own_dem || !pl->d_dem) {
        pl->d_dem = nullptr;
        SB_TRY(sb_rt_malloc((void**)&pl->d_dem, bytes));
        pl->own_dem = true;
    }
    SB_TRY(sb_rt_h2d(pl->d_dem, dem_host, bytes, pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    return update_curv_scale(pl);
}

int sb_set_dem_dev(sb_plan* pl, const double* dem_dev) {
    if (!pl || !dem_dev) return fail("sb_set_dem_dev: null");
    if (pl->own_dem && pl->d_dem) sb_rt_free(pl->d_dem);
    pl->own_dem = false;
    pl->d_dem = const_cast<double*>(dem_dev);
    return update_curv_scale(pl);
}

int sb_set_axes_host(sb_plan* pl, const double* x_host, const double* y_host) {
    if (!pl || !x_host || !y_host) return fail("sb_set_axes_host: null");
    if (!pl->d_x) SB_TRY(sb_rt_malloc((void**)&pl->d_x, (size_t)pl->nx * sizeof(double)));
    if (!pl->d_y) SB_TRY(sb_rt_malloc((void**)&pl->d_y, (size_t)pl->ny * sizeof(double)));
    SB_TRY(sb_rt_h2d(pl->d_x, x_host, (size_t)pl->nx * sizeof(double), pl->stream));
    SB_TRY(sb_rt_h2d(pl->d_y, y_host, (size_t)pl->ny * sizeof(double), pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_directional_laplacian(sb_plan* pl, const sb_angle* angle, double* out, int out_is_device) {
    if (!pl || !angle || !out) return fail("sb_directional_laplacian: null");
    if (!pl->d_dem) return fail("no DEM set");
    const long n = (long)pl->ny * pl->nx;
    double* dst = out;
    if (!out_is_device) {
        SB_OK(ensure(pl->raw, (size_t)n * 2 * sizeof(double)));
        dst = (double*)pl->raw.p;
    }
    sb::Angle a;
    a.ca = angle->cos_a; a.sa = angle->sin_a; a.ca2 = angle->cos2_a; a.sa2 = angle->sin2_a;
    SB_LAUNCH(sb::k_laplacian, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, pl->ny, pl->nx,
              (const double*)pl->d_dem, pl->dx, pl->dx2, pl->dy2, a, dst);
    SB_OK(check_launch(pl, "k_laplacian"));
    if (!out_is_device) return copy_out(pl, out, dst, (size_t)n, 0);
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_render_template(sb_plan* pl, const sb_template* tmpl, double* out, int out_is_device) {
    if (!pl || !tmpl || !out) return fail("sb_render_template: null");
    if (!pl->d_x || !pl->d_y) return fail("no axis vectors set");
    const long n = (long)pl->ny * pl->nx;
    double* dst = out;
    if (!out_is_device) {
        SB_OK(ensure(pl->raw, (size_t)n * 2 * sizeof(double)));
        dst = (double*)pl->raw.p;
    }
    sb::Tmpl t;
    std::memcpy(&t, tmpl, sizeof(t));
    SB_LAUNCH(sb::k_render_template, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, pl->ny, pl->nx, t,
              (const double*)pl->d_x, (const double*)pl->d_y, dst);
    SB_OK(check_launch(pl, "k_render_template"));
    if (!out_is_device) return copy_out(pl, out, dst, (size_t)n, 0);
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_match_template(sb_plan* pl, const sb_angle* angle, const sb_template* tmpl, double* amp, double* snr,
                      int out_is_device) {
    if (!pl || !angle || !tmpl || !amp || !snr) return fail("sb_match_template: null");
    const size_t n = (size_t)pl->ny * pl->nx;
    SweepOut so;
    if (out_is_device) {
        so.raw_amp = amp;
        so.raw_snr = snr;
    } else {
        SB_OK(ensure(pl->raw, n * 2 * sizeof(double)));
        so.raw_amp = (double*)pl->raw.p;
        so.raw_snr = so.raw_amp + n;
    }
    sb_template t = *tmpl;
    t.angle_id = 0;
    SB_OK(run_sweep(pl, angle, 1, &t, 1, so));
    if (!out_is_device) {
        SB_TRY(sb_rt_d2h(amp, so.raw_amp, n * sizeof(double), pl->stream));
        SB_TRY(sb_rt_d2h(snr, so.raw_snr, n * sizeof(double), pl->stream));
    }
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_match_template_raster(sb_plan* pl, const sb_angle* angle, const double* box_host, int sy_lo, int sy_hi,
                             int sx_lo, int sx_hi, double tscale, double* amp, double* snr, int out_is_device) {
    if (!pl || !angle || !box_host || !amp || !snr) return fail("sb_match_template_raster: null");
    if (sy_hi < sy_lo || sx_hi < sx_lo) return fail("sb_match_template_raster: empty box");
    const size_t cells = (size_t)(sy_hi - sy_lo + 1) * (size_t)(sx_hi - sx_lo + 1);
    SB_OK(ensure(pl->tbox, cells * sizeof(double)));
    SB_TRY(sb_rt_h2d(pl->tbox.p, box_host, cells * sizeof(double), pl->stream));
    sb_template t;
    std::memset(&t, 0, sizeof(t));
    t.cos_t = 1.0;
    t.sign = 1.0;
    t.tscale = tscale > 0.0 ? tscale : 1.0;
    t.kind = SB_KIND_RASTER;
    t.errmode = SB_ERRMASK_NONE;
    t.sy_lo = sy_lo; t.sy_hi = sy_hi; t.sx_lo = sx_lo; t.sx_hi = sx_hi;
    t.i_lo = 0; t.i_hi = pl->ny - 1; t.j_lo = 0; t.j_hi = pl->nx - 1;     // masks are the caller's
    return sb_match_template(pl, angle, &t, amp, snr, out_is_device);
}

int sb_best_reset(sb_plan* pl) {
    if (!pl) return fail("null plan");
    const long n = (long)pl->ny * pl->nx;
    SB_LAUNCH(sb::k_best_init, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, pl->d_bsnr, pl->d_bamp,
              pl->d_bidx);
    return check_launch(pl, "k_best_init");
}

int sb_sweep(sb_plan* pl, const sb_angle* angles, int n_angles, const sb_template* tmpls, int n_tmpls) {
    if (!pl || !angles || !tmpls) return fail("sb_sweep: null");
    return run_sweep(pl, angles, n_angles, tmpls, n_tmpls, SweepOut());
}

int sb_finalize(sb_plan* pl, const double* age_of_host, const double* angle_of_host, int n_idx, double* out4,
                int out_is_device) {
    if (!pl || !age_of_host || !angle_of_host || !out4 || n_idx <= 0) return fail("sb_finalize: bad arguments");
    const long n = (long)pl->ny * pl->nx;
    SB_OK(ensure(pl->tables, (size_t)2 * n_idx * sizeof(double)));
    double* d_age = (double*)pl->tables.p;
    double* d_ang = d_age + n_idx;
    SB_TRY(sb_rt_h2d(d_age, age_of_host, (size_t)n_idx * sizeof(double), pl->stream));
    SB_TRY(sb_rt_h2d(d_ang, angle_of_host, (size_t)n_idx * sizeof(double), pl->stream));
    double* dst = out4;
    if (!out_is_device) {
        SB_OK(ensure(pl->raw, (size_t)n * 4 * sizeof(double)));
        dst = (double*)pl->raw.p;
    }
    SB_LAUNCH(sb::k_finalize, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, (const float*)pl->d_bsnr,
              (const float*)pl->d_bamp, (const int*)pl->d_bidx, (const double*)d_age, (const double*)d_ang, dst);
    SB_OK(check_launch(pl, "k_finalize"));
    if (!out_is_device) return copy_out(pl, out4, dst, (size_t)n * 4, 0);
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_best_state(sb_plan* pl, float** snr_dev, float** amp_dev, int32_t** idx_dev) {
    if (!pl) return fail("null plan");
    if (snr_dev) *snr_dev = pl->d_bsnr;
    if (amp_dev) *amp_dev = pl->d_bamp;
    if (idx_dev) *idx_dev = pl->d_bidx;
    return 0;
}

int sb_best_pack(sb_plan* pl, unsigned long long* keys_dev) {
    if (!pl || !keys_dev) return fail("sb_best_pack: null");
    const long n = (long)pl->ny * pl->nx;
    SB_LAUNCH(sb::k_best_pack, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, (const float*)pl->d_bsnr,
              (const int*)pl->d_bidx, keys_dev);
    return check_launch(pl, "k_best_pack");
}

int sb_best_select(sb_plan* pl, const unsigned long long* gkeys_dev, float* amp_out_dev) {
    if (!pl || !gkeys_dev || !amp_out_dev) return fail("sb_best_select: null");
    const long n = (long)pl->ny * pl->nx;
    SB_LAUNCH(sb::k_best_select, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, (const float*)pl->d_bsnr,
              (const float*)pl->d_bamp, (const int*)pl->d_bidx, gkeys_dev, amp_out_dev);
    return check_launch(pl, "k_best_select");
}

int sb_best_unpack(sb_plan* pl, const unsigned long long* gkeys_dev, const float* amp_dev) {
    if (!pl || !gkeys_dev || !amp_dev) return fail("sb_best_unpack: null");
    const long n = (long)pl->ny * pl->nx;
    SB_LAUNCH(sb::k_best_unpack, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, gkeys_dev, amp_dev,
              pl->d_bsnr, pl->d_bamp, pl->d_bidx);
    return check_launch(pl, "k_best_unpack");
}

int sb_compare_host(sb_plan* pl, double* best4_host, const double* amp, const double* age, const double* angle,
                    const double* snr, double age_s, double angle_s) {
    if (!pl || !best4_host || !amp || !snr) return fail("sb_compare_host: null");
    const long n = (long)pl->ny * pl->nx;
    SB_OK(ensure(pl->raw, (size_t)n * 8 * sizeof(double)));
    double* d = (double*)pl->raw.p;
    SB_TRY(sb_rt_h2d(d, best4_host, (size_t)n * 4 * sizeof(double), pl->stream));
    SB_TRY(sb_rt_h2d(d + 4 * n, amp, (size_t)n * sizeof(double), pl->stream));
    SB_TRY(sb_rt_h2d(d + 5 * n, snr, (size_t)n * sizeof(double), pl->stream));
    const double* d_age = nullptr;
    const double* d_ang = nullptr;
    if (age) { SB_TRY(sb_rt_h2d(d + 6 * n, age, (size_t)n * sizeof(double), pl->stream)); d_age = d + 6 * n; }
    if (angle) { SB_TRY(sb_rt_h2d(d + 7 * n, angle, (size_t)n * sizeof(double), pl->stream)); d_ang = d + 7 * n; }
    SB_LAUNCH(sb::k_compare, dim3(div_up(n, 256)), dim3(256), 0, pl->stream, n, d, d + n, d + 2 * n, d + 3 * n,
              (const double*)(d + 4 * n), d_age, d_ang, (const double*)(d + 5 * n), age_s, angle_s);
    SB_OK(check_launch(pl, "k_compare"));
    return copy_out(pl, best4_host, d, (size_t)n * 4, 0);
}

int sb_debug_fft(sb_plan* pl, int n, int rows, const float* in_host, float* out_host, int inverse) {
    if (!pl || !in_host || !out_host || rows <= 0) return fail("sb_debug_fft: bad arguments");
    const float2* tw = nullptr;
    if (!is_pow2(n) || n < kMinFft || n > kMaxFftSupported) return fail("sb_debug_fft: unsupported length");
    SB_OK(twiddles<float>(pl, n, &tw));
    const size_t bytes = (size_t)rows * n * sizeof(float2);
    SB_OK(ensure(pl->raw, 2 * bytes));
    float2* d_in = (float2*)pl->raw.p;
    float2* d_out = d_in + (size_t)rows * n;
    SB_TRY(sb_rt_h2d(d_in, in_host, bytes, pl->stream));
    SB_OK(dispatch_n(n, [&](auto nn) {
        constexpr int N = decltype(nn)::value;
        using S = Shape<N, float>;
        auto kern = sb::k_fft_rows<N, float>;
        SB_ALLOW_SMEM(kern, S::smem);
        SB_LAUNCH(kern, dim3(div_up(rows, S::GP)), dim3(S::threads), S::smem, pl->stream, rows,
                  (const float2*)d_in, d_out, inverse, tw);
        return check_launch(pl, "k_fft_rows");
    }));
    SB_TRY(sb_rt_d2h(out_host, d_out, bytes, pl->stream));
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

int sb_sync(sb_plan* pl) {
    if (!pl) return fail("null plan");
    SB_TRY(sb_rt_sync(pl->stream));
    return 0;
}

}  // extern "C"
