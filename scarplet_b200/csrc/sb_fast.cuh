// Pipelined ("leapfrog") versions of the two per-template kernels, complex64 only.
//
// The simple kernels in sb_kernels.cuh run one transform at a time: every exchange
// between radix-16 stages costs two CTA barriers, and the loads that feed a stage
// (twiddles, spectra) are issued right where their values are needed.  ncu showed
// those kernels waiting, not working (46 % of warp stalls on loads, 9 % on barriers,
// issue slots 31 % busy).  Here every thread group carries TWO independent transforms
// a and b that take turns: while a computes a stage in registers, b's data rests in
// its own exchange buffer, and vice versa --
//
//     a.math  a.store | bar | b.math  b.store  a.load  a.twiddles | bar | a.math ...
//
// so there is ONE barrier per exchange, the twiddles and the exchanged data of the
// next stage are already in flight when the barrier opens, and only one transform's
// 16 elements are live in registers at any time.
//
//   k_conv_cols_f  a, b = the two fields (t * curv, M * curv^2) of one spectrum column
//   k_fit_rows_f   a, b = two templates of the batch, same raster row
#pragma once
#include "sb_kernels.cuh"

namespace sb {

using sbfft::TW;
#ifdef SB_EMU
using std::max;
using std::min;
#endif

// Developer ablation switches (scratch/ablate.sh): build with -DSB_ABLATE and set SB_DBG to
// switch parts of the kernels' memory traffic off and time what is left.  Compiled out of
// the product build.
#ifdef SB_ABLATE
#define SB_DBG_ON(flags, bit) (((flags) & (bit)) != 0)
#else
#define SB_DBG_ON(flags, bit) false
#endif

SB_CONSTEXPR int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// ---------------------------------------------------------------------------
// leapfrog driver.  Ctx provides, for phase P of stream F (0 = a, 1 = b):
//   twid<P>(w)       fetch the twiddles phase P will use
//   phase<P, F>(v,w) register work of the phase
//   store<P, F>(v)   scatter to the stream's exchange buffer (not after the last phase)
//   load<F>(v)       gather the next phase's elements
//   bar()            barrier over the threads that carry the two transforms
// On entry to Frog<P>: a has finished phase P and stored it, a barrier has passed,
// b holds its phase-P input in vb (and wb).
// ---------------------------------------------------------------------------
template <int P, int K, class Ctx>
struct Frog {
    SB_DEVICE static void run(Ctx& c, float2 (&va)[E], float2 (&vb)[E], float2 (&wa)[TW], float2 (&wb)[TW]) {
        c.template phase<P, 1>(vb, wb);
        if constexpr (P < K - 1) {
            c.template store<P, 1>(vb);
            c.template load<0>(va);
            c.template twid<P + 1>(wa);
            c.bar();
            c.template phase<P + 1, 0>(va, wa);
            if constexpr (P + 1 < K - 1) c.template store<P + 1, 0>(va);
            c.template load<1>(vb);
            c.template twid<P + 1>(wb);
            if constexpr (P + 1 < K - 1) c.bar();
            Frog<P + 1, K, Ctx>::run(c, va, vb, wa, wb);
        }
    }
};

template <int K, class Ctx>
SB_DEVICE void leapfrog(Ctx& c, float2 (&va)[E], float2 (&vb)[E]) {
    float2 wa[TW], wb[TW];
    c.template twid<0>(wa);
    c.template twid<0>(wb);
    c.template phase<0, 0>(va, wa);
    c.template store<0, 0>(va);
    c.bar();
    Frog<0, K, Ctx>::run(c, va, vb, wa, wb);
}

// ---------------------------------------------------------------------------
// k_conv_cols_f<Py, SPARSE>: grid (n_templates, ceil(KX / GP)) -- the template index
// runs fastest, so the CTAs that multiply by the same curvature-spectrum column are
// co-resident and the column comes out of L2 for all but the first of them.
// Phases of one field: forward stages 0..NST-1, the last one fused with the product
// (core.py:359 / :363) and inverse stage 0, then inverse stages 1..NST-1.
// SPARSE: the template support is shorter than T rows on either side of the origin,
// so each thread's only non-zero inputs are elements 0 and 15 (stage0_sparse2).
// ---------------------------------------------------------------------------
// PERSIST (k_conv_cols_p): the spectrum columns are in shared memory and the thread group has
// its own named barrier.
template <int N, bool SPARSE, bool PERSIST = false>
struct ConvCtx {
    static constexpr int NST = sbfft::num_stages(N);
    static constexpr int K = 2 * NST - 1;
    static constexpr int T = N / E;
    int t;
    float2 *smA, *smB;
    const float2* tw;
    const float2 *specA, *specB;      // curvature-spectrum columns of the two fields (+ t)
    float4* dst;                      // gbuf plane of this template (nullptr: inactive group)
    int kx, dly, out_ny, kpitch;
    int dbg;
    int bar_id;                       // PERSIST: named barrier of this thread group
    const float2* va_fin;             // SB_CONV_KEEP_A: field a's registers (the array handed to leapfrog)

    SB_DEVICE void bar() const {
#ifdef SB_ABL_NOBAR
        return;
#endif
        if constexpr (PERSIST) sb_bar(bar_id, T);
        else sb_sync();
    }
    SB_CONSTEXPR static int stage_of(int P) { return P < NST ? P : P - NST + 1; }

    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) {
        if (SB_DBG_ON(dbg, 4)) {
#pragma unroll
            for (int i = 0; i < TW; ++i) w[i] = make_float2(0.6f, 0.8f);
            return;
        }
        sbfft::load_tw<N, stage_of(P), float, PERSIST>(w, t, tw);   // PERSIST: the CTA's shared-memory copy
    }

    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        if constexpr (P == 0 && SPARSE) sbfft::stage0_sparse2<float>(v);
        else sbfft::stage_math<N, stage_of(P), float>(v, w);
        if constexpr (P == NST - 1) {
            const float2* spec = F == 0 ? specA : specB;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const float2 s = SB_DBG_ON(dbg, 2) ? make_float2(0.5f, 0.25f)
                                 : PERSIST ? spec[q * T] : sb_ld_stream(spec + q * T);
                const float2 pr = sbfft::cmul(v[q], s);
                v[q] = make_float2(pr.y, pr.x);                  // swap: inverse via forward
            }
            sbfft::stage_math<N, 0, float>(v, w);
        }
#define SB_CONV_KEEP_A 1    // field a's result waits in registers for field b (parking it in shared memory: +2 %)
#ifndef SB_CONV_KEEP_A
        if constexpr (P == K - 1 && F == 0) {
            // field a is done; park it in its own (now idle) exchange buffer, every thread in
            // the slots only it will read back, so that its registers are free for field b
#pragma unroll
            for (int q = 0; q < E; ++q) smA[t + q * T] = v[q];
        }
#endif
        if constexpr (P == K - 1 && F == 1) {
            if (dst) {
                // row m = t + q T lies at gbuf_index(t, kx) + q * T * kpitch (T is even): one pointer,
                // one add per store
                static_assert(kGbufRows == 2 && T % 2 == 0, "row-pair interleave");
                const long step = (long)T * kpitch;
                float4* p = dst + gbuf_index(t, kx, kpitch);
                const int io0 = t + dly;
#pragma unroll
                for (int q = 0; q < E; ++q) {
#ifdef SB_CONV_KEEP_A
                    const float2 a = va_fin[q];      // field a's result, still in the caller's registers
#else
                    const float2 a = smA[t + q * T];
#endif
                    if (SB_DBG_ON(dbg, 1) && v[q].x != 1.2345e-30f) continue;
                    const float4 out = gbuf_pack<float2, float4>(a, v[q]);
                    if (SB_DBG_ON(dbg, 128)) {      // timing only: same bytes, fully coalesced
                        sb_st_stream(dst + ((long)kx * N + t + q * T), out);
                        continue;
                    }
                    if (((io0 + q * T) & (N - 1)) < out_ny) sb_st_stream(p, out);
                    p += step;
                }
            }
        }
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        constexpr int S = (P == NST - 1) ? 0 : stage_of(P);
#ifdef SB_ABL_NOXCHG
        if (v[0].x != 1.2345e-30f) return;
#endif
        sbfft::stage_store<N, S, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) {
#ifdef SB_ABL_NOXCHG
        if (t >= 0) return;
#endif
        sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB);
    }
};

template <int N, bool SPARSE>
SB_GLOBAL SB_LAUNCH_BOUNDS((N / E > 256 ? N / E : 256), (N / E > 256 ? 1 : 2))
k_conv_cols_f(Geom g, const Tmpl* SB_RESTRICT tmpls, int tmpl_base, int angle_base,
              const float4* SB_RESTRICT trt, const float2* SB_RESTRICT fct, float4* SB_RESTRICT gbuf,
              const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int GP = (T > 256 ? T : 256) / T;
    constexpr int PL = sbfft::padded_len(N);
    const int grp = sb_tid() / T, t = sb_tid() % T;
    const int KX = g.Px / 2 + 1;
    const int p_loc = sb_bx();
    const int kx = sb_by() * GP + grp;
    const bool active = kx < KX;
    const int kxc = active ? kx : 0;
    const Tmpl* p = tmpls + tmpl_base + p_loc;
    const int sy_lo = p->sy_lo, sy_hi = p->sy_hi;
    const int a_loc = p->angle_id - angle_base;

    ConvCtx<N, SPARSE> c;
    c.t = t;
    c.smA = (float2*)sb_shared() + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw;
    c.specA = fct + (((long)a_loc * 2) * KX + kxc) * N + t;
    c.specB = c.specA + (long)KX * N;
    c.dst = active ? gbuf + (long)p_loc * N * g.kpitch : nullptr;
    c.kx = kx;
    c.dly = g.dly;
    c.out_ny = g.out_ny;
    c.kpitch = g.kpitch;
    c.dbg = g.dbg;
    // one 128-byte line of each spectrum column per thread (T lines per column)
    sb_prefetch_l2(c.specA + 15 * t);
    sb_prefetch_l2(c.specB + 15 * t);

    const float4* src = trt + ((long)p_loc * KX + kxc) * g.syp;
    float2 va[E], vb[E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
        va[q] = make_float2(0.f, 0.f);
        vb[q] = make_float2(0.f, 0.f);
        if (SPARSE && q != 0 && q != E - 1) continue;
        const int qy = t + q * T;
        const int s = qy < N / 2 ? qy : qy - N;
        if (active && s >= sy_lo && s <= sy_hi) {
            const float4 w = SB_DBG_ON(g.dbg, 8) ? make_float4(1.f, 2.f, 3.f, 4.f) : sb_ld_stream(src + (s - sy_lo));
            va[q] = make_float2(w.x, w.y);
            vb[q] = make_float2(w.z, w.w);
        }
    }
    c.va_fin = va;
    leapfrog<ConvCtx<N, SPARSE>::K>(c, va, vb);     // the last phase of field b writes gbuf
}

// ---------------------------------------------------------------------------
// k_conv_cols_p<Py, SPARSE>: persistent variant, grid (KX), 512 threads = G groups of T.
// A CTA owns one spectrum column kx for the whole batch of templates.  The batch is
// walked in runs of templates that share a search angle: the two curvature-spectrum
// columns of that angle are staged in shared memory ONCE and multiplied into every
// template of the run (k_conv_cols_f fetches them from L2 per template, right where the
// product needs them -- a third of its warp stalls); the twiddles stay in registers
// for the whole batch; the groups take the templates of a run round-robin, each behind
// its own named barrier.
// ---------------------------------------------------------------------------
constexpr int kConvPThreads = 512;
constexpr int kConvPMaxBatch = 64;

template <int N, bool SPARSE>
SB_GLOBAL SB_LAUNCH_BOUNDS(kConvPThreads, 1)
k_conv_cols_p(Geom g, const Tmpl* SB_RESTRICT tmpls, int tmpl_base, int cnt, int angle_base,
              const float4* SB_RESTRICT trt, const float2* SB_RESTRICT fct, float4* SB_RESTRICT gbuf,
              const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int G = kConvPThreads / T;
    constexpr int PL = sbfft::padded_len(N);
    static_assert(T >= 32 && G >= 1, "k_conv_cols_p: one group must be whole warps");
    typedef ConvCtx<N, SPARSE, true> Ctx;
    const int grp = sb_tid() / T, t = sb_tid() % T;
    const int KX = g.Px / 2 + 1;
    const int kx = sb_bx();
    float2* sm = (float2*)sb_shared();
    float2* spec_s = sm + (long)G * 2 * PL;               // [2][N]
    int* s_meta = (int*)(spec_s + 2 * N);                  // [kConvPMaxBatch][4]: sy_lo, sy_hi, angle_id
    float2* tw_s = (float2*)(s_meta + 4 * kConvPMaxBatch);  // compact twiddle table
    sbfft::copy_ctw<N, float>(tw_s, tw, sb_tid(), kConvPThreads);

    for (int i = sb_tid(); i < cnt; i += kConvPThreads) {
        const Tmpl* p = tmpls + tmpl_base + i;
        s_meta[4 * i + 0] = p->sy_lo;
        s_meta[4 * i + 1] = p->sy_hi;
        s_meta[4 * i + 2] = p->angle_id - angle_base;
    }

    Ctx c;
    c.t = t;
    c.smA = sm + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw_s;
    c.specA = spec_s + t;
    c.specB = spec_s + N + t;
    c.kx = kx;
    c.dly = g.dly;
    c.out_ny = g.out_ny;
    c.kpitch = g.kpitch;
    c.dbg = g.dbg;
    c.bar_id = 1 + grp;
    sb_sync();

    int s0 = 0;
#pragma unroll 1
    while (s0 < cnt) {
        const int a_loc = s_meta[4 * s0 + 2];
        int s1 = s0 + 1;
        while (s1 < cnt && s_meta[4 * s1 + 2] == a_loc) ++s1;
        if (s0 > 0) sb_sync();                              // the previous run's products are done
        {
            const float4* src = (const float4*)(fct + (((long)a_loc * 2) * KX + kx) * N);
            const float4* src2 = (const float4*)(fct + (((long)a_loc * 2 + 1) * KX + kx) * N);
            float4* dst = (float4*)spec_s;
            for (int i = sb_tid(); i < N / 2; i += kConvPThreads) {
                dst[i] = sb_ld_stream(src + i);
                dst[N / 2 + i] = sb_ld_stream(src2 + i);
            }
        }
        sb_sync();
        // SPARSE: a thread's only inputs are rows t and t - T of the template column; they are
        // fetched one template ahead, so that their latency hides behind the transform
        float4 in_lo = make_float4(0.f, 0.f, 0.f, 0.f), in_hi = in_lo;
        auto fetch_sparse = [&](int p_loc) {
            const int sy_lo = s_meta[4 * p_loc + 0], sy_hi = s_meta[4 * p_loc + 1];
            const float4* src = trt + ((long)p_loc * KX + kx) * g.syp;
            in_lo = make_float4(0.f, 0.f, 0.f, 0.f);
            in_hi = in_lo;
            if (t >= sy_lo && t <= sy_hi) in_lo = sb_ld_stream(src + (t - sy_lo));
            if (t - T >= sy_lo && t - T <= sy_hi) in_hi = sb_ld_stream(src + (t - T - sy_lo));
        };
        if (SPARSE && s0 + grp < s1) fetch_sparse(s0 + grp);
#pragma unroll 1
        for (int p_loc = s0 + grp; p_loc < s1; p_loc += G) {
            c.dst = gbuf + (long)p_loc * N * g.kpitch;
            float2 va[E], vb[E];
#pragma unroll
            for (int q = 0; q < E; ++q) {
                va[q] = make_float2(0.f, 0.f);
                vb[q] = make_float2(0.f, 0.f);
            }
            if (SPARSE) {
                va[0] = make_float2(in_lo.x, in_lo.y);
                vb[0] = make_float2(in_lo.z, in_lo.w);
                va[E - 1] = make_float2(in_hi.x, in_hi.y);
                vb[E - 1] = make_float2(in_hi.z, in_hi.w);
                if (p_loc + G < s1) fetch_sparse(p_loc + G);
            } else {
                const int sy_lo = s_meta[4 * p_loc + 0], sy_hi = s_meta[4 * p_loc + 1];
                const float4* src = trt + ((long)p_loc * KX + kx) * g.syp;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int qy = t + q * T;
                    const int s = qy < N / 2 ? qy : qy - N;
                    if (s >= sy_lo && s <= sy_hi) {
                        const float4 w = sb_ld_stream(src + (s - sy_lo));
                        va[q] = make_float2(w.x, w.y);
                        vb[q] = make_float2(w.z, w.w);
                    }
                }
            }
            c.va_fin = va;
            leapfrog<Ctx::K>(c, va, vb);                    // the last phase of field b writes gbuf
#ifndef SB_CONV_KEEP_A
            c.bar();                                        // parked field a has been read back
#endif
        }
        s0 = s1;
    }
}

// ---------------------------------------------------------------------------
// Fit kernel (complex64 fast path).  Per template and row: Hermitian-extended inverse row
// FFT of Gt + i Gm (real part xcorr, imaginary part T3), amplitude / SNR (core.py:360-367),
// edge mask (core.py:373-375), optional get_err_mask (core.py:369-371), running best-SNR
// select (core.py:198-243) in registers.  The raw-plane mode of match_template and the
// complex128 pipeline use k_fit_rows (sb_kernels.cuh).
// ---------------------------------------------------------------------------
constexpr int kFitMaxBatch = 64;

// X, T: raw outputs of the inverse transform (xcorr and T3 up to the exact power-of-two
// factors folded into FitT).  snr = |T1 / error| with error = (T3 - xcorr^2 / ts) / n + eps
// (core.py:362-367); the cancellation T3 - xcorr^2 / ts is evaluated with error-free products.
SB_DEVICE void fit_pixel_fast(float X, float T, const FitT& k, float& amp, float& snr) {
    amp = X * k.amp_k;                               // core.py:360
    const float pp = X * X;
    const float pe = fmaf(X, X, -pp);                // X*X = pp + pe exactly
    float num = fmaf(-pp, k.a_hi, T);
    num = fmaf(-pp, k.a_lo, num);
    num = fmaf(-pe, k.a_hi, num);
    const float t1 = pp * k.a_hi;                    // core.py:362
    const float err = fmaf(num, k.inv_n, k.eps_k);   // core.py:366
    snr = fabsf(sb_fdiv_fast(t1, err));              // core.py:367
}

// ---------------------------------------------------------------------------
// k_fit_rows_g<Px>: grid (Py / 2 / GP).  The two transforms
// a thread group pipelines are the two raster rows that share every 32-byte sector of the
// interleaved planes (gbuf_index), of ONE template: a thread fetches whole sectors with
// 256-bit loads and feeds one half to each row, so the planes cross the L2 -> SM crossbar
// in full sectors (a one-row kernel uses 16 bytes of every 32 it requests).  The running best keeps
// only the SNR in registers; amplitude and flat index go straight to the best state when a
// pixel improves (rare after the first templates of a sweep).
// ---------------------------------------------------------------------------
template <int N>
struct FitPairCtx {
    static constexpr int K = sbfft::num_stages(N);
    static constexpr int T = N / E;
    static constexpr int LOG2T = ilog2(T);
    int t;
    float2 *smA, *smB;
    const float2* tw;
    const float4* pair;               // the template's row pair: element kk of row F at pair[2 * kk + F]
    const FitT* s_fit;
    int slot;
    int giA, giB;                     // raster rows of the two streams
    bool actA, actB;
    int ox, m0, out_nx, nx;
    float* best_amp;
    int* best_idx;
    const int4* cross;                // get_err_mask column ranges per (angle, raster row), or null
    int ny;
    int dbg;
    float (&bsA)[E];
    float (&bsB)[E];
    float2 (&vb_in)[E];               // stream b's input, filled together with stream a's
    unsigned chg;                     // bit q (+16 for row B): the pixel's best SNR changed

    SB_DEVICE FitPairCtx(float (&a)[E], float (&b)[E], float2 (&vb)[E]) : bsA(a), bsB(b), vb_in(vb), chg(0u) {}

    SB_DEVICE void bar() const { sb_sync(); }

    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) { sbfft::load_tw<N, P, float, true>(w, t, tw); }

    SB_DEVICE unsigned mask(const FitT& k, int gi, bool act) const {
        if (!act || gi < k.i_lo || gi > k.i_hi) return 0u;
        int jl = k.j_lo, jh = k.j_hi;
        if (k.errmode != 0) {
            // snr[get_err_mask()] = 0 (core.py:369-371): a pixel whose SNR is 0 never wins the
            // fold, so inside a search the mask only narrows the row's candidate columns
            const int4 cr = cross[(long)k.angle_id * ny + gi];
            jl = max(jl, k.errmode == 1 ? cr.x : cr.z);
            jh = min(jh, k.errmode == 1 ? cr.y : cr.w);
        }
        const int lo = max(jl - ox, 0), hi = min(jh - ox, out_nx - 1);
        if (lo > hi) return 0u;
        const int qlo = max((lo - m0 + T - 1) >> LOG2T, 0);
        const int qhi = min((hi - m0) >> LOG2T, E - 1);
        return qlo <= qhi ? ((2u << qhi) - (1u << qlo)) : 0u;
    }

    // X(k) = Gt(k) + i Gm(k) and X(N-k) = conj Gt(k) + i conj Gm(k) come ready (and swapped: inverse
    // via forward) in the element (gbuf_pack)
    SB_DEVICE static float2 herm(const float4 g4, bool direct) {
        return direct ? make_float2(g4.x, g4.y) : make_float2(g4.z, g4.w);
    }

    template <int F> SB_DEVICE void epilogue(const float2 (&v)[E], float (&bs)[E]) {
        const int gi = F == 0 ? giA : giB;
        const FitT k = s_fit[slot];
        const unsigned mk = mask(k, gi, F == 0 ? actA : actB);
        float* pa = best_amp + ((long)gi * nx + ox);
        int* pi = best_idx + ((long)gi * nx + ox);
#pragma unroll
        for (int q = 0; q < E; ++q) {
            float amp, snr;
            fit_pixel_fast(v[q].y, v[q].x, k, amp, snr);
            snr = ((mk >> q) & 1u) ? snr : -1.f;              // edge-masked: never wins
            // first maximum wins (core.py:230-240); equal positive SNRs (in float32 mostly the
            // -90 / +90 degree pair) go to the lower flat index, whatever the batch order.
            // (A branch-free select over the 16 pixels was measured 3 % slower: it spills.)
            if (snr >= bs[q] && snr > 0.f) {
                const int jo = (m0 + q * T) & (N - 1);
                if (snr > bs[q] || k.idx < pi[jo]) {
                    bs[q] = snr;
                    pa[jo] = amp;
                    pi[jo] = k.idx;
                    chg |= 1u << (q + 16 * F);
                }
            }
        }
    }

#ifdef SB_F32X2
    // Two pixels (q, q + 8) per packed instruction: X = xcorr pair, Tp = T3 pair (fit_pixel_fast).
    // The window mask is applied to the candidate bits, not to the 16 values; improvements
    // (rare after the first templates of a sweep) are resolved after the common path.
    static constexpr bool SOA = K > 1;
    template <int F> SB_DEVICE void epilogue2(const sbfft::pk_t (&re)[8], const sbfft::pk_t (&im)[8], float (&bs)[E]) {
        using namespace sbfft;
        const int gi = F == 0 ? giA : giB;
        const FitT k = s_fit[slot];
        const unsigned mk = mask(k, gi, F == 0 ? actA : actB);
        float2 snr2[8], amp2[8];
        unsigned cand = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const pk_t X = im[j], Tp = re[j];
            amp2[j] = unpk(mul2(X, pk(k.amp_k, k.amp_k)));                    // core.py:360
            const pk_t pp = mul2(X, X);
            const float2 ppf = unpk(pp);
            const pk_t npp = pk(-ppf.x, -ppf.y);
            const float2 pef = unpk(fma2(X, X, npp));                          // X*X = pp + pe exactly
            pk_t num = fma2(npp, pk(k.a_hi, k.a_hi), Tp);
            num = fma2(npp, pk(k.a_lo, k.a_lo), num);
            num = fma2(pk(-pef.x, -pef.y), pk(k.a_hi, k.a_hi), num);
            const float2 t1 = unpk(mul2(pp, pk(k.a_hi, k.a_hi)));              // core.py:362
            const float2 err = unpk(fma2(num, pk(k.inv_n, k.inv_n), pk(k.eps_k, k.eps_k)));   // core.py:366
            const float s_lo = fabsf(sb_fdiv_fast(t1.x, err.x));               // core.py:367
            const float s_hi = fabsf(sb_fdiv_fast(t1.y, err.y));
            snr2[j] = make_float2(s_lo, s_hi);
            // (a NaN SNR -- NaN in the DEM -- never wins here; k_poison_windows marks those pixels)
            cand |= (s_lo >= bs[j] && s_lo > 0.f) ? (1u << j) : 0u;
            cand |= (s_hi >= bs[j + 8] && s_hi > 0.f) ? (1u << (j + 8)) : 0u;
        }
        cand &= mk;                                                            // edge-masked: never wins
        if (cand != 0u) {
            float* pa = best_amp + ((long)gi * nx + ox);
            int* pi = best_idx + ((long)gi * nx + ox);
            // first maximum wins (core.py:230-240); equal positive SNRs (in float32 mostly the
            // -90 / +90 degree pair) go to the lower flat index, whatever the batch order
#pragma unroll
            for (int q = 0; q < E; ++q) {
                if ((cand >> q) & 1u) {
                    const float snr = q < 8 ? snr2[q & 7].x : snr2[q & 7].y;
                    const float amp = q < 8 ? amp2[q & 7].x : amp2[q & 7].y;
                    const int jo = (m0 + q * T) & (N - 1);
                    if (snr > bs[q] || k.idx < pi[jo]) {
                        bs[q] = snr;
                        pa[jo] = amp;
                        pi[jo] = k.idx;
                        chg |= 1u << (q + 16 * F);
                    }
                }
            }
        }
    }
#else
    static constexpr bool SOA = false;
#endif

    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        if constexpr (P == 0 && F == 0) {
#pragma unroll
            for (int q = 0; q < E; ++q) {
                bool direct = q < E / 2;
                int kk = q < E / 2 ? t + q * T : N - (t + q * T);
                if (q == E / 2) { direct = t == 0; kk = direct ? N / 2 : N / 2 - t; }
                float4 ga, gb;
                if (SB_DBG_ON(dbg, 16)) { ga = make_float4(1.f, 2.f, 3.f, (float)q); gb = ga; }
                else sb_ld_sector(pair + kGbufRows * kk, ga, gb);
                v[q] = herm(ga, direct);
                vb_in[q] = herm(gb, direct);
            }
        }
#ifdef SB_F32X2
        if constexpr (P == K - 1 && SOA) {
            sbfft::pk_t re[8], im[8];
            sbfft::stage_math_soa<N, P>(v, w, re, im);
            if (SB_DBG_ON(dbg, 64) && v[0].x != 1.2345e-30f) return;
            if constexpr (F == 0) epilogue2<0>(re, im, bsA);
            else epilogue2<1>(re, im, bsB);
            return;
        }
#endif
        sbfft::stage_math<N, P, float>(v, w);
        if constexpr (P == K - 1) {
            if (SB_DBG_ON(dbg, 64) && v[0].x != 1.2345e-30f) return;
            if constexpr (F == 0) epilogue<0>(v, bsA);
            else epilogue<1>(v, bsB);
        }
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        sbfft::stage_store<N, P, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) { sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB); }
};

template <int N>
SB_GLOBAL SB_LAUNCH_BOUNDS((N / E > 256 ? N / E : 256), (N / E > 256 ? 1 : 2))
k_fit_rows_g(Geom g, int count, const int* SB_RESTRICT slots, const FitT* SB_RESTRICT fit,
             const float4* SB_RESTRICT gbuf, float* SB_RESTRICT best_snr, float* SB_RESTRICT best_amp, int* best_idx,
             const float2* SB_RESTRICT tw, const int4* SB_RESTRICT cross, long sub_stride) {
    // grid.y > 1 (small rasters, whose row pairs alone do not fill the GPU): the templates of the
    // launch are dealt round-robin to grid.y sub-streams; sub-stream j folds into copy j of the
    // best state (best_* + j * sub_stride; copy 0 is the state itself) and k_best_fold merges the
    // copies after the sweep.
    // `slots` (optional): the `count` batch slots this launch folds -- the templates of one best
    // state (template scale) inside a batch that mixes several; null: slots 0 .. count - 1.
    // best_*: the state's planes, offset so that raster row gi is at gi * nx (row slabs).
    // `cross` (templates with get_err_mask only): per (angle, raster row) the column ranges
    // with xr > 0 (.x .. .y) and xr < 0 (.z .. .w), k_err_cross.
    constexpr int T = N / E;
    constexpr int THREADS = T > 256 ? T : 256;
    constexpr int GP = THREADS / T;
    constexpr int PL = sbfft::padded_len(N);
    typedef FitPairCtx<N> Ctx;
    const int grp = sb_tid() / T, t = sb_tid() % T;
    float2* sm = (float2*)sb_shared();
    FitT* s_fit = (FitT*)(sm + (long)GP * 2 * PL);
    // [kFitMaxBatch] active templates, [kFitMaxBatch] flags (16-byte aligned: the compiler reads
    // them with vector loads), the count -- s_list[2 * kFitMaxBatch] --, [kFitMaxBatch] gbuf slots
    int* s_list = (int*)(s_fit + kFitMaxBatch);
    int* s_flag = s_list + kFitMaxBatch;
    int* s_gslot = s_flag + kFitMaxBatch + 2;
    float2* tw_s = (float2*)(s_gslot + kFitMaxBatch);
    sbfft::copy_ctw<N, float>(tw_s, tw, sb_tid(), THREADS);
    const int pairs = g.Py / 2;
    const int pr = min(sb_bx() * GP + grp, pairs - 1);    // row pair of the FFT domain
    const int ioA = (2 * pr + g.dly) & (g.Py - 1), ioB = (2 * pr + 1 + g.dly) & (g.Py - 1);
    const bool actA = sb_bx() * GP + grp < pairs && ioA < g.out_ny;
    const bool actB = sb_bx() * GP + grp < pairs && ioB < g.out_ny;

    // stage the batch's scalars; list the templates whose window meets one of this CTA's rows
    if (sb_tid() < count) {
        const int slot = slots ? slots[sb_tid()] : sb_tid();
        const FitT k = fit[slot];
        s_fit[sb_tid()] = k;
        s_gslot[sb_tid()] = slot;
        int flag = 0;
        for (int r = 0; r < 2 * GP && sb_tid() % sb_nby() == sb_by(); ++r) {
            const int m = 2 * sb_bx() * GP + r;
            const int io = (m + g.dly) & (g.Py - 1);
            if (m < g.Py && io < g.out_ny && g.oy + io >= k.i_lo && g.oy + io <= k.i_hi) flag = 1;
        }
        s_flag[sb_tid()] = flag;
    }
    sb_sync();
    if (sb_tid() < count) {
        int pos = 0;
        for (int i = 0; i < sb_tid(); ++i) pos += s_flag[i];
        if (s_flag[sb_tid()]) s_list[pos] = sb_tid();
        if (sb_tid() == count - 1) s_list[2 * kFitMaxBatch] = pos + s_flag[sb_tid()];
    }
    sb_sync();
    const int n_act = s_list[2 * kFitMaxBatch];

    float bsA[E], bsB[E];
    float2 va[E], vb[E];
    Ctx c(bsA, bsB, vb);
    c.t = t;
    c.smA = sm + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw_s;
    c.giA = g.oy + ioA;
    c.giB = g.oy + ioB;
    c.actA = actA;
    c.actB = actB;
    c.ox = g.ox;
    c.m0 = t + g.dlx;
    c.out_nx = g.out_nx;
    c.nx = g.nx;
    c.s_fit = s_fit;
    c.dbg = g.dbg;
    best_snr += sb_by() * sub_stride;
    best_amp += sb_by() * sub_stride;
    best_idx += sb_by() * sub_stride;
    c.best_amp = best_amp;
    c.best_idx = best_idx;
    c.cross = cross;
    c.ny = g.ny;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int jo = (t + q * T + g.dlx) & (N - 1);
        const bool in = jo < g.out_nx;
        bsA[q] = (actA && in) ? best_snr[(long)c.giA * g.nx + g.ox + jo] : 0.f;
        bsB[q] = (actB && in) ? best_snr[(long)c.giB * g.nx + g.ox + jo] : 0.f;
    }
    const long row_off = gbuf_index(2 * pr, 0, g.kpitch);
    const long tmpl_pitch = (long)g.Py * g.kpitch;
    // 128-byte lines that hold the row pair (with kGbufRows > 2 they are shared with the
    // neighbouring pairs, whose CTAs run at the same time)
    const int lines = (kGbufRows * (g.Px / 2 + 1) + 7) / 8;
    const long pf_off = gbuf_index(2 * pr - (2 * pr) % kGbufRows, 0, g.kpitch);

#pragma unroll 1
    for (int i = 0; i < n_act; ++i) {
        const int p = s_list[i];
        c.pair = gbuf + s_gslot[p] * tmpl_pitch + row_off;
        c.slot = p;
        if (i + 1 < n_act) {                                // next template's row pair towards L2
            const float4* nx0 = gbuf + s_gslot[s_list[i + 1]] * tmpl_pitch + pf_off;
            for (int l = t; l < lines; l += T) sb_prefetch_l2(nx0 + 8 * l);
        }
        // no barrier between templates: buffer A was last read before the final barrier,
        // buffer B is next written after the coming template's first barrier
        leapfrog<Ctx::K>(c, va, vb);
    }
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int jo = (t + q * T + g.dlx) & (N - 1);
        if ((c.chg >> q) & 1u) best_snr[(long)c.giA * g.nx + g.ox + jo] = bsA[q];
        if ((c.chg >> (q + 16)) & 1u) best_snr[(long)c.giB * g.nx + g.ox + jo] = bsB[q];
    }
}

// ---------------------------------------------------------------------------
// k_curv_rows_f<Px>: pipelined k_curv_rows (complex64): the two rows of a pair are the two
// transforms a thread group carries (one barrier per exchange); grid and layouts as k_curv_rows.
// ---------------------------------------------------------------------------
template <int N>
struct CurvCtx {
    static constexpr int K = sbfft::num_stages(N);
    static constexpr int T = N / E;
    int t;
    float2 *smA, *smB;
    const float2* tw;
    const float4* diffs;               // [pixel][dxx, dxy, dyy, -] float32
    float ca2, sc2, sa2;               // cos**2, 2 sin cos, sin**2 of the search angle
    int oy, ox, ny, nx, need_y_lo, need_x_lo, need_x_hi, split_x, poison, dbg;
    int row0;                          // global raster row of row 0 of `diffs` (row slab; 0 for a whole raster)
    float c2_scale;
    int rp;                            // row pair: rows 2 rp (stream a) and 2 rp + 1 (stream b)
    int need_rows;

    SB_DEVICE void bar() const { sb_sync(); }
    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) { sbfft::load_tw<N, P, float>(w, t, tw); }

    template <int F> SB_DEVICE void fill(float2 (&v)[E]) const {
        const int r = 2 * rp + F;
        const bool active = r < need_rows;
        int gi = wrap(oy + need_y_lo + (active ? r : 0), ny) - row0;
        gi += gi < 0 ? ny : 0;                   // row of the slab
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int qx = t + q * T;
            const int sx = qx < split_x ? qx : qx - N;
            float2 val = make_float2(0.f, 0.f);
            if (active && sx >= need_x_lo && sx <= need_x_hi) {
                const int gj = wrap_near(ox + sx, nx);
                const long o = (long)gi * nx + gj;
                // dem.py:103-104 on the float32 second differences: the transform rounds the
                // curvature to float32 anyway, and one 16-byte load replaces three 8-byte ones
                const float4 d = SB_DBG_ON(dbg, 256) ? make_float4((float)o * 1e-9f, 1.f, 2.f, 0.f) : sb_ldg(diffs + o);
                const float c = (d.x * ca2 - d.y * sc2) + d.z * sa2;
                val = make_float2(c, c * c * c2_scale);                    // curv, curv**2 (core.py:355)
            }
            if (poison && r == 0 && qx == 0) val = make_float2(NAN, NAN);
            v[q] = val;
        }
    }
    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        sbfft::stage_math<N, P, float>(v, w);       // both rows are filled before the pipeline starts
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        sbfft::stage_store<N, P, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) { sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB); }
};

// Hermitian split (sb_kernels.cuh) of the spectrum parked in natural order in `sm`
template <int N>
SB_DEVICE void split_from_shared(const float2* sm, int t, float4 (&h)[E / 2 + 1]) {
    constexpr int T = N / E;
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
        const int k = t + q * T;
        const float2 a = sm[sbfft::pad_index(k)];
        const float2 zp = sm[sbfft::pad_index((N - k) & (N - 1))];
        h[q] = make_float4(0.5f * (a.x + zp.x), 0.5f * (a.y - zp.y), 0.5f * (a.y + zp.y), -0.5f * (a.x - zp.x));
    }
    const float2 ny = sm[sbfft::pad_index(N / 2)];
    h[E / 2] = make_float4(ny.x, 0.f, ny.y, 0.f);
}

template <int N>
SB_GLOBAL SB_LAUNCH_BOUNDS((N / E > 256 ? N / E : 256), (N / E > 256 ? 1 : 2))
k_curv_rows_f(Geom g, const float4* SB_RESTRICT diffs, const Angle* SB_RESTRICT angles, int angle_base,
              float4* SB_RESTRICT cr, const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int GP = (T > 256 ? T : 256) / T;
    constexpr int PL = sbfft::padded_len(N);
    const int grp = sb_tid() / T, t = sb_tid() % T;
    const int KX = N / 2 + 1;
    const int a_loc = sb_bx();
    CurvCtx<N> c;
    c.t = t;
    c.smA = (float2*)sb_shared() + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw;
    c.diffs = diffs;
    c.oy = g.oy; c.ox = g.ox; c.ny = g.ny; c.nx = g.nx;
    c.need_y_lo = g.need_y_lo; c.need_x_lo = g.need_x_lo; c.need_x_hi = g.need_x_hi;
    c.split_x = g.split_x; c.poison = g.poison; c.c2_scale = (float)g.c2_scale; c.dbg = g.dbg;
    c.row0 = g.row0;
    {
        const Angle ang = angles[angle_base + a_loc];
        c.ca2 = (float)ang.ca2;
        c.sc2 = (float)(2.0 * ang.sa * ang.ca);
        c.sa2 = (float)ang.sa2;
    }
    c.rp = sb_by() * GP + grp;
    c.need_rows = g.need_y_hi - g.need_y_lo + 1;
    float2 va[E], vb[E];
    c.template fill<0>(va);                     // all 32 loads of the pair in flight together
    c.template fill<1>(vb);
    leapfrog<CurvCtx<N>::K>(c, va, vb);
    sb_sync();                                  // every thread is done with the exchange buffers
#pragma unroll
    for (int q = 0; q < E; ++q) {
        c.smA[sbfft::pad_index(t + q * T)] = va[q];
        c.smB[sbfft::pad_index(t + q * T)] = vb[q];
    }
    sb_sync();
    float4 h0[E / 2 + 1], h1[E / 2 + 1];
    split_from_shared<N>(c.smA, t, h0);
    split_from_shared<N>(c.smB, t, h1);
    if (2 * c.rp < c.need_rows)
        store_row_pair<N, float>(h0, h1, t, cr + (long)a_loc * KX * g.rpitch + 2 * c.rp, g.rpitch);
}

// ---------------------------------------------------------------------------
// Curvature spectra without one 2-D transform per orientation.  The directional Laplacian is
// linear in the three second differences (dem.py:103-104):
//     curv(a)   = c2 dxx - sc dxy + s2 dyy                      (c2 = cos^2 a, sc = 2 sin a cos a, s2 = sin^2 a)
//     curv(a)^2 = c2^2 dxx^2 + sc^2 dxy^2 + s2^2 dyy^2 - 2 c2 sc dxx dxy + 2 c2 s2 dxx dyy - 2 sc s2 dxy dyy
// and so is the Fourier transform: the spectra of the NINE planes (3 differences, 6 products)
// are computed once per FFT tile (five packed pairs through k_diff_rows_f / k_curv_cols) and
// every orientation's F[curv], F[curv^2] are combinations of them with nine scalars --
// k_combine_spectra, an elementwise pass at HBM speed -- instead of a row kernel and a column
// kernel per orientation.  Plane order (pair pp, field f -> plane 2 pp + f):
//   0 dxx   1 dxy   2 dyy   3 s dxx^2   4 s dxy^2   5 s dyy^2   6 s dxx dxy   7 s dxx dyy   8 s dxy dyy   9 unused
// (s = c2_scale, the power of two that brings the quadratic planes to the magnitude of the
// linear ones in the packed transforms).
// ---------------------------------------------------------------------------
constexpr int kDiffPairs = 5;
constexpr int kDiffPlanes = 9;

template <int N>
struct DiffCtx {
    static constexpr int K = sbfft::num_stages(N);
    static constexpr int T = N / E;
    int t;
    float2 *smA, *smB;
    const float2* tw;
    const float4* diffs;               // [pixel][dxx, dxy, dyy, -] float32
    int pp;                            // plane pair
    int oy, ox, ny, nx, need_y_lo, need_x_lo, need_x_hi, split_x, poison;
    int row0;
    float c2_scale;
    int rp;                            // row pair: rows 2 rp (stream a) and 2 rp + 1 (stream b)
    int need_rows;

    SB_DEVICE void bar() const { sb_sync(); }
    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) { sbfft::load_tw<N, P, float>(w, t, tw); }

    SB_DEVICE float2 planes(const float4 d) const {
        const float s = c2_scale;
        switch (pp) {
            case 0: return make_float2(d.x, d.y);
            case 1: return make_float2(d.z, d.x * d.x * s);
            case 2: return make_float2(d.y * d.y * s, d.z * d.z * s);
            case 3: return make_float2(d.x * d.y * s, d.x * d.z * s);
            default: return make_float2(d.y * d.z * s, 0.f);
        }
    }
    template <int F> SB_DEVICE void fill(float2 (&v)[E]) const {
        const int r = 2 * rp + F;
        const bool active = r < need_rows;
        int gi = wrap(oy + need_y_lo + (active ? r : 0), ny) - row0;
        gi += gi < 0 ? ny : 0;                   // row of the slab
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int qx = t + q * T;
            const int sx = qx < split_x ? qx : qx - N;
            float2 val = make_float2(0.f, 0.f);
            if (active && sx >= need_x_lo && sx <= need_x_hi) {
                const int gj = wrap_near(ox + sx, nx);
                val = planes(sb_ldg(diffs + (long)gi * nx + gj));
            }
            // a NaN in the DEM reaches every output pixel (dem.py:105 through the reference's fft2)
            if (poison && pp == 0 && r == 0 && qx == 0) val = make_float2(NAN, NAN);
            v[q] = val;
        }
    }
    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        sbfft::stage_math<N, P, float>(v, w);
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        sbfft::stage_store<N, P, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) { sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB); }
};

// grid (kDiffPairs, ceil(need_rows / 2 / GP)); output layout as k_curv_rows_f with the plane
// pair in the place of the orientation: cr[pp][kx][rpitch rows] float4
template <int N>
SB_GLOBAL SB_LAUNCH_BOUNDS((N / E > 256 ? N / E : 256), (N / E > 256 ? 1 : 2))
k_diff_rows_f(Geom g, const float4* SB_RESTRICT diffs, float4* SB_RESTRICT cr, const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int GP = (T > 256 ? T : 256) / T;
    constexpr int PL = sbfft::padded_len(N);
    const int grp = sb_tid() / T, t = sb_tid() % T;
    const int KX = N / 2 + 1;
    DiffCtx<N> c;
    c.t = t;
    c.smA = (float2*)sb_shared() + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw;
    c.diffs = diffs;
    c.pp = sb_bx();
    c.oy = g.oy; c.ox = g.ox; c.ny = g.ny; c.nx = g.nx;
    c.need_y_lo = g.need_y_lo; c.need_x_lo = g.need_x_lo; c.need_x_hi = g.need_x_hi;
    c.split_x = g.split_x; c.poison = g.poison; c.c2_scale = (float)g.c2_scale;
    c.row0 = g.row0;
    c.rp = sb_by() * GP + grp;
    c.need_rows = g.need_y_hi - g.need_y_lo + 1;
    float2 va[E], vb[E];
    c.template fill<0>(va);
    c.template fill<1>(vb);
    leapfrog<DiffCtx<N>::K>(c, va, vb);
    sb_sync();                                  // every thread is done with the exchange buffers
#pragma unroll
    for (int q = 0; q < E; ++q) {
        c.smA[sbfft::pad_index(t + q * T)] = va[q];
        c.smB[sbfft::pad_index(t + q * T)] = vb[q];
    }
    sb_sync();
    float4 h0[E / 2 + 1], h1[E / 2 + 1];
    split_from_shared<N>(c.smA, t, h0);
    split_from_shared<N>(c.smB, t, h1);
    if (2 * c.rp < c.need_rows)
        store_row_pair<N, float>(h0, h1, t, cr + (long)c.pp * KX * g.rpitch + 2 * c.rp, g.rpitch);
}

// per-orientation coefficients of the nine planes (host: float32 of the float64 cos / sin values)
struct SpecCoef { float c[kDiffPlanes + 3]; };      // padded to 48 bytes

// fct[a][field][kx][ky] = sum_p coef[a][p] * spec9[p][kx][ky] for the orientations a0 .. a0 + na - 1.
// One thread per pair of adjacent ky (float4): nine 16-byte loads, then two 16-byte stores per
// orientation.  n2 = planes' size in float4 units (KX * Py / 2).
SB_GLOBAL k_combine_spectra(long n2, int na, int a0, const SpecCoef* SB_RESTRICT coef, const float4* SB_RESTRICT spec9,
                            float4* SB_RESTRICT fct) {
    SpecCoef* sc = (SpecCoef*)sb_shared();
    for (int i = sb_tid(); i < na * (int)(sizeof(SpecCoef) / 4); i += 256) ((float*)sc)[i] = ((const float*)(coef + a0))[i];
    sb_sync();
    const long i = (long)sb_bx() * 256 + sb_tid();
    if (i >= n2) return;
    float4 p[kDiffPlanes];
#pragma unroll
    for (int k = 0; k < kDiffPlanes; ++k) p[k] = sb_ld_stream(spec9 + (long)k * n2 + i);
    // explicit fused multiply-adds in a fixed order: an orientation's spectra must not depend on
    // its position in the batch (the sharded searches are bit-identical to the single-GPU one)
#pragma unroll 1
    for (int a = 0; a < na; ++a) {
        const SpecCoef& w = sc[a];
        float4 A, B;
        A.x = fmaf(w.c[2], p[2].x, fmaf(w.c[1], p[1].x, w.c[0] * p[0].x));
        A.y = fmaf(w.c[2], p[2].y, fmaf(w.c[1], p[1].y, w.c[0] * p[0].y));
        A.z = fmaf(w.c[2], p[2].z, fmaf(w.c[1], p[1].z, w.c[0] * p[0].z));
        A.w = fmaf(w.c[2], p[2].w, fmaf(w.c[1], p[1].w, w.c[0] * p[0].w));
        B.x = w.c[3] * p[3].x; B.y = w.c[3] * p[3].y; B.z = w.c[3] * p[3].z; B.w = w.c[3] * p[3].w;
#pragma unroll
        for (int k = 4; k < kDiffPlanes; ++k) {
            B.x = fmaf(w.c[k], p[k].x, B.x); B.y = fmaf(w.c[k], p[k].y, B.y);
            B.z = fmaf(w.c[k], p[k].z, B.z); B.w = fmaf(w.c[k], p[k].w, B.w);
        }
        sb_st_stream(fct + (long)(2 * a) * n2 + i, A);
        sb_st_stream(fct + (long)(2 * a + 1) * n2 + i, B);
    }
}

// ---------------------------------------------------------------------------
// k_poison_windows: only launched when the DEM holds a NaN (dem.py:105).  Every FFT output is
// NaN then, and compare's arithmetic select (core.py:230-240: 0 * best + 0 * NaN) leaves NaN
// wherever a template is un-masked; the pipelined fit kernels never let a NaN win, so the
// pixels inside the windows of the batch's templates are marked here.  A NaN best SNR then
// loses no later comparison.
// ---------------------------------------------------------------------------
SB_GLOBAL k_poison_windows(Geom g, int count, const int* SB_RESTRICT slots, const FitT* SB_RESTRICT fit,
                           float* SB_RESTRICT best_snr) {
    const long i = (long)sb_bx() * 256 + sb_tid();
    if (i >= (long)g.out_ny * g.out_nx) return;
    const int gi = g.oy + (int)(i / g.out_nx), gj = g.ox + (int)(i % g.out_nx);
    bool hit = false;
    for (int p = 0; p < count && !hit; ++p) {
        const FitT k = fit[slots ? slots[p] : p];
        hit = gi >= k.i_lo && gi <= k.i_hi && gj >= k.j_lo && gj <= k.j_hi;
    }
    if (hit) best_snr[(long)gi * g.nx + gj] = NAN;
}

}  // namespace sb
