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
            sb_sync();
            c.template phase<P + 1, 0>(va, wa);
            if constexpr (P + 1 < K - 1) c.template store<P + 1, 0>(va);
            c.template load<1>(vb);
            c.template twid<P + 1>(wb);
            if constexpr (P + 1 < K - 1) sb_sync();
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
    sb_sync();
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
template <int N, bool SPARSE>
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

    SB_CONSTEXPR static int stage_of(int P) { return P < NST ? P : P - NST + 1; }

    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) {
        if (SB_DBG_ON(dbg, 4)) {
#pragma unroll
            for (int i = 0; i < TW; ++i) w[i] = make_float2(0.6f, 0.8f);
            return;
        }
        sbfft::load_tw<N, stage_of(P), float>(w, t, tw);
    }

    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        if constexpr (P == 0 && SPARSE) sbfft::stage0_sparse2<float>(v);
        else sbfft::stage_math<N, stage_of(P), float>(v, w);
        if constexpr (P == NST - 1) {
            const float2* spec = F == 0 ? specA : specB;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const float2 s = SB_DBG_ON(dbg, 2) ? make_float2(0.5f, 0.25f) : sb_ld_stream(spec + q * T);
                const float2 pr = sbfft::cmul(v[q], s);
                v[q] = make_float2(pr.y, pr.x);                  // swap: inverse via forward
            }
            sbfft::stage_math<N, 0, float>(v, w);
        }
        if constexpr (P == K - 1 && F == 0) {
            // field a is done; park it in its own (now idle) exchange buffer, every thread in
            // the slots only it will read back, so that its registers are free for field b
#pragma unroll
            for (int q = 0; q < E; ++q) smA[t + q * T] = v[q];
        }
        if constexpr (P == K - 1 && F == 1) {
            if (dst) {
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int io = (t + q * T + dly) & (N - 1);
                    const float2 a = smA[t + q * T];
                    if (SB_DBG_ON(dbg, 1) && v[q].x != 1.2345e-30f) continue;
                    if (io < out_ny) dst[gbuf_index(t + q * T, kx, kpitch)] = make_float4(a.y, a.x, v[q].y, v[q].x);
                }
            }
        }
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        constexpr int S = (P == NST - 1) ? 0 : stage_of(P);
        sbfft::stage_store<N, S, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) { sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB); }
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
    leapfrog<ConvCtx<N, SPARSE>::K>(c, va, vb);     // the last phase of field b writes gbuf
}

// ---------------------------------------------------------------------------
// k_fit_rows_f<Px>: grid (ceil(out_ny / GP)); every CTA walks the batch of templates
// two at a time.  Per template and row: Hermitian-extended inverse row FFT of
// Gt + i Gm (real part xcorr, imaginary part T3), amplitude / SNR (core.py:360-367),
// edge mask (core.py:373-375), running best-SNR select (core.py:198-243) in
// registers; the best state is read and written once per launch.
// Templates with get_err_mask (core.py:369-371) and the raw-plane mode of
// match_template use k_fit_rows.
// ---------------------------------------------------------------------------
constexpr int kFitMaxBatch = 64;

SB_DEVICE void fit_pixel_fast(float xraw, float traw, const FitT& k, float& amp, float& snr) {
    const float x = xraw * k.xn;                     // xcorr  (exact scaling)
    const float t3 = traw * k.tn;                    // T3     (exact scaling)
    amp = x * k.its_hi;                              // core.py:360
    const float pp = x * x;
    const float pe = fmaf(x, x, -pp);                // x*x = pp + pe exactly
    float num = fmaf(-pp, k.its_hi, t3);             // T3 - x^2/ts, error-free products
    num = fmaf(-pp, k.its_lo, num);
    num = fmaf(-pe, k.its_hi, num);
    const float t1 = x * amp;                        // core.py:362
    const float err = fmaf(num, k.inv_n, (float)kEps);   // core.py:366
    snr = fabsf(sb_fdiv_fast(t1, err));              // core.py:367
}

template <int N, bool PAIRED>
struct FitCtx {
    static constexpr int K = sbfft::num_stages(N);
    static constexpr int T = N / E;
    static constexpr int LOG2T = ilog2(T);
    int t;
    float2 *smA, *smB;
    const float2* tw;
    const float4 *rowA, *rowB;        // gbuf rows (nullptr: nothing to do for this stream)
    const FitT* s_fit;                // shared-memory copies of the batch's scalars
    int slotA, slotB;                 // batch slots of the two templates in flight
    int gi, ox, m0;                   // raster row, tile origin, t + dlx
    int out_nx, nx;
    bool row_active;
    int dbg;
    const int* best_idx;
    // running best per pixel: SNR, amplitude, and (one byte each) the batch slot of the
    // template that set them -- kNoSlot while the value read from the best state stands
    float (&bs)[E];
    float (&ba)[E];
    unsigned (&bw)[E / 4];
    static constexpr unsigned kNoSlot = 0xFFu;

    SB_DEVICE FitCtx(float (&s)[E], float (&a)[E], unsigned (&w)[E / 4]) : bs(s), ba(a), bw(w) {}

    SB_DEVICE unsigned slot_of(int q) const { return (bw[q >> 2] >> (8 * (q & 3))) & 0xFFu; }
    SB_DEVICE void set_slot(int q, unsigned slot) {
        bw[q >> 2] = (bw[q >> 2] & ~(0xFFu << (8 * (q & 3)))) | (slot << (8 * (q & 3)));
    }
    // flat index behind the current best of element q (rare path: exact SNR ties only)
    SB_DEVICE int current_idx(int q) const {
        const unsigned s = slot_of(q);
        if (s != kNoSlot) return s_fit[s].idx;
        const int jo = (m0 + q * T) & (N - 1);
        return best_idx[(long)gi * nx + ox + jo];
    }

    template <int P> SB_DEVICE void twid(float2 (&w)[TW]) {
        if (SB_DBG_ON(dbg, 32)) {
#pragma unroll
            for (int i = 0; i < TW; ++i) w[i] = make_float2(0.6f, 0.8f);
            return;
        }
        sbfft::load_tw<N, P, float>(w, t, tw);
    }

    // bit q set <=> column t + q*T of this row lies inside the template's window
    SB_DEVICE unsigned mask(const FitT& k) const {
        if (!row_active || gi < k.i_lo || gi > k.i_hi) return 0u;
        const int lo = max(k.j_lo - ox, 0), hi = min(k.j_hi - ox, out_nx - 1);
        const int qlo = max((lo - m0 + T - 1) >> LOG2T, 0);
        const int qhi = min((hi - m0) >> LOG2T, E - 1);
        return qlo <= qhi ? ((2u << qhi) - (1u << qlo)) : 0u;
    }

    template <int P, int F> SB_DEVICE void phase(float2 (&v)[E], const float2 (&w)[TW]) {
        if constexpr (P == 0) {
            // X(k) = Gt(k) + i Gm(k);  X(N-k) = conj Gt(k) + i conj Gm(k); stored swapped
            const float4* row = F == 0 ? rowA : rowB;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                bool direct = q < E / 2;
                int kk = q < E / 2 ? t + q * T : N - (t + q * T);
                if (q == E / 2) { direct = t == 0; kk = direct ? N / 2 : N / 2 - t; }
                // PAIRED: the CTA's other row group reads the other half of each sector
                if (row) g4 = SB_DBG_ON(dbg, 16) ? make_float4(1.f, 2.f, 3.f, (float)q)
                              : PAIRED ? sb_ld_shared_soon(row + 2 * kk) : sb_ld_stream(row + 2 * kk);
                v[q] = direct ? make_float2(g4.y + g4.z, g4.x - g4.w) : make_float2(g4.z - g4.y, g4.x + g4.w);
            }
        }
        sbfft::stage_math<N, P, float>(v, w);
        if constexpr (P == K - 1) {
            if ((F == 0 ? rowA : rowB) != nullptr) {
                const unsigned slot = F == 0 ? slotA : slotB;
                const FitT k = s_fit[slot];
                if (SB_DBG_ON(dbg, 64) && v[0].x != 1.2345e-30f) return;
                const unsigned mk = mask(k);
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    float amp, snr;
                    fit_pixel_fast(v[q].y, v[q].x, k, amp, snr);
                    // first maximum wins (core.py:230-240); equal positive SNRs resolve to the
                    // lower flat index so the result does not depend on batch order
                    const bool ok = (mk >> q) & 1u;
                    bool take = ok && snr > bs[q];
                    if (ok && snr == bs[q] && snr > 0.f) take = k.idx < current_idx(q);
                    if (take) {
                        bs[q] = snr;
                        ba[q] = amp;
                        set_slot(q, slot);
                    }
                }
            }
        }
    }
    template <int P, int F> SB_DEVICE void store(const float2 (&v)[E]) {
        sbfft::stage_store<N, P, float>(v, t, F == 0 ? smA : smB);
    }
    template <int F> SB_DEVICE void load(float2 (&v)[E]) { sbfft::stage_load<N, float>(v, t, F == 0 ? smA : smB); }
};

// MINT: threads per CTA when one row needs fewer.  512 puts the two rows that share every
// 32-byte sector of the interleaved planes into the same CTA.
template <int N, int MINT>
SB_GLOBAL SB_LAUNCH_BOUNDS((N / E > MINT ? N / E : MINT), ((N / E > MINT ? N / E : MINT) > 256 ? 1 : 2))
k_fit_rows_f(Geom g, int count, const FitT* SB_RESTRICT fit, const float4* SB_RESTRICT gbuf,
             float* SB_RESTRICT best_snr, float* SB_RESTRICT best_amp, int* best_idx,
             const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int THREADS = T > MINT ? T : MINT;
    constexpr int GP = THREADS / T;
    constexpr int PL = sbfft::padded_len(N);
    typedef FitCtx<N, (GP > 1)> Ctx;
    const int grp = sb_tid() / T, t = sb_tid() % T;
    float2* sm = (float2*)sb_shared();
    FitT* s_fit = (FitT*)(sm + (long)GP * 2 * PL);
    int* s_list = (int*)(s_fit + kFitMaxBatch);          // [kFitMaxBatch] active templates, then their count
    int* s_flag = s_list + kFitMaxBatch + 1;             // [kFitMaxBatch]
    const int io = sb_bx() * GP + grp;
    const bool active = io < g.out_ny;
    const int gi = g.oy + io;
    const int cta_lo = g.oy + sb_bx() * GP;
    const int cta_hi = min(cta_lo + GP - 1, g.oy + g.out_ny - 1);

    // stage the batch's scalars; list the templates whose window meets this CTA's rows
    if (sb_tid() < count) {
        const FitT k = fit[sb_tid()];
        s_fit[sb_tid()] = k;
        s_flag[sb_tid()] = !(cta_hi < k.i_lo || cta_lo > k.i_hi);
    }
    sb_sync();
    if (sb_tid() < count) {
        int pos = 0;
        for (int i = 0; i < sb_tid(); ++i) pos += s_flag[i];
        if (s_flag[sb_tid()]) s_list[pos] = sb_tid();
        if (sb_tid() == count - 1) s_list[kFitMaxBatch] = pos + s_flag[sb_tid()];
    }
    sb_sync();
    const int n_act = s_list[kFitMaxBatch];

    float bs[E], ba[E];
    unsigned bw[E / 4];
    Ctx c(bs, ba, bw);
    c.t = t;
    c.smA = sm + (long)grp * 2 * PL;
    c.smB = c.smA + PL;
    c.tw = tw;
    c.gi = gi;
    c.ox = g.ox;
    c.m0 = t + g.dlx;
    c.out_nx = g.out_nx;
    c.nx = g.nx;
    c.row_active = active;
    c.dbg = g.dbg;
    c.s_fit = s_fit;
    c.best_idx = best_idx;
#pragma unroll
    for (int q = 0; q < E / 4; ++q) bw[q] = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int jo = (t + q * T + g.dlx) & (N - 1);
        bs[q] = 0.f; ba[q] = 0.f;
        if (active && jo < g.out_nx) {
            const long o = (long)gi * g.nx + g.ox + jo;
            bs[q] = best_snr[o];
            ba[q] = best_amp[o];
        }
    }
    // this row's half of its interleaved row pair (gbuf_index)
    const long row_off = gbuf_index(((active ? io : 0) - g.dly) & (g.Py - 1), 0, g.kpitch);
    const long tmpl_pitch = (long)g.Py * g.kpitch;
    const int lines = (g.Px / 2 + 1 + 3) / 4;              // 128-byte lines the row's elements lie in

#pragma unroll 1
    for (int i = 0; i < n_act; i += 2) {
        const int pa = s_list[i];
        const int pb = i + 1 < n_act ? s_list[i + 1] : -1;
        c.rowA = active ? gbuf + pa * tmpl_pitch + row_off : nullptr;
        c.rowB = (active && pb >= 0) ? gbuf + pb * tmpl_pitch + row_off : nullptr;
        c.slotA = pa;
        c.slotB = pb >= 0 ? pb : pa;
        if (active && i + 2 < n_act) {                      // next pair's rows towards L2
            const float4* nx0 = gbuf + s_list[i + 2] * tmpl_pitch + row_off;
            for (int l = t; l < lines; l += T) sb_prefetch_l2(nx0 + 8 * l);
            if (i + 3 < n_act) {
                const float4* nx1 = gbuf + s_list[i + 3] * tmpl_pitch + row_off;
                for (int l = t; l < lines; l += T) sb_prefetch_l2(nx1 + 8 * l);
            }
        }
        float2 va[E], vb[E];
        // no barrier between pairs: buffer A was last read before the pair's final barrier,
        // buffer B is next written after the coming pair's first barrier
        leapfrog<Ctx::K>(c, va, vb);
    }
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int jo = (t + q * T + g.dlx) & (N - 1);
        const unsigned slot = c.slot_of(q);
        if (active && jo < g.out_nx && slot != Ctx::kNoSlot) {
            const long o = (long)gi * g.nx + g.ox + jo;
            best_snr[o] = bs[q];
            best_amp[o] = ba[q];
            best_idx[o] = s_fit[slot].idx;
        }
    }
}

}  // namespace sb
