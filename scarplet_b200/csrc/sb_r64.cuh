// Length-4096 FFT as 64 x 64: a group of 64 threads (two warps) carries one transform, 64
// elements per thread in registers, ONE shared-memory exchange per transform (the radix-16
// core of sb_fft.cuh needs two), 26 % fewer floating-point instructions per point (two
// radix-64 passes with one set of twiddles instead of three radix-16 passes with two), and a
// barrier that spans two warps instead of a CTA.  The price is registers: 64 complex values
// live per thread, so a CTA is 256 threads = 4 independent transforms per SM.
//
//   thread t holds x[t + 64 q], q = 0..63                       (natural order, like sb_fft.cuh)
//   A[t][k1]   = sum_q x[t + 64 q] W64^(q k1)                    radix-64 pass in registers
//   A'[t][k1]  = A[t][k1] * W4096^(t k1)                         twiddles
//   exchange   : thread k1 gathers A'[n2][k1], n2 = 0..63       one transpose through shared memory
//   X[k1 + 64 k2] = sum_n2 A'[n2][k1] W64^(n2 k2)                radix-64 pass in registers
//
// A radix-64 pass is 8 x 8: DFT-8 over the high digit, the constant W64^(q2 u1) rotations,
// DFT-8 over the low digit.  Done in place it leaves its outputs digit-transposed in the
// register array (logical index u at slot(u) = 8 (u & 7) + (u >> 3)); the second flavour takes
// its inputs digit-transposed and leaves natural order.  All indices are compile-time, so the
// permutation costs nothing.
#pragma once
#include "sb_kernels.cuh"

namespace sb64 {

constexpr int R = 64;              // radix = threads per transform = elements per thread
constexpr int N = R * R;           // 4096
constexpr int kPitch = R + 1;      // exchange buffer: row n2 at n2 * 65 (odd pitch: conflict-free both ways)
constexpr int kXchg = R * kPitch;  // float2 elements of one exchange buffer
constexpr int kTwRows = 14;        // W4096^(t u) for u = 1..7 and u = 8, 16, .., 56

SB_CONSTEXPR int slot(int u) { return 8 * (u & 7) + (u >> 3); }

// cos(2 pi k / 64), k = 0..16
SB_CONSTEXPR float cos64_q(int k) {
    constexpr float c[17] = {1.f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f,
                             0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f,
                             0.77301045336273699f, 0.70710678118654757f, 0.63439328416364549f,
                             0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f,
                             0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.f};
    return c[k];
}
SB_CONSTEXPR float cos64(int k) {          // any k
    k &= 63;
    return k <= 16 ? cos64_q(k) : k <= 32 ? -cos64_q(32 - k) : k <= 48 ? -cos64_q(k - 32) : cos64_q(64 - k);
}
SB_CONSTEXPR float sin64(int k) { return cos64(k - 16); }

// x *= W64^k = exp(-2 pi i k / 64), k a compile-time constant
template <int K>
SB_DEVICE float2 mul_w64(float2 a) {
    constexpr int k = K & 63;
    if constexpr (k == 0) return a;
    else if constexpr (k == 16) return make_float2(a.y, -a.x);
    else if constexpr (k == 32) return make_float2(-a.x, -a.y);
    else if constexpr (k == 48) return make_float2(-a.y, a.x);
    else return sbfft::cmul(a, make_float2(cos64(k), -sin64(k)));
}

// One radix-64 pass over the thread's registers.
//   PERM_IN = false: logical input q at x[q], logical output u at x[slot(u)]
//   PERM_IN = true : logical input q at x[slot(q)], logical output u at x[u]
template <bool PERM_IN>
SB_DEVICE void dft64(float2 (&x)[R]) {
    // step 1: DFT-8 over the high input digit q1 for every low digit q2;  y[q2][u1] replaces it
#pragma unroll
    for (int q2 = 0; q2 < 8; ++q2) {
        float2 v[8];
#pragma unroll
        for (int q1 = 0; q1 < 8; ++q1) v[q1] = x[PERM_IN ? 8 * q2 + q1 : 8 * q1 + q2];
        sbfft::Dft<float, 8>::run(v);
#pragma unroll
        for (int u1 = 0; u1 < 8; ++u1) x[PERM_IN ? 8 * q2 + u1 : 8 * u1 + q2] = v[u1];
    }
    // steps 2 + 3: rotate by W64^(q2 u1), then DFT-8 over q2 for every u1
#pragma unroll
    for (int u1 = 0; u1 < 8; ++u1) {
        float2 v[8];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) v[q2] = x[PERM_IN ? 8 * q2 + u1 : 8 * u1 + q2];
        // constant rotations (the compiler sees q2 * u1 as a constant in the unrolled code)
        if (u1 == 1) { v[1] = mul_w64<1>(v[1]); v[2] = mul_w64<2>(v[2]); v[3] = mul_w64<3>(v[3]); v[4] = mul_w64<4>(v[4]); v[5] = mul_w64<5>(v[5]); v[6] = mul_w64<6>(v[6]); v[7] = mul_w64<7>(v[7]); }
        if (u1 == 2) { v[1] = mul_w64<2>(v[1]); v[2] = mul_w64<4>(v[2]); v[3] = mul_w64<6>(v[3]); v[4] = mul_w64<8>(v[4]); v[5] = mul_w64<10>(v[5]); v[6] = mul_w64<12>(v[6]); v[7] = mul_w64<14>(v[7]); }
        if (u1 == 3) { v[1] = mul_w64<3>(v[1]); v[2] = mul_w64<6>(v[2]); v[3] = mul_w64<9>(v[3]); v[4] = mul_w64<12>(v[4]); v[5] = mul_w64<15>(v[5]); v[6] = mul_w64<18>(v[6]); v[7] = mul_w64<21>(v[7]); }
        if (u1 == 4) { v[1] = mul_w64<4>(v[1]); v[2] = mul_w64<8>(v[2]); v[3] = mul_w64<12>(v[3]); v[4] = mul_w64<16>(v[4]); v[5] = mul_w64<20>(v[5]); v[6] = mul_w64<24>(v[6]); v[7] = mul_w64<28>(v[7]); }
        if (u1 == 5) { v[1] = mul_w64<5>(v[1]); v[2] = mul_w64<10>(v[2]); v[3] = mul_w64<15>(v[3]); v[4] = mul_w64<20>(v[4]); v[5] = mul_w64<25>(v[5]); v[6] = mul_w64<30>(v[6]); v[7] = mul_w64<35>(v[7]); }
        if (u1 == 6) { v[1] = mul_w64<6>(v[1]); v[2] = mul_w64<12>(v[2]); v[3] = mul_w64<18>(v[3]); v[4] = mul_w64<24>(v[4]); v[5] = mul_w64<30>(v[5]); v[6] = mul_w64<36>(v[6]); v[7] = mul_w64<42>(v[7]); }
        if (u1 == 7) { v[1] = mul_w64<7>(v[1]); v[2] = mul_w64<14>(v[2]); v[3] = mul_w64<21>(v[3]); v[4] = mul_w64<28>(v[4]); v[5] = mul_w64<35>(v[5]); v[6] = mul_w64<42>(v[6]); v[7] = mul_w64<49>(v[7]); }
        sbfft::Dft<float, 8>::run(v);
#pragma unroll
        for (int u2 = 0; u2 < 8; ++u2) x[PERM_IN ? 8 * u2 + u1 : 8 * u1 + u2] = v[u2];
    }
}

// The same pass when only the inputs q = 0, 1, 2 and 61, 62, 63 can be non-zero (a template
// column: support shorter than 192 rows either side of the origin).  Natural in, permuted out.
// Step 1 degenerates: for q2 = 0, 1, 2 only q1 = 0 contributes (y[q2][u1] = x[q2]); for
// q2 = 5, 6, 7 only q1 = 7 does (y[q2][u1] = x[56 + q2] W8^(7 u1)); q2 = 3, 4 are zero.
SB_DEVICE void dft64_sparse6(float2 (&x)[R]) {
    const float h = 0.70710678118654752440f;
    float2 lo[3] = {x[0], x[1], x[2]};
    float2 hi[3] = {x[61], x[62], x[63]};
#pragma unroll
    for (int u1 = 0; u1 < 8; ++u1) {
        float2 v[8];
#pragma unroll
        for (int q2 = 0; q2 < 3; ++q2) v[q2] = lo[q2];
        v[3] = make_float2(0.f, 0.f);
        v[4] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q2 = 5; q2 < 8; ++q2) {
            // W8^(7 u1) = W8^(-u1) = exp(+2 pi i u1 / 8)
            const float2 a = hi[q2 - 5];
            float2 r;
            switch (u1) {
                case 0: r = a; break;
                case 1: r = make_float2(h * (a.x - a.y), h * (a.x + a.y)); break;
                case 2: r = make_float2(-a.y, a.x); break;
                case 3: r = make_float2(-h * (a.x + a.y), h * (a.x - a.y)); break;
                case 4: r = make_float2(-a.x, -a.y); break;
                case 5: r = make_float2(h * (a.y - a.x), -h * (a.x + a.y)); break;
                case 6: r = make_float2(a.y, -a.x); break;
                default: r = make_float2(h * (a.x + a.y), h * (a.y - a.x)); break;
            }
            v[q2] = r;
        }
        if (u1 == 1) { v[1] = mul_w64<1>(v[1]); v[2] = mul_w64<2>(v[2]); v[5] = mul_w64<5>(v[5]); v[6] = mul_w64<6>(v[6]); v[7] = mul_w64<7>(v[7]); }
        if (u1 == 2) { v[1] = mul_w64<2>(v[1]); v[2] = mul_w64<4>(v[2]); v[5] = mul_w64<10>(v[5]); v[6] = mul_w64<12>(v[6]); v[7] = mul_w64<14>(v[7]); }
        if (u1 == 3) { v[1] = mul_w64<3>(v[1]); v[2] = mul_w64<6>(v[2]); v[5] = mul_w64<15>(v[5]); v[6] = mul_w64<18>(v[6]); v[7] = mul_w64<21>(v[7]); }
        if (u1 == 4) { v[1] = mul_w64<4>(v[1]); v[2] = mul_w64<8>(v[2]); v[5] = mul_w64<20>(v[5]); v[6] = mul_w64<24>(v[6]); v[7] = mul_w64<28>(v[7]); }
        if (u1 == 5) { v[1] = mul_w64<5>(v[1]); v[2] = mul_w64<10>(v[2]); v[5] = mul_w64<25>(v[5]); v[6] = mul_w64<30>(v[6]); v[7] = mul_w64<35>(v[7]); }
        if (u1 == 6) { v[1] = mul_w64<6>(v[1]); v[2] = mul_w64<12>(v[2]); v[5] = mul_w64<30>(v[5]); v[6] = mul_w64<36>(v[6]); v[7] = mul_w64<42>(v[7]); }
        if (u1 == 7) { v[1] = mul_w64<7>(v[1]); v[2] = mul_w64<14>(v[2]); v[5] = mul_w64<35>(v[5]); v[6] = mul_w64<42>(v[6]); v[7] = mul_w64<49>(v[7]); }
        sbfft::Dft<float, 8>::run(v);
#pragma unroll
        for (int u2 = 0; u2 < 8; ++u2) x[8 * u1 + u2] = v[u2];
    }
}

// host: table [kTwRows][64] of W4096^(t u): rows 0..6 for u = 1..7, rows 7..13 for u = 8, 16, .., 56
inline void fill_twiddles64(float2* out) {
    for (int r = 0; r < kTwRows; ++r) {
        const int u = r < 7 ? r + 1 : 8 * (r - 6);
        for (int t = 0; t < R; ++t) {
            const double a = -2.0 * M_PI * (double)u * (double)t / (double)N;
            out[r * R + t].x = (float)cos(a);
            out[r * R + t].y = (float)sin(a);
        }
    }
}

// x[logical u] *= W4096^(t u); PERM: logical u sits at x[slot(u)].  tw: the table (shared memory).
// W^(8 a + b) = W^(8 a) * W^b: 14 table reads, 49 derived products (1.5 ulp, as the radix-16 diet).
template <bool PERM>
SB_DEVICE void twiddle64(float2 (&x)[R], int t, const float2* SB_RESTRICT tw) {
    float2 wl[8];
#pragma unroll
    for (int b = 1; b < 8; ++b) wl[b] = tw[(b - 1) * R + t];
#pragma unroll
    for (int b = 1; b < 8; ++b) x[PERM ? slot(b) : b] = sbfft::cmul(x[PERM ? slot(b) : b], wl[b]);
#pragma unroll
    for (int a = 1; a < 8; ++a) {
        const float2 wh = tw[(6 + a) * R + t];
        x[PERM ? slot(8 * a) : 8 * a] = sbfft::cmul(x[PERM ? slot(8 * a) : 8 * a], wh);
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            const int u = 8 * a + b;
            x[PERM ? slot(u) : u] = sbfft::cmul(x[PERM ? slot(u) : u], sbfft::cmul(wh, wl[b]));
        }
    }
}

// transpose through the group's exchange buffer: thread t gives logical u (at slot(u) if PERM),
// takes logical q into x[q].  bar(): barrier over the group's 64 threads.
template <bool PERM, class Bar>
SB_DEVICE void exchange64(float2 (&x)[R], int t, float2* sm, Bar bar) {
    bar();                                   // every thread is done reading the previous contents
#pragma unroll
    for (int u = 0; u < R; ++u) sm[t * kPitch + u] = x[PERM ? slot(u) : u];
    bar();
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = sm[q * kPitch + t];
}

// Forward FFT of length 4096 over the group's registers: natural in (x[q] = element t + 64 q),
// digit-transposed out (element t + 64 u at x[slot(u)]).
template <class Bar>
SB_DEVICE void forward4096(float2 (&x)[R], int t, float2* sm, const float2* SB_RESTRICT tw, Bar bar) {
    dft64<false>(x);
    twiddle64<true>(x, t, tw);
    exchange64<true>(x, t, sm, bar);
    dft64<false>(x);
}

struct GroupBar {         // named barrier over one group of 64 threads (ids 1..15)
    int id;
    SB_DEVICE void operator()() const { sb_bar(id, R); }
};

// unit-test kernel: batched FFT of length 4096 by the radix-64 core; 256 threads = 4 rows per CTA
SB_GLOBAL SB_LAUNCH_BOUNDS(256, 1)
k_fft4096_r64(int rows, const float2* SB_RESTRICT in, float2* SB_RESTRICT outp, int inverse,
              const float2* SB_RESTRICT tw64) {
    const int grp = sb_tid() / R, t = sb_tid() % R;
    float2* sm = (float2*)sb_shared();
    float2* tw_s = sm + 4 * kXchg;
    for (int i = sb_tid(); i < kTwRows * R; i += 256) tw_s[i] = tw64[i];
    sb_sync();
    const int r = sb_bx() * 4 + grp;
    const bool active = r < rows;
    float2 x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const float2 a = active ? in[(long)r * N + t + R * q] : make_float2(0.f, 0.f);
        x[q] = inverse ? make_float2(a.y, a.x) : a;
    }
    forward4096(x, t, sm + grp * kXchg, tw_s, GroupBar{1 + grp});
    if (active) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const float2 v = x[slot(u)];
            outp[(long)r * N + t + R * u] = inverse ? make_float2(v.y, v.x) : v;
        }
    }
}


// ---------------------------------------------------------------------------
// k_conv_cols_r<SPARSE>: the column kernel (core.py:359, 363: ft * fc, fm2 * fc2 and the
// column half of ifft2) for Py = 4096 on the radix-64 core.  Persistent: grid (KX), a CTA owns
// one spectrum column for the whole batch; 256 threads = 4 groups of 64 = 2 pairs.  A PAIR
// takes one template at a time, its two groups transform the two fields (t * curv, M * curv^2)
// side by side, each behind its own two-warp barrier:
//
//   template column (6 non-zero rows per thread when SPARSE) -> forward 4096 -> x curvature
//   spectrum column (staged in shared memory once per run of same-angle templates) -> inverse
//   4096 -> both fields of a row leave in one 16-byte store (gbuf layout of sb_kernels.cuh).
//
// For the combined store the groups swap halves through their (by then idle) exchange
// buffers: group a keeps rows u < 32 and parks u >= 32, group b the other way round; one
// pair-wide barrier; each group stores its half with the partner's values beside its own.
// SPARSE: every template of the batch has |row offset| < 192, so a thread's non-zero inputs
// are q = 0, 1, 2 and 61, 62, 63 (dft64_sparse6).
// ---------------------------------------------------------------------------
// rows u = 32 HALF .. 32 HALF + 31 of a finished column (row m = t + 64 u at slot(u)) into the
// group's exchange buffer, for the partner group to pick up (register indices are static)
template <int HALF>
SB_DEVICE void park_half(const float2 (&x)[R], int t, float2* xch) {
#pragma unroll
    for (int u = 0; u < R / 2; ++u) xch[(u + HALF * (R / 2)) * R + t] = x[slot(u + HALF * (R / 2))];
}
// store rows u = 32 HALF .. of both fields: mine from registers, the partner's from its buffer.
// FIELD: which of the two fields is mine.  Row m = t + 64 u lies at gbuf_index(t, kx) + u * 64 * kpitch
// (m / 2 = t / 2 + 32 u): one pointer, one add per store; in a periodic domain every row is stored.
template <int HALF, int FIELD>
SB_DEVICE void store_half(const float2 (&x)[R], int t, const float2* xch_partner, float4* dst, int kx, int kpitch,
                          int dly, int out_ny) {
    static_assert(sb::kGbufRows == 2, "store_half: row-pair interleave");
    const long step = (long)R * kpitch;
    float4* p = dst + sb::gbuf_index(t, kx, kpitch) + (long)(HALF * (R / 2)) * step;
    if (out_ny >= N) {
#pragma unroll
        for (int u = 0; u < R / 2; ++u) {
            const int uu = u + HALF * (R / 2);
            const float2 mine = x[slot(uu)];
            const float2 other = xch_partner[uu * R + t];
            sb_st_stream(p, sb::gbuf_pack<float2, float4>(FIELD == 0 ? mine : other, FIELD == 0 ? other : mine));
            p += step;
        }
    } else {
        const int io0 = t + dly + HALF * (R / 2) * R;
#pragma unroll
        for (int u = 0; u < R / 2; ++u) {
            const int uu = u + HALF * (R / 2);
            const float2 mine = x[slot(uu)];
            const float2 other = xch_partner[uu * R + t];
            const float4 v = sb::gbuf_pack<float2, float4>(FIELD == 0 ? mine : other, FIELD == 0 ? other : mine);
            if (((io0 + u * R) & (N - 1)) < out_ny) sb_st_stream(p, v);
            p += step;
        }
    }
}

constexpr int kConvRThreads = 256;
constexpr int kConvRMaxBatch = 64;
constexpr size_t kConvRSmem = (size_t)(4 * kXchg + 2 * N + kTwRows * R) * sizeof(float2) + kConvRMaxBatch * 4 * sizeof(int);

template <bool SPARSE>
SB_GLOBAL SB_LAUNCH_BOUNDS(kConvRThreads, 1)
k_conv_cols_r(sb::Geom g, const sb::Tmpl* SB_RESTRICT tmpls, int tmpl_base, int cnt, int angle_base,
              const float4* SB_RESTRICT trt, const float2* SB_RESTRICT fct, float4* SB_RESTRICT gbuf,
              const float2* SB_RESTRICT tw64) {
    const int grp = sb_tid() / R, t = sb_tid() % R;
    const int pair = grp >> 1, f = grp & 1;                // field of this group: 0 = t * curv, 1 = M * curv^2
    const int KX = g.Px / 2 + 1;
    const int kx = sb_bx();
    float2* sm = (float2*)sb_shared();
    float2* xch = sm + grp * kXchg;                        // this group's exchange buffer
    float2* xch_partner = sm + (grp ^ 1) * kXchg;
    float2* spec_s = sm + 4 * kXchg;                       // [2][N]
    float2* tw_s = spec_s + 2 * N;                         // [kTwRows][R]
    int* s_meta = (int*)(tw_s + kTwRows * R);              // [kConvRMaxBatch][4]: sy_lo, sy_hi, angle
    for (int i = sb_tid(); i < kTwRows * R; i += kConvRThreads) tw_s[i] = tw64[i];
    for (int i = sb_tid(); i < cnt; i += kConvRThreads) {
        const sb::Tmpl* p = tmpls + tmpl_base + i;
        s_meta[4 * i + 0] = p->sy_lo;
        s_meta[4 * i + 1] = p->sy_hi;
        s_meta[4 * i + 2] = p->angle_id - angle_base;
    }
    const GroupBar gbar{1 + grp};
    const int pbar = 5 + pair;
    const float2* spec = spec_s + f * N + t;
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
            for (int i = sb_tid(); i < N / 2; i += kConvRThreads) {
                dst[i] = sb_ld_stream(src + i);
                dst[N / 2 + i] = sb_ld_stream(src2 + i);
            }
        }
        sb_sync();
        // SPARSE: rows t, t + 64, t + 128 and t - 192, t - 128, t - 64 of the template column,
        // fetched one template ahead so that their latency hides behind the transforms
        float2 in[6];
        auto fetch_sparse = [&](int p_loc) {
            const int sy_lo = s_meta[4 * p_loc + 0], sy_hi = s_meta[4 * p_loc + 1];
            const float2* src = (const float2*)(trt + ((long)p_loc * KX + kx) * g.syp) + f;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int srow = j < 3 ? t + R * j : t + R * (j - 6);
                in[j] = make_float2(0.f, 0.f);
                if (srow >= sy_lo && srow <= sy_hi) in[j] = sb_ld_stream(src + 2 * (srow - sy_lo));
            }
        };
        if (SPARSE && s0 + pair < s1) fetch_sparse(s0 + pair);
#pragma unroll 1
        for (int p_loc = s0 + pair; p_loc < s1; p_loc += 2) {
            float2 x[R];
            if (SPARSE) {
#pragma unroll
                for (int q = 0; q < R; ++q) x[q] = make_float2(0.f, 0.f);
                x[0] = in[0]; x[1] = in[1]; x[2] = in[2];
                x[61] = in[3]; x[62] = in[4]; x[63] = in[5];
                if (p_loc + 2 < s1) fetch_sparse(p_loc + 2);
                dft64_sparse6(x);
            } else {
                const int sy_lo = s_meta[4 * p_loc + 0], sy_hi = s_meta[4 * p_loc + 1];
                const float2* src = (const float2*)(trt + ((long)p_loc * KX + kx) * g.syp) + f;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int qy = t + R * q;
                    const int srow = qy < N / 2 ? qy : qy - N;
                    x[q] = make_float2(0.f, 0.f);
                    if (srow >= sy_lo && srow <= sy_hi) x[q] = sb_ld_stream(src + 2 * (srow - sy_lo));
                }
                dft64<false>(x);
            }
            // forward: twiddles, transpose, second pass (logical u at slot(u))
            twiddle64<true>(x, t, tw_s);
            sb_bar(pbar, 2 * R);                            // the partner has read what I parked for the last template
            exchange64<true>(x, t, xch, gbar);
            dft64<false>(x);
            // spectrum product (core.py:359 / :363); swap: inverse via forward
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const float2 pr = sbfft::cmul(x[slot(u)], spec[R * u]);
                x[slot(u)] = make_float2(pr.y, pr.x);
            }
            // inverse: digit-transposed in, natural out; twiddles; transpose; last pass
            dft64<true>(x);
            twiddle64<false>(x, t, tw_s);
            exchange64<false>(x, t, xch, gbar);
            dft64<false>(x);                                // row m = t + 64 u at slot(u)
            // swap halves with the partner group and store both fields of a row together
            gbar();                                         // my group is done reading my exchange buffer
            if (f == 0) park_half<1>(x, t, xch); else park_half<0>(x, t, xch);
            sb_bar(pbar, 2 * R);
            float4* dst = gbuf + (long)p_loc * N * g.kpitch;
            if (f == 0) store_half<0, 0>(x, t, xch_partner, dst, kx, g.kpitch, g.dly, g.out_ny);
            else store_half<1, 1>(x, t, xch_partner, dst, kx, g.kpitch, g.dly, g.out_ny);
        }
        // a pair that ran out of templates must still meet its partner's barriers: none are
        // pending here, because both groups of a pair walk the same templates
        s0 = s1;
    }
}

}  // namespace sb64
