// Register-resident Stockham FFT for power-of-two lengths (complex64).
//
// One FFT of length N is carried by a group of T = N / E threads; thread t holds
// E = 16 elements x[t + q*T], q = 0..E-1, in registers.  Every stage is a set of
// radix-R butterflies (R <= E, R in {2,4,8,16}) done entirely in registers; between
// stages the data is exchanged through a padded shared-memory buffer.  Input and
// output are both in natural order with the same thread/register layout, so
// global loads and stores are coalesced across the group and no bit reversal is
// ever materialised.  Twiddles come from per-stage tables laid out [u][k] so that
// a warp reads consecutive addresses.
//
// The inverse transform is obtained by swapping real and imaginary parts before
// and after a forward transform (unnormalised).
#pragma once
#include "sb_rt.h"

namespace sbfft {

constexpr int E = 16;  // elements per thread

SB_HOSTDEV int pad_index(int i) { return i + (i >> 4); }
SB_CONSTEXPR int padded_len(int n) { return n + (n >> 4); }

// ---- stage plan -----------------------------------------------------------
// radices for N = 2^k, 128 <= N <= 8192: as many 16s as fit, then the rest.
SB_CONSTEXPR int num_stages(int N) {
    int s = 0;
    while (N > 1) { N = (N >= 16) ? N / 16 : 1; ++s; }
    return s;
}
SB_CONSTEXPR int stage_radix(int N, int s) {
    int r = 16;
    for (int i = 0; i <= s; ++i) { r = (N >= 16) ? 16 : N; N /= r; }
    return r;
}
SB_CONSTEXPR int stage_ns(int N, int s) {   // product of the radices before stage s
    int ns = 1;
    for (int i = 0; i < s; ++i) ns *= stage_radix(N, i);
    return ns;
}
// offset (in float2) of stage s in the twiddle table; stage 0 has no twiddles
SB_CONSTEXPR int twiddle_offset(int N, int s) {
    int off = 0;
    for (int i = 1; i < s; ++i) off += (stage_radix(N, i) - 1) * stage_ns(N, i);
    return off;
}
SB_CONSTEXPR int twiddle_count(int N) { return twiddle_offset(N, num_stages(N)); }

// host: fill the table for length N (float64 sincos rounded once to float32)
inline void fill_twiddles(int N, float2* out) {
    int ns = 1, off = 0;
    for (int s = 0; N > ns; ++s) {
        int rest = N / ns;
        int r = rest >= 16 ? 16 : rest;
        if (s > 0) {
            for (int u = 1; u < r; ++u)
                for (int k = 0; k < ns; ++k) {
                    double a = -2.0 * M_PI * (double)u * (double)k / ((double)ns * (double)r);
                    out[off + (u - 1) * ns + k] = make_float2((float)cos(a), (float)sin(a));
                }
            off += (r - 1) * ns;
        }
        ns *= r;
    }
}

// ---- small DFTs in registers ------------------------------------------------
SB_DEVICE float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SB_DEVICE float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SB_DEVICE float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
SB_DEVICE float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

template <int R> struct Dft;

template <> struct Dft<2> {
    SB_DEVICE static void run(float2* x) {
        float2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};

template <> struct Dft<4> {
    SB_DEVICE static void run(float2* x) {
        float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
        float2 t2 = cadd(x[1], x[3]), t3 = mul_mi(csub(x[1], x[3]));
        x[0] = cadd(t0, t2);
        x[2] = csub(t0, t2);
        x[1] = cadd(t1, t3);
        x[3] = csub(t1, t3);
    }
};

template <> struct Dft<8> {
    SB_DEVICE static void run(float2* x) {
        const float h = 0.70710678118654752440f;
        float2 e[4] = {x[0], x[2], x[4], x[6]};
        float2 o[4] = {x[1], x[3], x[5], x[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        // o[k] *= W8^k
        o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        o[2] = mul_mi(o[2]);
        o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x[k] = cadd(e[k], o[k]);
            x[k + 4] = csub(e[k], o[k]);
        }
    }
};

template <> struct Dft<16> {
    SB_DEVICE static void run(float2* x) {
        const float h = 0.70710678118654752440f;
        const float c1 = 0.92387953251128675613f;   // cos(pi/8)
        const float s1 = 0.38268343236508977173f;   // sin(pi/8)
        float2 e[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
        Dft<8>::run(e);
        Dft<8>::run(o);
        // o[k] *= W16^k = exp(-i*pi*k/8)
        o[1] = cmul(o[1], make_float2(c1, -s1));
        o[2] = make_float2(h * (o[2].x + o[2].y), h * (o[2].y - o[2].x));
        o[3] = cmul(o[3], make_float2(s1, -c1));
        o[4] = mul_mi(o[4]);
        o[5] = cmul(o[5], make_float2(-s1, -c1));
        o[6] = make_float2(h * (o[6].y - o[6].x), -h * (o[6].x + o[6].y));
        o[7] = cmul(o[7], make_float2(-c1, -s1));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = cadd(e[k], o[k]);
            x[k + 8] = csub(e[k], o[k]);
        }
    }
};

// ---- one Stockham stage ----------------------------------------------------
// v[q] = x[t + q*T] on entry.  On exit the same layout holds the stage output
// (after the exchange for all but the last stage).
template <int N, int S>
SB_DEVICE void stage(float2 (&v)[E], int t, float2* sm, const float2* SB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int R = stage_radix(N, S);
    constexpr int NS = stage_ns(N, S);
    constexpr int B = E / R;                 // butterflies per thread
    constexpr bool LAST = (NS * R == N);
    constexpr int TWO = twiddle_offset(N, S);
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int j = t + m * T;             // butterfly index in [0, N/R)
        float2 x[R];
#pragma unroll
        for (int u = 0; u < R; ++u) x[u] = v[m + u * B];
        if (NS > 1) {
            const int k = j & (NS - 1);
#pragma unroll
            for (int u = 1; u < R; ++u) x[u] = cmul(x[u], sb_ldg(tw + TWO + (u - 1) * NS + k));
        }
        Dft<R>::run(x);
        if (LAST) {
#pragma unroll
            for (int u = 0; u < R; ++u) v[m + u * B] = x[u];
        } else {
            const int base = (j / NS) * (NS * R) + (j & (NS - 1));
#pragma unroll
            for (int u = 0; u < R; ++u) sm[pad_index(base + u * NS)] = x[u];
        }
    }
    if (!LAST) {
        sb_sync();
#pragma unroll
        for (int q = 0; q < E; ++q) v[q] = sm[pad_index(t + q * T)];
        sb_sync();
    }
}

template <int N, int S = 0>
struct Stages {
    SB_DEVICE static void run(float2 (&v)[E], int t, float2* sm, const float2* SB_RESTRICT tw) {
        stage<N, S>(v, t, sm, tw);
        if constexpr (S + 1 < num_stages(N)) Stages<N, S + 1>::run(v, t, sm, tw);
    }
};

// Forward FFT of length N over the group's registers.  `sm` is the group's
// private exchange buffer of padded_len(N) float2.  All threads of the CTA must
// call this together (it contains CTA-wide barriers).
template <int N>
SB_DEVICE void forward(float2 (&v)[E], int t, float2* sm, const float2* SB_RESTRICT tw) {
    Stages<N, 0>::run(v, t, sm, tw);
}

SB_DEVICE void swap_all(float2 (&v)[E]) {
#pragma unroll
    for (int q = 0; q < E; ++q) v[q] = make_float2(v[q].y, v[q].x);
}

}  // namespace sbfft
