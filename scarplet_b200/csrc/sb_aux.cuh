// Kernels either side of the match path: the curvature noise level of a
// DEM (dem.py:152-179) and nodata filling (dem.py:388-414).  float64 throughout; none of them
// is on the search's hot path.
#pragma once
#include "sb_kernels.cuh"

namespace sb {

// scipy.ndimage 'reflect' boundary (d c b a | a b c d | d c b a), any distance
SB_DEVICE int reflect_index(int i, int n) {
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

// Separable Gaussian low-pass of dem.py:172 (`ndimage.gaussian_filter(del2z, 100)`), first along
// axis 0.  The directional Laplacian is linear in the three second differences
// (dem.py:103-104), and so is the filter: the planes dxx, dxy, dyy are filtered once instead
// of once per direction.  Plane 3 is the NaN indicator of the DEM (dem.py:105): filtered with
// the same strictly positive weights it is > 0 exactly where scipy's filter output is NaN.
// grid (ceil(n / 256), 4); `diffs` = [dxx][dxy][dyy] planes, out = 4 planes.
SB_GLOBAL k_gauss_axis0(int ny, int nx, int radius, const double* SB_RESTRICT w, const double* SB_RESTRICT diffs,
                        const double* SB_RESTRICT dem, double* SB_RESTRICT out) {
    const long n = (long)ny * nx;
    const long i = (long)sb_bx() * 256 + sb_tid();
    if (i >= n) return;
    const int p = sb_by();
    const int r = (int)(i / nx), c = (int)(i % nx);
    double acc = 0.0;
    for (int k = -radius; k <= radius; ++k) {
        const long o = (long)reflect_index(r + k, ny) * nx + c;
        double v;
        if (p < 3) {
            v = sb_ldg(diffs + (long)p * n + o);
            if (v != v) v = 0.0;                 // the NaN cell itself: flagged through plane 3
        } else {
            const double z = sb_ldg(dem + o);
            v = z != z ? 1.0 : 0.0;
        }
        acc += sb_ldg(w + k + radius) * v;
    }
    out[(long)p * n + i] = acc;
}

SB_GLOBAL k_gauss_axis1(int ny, int nx, int radius, const double* SB_RESTRICT w, const double* SB_RESTRICT in,
                        double* SB_RESTRICT out) {
    const long n = (long)ny * nx;
    const long i = (long)sb_bx() * 256 + sb_tid();
    if (i >= n) return;
    const int p = sb_by();
    const int r = (int)(i / nx), c = (int)(i % nx);
    const double* row = in + (long)p * n + (long)r * nx;
    double acc = 0.0;
    for (int k = -radius; k <= radius; ++k) acc += sb_ldg(w + k + radius) * sb_ldg(row + reflect_index(c + k, nx));
    out[(long)p * n + i] = acc;
}

// Sums over the pixels whose filter window holds no NaN of the high-passed second differences
// h = d - lowpass(d) and of their products:
// [count, hxx, hxy, hyy, hxx^2, hxy^2, hyy^2, hxx hxy, hxx hyy, hxy hyy]; per-block partials,
// fixed order.  mean and standard deviation of the high-passed directional curvature follow
// for every direction as quadratic forms (host side, dem.py:173-175).
SB_GLOBAL k_noise_moments(long n, const double* SB_RESTRICT diffs, const double* SB_RESTRICT dem,
                          const double* SB_RESTRICT low, double* SB_RESTRICT partial) {
    double* sd = (double*)sb_shared();            // [10][256]
    double acc[10];
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    for (long i = (long)sb_bx() * 256 + sb_tid(); i < n; i += (long)sb_nbx() * 256) {
        if (low[3 * n + i] > 0.0) continue;
        const double z = sb_ldg(dem + i);
        if (z != z) continue;
        const double a = diffs[i] - low[i], b = diffs[n + i] - low[n + i], c = diffs[2 * n + i] - low[2 * n + i];
        acc[0] += 1.0; acc[1] += a; acc[2] += b; acc[3] += c;
        acc[4] += a * a; acc[5] += b * b; acc[6] += c * c;
        acc[7] += a * b; acc[8] += a * c; acc[9] += b * c;
    }
    for (int k = 0; k < 10; ++k) sd[k * 256 + sb_tid()] = acc[k];
    sb_sync();
    for (int s = 128; s > 0; s >>= 1) {
        if (sb_tid() < s)
            for (int k = 0; k < 10; ++k) sd[k * 256 + sb_tid()] += sd[k * 256 + sb_tid() + s];
        sb_sync();
    }
    if (sb_tid() < 10) partial[(long)sb_bx() * 10 + sb_tid()] = sd[sb_tid() * 256];
}

// One pass of nodata filling (replaces rasterio.fill.fillnodata at dem.py:406-408): a NaN cell
// becomes the inverse-distance-weighted mean of the nearest valid cell along each of the eight
// row / column / diagonal rays, searched up to `max_dist` cells; cells with no valid neighbour
// within reach stay NaN and are counted (the caller iterates like dem.py:402-411).
SB_GLOBAL k_fill_nodata(int ny, int nx, const double* SB_RESTRICT in, double* SB_RESTRICT out, double max_dist,
                        long* SB_RESTRICT left) {
    const long i = (long)sb_bx() * 256 + sb_tid();
    if (i >= (long)ny * nx) return;
    const double z = in[i];
    if (z == z) { out[i] = z; return; }
    const int r = (int)(i / nx), c = (int)(i % nx);
    double wsum = 0.0, vsum = 0.0;
    for (int d = 0; d < 8; ++d) {
        const int dr = d < 3 ? -1 : d < 5 ? 0 : 1;
        const int dc = (d == 0 || d == 3 || d == 5) ? -1 : (d == 1 || d == 6) ? 0 : 1;
        const double step = (dr != 0 && dc != 0) ? 1.4142135623730951 : 1.0;
        for (int s = 1; s * step <= max_dist; ++s) {
            const int rr = r + s * dr, cc = c + s * dc;
            if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) break;
            const double v = in[(long)rr * nx + cc];
            if (v == v) {
                const double wgt = 1.0 / (s * step);
                wsum += wgt;
                vsum += wgt * v;
                break;
            }
        }
    }
    if (wsum > 0.0) {
        out[i] = vsum / wsum;
    } else {
        out[i] = z;
#ifdef SB_EMU
        __atomic_fetch_add(left + (i & 1023), 1L, __ATOMIC_RELAXED);
#else
        atomicAdd((unsigned long long*)(left + (i & 1023)), 1ULL);
#endif
    }
}

}  // namespace sb
