"""In-memory DEM container as the hot path consumes it (dem.py:203-218, 351-372 of the
reference): ``_griddata`` (float64, ny x nx) and ``_georef_info.dx / .dy``.  Reading rasters
(GDAL / rasterio) is outside the scope of this package; any object with those two
attributes — including the reference's own ``DEMGrid`` — is accepted by ``core``.  The
methods either side of the match path are here, on the GPU where they compute:

* ``_calculate_directional_laplacian``       dem.py:68-107
* ``_estimate_curvature_noiselevel``         dem.py:152-179
* ``_fill_nodata``                           dem.py:388-414 (device fill in place of rasterio)
* ``save``                                   dem.py:291-306 (dependency-free GeoTIFF writer)
"""
import numpy as np


class GeorefInfo(object):
    def __init__(self, dx=1.0, dy=None, nx=None, ny=None):
        self.dx = dx
        self.dy = dx if dy is None else dy
        self.nx = nx
        self.ny = ny
        self.geo_transform = None
        self.projection = None
        self.xllcenter = 0
        self.yllcenter = 0


class DEMGrid(object):
    """Elevation grid held in host memory; curvature is computed on the GPU."""

    def __init__(self, data=None, dx=1.0, dy=None):
        if data is None:
            data = np.empty((0, 0))
        self._griddata = np.array(data, dtype=np.float64)
        ny, nx = self._griddata.shape
        self._georef_info = GeorefInfo(dx, dy, nx, ny)
        self.shape = self._griddata.shape
        self.nodata_value = np.nan
        self.is_interpolated = False

    def _plan(self):
        from .engine import Plan
        ny, nx = self._griddata.shape
        return Plan(ny, nx, self._georef_info.dx, self._georef_info.dy)

    def _calculate_directional_laplacian(self, alpha):
        """dem.py:68-107, evaluated by the CUDA stencil in float64 (bit-exact with the
        NumPy reference).  Unlike the reference the input grid is not modified."""
        with self._plan() as plan:
            plan.set_dem(self._griddata)
            return plan.directional_laplacian(alpha)

    def _calculate_laplacian(self):
        """dem.py:62-66"""
        return self._calculate_directional_laplacian(0)

    def _estimate_curvature_noiselevel(self, sigma=100, truncate=4.0):
        """Mean and standard deviation of the high-passed directional curvature for 180
        directions (dem.py:152-179): ``highpass = del2z - gaussian_filter(del2z, 100)``,
        ``nanmean`` / ``nanstd`` per direction.  Returns ``(angles, mean, sd)`` like the
        reference (``mean`` and ``sd`` are lists of 180 floats).

        The directional Laplacian is linear in the second differences (dem.py:103-104) and so
        is the Gaussian filter: the device filters dxx, dxy, dyy once and reduces them to ten
        moments; the 180 directions are quadratic forms of those (instead of 180 Laplacians
        and 180 filters)."""
        angles = np.linspace(0, np.pi, num=180)                      # dem.py:166
        z = self._griddata
        has_nan = bool(np.isnan(z).any())
        with self._plan() as plan:
            plan.set_dem(z)
            moments = [plan.curvature_noise_moments(sigma, truncate)]
            if has_nan:
                # The reference's Laplacian zero-fills nodata cells of the grid in place on its
                # first call (dem.py:85-86): only the first direction sees them as NaN
                # (dem.py:105), the other 179 see zeros.  Same values here, grid untouched.
                plan.set_dem(np.where(np.isnan(z), 0.0, z))
                moments.append(plan.curvature_noise_moments(sigma, truncate))

        def stats(m, ang):
            if m[0] <= 0:
                nan = np.full(len(ang), np.nan)
                return nan, nan
            s_xx, s_xy, s_yy, q_xx, q_xy, q_yy, q_xx_xy, q_xx_yy, q_xy_yy = m[1:] / m[0]
            c2, s2 = np.cos(ang) ** 2, np.sin(ang) ** 2              # dem.py:103-104
            sc = 2 * np.sin(ang) * np.cos(ang)
            mean = s_xx * c2 - s_xy * sc + s_yy * s2
            second = (q_xx * c2 ** 2 + q_xy * sc ** 2 + q_yy * s2 ** 2
                      - 2 * q_xx_xy * c2 * sc + 2 * q_xx_yy * c2 * s2 - 2 * q_xy_yy * sc * s2)
            return mean, np.sqrt(np.maximum(second - mean ** 2, 0.0))

        mean, sd = stats(moments[-1], angles)
        if has_nan:
            m0, s0 = stats(moments[0], angles[:1])
            mean = np.concatenate([m0, mean[1:]])
            sd = np.concatenate([s0, sd[1:]])
        return angles, [float(v) for v in mean], [float(v) for v in sd]

    def _fill_nodata(self):
        """Fill nodata cells by interpolation, in place (dem.py:388-414).  The reference
        wraps ``rasterio.fill.fillnodata`` (GDAL's inverse-distance fill); here each pass is
        a CUDA kernel: a nodata cell becomes the inverse-distance-weighted mean of the
        nearest valid cell along the eight row / column / diagonal rays, with the
        reference's search distance (half the longest nodata run count, dem.py:403-405) and
        its loop until nothing is left (dem.py:402)."""
        if ~np.isnan(self.nodata_value):                             # dem.py:392-395
            self._griddata[self._griddata == self.nodata_value] = np.nan
        self.nodata_mask = np.isnan(self._griddata)
        num = int(self.nodata_mask.sum())
        if num and num < self._griddata.size:
            z = np.ascontiguousarray(self._griddata, dtype=np.float64)
            with self._plan() as plan:
                prev = None
                while num > 0 and num != prev:                       # dem.py:402
                    mask = np.isnan(z)
                    col_nodata = mask.sum(axis=0).max()              # dem.py:403-405
                    row_nodata = mask.sum(axis=1).max()
                    dist = max(row_nodata, col_nodata) / 2
                    prev = num
                    num = plan.fill_nodata_pass(z, max(dist, 1.0))
            self._griddata = z
        self.is_interpolated = True

    def save(self, filename, epsg=None):
        """Save the grid as a single-band Float32 GeoTIFF (dem.py:291-306)."""
        from .geotiff import write_geotiff, _transform_of
        return write_geotiff(filename, self._griddata, geo_transform=_transform_of(self._georef_info), epsg=epsg)
