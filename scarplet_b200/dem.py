"""In-memory DEM container as the hot path consumes it (dem.py:203-218, 351-372 of the
reference): ``_griddata`` (float64, ny x nx) and ``_georef_info.dx / .dy``.  File I/O
(GDAL / rasterio) is outside the scope of this package; any object with those two
attributes — including the reference's own ``DEMGrid`` — is accepted by ``core``."""
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
        self.is_interpolated = False

    def _calculate_directional_laplacian(self, alpha):
        """dem.py:68-107, evaluated by the CUDA stencil in float64 (bit-exact with the
        NumPy reference).  Unlike the reference the input grid is not modified."""
        from .engine import Plan
        ny, nx = self._griddata.shape
        with Plan(ny, nx, self._georef_info.dx, self._georef_info.dy) as plan:
            plan.set_dem(self._griddata)
            return plan.directional_laplacian(alpha)

    def _calculate_laplacian(self):
        """dem.py:62-66"""
        return self._calculate_directional_laplacian(0)
