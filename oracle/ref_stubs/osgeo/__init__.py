"""Import stand-in for GDAL's python bindings: enough for DEMGrid.load on
the reference's bundled single-band TIFFs (read with PIL)."""
import types as _types

import numpy as _np

gdalconst = _types.ModuleType("osgeo.gdalconst")
gdalconst.GDT_Float32 = 6
osr = _types.ModuleType("osgeo.osr")
ogr = _types.ModuleType("osgeo.ogr")


class _SpatialReference(object):
    def ImportFromEPSG(self, code):
        self.code = code

    def ExportToWkt(self):
        return "EPSG:%s" % getattr(self, "code", None)


osr.SpatialReference = _SpatialReference


class _Band(object):
    def __init__(self, arr):
        self._arr = arr

    def GetNoDataValue(self):
        return None

    def ReadAsArray(self):
        return self._arr


class _Dataset(object):
    def __init__(self, filename):
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        im = Image.open(filename)
        self._arr = _np.array(im)
        tags = getattr(im, "tag_v2", {})
        dx = dy = 1.0
        ulx = uly = 0.0
        if 33550 in tags:                       # ModelPixelScale
            dx, dy = float(tags[33550][0]), float(tags[33550][1])
        elif 34264 in tags:                     # ModelTransformation
            m = tags[34264]
            dx, dy = float(m[0]), float(m[5])
            ulx, uly = float(m[3]), float(m[7])
        if 33922 in tags:                       # ModelTiepoint
            ulx, uly = float(tags[33922][3]), float(tags[33922][4])
        self._gt = (ulx, dx, 0.0, uly, 0.0, dy)
        self.RasterYSize, self.RasterXSize = self._arr.shape

    def GetRasterBand(self, i):
        return _Band(self._arr)

    def GetGeoTransform(self):
        return self._gt

    def GetProjection(self):
        return ""


gdal = _types.ModuleType("osgeo.gdal")
gdal.Open = _Dataset
