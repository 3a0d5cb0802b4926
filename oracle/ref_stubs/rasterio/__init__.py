"""Import stand-in for rasterio."""
from . import fill  # noqa: F401
