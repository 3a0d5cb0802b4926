"""Import stand-in for pyfftw (see ../README.md)."""
from . import interfaces  # noqa: F401
