"""Import stand-in for matplotlib (plotting is off the hot path)."""
from . import pyplot  # noqa: F401
