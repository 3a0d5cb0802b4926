"""Import the UNMODIFIED reference package from /root/reference.

TEST INFRASTRUCTURE ONLY; works only in the build container (the GPU box has no
/root/reference).  Puts the stand-in modules of ``ref_stubs/`` (numexpr, pyfftw,
matplotlib, osgeo, rasterio — none installable here) ahead on ``sys.path`` and
imports ``scarplet`` from the read-only reference checkout.  Used by
``tests/golden/make_golden.py`` to generate fixtures and by
``tests/test_oracle_vs_reference.py`` to validate the NumPy restatement.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SCARPLET_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scarplet"))


def import_reference():
    """Returns the reference's ``scarplet`` module (core API star-imported)."""
    if not available():
        raise ImportError("reference checkout not found at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True      # the reference tree is read-only
    import scarplet                      # noqa: E402
    return scarplet


def make_grid(z, dx=1.0, dy=None):
    """Hand-built reference ``DEMGrid`` (as scarplet/tests/test_core.py:104-121)."""
    import numpy as np
    import_reference()
    from scarplet.dem import DEMGrid
    g = DEMGrid()
    g._griddata = np.array(z, dtype=np.float64)
    g._georef_info.dx = dx
    g._georef_info.dy = dx if dy is None else dy
    g._georef_info.ny, g._georef_info.nx = g._griddata.shape
    return g
