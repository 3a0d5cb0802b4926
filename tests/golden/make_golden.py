"""Generates the fixtures under tests/golden/ (run in the build container only).

Two sources, both from the read-only reference checkout at /root/reference:

1. the reference's OWN golden files and test inputs (scarplet/tests/results/*.npy,
   scarplet/tests/data/*.tif), re-packed compactly: inputs as float32 .npy (the TIFFs
   are Float32), goldens as compressed .npz; the five 3.2 MB Laplacian goldens are
   kept as SHA-256 digests of their float64 bytes plus two corner crops (the oracle
   and the CUDA stencil reproduce them bit for bit, so a digest is a complete check);
2. outputs of the UNMODIFIED reference imported here through oracle/ref_import.py
   (numpy.fft / NumPy standing in for pyfftw / numexpr) on small seeded DEMs:
   match_template for every built-in template family, and compare's tie semantics.

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

REF_TESTS = os.path.join(ref_import.REFERENCE_ROOT, "scarplet", "tests")


def read_tiff(path):
    from PIL import Image
    return np.array(Image.open(path))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def seeded_dem(ny, nx, seed):
    """Small rough surface with a scarp across it (float32-valued, like a GDAL load)."""
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.fftfreq(nx)[None, :]
    k = np.sqrt(kx ** 2 + ky ** 2)
    k[0, 0] = 1
    spec = (rng.standard_normal((ny, nx)) + 1j * rng.standard_normal((ny, nx))) * k ** -2.0
    spec[0, 0] = 0
    z = np.real(np.fft.ifft2(spec))
    z = 3.0 * z / z.std()
    y, x = np.mgrid[0:ny, 0:nx]
    xr = (x - x.mean()) * np.cos(0.4) + (y - y.mean()) * np.sin(0.4)
    from scipy.special import erf
    z = z + 100 + 0.05 * x + 1.5 * erf(xr / (2 * np.sqrt(3.)))
    return z.astype(np.float32)


def main():
    sl = ref_import.import_reference()
    from scarplet import WindowedTemplate as WT

    meta = {}
    # ---- 1. the reference's own goldens ---------------------------------
    syn = read_tiff(os.path.join(REF_TESTS, "data", "synthetic.tif"))
    fz = read_tiff(os.path.join(REF_TESTS, "data", "faultzone.tif"))
    assert syn.dtype == np.float32 and fz.dtype == np.float32
    np.save(os.path.join(HERE, "synthetic_dem_f32.npy"), syn)
    np.savez_compressed(os.path.join(HERE, "faultzone_dem_f32.npz"), z=fz)
    meta["synthetic"] = {"dx": 1.0, "dy": 1.0, "shape": list(syn.shape)}
    meta["faultzone"] = {"dx": 2.0, "dy": 2.0, "shape": list(fz.shape)}

    res = os.path.join(REF_TESTS, "results")
    np.savez_compressed(
        os.path.join(HERE, "reference_goldens.npz"),
        scarp_template=np.load(os.path.join(res, "scarp_template.npy")),
        channel_template=np.load(os.path.join(res, "channel_template.npy")),
        synthetic_match1=np.load(os.path.join(res, "synthetic_match1.npy")),
        synthetic_match2=np.load(os.path.join(res, "synthetic_match2.npy")))
    m3 = np.load(os.path.join(res, "synthetic_match3.npy"), allow_pickle=True)
    meta["synthetic_match3"] = {"amp_all_zero": bool((m3[0] == 0).all()), "age": float(m3[1]),
                                "angle": float(m3[2]), "snr_all_zero": bool((m3[3] == 0).all()),
                                "shape": list(m3[0].shape)}
    lap = {}
    crops = {}
    for name, alpha in (("faultzone_del2z", 0.0), ("faultzone_del2z_-90", -np.pi / 2),
                        ("faultzone_del2z_-45", -np.pi / 4), ("faultzone_del2z_45", np.pi / 4),
                        ("faultzone_del2z_90", np.pi / 2)):
        g = np.load(os.path.join(res, name + ".npy"))
        lap[name] = {"alpha": alpha, "sha256": sha(g), "shape": list(g.shape)}
        crops[name + "_tl"] = g[:96, :96]
        crops[name + "_br"] = g[-96:, -96:]
    meta["laplacian"] = lap
    np.savez_compressed(os.path.join(HERE, "laplacian_crops.npz"), **crops)

    # ---- 2. outputs of the unmodified reference on seeded inputs -----------------
    cases = {}
    dem_a = seeded_dem(72, 96, 11)
    dem_b = seeded_dem(65, 81, 12)
    np.savez_compressed(os.path.join(HERE, "seeded_dems_f32.npz"), a=dem_a, b=dem_b)
    specs = [
        ("scarp_a", "a", "Scarp", 10, 3.0, 0.3, 1.0),
        ("scarp_a_neg90", "a", "Scarp", 10, 3.0, -np.pi / 2, 1.0),
        ("scarp_b_odd", "b", "Scarp", 8, 5.0, -0.9, 2.0),
        ("channel_a", "a", "Channel", 6, 0.15, 0.5, 1.0),
        ("ricker_b_odd", "b", "Ricker", 5, 0.1, -1.2, 1.0),
        ("right_upper_a", "a", "RightFacingUpperBreakScarp", 10, 3.0, 0.2, 1.0),
        ("left_upper_b", "b", "LeftFacingUpperBreakScarp", 8, 4.0, -0.6, 1.0),
    ]
    listing = []
    for name, which, cls, scale, age, angle, de in specs:
        z = (dem_a if which == "a" else dem_b).astype(np.float64)
        grid = ref_import.make_grid(z, de, de)
        amp, _, _, snr = sl.match_template(grid, getattr(WT, cls), scale, age, angle)
        cases[name + "_amp"] = amp
        cases[name + "_snr"] = snr
        listing.append({"name": name, "dem": which, "template": cls, "scale": scale, "age": age,
                        "angle": angle, "de": de})
    meta["match_template_cases"] = listing
    # single-age orientation search through the reference's Pool path
    grid = ref_import.make_grid(dem_a.astype(np.float64), 1.0, 1.0)
    cases["search_a"] = sl.calculate_best_fit_parameters(grid, WT.Scarp, 10, 3.0)
    meta["search_a"] = {"dem": "a", "template": "Scarp", "scale": 10, "age": 3.0, "de": 1.0}
    # compare(): strict selects, exact tie resets the pixel (core.py:230-240)
    r1 = (np.array([[1., 2.], [3., 4.]]), 10., 0.1, np.array([[1., 5.], [2., 0.]]))
    r2 = (np.array([[5., 6.], [7., 8.]]), 20., 0.2, np.array([[1., 4.], [3., 0.]]))
    r3 = (np.array([[9., 9.], [9., 9.]]), 30., 0.3, np.array([[.5, 4.], [3., 1.]]))
    out = sl.compare([r1, r2, r3], 2, 2)
    cases["compare_out"] = np.stack(out)
    np.savez_compressed(os.path.join(HERE, "reference_runs.npz"), **cases)

    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(HERE)):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, fn)), fn))


if __name__ == "__main__":
    main()
