"""The CPU oracle against the reference's own golden files and against outputs of the
unmodified reference (tests/golden/, made by tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import scarplet_oracle as O

KINDS = {"Scarp": O.SCARP, "Channel": O.RICKER, "Ricker": O.RICKER,
         "RightFacingUpperBreakScarp": O.RIGHT_UPPER, "LeftFacingUpperBreakScarp": O.LEFT_UPPER}


def test_templates_match_reference_goldens(golden):
    g = golden.npz("reference_goldens.npz")
    assert np.array_equal(O.template_array(O.SCARP, 100, 10, 0, 100, 100, 1), g["scarp_template"])
    assert np.allclose(O.template_array(O.RICKER, 100, 0.1, 0, 100, 100, 1), g["channel_template"],
                       rtol=1e-15, atol=1e-26)


def test_laplacian_goldens_bit_exact(golden):
    z = golden.faultzone_dem
    crops = golden.npz("laplacian_crops.npz")
    for name, info in golden.meta["laplacian"].items():
        out = O.directional_laplacian(z, 2.0, 2.0, info["alpha"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == info["sha256"], name
        assert np.array_equal(out[:96, :96], crops[name + "_tl"])
        assert np.array_equal(out[-96:, -96:], crops[name + "_br"])
    assert np.array_equal(O.laplacian(z, 2.0, 2.0), O.directional_laplacian(z, 2.0, 2.0, 0))


def test_laplacian_does_not_mutate_and_carries_nan():
    rng = np.random.default_rng(1)
    z = rng.standard_normal((20, 30))
    z[3, 4] = np.nan
    keep = z.copy()
    out = O.directional_laplacian(z, 1.0, 1.0, 0.3)
    assert np.array_equal(z, keep, equal_nan=True)
    assert np.isnan(out[3, 4]) and np.isnan(out).sum() == 1


def test_match_single_age_golden(golden):
    """scarplet/tests/test_core.py:44-61"""
    res = O.match(golden.synthetic_dem, 1.0, 1.0, O.SCARP, scale=100, age=10,
                  ang_max=np.pi / 2, ang_min=-np.pi / 2, processes=4)
    gold = golden.npz("reference_goldens.npz")["synthetic_match2"]
    for i in range(4):
        assert np.allclose(res[i], gold[i])


def test_match_template_all_masked_golden(golden):
    """scarplet/tests/test_core.py:63-82"""
    amp, age, angle, snr = O.match_template(golden.synthetic_dem, 1.0, 1.0, O.SCARP, 100, 10, 0)
    m3 = golden.meta["synthetic_match3"]
    assert (amp == 0).all() and (snr == 0).all() and age == m3["age"] and angle == m3["angle"]


@pytest.mark.slow
def test_match_age_sweep_golden(golden):
    """scarplet/tests/test_core.py:24-42 (35 ages x 181 angles, ~1 min on 4 cores)"""
    res = O.match(golden.synthetic_dem, 1.0, 1.0, O.SCARP, scale=100, ang_max=np.pi / 2,
                  ang_min=-np.pi / 2, processes=8)
    gold = golden.npz("reference_goldens.npz")["synthetic_match1"]
    for i in range(4):
        assert np.allclose(res[i], gold[i])


def test_oracle_equals_reference_runs(golden):
    runs = golden.npz("reference_runs.npz")
    for case in golden.meta["match_template_cases"]:
        z = golden.seeded_dem(case["dem"])
        amp, _, _, snr = O.match_template(z, case["de"], case["de"], KINDS[case["template"]],
                                          case["scale"], case["age"], case["angle"])
        assert np.array_equal(amp, runs[case["name"] + "_amp"]), case["name"]
        assert np.array_equal(snr, runs[case["name"] + "_snr"]), case["name"]
    info = golden.meta["search_a"]
    res = O.calculate_best_fit_parameters(golden.seeded_dem("a"), info["de"], info["de"], O.SCARP,
                                          info["scale"], info["age"], processes=4)
    assert np.array_equal(res, runs["search_a"])


def test_compare_semantics(golden):
    r1 = (np.array([[1., 2.], [3., 4.]]), 10., 0.1, np.array([[1., 5.], [2., 0.]]))
    r2 = (np.array([[5., 6.], [7., 8.]]), 20., 0.2, np.array([[1., 4.], [3., 0.]]))
    r3 = (np.array([[9., 9.], [9., 9.]]), 30., 0.3, np.array([[.5, 4.], [3., 1.]]))
    out = np.stack(O.compare([r1, r2, r3], 2, 2))
    assert np.array_equal(out, golden.npz("reference_runs.npz")["compare_out"])
    # exact tie -> reset to zero, then refilled by a later, lower SNR (core.py:230-240)
    assert out[3, 0, 0] == 0.5 and out[0, 0, 0] == 9.0
