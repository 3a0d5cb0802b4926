"""Shared parity metrics (north_star tolerances)."""
import numpy as np

AMP_SNR_RTOL = 1e-4        # north_star: amplitude and SNR relative error <= 1e-4
INDEX_AGREEMENT = 0.999    # north_star: best angle/age identical on >= 99.9 % of valid pixels


def stack_report(res, ref, odd_template=True):
    """Compare two [amp, age, angle, snr] stacks.

    Index agreement is counted modulo the +-90 degree degeneracy (the first and last
    orientation are the same line; for odd templates the amplitude flips sign), as the
    reference itself resolves that pair by 1e-11-level noise."""
    res = np.asarray(res)
    ref = np.asarray(ref)
    amp, age, ang, snr = res
    ramp, rage, rang, rsnr = ref
    valid = rsnr > 0
    nvalid = int(valid.sum())
    same_age = np.isclose(age, rage, rtol=1e-12, atol=0)
    same = np.isclose(ang, rang, rtol=0, atol=1e-12) & same_age
    a_lo, a_hi = rang[valid].min() if nvalid else 0, rang[valid].max() if nvalid else 0
    degenerate = np.zeros_like(valid)
    if nvalid and np.isclose(a_hi - a_lo, np.pi):
        degenerate = ((np.isclose(ang, a_lo) & np.isclose(rang, a_hi)) |
                      (np.isclose(ang, a_hi) & np.isclose(rang, a_lo))) & same_age
    agree = (same | degenerate) & valid
    sign = np.where(degenerate & odd_template, -1.0, 1.0)
    with np.errstate(divide='ignore', invalid='ignore'):
        snr_rel = np.abs(snr - rsnr) / np.abs(rsnr)
        amp_rel = np.abs(sign * amp - ramp) / np.abs(ramp)
    # Pixels the reference zeroed through compare's exact-tie reset (core.py:230-240): the
    # reference holds 0 while the first-wins device fold keeps the tied maximum.  For the
    # +-90 degree pair the tie is exact in ~7 % of the pixels whose best fit lies there
    # (SURVEY.md 8a-3), so those show up as "mask" differences that are not mask errors.
    tie_reset = (rsnr == 0) & (snr > 0)
    if nvalid and np.isclose(a_hi - a_lo, np.pi):
        tie_reset &= (np.isclose(ang, a_lo) | np.isclose(ang, a_hi))
    return {
        "valid": nvalid,
        "tie_reset_pixels": int(tie_reset.sum()),
        "mask_mismatch_unexplained": int((((snr > 0) != valid) & ~tie_reset).sum()),
        "mask_equal": bool(((snr > 0) == valid).all()),
        "mask_mismatch": int(((snr > 0) != valid).sum()),
        "index_agreement": float(agree.sum() / max(nvalid, 1)),
        "snr_rel_max": float(snr_rel[agree].max()) if agree.any() else 0.0,
        "snr_rel_p50": float(np.median(snr_rel[agree])) if agree.any() else 0.0,
        "amp_rel_p999": float(np.quantile(amp_rel[agree], 0.999)) if agree.any() else 0.0,
        "amp_rel_max": float(amp_rel[agree].max()) if agree.any() else 0.0,
        "frac_snr_over_tol": float((snr_rel[agree] > AMP_SNR_RTOL).mean()) if agree.any() else 0.0,
        "frac_amp_over_tol": float((amp_rel[agree] > AMP_SNR_RTOL).mean()) if agree.any() else 0.0,
        "disagree_snr_gap_max": float(snr_rel[valid & ~agree].max()) if (valid & ~agree).any() else 0.0,
    }
