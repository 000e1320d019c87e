import numpy as np


def relrms(a, b):
    """Relative RMS of (a - b) against b over pixels finite in both."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = np.isfinite(a) & np.isfinite(b)
    return float(np.sqrt(np.mean((a[ok] - b[ok]) ** 2)) / np.sqrt(np.mean(b[ok] ** 2)))


def golden_diff(case):
    """(full-resolution golden DIFF or its float32 copy, float64 ::8 subsample or None)."""
    if 'REFRUN_DIFF' in case:
        return case['REFRUN_DIFF'], None
    return case['REFRUN_DIFF_f32'].astype(np.float64), case['REFRUN_DIFF_sub8']
