"""Band derivation helpers of the reference (``biseqt/blot.py:78-160``): the
step immediately before the banded DP path in every real caller
(``experiments/blot_stats.py:453-456``).  Host-side, vectorised."""
import numpy as np
from scipy.special import erfcinv


def wall_to_wall_distance(len0, len1, diag):
    """L = min(l0 - d, l1) + min(d, 0)   (blot.py:78-88); arrays welcome."""
    return np.minimum(np.asarray(len0) - diag, len1) + np.minimum(diag, 0)


def expected_overlap_len(len0, len1, diag, gap_prob):
    """K = ceil(2 / (2 - g) * L)   (blot.py:91-111)."""
    L = wall_to_wall_distance(len0, len1, diag)
    expected_len = (2. / (2 - gap_prob)) * L
    assert np.all(expected_len >= 0)
    out = np.ceil(expected_len).astype(np.int64)
    return int(out) if out.ndim == 0 else out


def band_radius(expected_len, gap_prob, sensitivity):
    """r = ceil(erfcinv(1 - sensitivity) * sqrt(2 g K)), at least 1 (blot.py:115-141)."""
    assert 0 < gap_prob < 1 and 0 < sensitivity < 1
    epsilon = 1. - sensitivity
    C = erfcinv(epsilon) * np.sqrt(2 * gap_prob)
    radius = C * np.sqrt(expected_len)
    return max(1, int(np.ceil(radius)))


def band_radii(expected_lens, gap_prob, sensitivity):
    """Bulk :func:`band_radius` (blot.py:144-160), vectorised."""
    assert 0 < gap_prob < 1 and 0 < sensitivity < 1
    epsilon = 1. - sensitivity
    C = erfcinv(epsilon) * np.sqrt(2 * gap_prob)
    K = np.asarray(expected_lens, dtype=np.float64)
    return np.maximum(1, np.ceil(C * np.sqrt(K)).astype(np.int64))
