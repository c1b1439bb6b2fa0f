"""biseqt_b200.band against the laws the reference's own tests check for its band helpers
(tests/test_blot.py:35-108 of the reference, re-expressed for py3: monotonicity of
expected_overlap_len over diagonals / gap probability, the sqrt(K) law of band_radius, and the
vectorised forms against the scalar ones) plus the closed forms of biseqt/blot.py:78-160."""
from math import ceil, sqrt

import numpy as np
from scipy.special import erfcinv

from biseqt_b200.band import band_radii, band_radius, expected_overlap_len, wall_to_wall_distance


def test_expected_overlap_len():
    n, gap = 50, .1
    ds = range(-n, n + 1)
    aln_lens = np.array([expected_overlap_len(n, n, d, gap) for d in ds])
    assert np.all(np.diff(aln_lens[0:n]) >= 0)      # grows as d goes from -|T| to 0
    assert aln_lens[n] >= n                          # at diagonal 0 at least n
    assert np.all(np.diff(aln_lens[n:]) <= 0)        # shrinks as d goes from 0 to |S|
    gaps = [i * .05 for i in range(6)]
    assert np.all(np.diff([expected_overlap_len(n, n, 0, g) for g in gaps]) >= 0)
    N = [i * 100 for i in range(1, 5)]
    assert np.all(np.array([expected_overlap_len(n, N, 0, g) for g in gaps]) < 70)   # the shorter sequence rules
    # closed form, blot.py:78-111, and the vectorised call
    for (l0, l1, d, g) in ((50, 70, -20, .1), (50, 70, 10, .25), (1000, 900, 300, .15)):
        L = min(l0 - d, l1) + min(d, 0)
        assert wall_to_wall_distance(l0, l1, d) == L
        assert expected_overlap_len(l0, l1, d, g) == int(ceil(2. / (2 - g) * L))
    d = np.arange(-40, 41)
    vec = expected_overlap_len(50, np.full(len(d), 60), d, .1)
    assert list(vec) == [expected_overlap_len(50, 60, int(x), .1) for x in d]


def test_band_radius():
    gap_prob, sensitivity = .1, 1 - 1e-3
    Ks = [i * 200 for i in range(1, 10)]
    ratios = np.array([band_radius(K, gap_prob, sensitivity) / np.sqrt(K) for K in Ks])
    assert np.allclose(ratios - ratios[0], 0, atol=1e-1)    # r ~ sqrt(K)
    n = 50
    assert np.all(np.diff([band_radius(n, i * .05, .99) for i in range(1, 7)]) >= 0)       # grows with g
    assert np.all(np.diff([band_radius(n, gap_prob, 1 - i * .05) for i in range(1, 7)]) <= 0)   # shrinks with eps
    assert band_radius(54055, .15, .99) == int(ceil(erfcinv(.01) * sqrt(2 * .15 * 54055)))   # r = 232
    assert band_radius(54055, .15, .99) == 232
    assert band_radius(0, .1, .99) == 1                      # at least 1 (blot.py:141)


def test_band_radii():
    n, g, s = 50, .1, .99
    radii = band_radii(range(n), gap_prob=g, sensitivity=s)
    assert len(radii) == n and np.all(np.diff(radii) >= 0)
    assert list(radii) == [band_radius(K, g, s) for K in range(n)]
    lens = [expected_overlap_len(n, n, diag, g) for diag in range(-n, n)]
    radii = band_radii(lens, g, s)
    assert len(radii) == 2 * n
    assert np.all(np.diff(radii[0:n]) >= 0) and np.all(np.diff(radii[n:]) <= 0)
