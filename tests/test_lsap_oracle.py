"""The LSAP restatement (oracle/lsap_oracle.py) is pinned against scipy itself: identical assignments on random,
rectangular (both orientations) and tie-heavy integer matrices."""
import numpy as np
from scipy.optimize import linear_sum_assignment as scipy_lsa

from oracle.lsap_oracle import linear_sum_assignment as oracle_lsa


def _check(c):
    r, k = scipy_lsa(c)
    ro, ko = oracle_lsa(c.tolist())
    assert list(r) == ro and list(k) == ko, (c.shape, list(r), ro, list(k), ko)


def test_random_float32_costs_all_shapes():
    rng = np.random.default_rng(0)
    for nr, nc in [(100, 40), (100, 1), (40, 100), (7, 7), (1, 9), (300, 25), (60, 60), (100, 37)]:
        for _ in range(6):
            _check(rng.standard_normal((nr, nc)).astype(np.float32).astype(np.float64))


def test_tie_heavy_integer_costs():
    rng = np.random.default_rng(1)
    for _ in range(300):
        nr, nc = rng.integers(1, 13), rng.integers(1, 13)
        _check(rng.integers(0, 3, size=(nr, nc)).astype(np.float64))
    _check(np.zeros((6, 6)))       # constant matrix -> identity (scipy gh-11602)
    _check(np.ones((9, 4)))
    _check(np.ones((4, 9)))
