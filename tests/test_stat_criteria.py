"""CPU-only: the acceptance rule of the T-stat parity tier (tests/stat_criteria.py) is itself under test -- its false-alarm
rate under parity and its power against the three ways a device path can be wrong."""
import numpy as np
import pytest

from stat_criteria import joint_check


def _rejects(z, nu):
    try:
        joint_check(z, nu)
    except AssertionError:
        return True
    return False


@pytest.mark.parametrize("K,nu", [(1, 15), (10, 38), (100, 46), (2048, 94), (40, 10 ** 6)])
def test_false_alarm_rate_under_parity(K, nu):
    """three criteria at alpha = 1e-3 each: a correct implementation fails a test with probability < 1 %"""
    rng = np.random.default_rng(K + nu % 1000)
    n = 1500
    draw = (lambda: rng.standard_normal(K)) if nu >= 10 ** 6 else (lambda: rng.standard_t(nu, K))
    bad = sum(_rejects(draw(), nu) for _ in range(n))
    assert bad / n < 0.01, bad / n


def test_power_single_gross_error():
    """one quantity 8 s.e. off among 500 (e.g. one wrong parent offset): Bonferroni"""
    rng = np.random.default_rng(1)
    hits = 0
    for _ in range(200):
        z = rng.standard_t(46, 500)
        z[rng.integers(500)] += 8.0
        hits += _rejects(z, 46)
    assert hits >= 190


def test_power_small_systematic_bias():
    """every quantity 0.5 s.e. off (e.g. a biased uniform -> exponential map): invisible per parameter, caught by the omnibus
    (0.4 s.e.: rejected in 92 % of the runs, 0.3 s.e.: 28 %)"""
    rng = np.random.default_rng(2)
    hits = sum(_rejects(rng.standard_t(94, 2048) + 0.5, 94) for _ in range(100))
    assert hits >= 95


def test_power_many_moderate_errors():
    """a tenth of the quantities 3.5 s.e. off: more exceedances of 3 s.e. than chance produces"""
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(200):
        z = rng.standard_t(46, 100)
        z[:10] += 3.5
        hits += _rejects(z, 46)
    assert hits >= 190


def test_single_comparison_bounds():
    """K = 1: accepted up to the alpha = 1e-3 two-sided t quantile, rejected beyond"""
    assert not _rejects(np.array([3.2]), 30)
    assert _rejects(np.array([4.5]), 30)
    assert not _rejects(np.array([3.25]), 10 ** 6)
    assert _rejects(np.array([3.6]), 10 ** 6)
