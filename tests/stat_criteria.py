"""Statistical acceptance criteria of the T-stat parity tier (independent RNG streams: Philox on the device, mt19937
in the oracle), stated once instead of ad-hoc multipliers.

Setting: K quantities are estimated on both sides from n_g and n_o replicate seeds; z_k = (mean_g - mean_o) / sqrt(se_g^2
+ se_o^2).  Under parity every z_k is approximately Student-t with nu = n_g + n_o - 2 degrees of freedom (the replicate
means are sums over many samples); against an exact answer nu = n - 1, for large-sample frequencies nu = 1e6 (normal).

  north_star (BASELINE.json): "agree with the reference within 3 Monte Carlo standard errors".

A single comparison of a true null exceeds 3 s.e. with probability P(|t_nu| > 3) = 0.27 % (normal) .. 0.9 % (nu = 15), and
a test compares K = 1 .. 2000 quantities at once, so "every |z_k| <= 3" cannot be the rule: it would fail correct code
with probability 1 - (1 - p_3)^K (5 % at K = 10, certainty at K = 2000).  The rule is family-wise, alpha = 1e-3 per test
(the suite holds ~60 such tests: < 6 % chance of one false alarm over the whole suite), and has three parts, ALL of
which must hold:

  exceedances of 3 s.e.  X = #{k: |z_k| > 3} ~ Binomial(K, P(|t_nu| > 3)) under parity; allowed: X <= x_max, the smallest
                         integer with P(X > x_max) <= alpha.  (K = 1: one exceedance is tolerated, but only up to the
                         Bonferroni bound below, i.e. |z| <= t_nu^{-1}(1 - alpha/2) = 3.3 (normal) .. 4.1 (nu = 15).)
  Bonferroni             max_k |z_k| <= t_nu^{-1}(1 - alpha / (2K)): no single comparison may be further out than the most
                         extreme of K null comparisons would be with probability alpha.
  omnibus                sum_k z_k^2 <= a chi2_b^{-1}(1 - alpha) with the scaled chi-square that matches the first two
                         moments of a sum of K squared t_nu variables, E[t^2] = nu/(nu-2), Var[t^2] = 2 nu^2 (nu-1) /
                         ((nu-2)^2 (nu-4)):  a = Var/(2 E), b = 2 K E^2 / Var  (nu -> inf: the chi-square with K degrees of
                         freedom).  Catches a small systematic bias spread over many quantities, which neither of the
                         first two sees.

Seeds are fixed, so a pass/fail is reproducible.  With MCCBN_STAT_LOG=<file> every check appends its statistics (K, nu,
number of exceedances of 3 s.e., max |z|, the bounds) to that file: profiles/ keeps the table of a full GPU run, so the
margin of every parity comparison -- not just pass/fail -- is on record.
"""
import math
import os

import numpy as np
from scipy import stats

ALPHA = 1e-3


def replicate_z(g, o):
    """g, o: (n_rep, K) replicate estimates.  Returns z (K,), degrees of freedom."""
    g, o = np.asarray(g, float), np.asarray(o, float)
    if g.ndim == 1:
        g, o = g[:, None], o[:, None]
    vg, vo = g.var(axis=0, ddof=1) / g.shape[0], o.var(axis=0, ddof=1) / o.shape[0]
    diff = g.mean(axis=0) - o.mean(axis=0)
    se = np.sqrt(vg + vo)
    # quantities that are deterministic on both sides (e.g. the expected waiting time of an event whose weight does not
    # depend on the draw) must agree to fp32 sampling accuracy
    scale = np.maximum(np.abs(o.mean(axis=0)), 1e-300)
    z = np.where(se > 1e-7 * scale, diff / np.maximum(se, 1e-300), np.where(np.abs(diff) <= 2e-6 * scale, 0.0, np.inf))
    return z, g.shape[0] + o.shape[0] - 2


def joint_check(z, nu, label="", alpha=ALPHA):
    """Asserts the three criteria of the module docstring; returns a dict with the observed statistics."""
    z = np.asarray(z, float).ravel()
    z = z[~np.isnan(z)]
    K = z.size
    assert K > 0
    p3 = 2.0 * stats.t.sf(3.0, nu)
    x_max = int(stats.binom.isf(alpha, K, p3))  # smallest x with P(X > x) <= alpha
    n_exc = int(np.sum(np.abs(z) > 3.0))
    bonf = float(stats.t.isf(alpha / (2.0 * K), nu))
    zmax = float(np.max(np.abs(z)))
    m2 = nu / (nu - 2.0)
    v2 = 2.0 * nu * nu * (nu - 1.0) / ((nu - 2.0) ** 2 * (nu - 4.0))
    a, b = v2 / (2.0 * m2), 2.0 * K * m2 * m2 / v2
    chi_bound = float(a * stats.chi2.isf(alpha, b))
    chi = float(np.sum(z * z))
    info = dict(K=K, nu=nu, n_exceed_3se=n_exc, allowed_exceed=x_max, max_abs_z=zmax, bonferroni=bonf, chi2=chi,
                chi2_bound=chi_bound)
    log = os.environ.get("MCCBN_STAT_LOG")
    if log:
        with open(log, "a") as f:
            f.write("%-60s K=%-5d nu=%-7d exceed_3se=%d (allowed %d)  max|z|=%.2f (Bonferroni %.2f)  sum z^2=%.1f (bound %.1f)\n"
                    % (label, K, nu, n_exc, x_max, zmax, bonf, chi, chi_bound))
    assert n_exc <= x_max, (label, "exceedances of 3 s.e.", info)
    assert zmax <= bonf, (label, "Bonferroni", info)
    assert chi <= chi_bound, (label, "omnibus", info)
    return info
