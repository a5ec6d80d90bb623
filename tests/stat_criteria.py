"""Statistical acceptance criteria of the T-stat parity tier (independent RNG streams: Philox on the device, mt19937
in the oracle), stated once instead of ad-hoc multipliers.

Setting: K quantities are estimated on both sides from n_g and n_o replicate seeds; z_k = (mean_g - mean_o) / sqrt(se_g^2
+ se_o^2).  Under parity every z_k is approximately Student-t with nu = n_g + n_o - 2 degrees of freedom (the replicate
means are sums over many samples).

  north_star (BASELINE.json): "agree with the reference within 3 Monte Carlo standard errors".

  per-parameter 3 s.e.   |z_k| <= 3 for every k, EXCEPT that with K simultaneous comparisons the number of exceedances is
                         itself random: X ~ Binomial(K, P(|t_nu| > 3)).  The test allows X <= x_max with x_max the smallest
                         integer with P(X > x_max) <= alpha (alpha = 1e-3; for K <= 7 that is x_max = 0, i.e. the plain
                         3 s.e. rule).
  Bonferroni             max_k |z_k| <= t_nu^{-1}(1 - alpha / (2K)): no single comparison may be further out than the most
                         extreme of K null comparisons would be with probability alpha.
  omnibus (chi-square)   sum_k z_k^2 <= K E[t^2] + z_{1-alpha} sqrt(K Var[t^2]), E[t^2] = nu/(nu-2),
                         Var[t^2] = 2 nu^2 (nu-1) / ((nu-2)^2 (nu-4)): catches a small systematic bias spread over many
                         quantities, which neither of the first two sees.

A parity test passes when all three hold.  Seeds are fixed, so a pass/fail is reproducible.
"""
import math

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
    chi_bound = K * m2 + stats.norm.isf(alpha) * math.sqrt(K * v2)
    chi = float(np.sum(z * z))
    info = dict(K=K, nu=nu, n_exceed_3se=n_exc, allowed_exceed=x_max, max_abs_z=zmax, bonferroni=bonf, chi2=chi,
                chi2_bound=chi_bound)
    assert n_exc <= x_max, (label, "per-parameter 3 s.e.", info)
    assert zmax <= bonf, (label, "Bonferroni", info)
    assert chi <= chi_bound, (label, "omnibus", info)
    return info
