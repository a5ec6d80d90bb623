"""Model-derived exact answers for the H-CBN2 model -- test infrastructure.

The reference ships no golden vectors for this path; these closed forms are what pins the oracle (and through
it the CUDA path) to the *model* the reference implements:

  Z_j ~ Exp(lambda_j), T_j = Z_j + max_{u in pa(j)} T_u, X_j = [T_j <= T_s], P(Y|X) = eps^d (1-eps)^(p-d).

K2  T_s ~ Exp(lambda_s): the hidden genotype performs a race of exponentials on the lattice of compatible
    genotypes.  Exit(x) = {k not in x : pa(k) subset of x}, Lambda(x) = sum_{k in Exit(x)} lambda_k,
    reach(empty) = 1, reach(x) = sum_{j in x, j in Exit(x\\j)} reach(x\\j) lambda_j / (lambda_s + Lambda(x\\j)),
    P(X = x) = reach(x) lambda_s / (lambda_s + Lambda(x)).
K2' T_s = t given: P(X(t) = x) = [exp(Q t)]_{empty, x} with generator Q over compatible genotypes.
K1  empty poset, T_s = t given: events independent, P(X_j = 1) = 1 - exp(-lambda_j t).
"""
import itertools
import math

import numpy as np


def _parent_masks(poset):
    poset = np.asarray(poset)
    p = poset.shape[0]
    return [sum(1 << int(u) for u in np.nonzero(poset[:, j])[0]) for j in range(p)]


def compatible_states(poset):
    p = np.asarray(poset).shape[0]
    pm = _parent_masks(poset)
    return [x for x in range(1 << p) if all((pm[j] & ~x) == 0 for j in range(p) if (x >> j) & 1)]


def genotype_probs_exp(poset, lambdas, lambda_s):
    """K2: dict {x (bitmask over event ids): P(X = x)} for T_s ~ Exp(lambda_s)."""
    p = len(lambdas)
    pm = _parent_masks(poset)
    states = sorted(compatible_states(poset), key=lambda x: bin(x).count("1"))
    comp = set(states)

    def exit_rate(x):
        return sum(lambdas[k] for k in range(p) if not (x >> k) & 1 and (pm[k] & ~x) == 0)

    reach = {0: 1.0}
    for x in states:
        if x == 0:
            continue
        r = 0.0
        for j in range(p):
            if (x >> j) & 1:
                y = x & ~(1 << j)
                if y in comp and (pm[j] & ~y) == 0:
                    r += reach[y] * lambdas[j] / (lambda_s + exit_rate(y))
        reach[x] = r
    return {x: reach[x] * lambda_s / (lambda_s + exit_rate(x)) for x in states}


def genotype_probs_time(poset, lambdas, t):
    """K2': dict {x: P(X(t) = x)} for a fixed sampling time t (matrix exponential on the lattice)."""
    from scipy.linalg import expm
    p = len(lambdas)
    pm = _parent_masks(poset)
    states = compatible_states(poset)
    idx = {x: i for i, x in enumerate(states)}
    Q = np.zeros((len(states), len(states)))
    for x in states:
        for k in range(p):
            if not (x >> k) & 1 and (pm[k] & ~x) == 0:
                Q[idx[x], idx[x | (1 << k)]] += lambdas[k]
        Q[idx[x], idx[x]] = -Q[idx[x]].sum()
    P = expm(Q * t)[idx[0]]
    return {x: P[idx[x]] for x in states}


def obs_prob_and_dist(px, y_mask, eps, p):
    """P(Y) and E[d(X,Y) | Y] from a genotype distribution px = {x: P(x)}."""
    py, pd = 0.0, 0.0
    for x, q in px.items():
        d = bin(x ^ y_mask).count("1")
        w = (eps ** d) * ((1 - eps) ** (p - d))
        py += q * w
        pd += q * w * d
    return py, pd / py


def mask_of(genotype):
    return sum(1 << j for j, g in enumerate(np.asarray(genotype).ravel()) if g)


def exact_obs_loglik(obs, poset, lambdas, eps, lambda_s=1.0, times=None, weights=None):
    obs = np.asarray(obs)
    N, p = obs.shape
    weights = np.ones(N) if weights is None else weights
    ll, ed = 0.0, np.zeros(N)
    px = genotype_probs_exp(poset, lambdas, lambda_s) if times is None else None
    for i in range(N):
        if times is not None:
            px = genotype_probs_time(poset, lambdas, times[i])
        py, d = obs_prob_and_dist(px, mask_of(obs[i]), eps, p)
        ll += weights[i] * math.log(py)
        ed[i] = d
    return ll, ed


def k1_empty_poset(y, lambdas, eps, t):
    """K1: P(Y | t) and E[d | Y, t] on the empty poset with sampling time t."""
    F = 1 - np.exp(-np.asarray(lambdas) * t)
    y = np.asarray(y).astype(bool)
    # per event: P(Y_j = y_j) = F(1-eps) + (1-F)eps for y_j = 1 ; F eps + (1-F)(1-eps) for y_j = 0
    p1 = np.where(y, F * (1 - eps) + (1 - F) * eps, F * eps + (1 - F) * (1 - eps))
    # P(X_j != Y_j | Y_j)
    pm = np.where(y, (1 - F) * eps, F * eps) / p1
    return float(np.prod(p1)), float(pm.sum())


def all_genotypes(p):
    return np.array(list(itertools.product([0, 1], repeat=p)), dtype=np.int32)


def forward_weight_moments(px, y_mask, eps, p):
    """E[w] and E[w^2] of the forward scheme's importance weight w = eps^d (1-eps)^(p-d), d = d(X, Y), X ~ px.
    The reference's diagnostic (utils/hcbn_test.r:415-590) is the effective sample size (sum w)^2 / sum w^2, whose
    large-L limit is L E[w]^2 / E[w^2]."""
    m1 = m2 = 0.0
    for x, q in px.items():
        d = bin(x ^ y_mask).count("1")
        w = eps ** d * (1 - eps) ** (p - d)
        m1 += q * w
        m2 += q * w * w
    return m1, m2


def brute_force_posterior(poset, lambdas, eps, M, seed=0, lambda_s=1.0):
    """The reference's validation method T1 (utils/hcbn_test.r:101-193, hcbn_functions.r:163-203): simulate M draws of
    the generative model with noise (R/cbn_sampler.r:93-135) in plain numpy and tabulate, per observed genotype y,
    the frequency P(Y = y) and the conditional means E[Z_j | Y = y], E[d(X, Y) | Y = y] with their standard errors.
    Independent of the oracle and of the device code.  Returns {y_mask: dict(n, prob, Z, Z_se, d, d_se)}."""
    poset = np.asarray(poset)
    lam = np.asarray(lambdas, float)
    p = lam.size
    rng = np.random.default_rng(seed)
    Z = rng.exponential(1.0, (M, p)) / lam
    Ts = rng.exponential(1.0 / lambda_s, M)
    # topological order by repeated relaxation (p is small)
    indeg = poset.sum(axis=0).astype(int)
    order, ready = [], [j for j in range(p) if indeg[j] == 0]
    while ready:
        j = ready.pop(0)
        order.append(j)
        for c in np.nonzero(poset[j])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(int(c))
    T = np.zeros((M, p))
    for j in order:
        par = np.nonzero(poset[:, j])[0]
        T[:, j] = Z[:, j] + (T[:, par].max(axis=1) if par.size else 0.0)
    X = T <= Ts[:, None]
    Y = X ^ (rng.random((M, p)) < eps)
    d = (X ^ Y).sum(axis=1)
    key = (Y * (1 << np.arange(p))).sum(axis=1)
    out = {}
    for y in np.unique(key):
        sel = key == y
        n = int(sel.sum())
        if n < 2:
            continue
        out[int(y)] = dict(n=n, prob=n / M, Z=Z[sel].mean(axis=0), Z_se=Z[sel].std(axis=0, ddof=1) / math.sqrt(n),
                           d=float(d[sel].mean()), d_se=float(d[sel].std(ddof=1) / math.sqrt(n)))
    return out
