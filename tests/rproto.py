"""A SECOND, independent statement of the H-CBN2 E-step / M-step mathematics: the reference's pure-R prototype
(/root/reference/hcbn_functions.r), restated in numpy.  The oracle (oracle/) follows the C++ sources; this module
follows the R prototype the authors validated the C++ against.  Cross-checking the two on random inputs
(tests/test_second_statement.py) pins the oracle by two independent readings of the reference instead of one.
Plain numpy, independent of the oracle and of the device code.  Citations: hcbn_functions.r line numbers.

R draws (rexp, runif inside rtexp) are replaced by inverse-CDF transforms of SUPPLIED uniforms so that both statements
can be evaluated on identical randomness.
"""
import numpy as np


def hamming_dist(x, y):
    """hamming.dist :7-11"""
    return int(np.sum(np.abs(np.asarray(x, int) - np.asarray(y, int))))


def get_parents(poset):
    poset = np.asarray(poset)
    return [np.nonzero(poset[:, j])[0] for j in range(poset.shape[0])]


def topological_sort(poset):
    """my.topological.sort (R/common.r): any valid order gives the same times"""
    poset = np.asarray(poset)
    p = poset.shape[0]
    indeg = poset.sum(axis=0).astype(int).copy()
    order, ready = [], [j for j in range(p) if indeg[j] == 0]
    while ready:
        j = ready.pop(0)
        order.append(j)
        for c in np.nonzero(poset[j])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(int(c))
    return order


def rtexp_from_uniform(x, rate, u):
    """rtexp :491-495 with rand_num = u"""
    return -np.log(1.0 - u * (1.0 - np.exp(-rate * x))) / rate if x > 0 else 0.0


def sample_mutation_times(genotype, poset, lambdas, sampling_time, u):
    """sample.mutation.times :503-547 (conditional proposal given a hidden genotype) from supplied uniforms u[j].
    Returns (log proposal density, Tdiff)."""
    genotype = np.asarray(genotype)
    p = genotype.size
    parents = get_parents(poset)
    Tdiff, Tsum, dens = np.zeros(p), np.zeros(p), 0.0
    for j in topological_sort(poset):
        mx = 0.0
        for par in parents[j]:
            if Tsum[par] > mx:
                mx = Tsum[par]
        lam = lambdas[j]
        if genotype[j] == 1:
            c = sampling_time - mx
            Z = rtexp_from_uniform(c, lam, u[j])                       # :534
            Tsum[j] = mx + Z
            # dexp(Z, log = TRUE) - pexp(c, log.p = TRUE)                 :536-537
            dens += (np.log(lam) - lam * Z) - np.log(1.0 - np.exp(-lam * c))
        else:
            Z = -np.log(1.0 - u[j]) / lam                              # rexp(1, rate) :540, inverse CDF
            Tsum[j] = max(sampling_time, mx) + Z
            dens += np.log(lam) - lam * Z                              # :542
        Tdiff[j] = Tsum[j] - mx                                        # :544
    return dens, Tdiff


def log_cbn_density(Tdiff, rate):
    """log.cbn.density :551-554"""
    return float(np.sum(np.log(rate)) - np.sum(np.asarray(rate) * np.asarray(Tdiff)))


def forward_weights(genotype, poset, eps, Z, T_sampling):
    """forward branch of importance.weight_ :590-604 on supplied waiting times (sample_genotypes of R/cbn_sampler.r
    :104-119 for the hidden genotypes): w = eps^d (1 - eps)^(p - d)."""
    Z = np.asarray(Z, float)
    L, p = Z.shape
    parents = get_parents(poset)
    T = np.zeros((L, p))
    for j in topological_sort(poset):
        T[:, j] = Z[:, j] + (T[:, parents[j]].max(axis=1) if len(parents[j]) else 0.0)
    X = (T <= np.asarray(T_sampling, float)[:, None]).astype(int)
    d = np.array([hamming_dist(X[l], genotype) for l in range(L)])
    w = eps ** d * (1 - eps) ** (p - d)
    return dict(w=w, dist=d, X=X, T_sum=T)


def m_step(expected_dist, expected_Tdiff, max_lambda=1e6):
    """M step of MCEM.hcbn_ :1391-1398: eps = mean(expected.dist) / p, lambda = 1 / colMeans(expected.Tdiff), capped"""
    expected_Tdiff = np.asarray(expected_Tdiff, float)
    p = expected_Tdiff.shape[1]
    eps = float(np.mean(expected_dist)) / p
    lambdas = 1.0 / expected_Tdiff.mean(axis=0)
    lambdas = np.where(lambdas > max_lambda, max_lambda, lambdas)
    return lambdas, eps


def complete_loglikelihood(lambdas, eps, Tdiff, dist):
    """complete.loglikelihood_ :64-79"""
    lambdas, Tdiff, dist = np.asarray(lambdas, float), np.asarray(Tdiff, float), np.asarray(dist, float)
    p, N = lambdas.size, dist.size
    ll = N * np.sum(np.log(lambdas)) - np.sum(lambdas[None, :] * Tdiff)
    if eps == 0:
        if not np.all(dist == 0):
            me = np.finfo(float).eps
            ll += np.log(eps + me) * dist.sum() + np.log(1 - (eps + me)) * np.sum(p - dist)
    else:
        ll += np.log(eps) * dist.sum() + np.log(1 - eps) * np.sum(p - dist)
    return float(ll)
