"""Synthetic-input generators: the benchmark/test input specification of the reference, restated in numpy.

* random_poset        -- R/common.r:48-86 (random_posets: ceil(choose(p,2)*density) strictly-upper-triangular
                         cover candidates drawn without replacement, then transitive reduction :107-119)
* sample_genotypes    -- R/cbn_sampler.r:93-135 (forward simulation of the CBN + i.i.d. bit-flip noise eps)
* topological_sort    -- R/common.r my.topological.sort (any valid order is equivalent)

R's RNG stream cannot be reproduced here (no R in the image) and does not need to be: inputs only have to
follow the same distribution.  numpy's default_rng(seed) is the generator.
"""
import math

import numpy as np


def transitive_closure(A):
    A = (np.asarray(A) != 0)
    p = A.shape[0]
    R = A.copy()
    for k in range(p):
        R |= np.outer(R[:, k], R[k, :])
    return R.astype(np.int32)


def transitive_reduction(A):
    """Cover relations of the DAG A (edge i->j is dropped when j is reachable from i by a longer path)."""
    A = (np.asarray(A) != 0)
    C = transitive_closure(A) != 0
    # i->j redundant iff exists k: i->k (closure) and k->j (closure)
    two = (C.astype(np.int64) @ C.astype(np.int64)) > 0
    return (A & ~two).astype(np.int32)


def random_poset(p, graph_density=0.15, trans_reduced=True, rng=None):
    rng = np.random.default_rng(rng)
    n_pairs = p * (p - 1) // 2
    n_edges = int(math.ceil(n_pairs * graph_density))
    poset = np.zeros((p, p), np.int32)
    iu = np.triu_indices(p, k=1)
    pick = rng.choice(n_pairs, size=n_edges, replace=False)
    poset[iu[0][pick], iu[1][pick]] = 1
    if trans_reduced:
        poset = transitive_reduction(poset)
    return np.asfortranarray(poset)


def topological_sort(poset):
    poset = np.asarray(poset)
    p = poset.shape[0]
    indeg = poset.sum(axis=0).astype(int)
    order, ready = [], [j for j in range(p) if indeg[j] == 0]
    while ready:
        j = ready.pop(0)
        order.append(j)
        for c in np.nonzero(poset[j])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(int(c))
    if len(order) != p:
        raise ValueError("Poset does not correspond to an acyclic graph")
    return order


def sample_genotypes(n, poset, sampling_param, lambdas, eps=0.0, rng=None):
    """R/cbn_sampler.r:93-135 with sampling_fn = sampling.expo: T_s ~ Exp(rate = 1/sampling_param)."""
    rng = np.random.default_rng(rng)
    poset = np.asarray(poset)
    lambdas = np.asarray(lambdas, dtype=np.float64)
    p = lambdas.size
    T_sampling = rng.exponential(sampling_param, size=n)
    T_events = rng.exponential(1.0, size=(n, p)) / lambdas[None, :]
    T_sum = np.zeros((n, p))
    for e in topological_sort(poset):
        parents = np.nonzero(poset[:, e])[0]
        if parents.size == 0:
            T_sum[:, e] = T_events[:, e]
        else:
            T_sum[:, e] = T_events[:, e] + T_sum[:, parents].max(axis=1)
    hidden = (T_sum <= T_sampling[:, None]).astype(np.int32)
    flip = rng.random((n, p)) < eps
    obs = np.where(flip, 1 - hidden, hidden).astype(np.int32)
    return dict(n=n, p=p, T_sampling=T_sampling, obs_events=np.asfortranarray(obs),
                hidden_genotypes=np.asfortranarray(hidden), T_events=T_events, T_sum_events=T_sum,
                lambdas=lambdas, eps=eps, sampling_param=sampling_param)


def make_config(p, N, eps=0.05, seed=10, lambda_s=1.0, density=0.15):
    """One synthetic problem per SURVEY.md section 8(d): poset, true lambdas, noisy genotypes, start values."""
    rng = np.random.default_rng(seed)
    poset = random_poset(p, density, rng=rng)
    lambdas = rng.uniform(lambda_s / 3.0, 3.0 * lambda_s, size=p)
    sim = sample_genotypes(N, poset, 1.0 / lambda_s, lambdas, eps=eps, rng=rng)
    lambda0 = rng.uniform(lambda_s / 3.0, 3.0 * lambda_s, size=p)
    return dict(poset=poset, lambdas=lambdas, obs=sim["obs_events"], times=sim["T_sampling"], lambda0=lambda0,
                eps=eps, lambda_s=lambda_s, hidden=sim["hidden_genotypes"])
