"""Forward simulators of the generative model on the device (`_sample_genotypes`, `_sample_times`,
`_generate_genotypes`, src/mcem_hcbn.cpp:1012-1120) against the oracle and the exact genotype probabilities."""
import numpy as np
import pytest

import exact
import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check

NU_NORMAL = 10 ** 6  # large-sample z statistics (N >= 3000 draws): the joint rule of stat_criteria with normal tails

pytestmark = pytest.mark.gpu


def _consistent(poset, lam, r):
    """structural identities every draw satisfies exactly: T_j = Z_j + max parents, X_j = [T_j <= T_s]"""
    p = poset.shape[0]
    order = synth.topological_sort(poset)
    Z = r["Tdiff"]
    T = np.zeros_like(Z)
    for j in order:
        par = np.nonzero(poset[:, j])[0]
        T[:, j] = Z[:, j] + (T[:, par].max(axis=1) if par.size else 0.0)
    return T


def test_sample_times_structure_and_marginals():
    c = synth.make_config(9, 5, seed=3, density=0.3)
    N = 20000
    r = mc.sample_times(N, c["poset"], c["lambdas"], seed=5)
    assert np.all(r["Tdiff"] > 0)
    np.testing.assert_allclose(r["mutation_times"], _consistent(c["poset"], c["lambdas"], r), rtol=1e-14)
    # waiting times are Exp(lambda_j): mean 1/lambda, variance 1/lambda^2
    z = (r["Tdiff"].mean(axis=0) - 1 / c["lambdas"]) / (1 / c["lambdas"] / np.sqrt(N))
    joint_check(z, NU_NORMAL, "Exp(lambda) means")
    # reproducible, seed-dependent
    assert np.array_equal(mc.sample_times(100, c["poset"], c["lambdas"], seed=5)["Tdiff"], r["Tdiff"][:100])
    assert not np.array_equal(mc.sample_times(100, c["poset"], c["lambdas"], seed=6)["Tdiff"], r["Tdiff"][:100])


def test_sample_genotypes_matches_exact_distribution(oracle):
    c = synth.make_config(6, 5, seed=11, density=0.3)
    N = 40000
    r = mc.sample_genotypes(N, c["poset"], c["lambdas"], lambda_s=1.0, seed=9)
    T = _consistent(c["poset"], c["lambdas"], r)
    assert np.array_equal(r["samples"], (T <= r["sampling_time"][:, None]).astype(np.int32))
    # every sampled genotype is compatible with the poset
    assert mc.compatible_observations(r["samples"], c["poset"])["num_compatible"] == N
    # genotype frequencies against the exact lattice recursion (K2), chi-square-like z per genotype
    pd = exact.genotype_probs_exp(c["poset"], c["lambdas"], 1.0)
    px = np.array([pd.get(x, 0.0) for x in range(64)])
    codes = (r["samples"] << np.arange(6)).sum(axis=1)
    freq = np.bincount(codes, minlength=64) / N
    se = np.sqrt(np.maximum(px * (1 - px), 1e-12) / N)
    joint_check(((freq - px) / se)[px > 1e-4], NU_NORMAL, "genotype frequencies vs K2")
    assert freq[px == 0].sum() == 0
    # same marginals as the oracle's mt19937 draw
    o_samples, _, _ = oracle.sample_genotypes(N, c["poset"], c["lambdas"], seed=4)
    fo = o_samples.mean(axis=0)
    fg = r["samples"].mean(axis=0)
    joint_check((fg - fo) / np.sqrt(fg * (1 - fg) / N + fo * (1 - fo) / N), NU_NORMAL, "marginals vs oracle")


def test_sample_genotypes_with_given_times_and_generate_genotypes(oracle):
    c = synth.make_config(8, 5, seed=2, density=0.3)
    N = 3000
    times = np.random.default_rng(0).uniform(0.2, 3.0, N)
    r = mc.sample_genotypes(N, c["poset"], c["lambdas"], times=times, seed=3)
    T = _consistent(c["poset"], c["lambdas"], r)
    assert np.array_equal(r["sampling_time"], times)
    assert np.array_equal(r["samples"], (T <= times[:, None]).astype(np.int32))
    # generate_genotypes from supplied event times and sampling times: pure comparison, bit-exact against the oracle
    g = mc.generate_genotypes(T, c["poset"], times=times)
    o_samples, _ = oracle.generate_genotypes(T, c["poset"], times=times)
    assert np.array_equal(g["samples"], o_samples) and np.array_equal(g["samples"], r["samples"])
    # without times: T_s ~ Exp(lambda_s)
    g2 = mc.generate_genotypes(T, c["poset"], lambda_s=2.0, seed=8)
    assert np.array_equal(g2["samples"], (T <= g2["sampling_time"][:, None]).astype(np.int32))
    z = (g2["sampling_time"].mean() - 0.5) / (0.5 / np.sqrt(N))
    joint_check(np.array([z]), NU_NORMAL, "T_s ~ Exp(lambda_s)")
