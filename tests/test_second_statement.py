"""CPU-only: the oracle against a second, independent statement of the mathematics -- the reference's pure-R prototype
hcbn_functions.r restated in numpy (tests/rproto.py; SURVEY 8c names :578-604, :503-554, :56-79, :1391-1398 as the
cross-check).  The reference ships no golden vectors for this path; agreement of the C++-derived oracle with the
R-derived statement on random inputs (<= 1e-12) is the strongest pin available without R."""
import numpy as np
import pytest

import rproto
from mccbn_b200 import synth


@pytest.mark.parametrize("p,seed", [(4, 0), (9, 1), (16, 2), (30, 3)])
def test_conditional_proposal_and_density(oracle, p, seed):
    """generate_mutation_times + cbn_density_log (src/add_remove.cpp:156-216) vs sample.mutation.times + log.cbn.density
    (hcbn_functions.r:503-554) on the same genotypes, sampling times and uniforms."""
    rng = np.random.default_rng(seed)
    poset = synth.random_poset(p, 0.25, rng=rng)
    lam = rng.uniform(0.3, 3.0, p)
    L = 40
    # compatible hidden genotypes (forward simulation), so that every cutoff T_s - max_parents is positive
    sim = synth.sample_genotypes(L, poset, 1.0, lam, eps=0.0, rng=rng)
    X = sim["hidden_genotypes"]
    Ts = sim["T_sampling"] + rng.uniform(0.0, 0.5, L)  # later sampling keeps X feasible
    u = rng.random((L, p))
    o = oracle.conditional_from_uniforms(X, poset, lam, u, Ts)
    for l in range(L):
        dens, Tdiff = rproto.sample_mutation_times(X[l], poset, lam, Ts[l], u[l])
        np.testing.assert_allclose(o["Tdiff"][l], Tdiff, rtol=1e-10, atol=1e-13)
        assert abs(o["log_q"][l] - dens) <= 1e-10 * max(1.0, abs(dens))
        assert abs(o["log_pz"][l] - rproto.log_cbn_density(Tdiff, lam)) <= 1e-10 * max(1.0, abs(o["log_pz"][l]))
        assert not o["incompatible"][l]
    np.testing.assert_allclose(oracle.cbn_density_log(o["Tdiff"], lam),
                               [rproto.log_cbn_density(o["Tdiff"][l], lam) for l in range(L)], rtol=1e-12)


@pytest.mark.parametrize("p,seed", [(5, 0), (12, 1), (30, 2)])
def test_forward_weights(oracle, p, seed):
    """forward branch of importance_weight (src/mcem_hcbn.cpp:371-388) vs importance.weight_ (hcbn_functions.r:590-604)"""
    rng = np.random.default_rng(seed)
    poset = synth.random_poset(p, 0.2, rng=rng)
    lam = rng.uniform(0.3, 3.0, p)
    L = 300
    Z = rng.exponential(1.0, (L, p)) / lam
    Ts = rng.exponential(1.0, L)
    y = rng.integers(0, 2, p)
    for eps in (0.01, 0.2, 0.7):
        o = oracle.forward_from_times(y, poset, eps, Z, Ts)
        r = rproto.forward_weights(y, poset, eps, Z, Ts)
        assert np.array_equal(o["X"], r["X"]) and np.array_equal(o["dist"], r["dist"])
        np.testing.assert_allclose(o["T_sum"], r["T_sum"], rtol=1e-15)
        np.testing.assert_allclose(o["w"], r["w"], rtol=1e-12)


@pytest.mark.parametrize("p,N,seed", [(6, 25, 0), (20, 60, 1)])
def test_m_step(oracle, p, N, seed):
    """E-step totals + M-step (src/mcem_hcbn.cpp:683-709) vs MCEM.hcbn_ (hcbn_functions.r:1391-1398), unit weights"""
    c = synth.make_config(p, N, eps=0.08, seed=seed, density=0.25)
    rng = np.random.default_rng(seed + 10)
    L = 200
    Z = rng.exponential(1.0, (L, p)) / c["lambdas"]
    Ts = rng.exponential(1.0, L)
    o = oracle.estep_forward_from_times(c["obs"], c["poset"], c["lambdas"], 0.08, Z, Ts)
    # per-observation expectations from the prototype's weights
    ed, eT = np.zeros(N), np.zeros((N, p))
    for i in range(N):
        r = rproto.forward_weights(c["obs"][i], c["poset"], 0.08, Z, Ts)
        ed[i] = np.sum(r["w"] * r["dist"]) / np.sum(r["w"])
        eT[i] = (r["w"][:, None] * Z).sum(axis=0) / np.sum(r["w"])
    np.testing.assert_allclose(o["dist"], ed, rtol=1e-11)
    np.testing.assert_allclose(o["Tdiff"], eT, rtol=1e-11)
    lam_new, eps_new = rproto.m_step(ed, eT)
    np.testing.assert_allclose(o["lambda_new"], lam_new, rtol=1e-11)
    assert abs(o["eps_new"] - eps_new) <= 1e-12


def test_complete_loglikelihood(oracle):
    rng = np.random.default_rng(5)
    N, p = 30, 7
    lam = rng.uniform(0.5, 2.0, p)
    Tdiff = rng.exponential(1.0, (N, p))
    dist = rng.uniform(0.0, 2.0, N)
    for eps, d in ((0.05, dist), (0.0, dist), (0.0, np.zeros(N))):
        a = oracle.complete_log_likelihood(lam, eps, Tdiff, d, np.ones(N))
        b = rproto.complete_loglikelihood(lam, eps, Tdiff, d)
        assert abs(a - b) <= 1e-10 * max(1.0, abs(b)), (eps, a, b)
