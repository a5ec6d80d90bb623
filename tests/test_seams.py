"""The reference's small export seams against the oracle.  Host-arithmetic seams run without a GPU
(`_scale_path_to_mutation`, `_cbn_density_log`, `_complete_log_likelihood`, `_draw_sample`); the sampling seams
(`_coin_tossing`, `_generate_mutation_times`) need the device."""
import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import synth


def test_scale_path_density_and_complete_loglik(oracle):
    rng = np.random.default_rng(1)
    for p in (1, 6, 17, 32):
        poset = synth.random_poset(p, 0.25, rng=rng)
        lam = rng.uniform(0.3, 3.0, p)
        np.testing.assert_allclose(mc.scale_path_to_mutation(poset, lam), oracle.scale_path_to_mutation(poset, lam),
                                   rtol=1e-15)
        N = 37
        T = rng.exponential(1.0, (N, p)) / lam
        np.testing.assert_allclose(mc.cbn_density_log(T, lam), oracle.cbn_density_log(T, lam), rtol=1e-13)
        d = rng.integers(0, p + 1, N).astype(float)
        w = rng.uniform(0.5, 2.0, N)
        for eps in (0.07, 0.0):
            a = mc.complete_log_likelihood(lam, eps, T, d, w)
            b = oracle.complete_log_likelihood(lam, eps, T, d, w)
            assert abs(a - b) <= 1e-12 * abs(b), (p, eps, a, b)


def test_draw_sample_bit_exact(oracle):
    """same generator, same seeding chain, same std::discrete_distribution as the reference: identical draws"""
    rng = np.random.default_rng(2)
    for p in (5, 12, 20):
        poset = synth.random_poset(p, 0.3, rng=rng)
        q = rng.uniform(0.2, 4.0, p)
        comp = synth.sample_genotypes(6, poset, 1.0, rng.uniform(0.5, 2, p), 0.0, rng)["hidden_genotypes"]
        gens = [g for g in comp if 0 < g.sum() < p] + [rng.integers(0, 2, p) for _ in range(6)]
        for g in gens:
            if g.sum() in (0, p):
                continue
            for move in (0, 1, 2):
                for seed in (1, 7, 123):
                    a, qa = mc.draw_sample(g, poset, move, q, seed=seed)
                    b, qb = oracle.draw_sample(g, poset, move, q, seed=seed)
                    assert np.array_equal(a, b) and qa == qb, (p, move, seed)


@pytest.mark.gpu
def test_generate_mutation_times_reference_stream(oracle):
    """uniforms from the reference's generator in the reference's order, fp64 arithmetic on the device: <= 1e-10"""
    rng = np.random.default_rng(3)
    for p, times in ((7, False), (16, True), (25, False)):
        c = synth.make_config(p, 5, seed=p, density=0.25)
        N = 400
        X = synth.sample_genotypes(N, c["poset"], 1.0, c["lambdas"], 0.0, rng)["hidden_genotypes"].astype(np.int32)
        t = rng.uniform(0.3, 3.0, N) if times else None
        a = mc.generate_mutation_times(X, c["poset"], c["lambdas"], times=t, lambda_s=1.5, seed=11)
        b = oracle.generate_mutation_times(X, c["poset"], c["lambdas"], times=t, lambda_s=1.5, seed=11)
        np.testing.assert_allclose(a["sampling_time"], b["sampling_time"], rtol=1e-15)
        np.testing.assert_allclose(a["Tdiff"], b["Tdiff"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(a["proposal_density"], b["proposal_density"], rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_coin_tossing_frequencies(oracle):
    g = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 0], np.int32)
    L = 50000
    for eps in (0.0, 0.05, 0.3, 1.0):
        x = mc.coin_tossing(eps, L, g, seed=4)
        assert x.shape == (L, g.size) and set(np.unique(x)) <= {0, 1}
        flips = (x != g[None, :]).mean(axis=0)
        se = max(np.sqrt(eps * (1 - eps) / L), 1e-9)
        assert np.abs(flips - eps).max() < 4.5 * se
        o = oracle.coin_tossing(eps, L, g, seed=4)
        assert abs(flips.mean() - (o != g[None, :]).mean()) < 4.5 * np.sqrt(2) * se / np.sqrt(g.size) + 1e-12
    # independent across columns
    x = mc.coin_tossing(0.3, L, g, seed=5) != g[None, :]
    cc = np.corrcoef(x.T.astype(float))
    assert np.abs(cc - np.eye(g.size)).max() < 5 / np.sqrt(L)
