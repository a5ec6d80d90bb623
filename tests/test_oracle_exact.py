"""Pins the CPU oracle (oracle/mccbn_oracle.cpp) against model-derived exact answers and
the reference's own validation method (IS estimate vs exact/brute-force probabilities, utils/hcbn_test.r T1).
CPU only.  3-sigma bands use replicate seeds for the standard error."""
import math

import numpy as np
import pytest

import exact
from mccbn_b200 import synth


def _small_problem(p=6, N=40, eps=0.08, seed=3):
    cfg = synth.make_config(p, N, eps=eps, seed=seed, density=0.3)
    return cfg


def test_rng_chain_is_exponential(oracle):
    # StdRng seeding chain + rtexp(cutoff=inf): mean/var of Exp(rate), and reproducible per seed
    a = oracle.rexp(200000, 2.5, 42)
    b = oracle.rexp(200000, 2.5, 42)
    assert np.array_equal(a, b)
    assert abs(a.mean() - 0.4) < 4 * 0.4 / math.sqrt(a.size)
    assert abs(a.var() - 0.16) < 0.01
    assert not np.array_equal(a[:10], oracle.rexp(10, 2.5, 43))


def test_flatten_and_cycles(oracle):
    poset = np.zeros((4, 4), np.int32)
    poset[0, 1] = poset[1, 2] = poset[0, 3] = 1
    topo, mask, cyc = oracle.flatten_poset(poset)
    assert not cyc
    pos = {int(e): k for k, e in enumerate(topo)}
    assert pos[0] < pos[1] < pos[2] and pos[0] < pos[3]
    assert list(mask) == [0, 1, 2, 1]
    poset[2, 0] = 1
    assert oracle.flatten_poset(poset)[2]
    with pytest.raises(oracle.OracleError) as e:
        oracle.num_compatible_observations(np.zeros((2, 4), np.int32), poset)
    assert e.value.code == oracle.ERR_NOT_ACYCLIC
    assert "acyclic" in str(e.value)


def test_compatibility_counts_bruteforce(oracle):
    rng = np.random.default_rng(0)
    for p in (3, 7, 12):
        poset = synth.random_poset(p, 0.3, rng=rng)
        obs = rng.integers(0, 2, size=(200, p)).astype(np.int32)
        comp = np.array([all(obs[i, u] or not obs[i, v] for u, v in zip(*np.nonzero(poset))) for i in range(200)])
        assert np.array_equal(oracle.compatible_observations(obs, poset), comp.astype(np.int32))
        assert oracle.num_compatible_observations(obs, poset) == comp.sum()
        inc = sum(int(((obs[:, u] == 0) & (obs[:, v] == 1)).sum()) for u, v in zip(*np.nonzero(poset)))
        assert oracle.num_incompatible_events(obs, poset) == inc


def test_k4_neighbors(oracle):
    # K4: ring k has C(p,k) rows, each at Hamming distance exactly k, all distinct; row 0 is Y
    rng = np.random.default_rng(1)
    for p, k in ((5, 1), (6, 2), (7, 3), (4, 4)):
        y = rng.integers(0, 2, size=p)
        rows, ring = oracle.neighbors(p, k, y)
        assert rows.shape[0] == sum(math.comb(p, i) for i in range(k + 1))
        assert np.array_equal(rows[0], y)
        d = (rows != y[None, :]).sum(axis=1)
        assert np.array_equal(d, ring)
        assert len({tuple(r) for r in rows}) == rows.shape[0]
        for i in range(k + 1):
            assert (ring == i).sum() == math.comb(p, i)
    # documented order for k = 1: row r (r >= 1) flips column p - r
    rows, _ = oracle.neighbors(5, 1, np.zeros(5, np.int32))
    assert [int(np.nonzero(r)[0][0]) for r in rows[1:]] == [4, 3, 2, 1, 0]


def test_forward_from_times_matches_numpy(oracle):
    cfg = _small_problem()
    rng = np.random.default_rng(5)
    p, L = cfg["poset"].shape[0], 300
    Z = rng.exponential(1.0, (L, p)) / cfg["lambdas"]
    Ts = rng.exponential(1.0, L)
    y = cfg["obs"][0]
    r = oracle.forward_from_times(y, cfg["poset"], 0.07, Z, Ts)
    T = np.zeros((L, p))
    for e in synth.topological_sort(cfg["poset"]):
        par = np.nonzero(cfg["poset"][:, e])[0]
        T[:, e] = Z[:, e] + (T[:, par].max(axis=1) if par.size else 0.0)
    X = (T <= Ts[:, None]).astype(np.int32)
    assert np.array_equal(r["T_sum"], T)
    assert np.array_equal(r["X"], X)
    d = (X != y[None, :]).sum(axis=1)
    assert np.array_equal(r["dist"], d)
    np.testing.assert_allclose(r["w"], 0.07 ** d * 0.93 ** (p - d), rtol=1e-14)


def test_k3_density_identity(oracle):
    # K3: log P(Z) - log q = sum_{X_j=1} log F_j(c_j) - sum_{X_j=0} lambda_j max(0, T_s - m_j)  (= log P(X | T_s, path))
    cfg = _small_problem(p=7)
    rng = np.random.default_rng(6)
    p, L = 7, 200
    sim = synth.sample_genotypes(L, cfg["poset"], 1.0, cfg["lambdas"], eps=0.0, rng=rng)
    X = sim["hidden_genotypes"]  # compatible by construction
    u = rng.random((L, p))
    Ts = rng.exponential(1.0, L) + 0.05
    r = oracle.conditional_from_uniforms(X, cfg["poset"], cfg["lambdas"], u, Ts)
    assert not r["incompatible"].any()
    Z = r["Tdiff"]
    assert (Z >= 0).all()
    T = np.zeros((L, p))
    logw = np.zeros(L)
    for e in synth.topological_sort(cfg["poset"]):
        par = np.nonzero(cfg["poset"][:, e])[0]
        m = T[:, par].max(axis=1) if par.size else np.zeros(L)
        T[:, e] = m + Z[:, e]
        c = Ts - m
        lam_e = cfg["lambdas"][e]
        # X_j = 1: log F(c);  X_j = 0: log S(max(0, T_s - m)) = -lambda max(0, c)
        logw += np.where(X[:, e] == 1, np.log1p(-np.exp(-lam_e * np.where(X[:, e] == 1, c, 1.0))),
                         -lam_e * np.maximum(c, 0.0))
    np.testing.assert_allclose(r["log_pz"] - r["log_q"], logw, rtol=1e-9, atol=1e-11)
    # implied genotype of the generated times equals X
    Tsum = np.zeros((L, p))
    for e in synth.topological_sort(cfg["poset"]):
        par = np.nonzero(cfg["poset"][:, e])[0]
        m = Tsum[:, par].max(axis=1) if par.size else np.zeros(L)
        Tsum[:, e] = m + Z[:, e]
    assert np.array_equal((Tsum <= Ts[:, None]).astype(np.int32), X)
    # an incompatible genotype is flagged (child without its parent)
    u_e, v_e = [int(a[0]) for a in np.nonzero(cfg["poset"])]
    bad = np.zeros((1, p), np.int32)
    bad[0, v_e] = 1
    rb = oracle.conditional_from_uniforms(bad, cfg["poset"], cfg["lambdas"], u[:1], Ts[:1])
    assert rb["incompatible"][0] == 1


def _replicate(fn, seeds):
    v = np.array([fn(s) for s in seeds])
    return v.mean(axis=0), v.std(axis=0, ddof=1) / math.sqrt(len(seeds))


@pytest.mark.parametrize("sampling,kw", [("forward", {}), ("backward", dict(neighborhood_dist=6)),
                                         ("pool", {}), ("bernoulli", {}), ("add-remove", {})])
def test_k2_obs_loglik_all_schemes(oracle, sampling, kw):
    """Every proposal is an unbiased estimator of P(Y_i) (Appendix C); compare the IS estimate of P(Y) with K2.
    (backward uses the full Hamming ball k = p so that it is unbiased; bernoulli/add-remove have restricted support
    and are only checked to be a lower bound on average.)"""
    cfg = _small_problem(p=6, N=12, eps=0.1, seed=11)
    p = 6
    lam, poset, obs = cfg["lambdas"], cfg["poset"], cfg["obs"]
    px = exact.genotype_probs_exp(poset, lam, 1.0)
    assert abs(sum(px.values()) - 1.0) < 1e-12
    exact_py = np.array([exact.obs_prob_and_dist(px, exact.mask_of(obs[i]), 0.1, p)[0] for i in range(obs.shape[0])])
    L = 2000 if sampling != "pool" else 300

    def est(seed):
        r = oracle.importance_weight(obs, L, poset, lam, 0.1, sampling=sampling, seed=seed, **kw)
        Leff = r["L_eff"] if sampling == "backward" else (L if sampling != "pool" else L)
        return r["w"] / Leff

    mean, se = _replicate(est, range(100, 124))
    z = (mean - exact_py) / np.maximum(se, 1e-300)
    if sampling in ("forward", "backward", "pool"):
        assert np.abs(z).max() < 4.0, (sampling, z)
    else:
        # restricted-support proposals: estimate of P(Y) is biased low, never high
        assert (z < 4.0).all(), (sampling, z)
        assert (mean > 0).all()


def test_k2_expected_distance_forward(oracle):
    cfg = _small_problem(p=6, N=10, eps=0.1, seed=12)
    lam, poset, obs = cfg["lambdas"], cfg["poset"], cfg["obs"]
    _, ed = exact.exact_obs_loglik(obs, poset, lam, 0.1)

    def est(seed):
        return oracle.importance_weight(obs, 4000, poset, lam, 0.1, sampling="forward", seed=seed)["dist"]

    mean, se = _replicate(est, range(200, 224))
    assert np.abs((mean - ed) / se).max() < 4.0


def test_k1_empty_poset_given_times(oracle):
    p, N = 5, 8
    rng = np.random.default_rng(8)
    lam = rng.uniform(0.5, 2.0, p)
    poset = np.zeros((p, p), np.int32)
    obs = rng.integers(0, 2, (N, p)).astype(np.int32)
    times = rng.uniform(0.3, 2.0, N)
    ex = np.array([exact.k1_empty_poset(obs[i], lam, 0.05, times[i]) for i in range(N)])

    def est(seed):
        r = oracle.importance_weight(obs, 3000, poset, lam, 0.05, times=times, sampling="forward", seed=seed)
        # dist * sum w is an unbiased estimate of L P(Y) E[d | Y]; `dist` alone is a ratio estimator (O(1/L) bias)
        return np.concatenate([r["w"] / 3000, r["dist"] * r["w"] / 3000])

    mean, se = _replicate(est, range(300, 324))
    target = np.concatenate([ex[:, 0], ex[:, 0] * ex[:, 1]])
    assert np.abs((mean - target) / se).max() < 4.0
    # K2' agrees with K1 on the empty poset
    for i in range(N):
        px = exact.genotype_probs_time(poset, lam, times[i])
        py, d = exact.obs_prob_and_dist(px, exact.mask_of(obs[i]), 0.05, p)
        assert abs(py - ex[i, 0]) < 1e-12 and abs(d - ex[i, 1]) < 1e-10


def test_estep_from_times_consistency(oracle):
    """ref_estep_forward_from_times == importance_weight arithmetic done by hand in numpy (fp64, 1e-12)."""
    cfg = _small_problem(p=6, N=15, eps=0.1, seed=13)
    rng = np.random.default_rng(14)
    p, L = 6, 500
    lam, poset, obs = cfg["lambdas"], cfg["poset"], cfg["obs"]
    Z = rng.exponential(1.0, (L, p)) / lam
    Ts = rng.exponential(1.0, L)
    r = oracle.estep_forward_from_times(obs, poset, lam, 0.1, Z, Ts)
    T = np.zeros((L, p))
    for e in synth.topological_sort(poset):
        par = np.nonzero(poset[:, e])[0]
        T[:, e] = Z[:, e] + (T[:, par].max(axis=1) if par.size else 0.0)
    X = (T <= Ts[:, None])
    ll = 0.0
    S = np.zeros(p)
    D = 0.0
    for i in range(obs.shape[0]):
        d = (X != obs[i].astype(bool)[None, :]).sum(axis=1)
        w = 0.1 ** d * 0.9 ** (p - d)
        ll += math.log(w.sum() / L)
        D += (w * d).sum() / w.sum()
        S += (Z * w[:, None]).sum(axis=0) / w.sum()
    assert abs(r["obs_llhood"] - ll) < 1e-10 * abs(ll)
    np.testing.assert_allclose(r["Tdiff_colsum"], S, rtol=1e-12)
    np.testing.assert_allclose(r["lambda_new"], obs.shape[0] / S, rtol=1e-12)
    assert abs(r["eps_new"] - D / (obs.shape[0] * p)) < 1e-14


def test_mcem_recovers_parameters(oracle):
    """Reference validation method (utils/hcbn_simulation_test.r:81): recovery error on simulated data."""
    cfg = synth.make_config(5, 400, eps=0.05, seed=21, density=0.3)
    r = oracle.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], L=200, eps=0.1, max_iter=60, update_step_size=20,
                         thrds=oracle.num_threads(), seed=5, return_trace=True)
    rel = np.abs(r["lambda_"] - cfg["lambdas"]).mean() / cfg["lambdas"].mean()
    assert rel < 0.35, rel
    assert abs(r["eps"] - 0.05) < 0.03
    assert r["trace"].shape[1] == 7 and 20 <= r["n_iter"] <= 60
    # batch-average rule: returned values are means of the last batch of iterates
    n = r["n_iter"]
    last = r["trace"][n - (n % 20 or 20):n] if n == 60 else r["trace"][n - 20:n]
    np.testing.assert_allclose(r["lambda_"], last[:, 2:].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(r["llhood"], last[:, 0].mean(), rtol=1e-12)


def test_zero_weight_error(oracle):
    cfg = _small_problem(p=6, N=5)
    with pytest.raises(oracle.OracleError) as e:
        # backward with L < 1 + p  =>  reps = 0 => no samples => "all samples have weight 0"
        oracle.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], L=3, eps=0.1, sampling="backward", max_iter=2,
                         update_step_size=1)
    assert e.value.code == oracle.ERR_ZERO_WEIGHT


def test_initialize_lambda_counts(oracle):
    cfg = synth.make_config(6, 300, eps=0.05, seed=31, density=0.3)
    lam, eps0 = oracle.initialize_lambda(cfg["obs"], cfg["poset"])
    obs, poset = cfg["obs"], cfg["poset"]
    clo = synth.transitive_closure(poset)
    for j in range(6):
        anc = np.nonzero(clo[:, j])[0]
        ok = obs[:, anc].all(axis=1) if anc.size else np.ones(obs.shape[0], bool)
        count = np.float32(1 + ok.sum())
        mutated = np.float32(np.float32(1e-7) + np.float32((ok & (obs[:, j] == 1)).sum()))
        aux = np.float32(mutated / count)
        expect = float(aux) * 1.0 / (1.0 - float(aux))
        assert lam[j] == pytest.approx(expect, rel=1e-12)
    assert eps0 == pytest.approx(oracle.num_incompatible_events(obs, poset) / (300 * 6))


# ---- OT-CBN callers of the conditional proposal (src/mcem.cpp:214-276, R/mcem_genotype_probabilities.r) --------------
def _otcbn_problem(seed=5, p=6):
    cfg = synth.make_config(p, 30, eps=0.0, seed=seed, density=0.3)
    comp = exact.compatible_states(cfg["poset"])
    # a compatible genotype with about half of the events
    x = sorted(comp, key=lambda s: abs(bin(s).count("1") - p // 2))[0]
    g = np.array([(x >> j) & 1 for j in range(p)], np.int32)
    return cfg, g, x


@pytest.mark.parametrize("avail", [True, False])
def test_otcbn_draw_hidden_vars_estimates_genotype_probability(oracle, avail):
    """E_q[P(Z)/q(Z)] = P(X = genotype): the identity genotype_probability_fast relies on (K2 / K2')."""
    cfg, g, x = _otcbn_problem()
    lam, poset, t = cfg["lambdas"], cfg["poset"], 0.9
    px = exact.genotype_probs_time(poset, lam, t) if avail else exact.genotype_probs_exp(poset, lam, 1.0)
    est = []
    for s in range(1, 13):
        r = oracle.draw_hidden_vars_samples(4000, g, t, lam, poset, 1.0, avail, seed=s)
        assert r["T"].shape == (4000, len(lam)) and (r["T"] >= 0).all()
        logp = (np.log(lam)[None, :] - r["T"] * lam[None, :]).sum(axis=1)
        est.append(np.exp(logp - r["densities"]).mean())
    est = np.array(est)
    se = est.std(ddof=1) / math.sqrt(est.size)
    assert abs(est.mean() - px[x]) < 4 * se, (est.mean(), px[x], se)


def test_otcbn_proposal_density_from_output(oracle):
    """the density recomputed from the sampler's OUTPUT equals the density it reported (time given)"""
    cfg, g, _ = _otcbn_problem(seed=8, p=7)
    r = oracle.draw_hidden_vars_samples(500, g, 1.3, cfg["lambdas"], cfg["poset"], 1.0, True, seed=2)
    d = oracle.otcbn_proposal_density(g, np.full(500, 1.3), cfg["lambdas"], cfg["poset"], r["T"])
    np.testing.assert_allclose(d, r["densities"], rtol=1e-10, atol=1e-10)


def test_otcbn_loglike_importance_sampling_exact(oracle):
    """loglike_importance_sampling (R/mcem_genotype_probabilities.r:10-23) against the exact log-probabilities"""
    cfg, _, _ = _otcbn_problem(seed=11)
    poset, lam = cfg["poset"], cfg["lambdas"]
    ok = oracle.compatible_observations(cfg["obs"], poset).astype(bool)
    obs, times = cfg["obs"][ok][:8], cfg["times"][ok][:8]
    wt = np.linspace(0.5, 2.0, obs.shape[0])
    r = oracle.loglike_importance_sampling(poset, lam, obs, times, 20000, wt, 1.0, True, seed=3)
    ex = np.array([wt[i] * math.log(exact.genotype_probs_time(poset, lam, times[i])[exact.mask_of(obs[i])])
                   for i in range(obs.shape[0])])
    np.testing.assert_allclose(r["probs"], ex, atol=0.05)
    assert abs(r["approx_loglike"] - r["probs"].sum()) < 1e-12


def test_effective_sample_size_forward(oracle):
    """The reference's own diagnostic (utils/hcbn_test.r T4-T6): ESS = (sum w)^2 / sum w^2 per genotype, against its
    exact large-L limit L E[w]^2 / E[w^2] from the genotype distribution (K2)."""
    cfg = _small_problem(p=6, N=12, eps=0.1, seed=4)
    L = 60000
    r = oracle.importance_weight(cfg["obs"], L, cfg["poset"], cfg["lambdas"], 0.1, thrds=oracle.num_threads(), seed=5)
    px = exact.genotype_probs_exp(cfg["poset"], cfg["lambdas"], 1.0)
    ess = r["w"] ** 2 / r["w_sqrt"]
    for i in range(cfg["obs"].shape[0]):
        m1, m2 = exact.forward_weight_moments(px, exact.mask_of(cfg["obs"][i]), 0.1, 6)
        assert abs(ess[i] / (L * m1 * m1 / m2) - 1) < 0.05, (i, ess[i], L * m1 * m1 / m2)
        assert 1.0 <= ess[i] <= L


def _t1_targets(bf, k=6):
    """the k most frequent observed genotypes of a brute-force table"""
    return sorted(bf, key=lambda y: -bf[y]["n"])[:k]


def test_t1_brute_force_posterior_means(oracle):
    """Reference validation T1: importance-sampling estimates of P(Y), E[Z_j | Y] and E[d | Y] against brute-force
    frequencies / conditional means over simulated noisy genotypes (an estimator-independent answer for E[Z | Y])."""
    cfg = _small_problem(p=5, N=4, eps=0.1, seed=6)
    poset, lam, eps, p = cfg["poset"], cfg["lambdas"], 0.1, 5
    bf = exact.brute_force_posterior(poset, lam, eps, 1500000, seed=1)
    ys = _t1_targets(bf)
    obs = np.array([[(y >> j) & 1 for j in range(p)] for y in ys], np.int32)
    L = 40000
    reps = [oracle.importance_weight(obs, L, poset, lam, eps, thrds=oracle.num_threads(), seed=s) for s in range(1, 9)]
    for i, y in enumerate(ys):
        b = bf[y]
        Zs = np.array([r["Tdiff"][i] for r in reps])
        ds = np.array([r["dist"][i] for r in reps])
        ps = np.array([r["w"][i] / L for r in reps])
        se_Z = np.sqrt(Zs.var(axis=0, ddof=1) / len(reps) + b["Z_se"] ** 2)
        assert np.abs((Zs.mean(axis=0) - b["Z"]) / se_Z).max() < 4.5, (y, Zs.mean(axis=0), b["Z"])
        assert abs(ds.mean() - b["d"]) < 4.5 * math.sqrt(ds.var(ddof=1) / len(reps) + b["d_se"] ** 2)
        se_p = math.sqrt(ps.var(ddof=1) / len(reps) + b["prob"] * (1 - b["prob"]) / 1500000)
        assert abs(ps.mean() - b["prob"]) < 4.5 * se_p
