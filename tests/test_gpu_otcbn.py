"""OT-CBN callers of the conditional proposal: `drawHiddenVarsSamples`, `expectedTimeDifference`
(src/mcem.cpp:214-276) and the batched `loglike_importance_sampling` (R/mcem_genotype_probabilities.r:10-23).

T-int : support of the draws (mutated events before the sampling time, the others after it)
T-fp  : the reported log proposal density equals the oracle's density of the same draws (<= 1e-10)
T-stat: importance estimates against the exact genotype probabilities (K2 / K2') and against the oracle's mt19937
        stream, within 4 s.e. over replicate seeds
"""
import math

import numpy as np
import pytest

import exact
import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check

pytestmark = pytest.mark.gpu


def _problem(seed=5, p=6):
    cfg = synth.make_config(p, 40, eps=0.0, seed=seed, density=0.3)
    comp = exact.compatible_states(cfg["poset"])
    x = sorted(comp, key=lambda s: abs(bin(s).count("1") - p // 2))[0]
    g = np.array([(x >> j) & 1 for j in range(p)], np.int32)
    return cfg, g, x


def _t_sum(T, poset):
    """accumulated event times from waiting times (N x p) -- T_sum[e] = max parents + T[e]"""
    topo, pm = mc.flatten_poset(poset)
    S = np.zeros_like(T)
    for e in topo:
        par = [u for u in range(T.shape[1]) if (int(pm[e]) >> u) & 1]
        S[:, e] = (S[:, par].max(axis=1) if par else 0.0) + T[:, e]
    return S


@pytest.mark.parametrize("avail", [True, False])
def test_draw_hidden_vars_samples(oracle, avail):
    cfg, g, x = _problem()
    lam, poset, t = cfg["lambdas"], cfg["poset"], 0.9
    N = 5001  # ragged
    r = mc.draw_hidden_vars_samples(N, g, t, lam, poset, 1.0, avail, seed=4)
    T, dens, ts = r["T"], r["densities"], r["time"]
    assert T.shape == (N, len(lam)) and np.isfinite(T).all() and (T >= 0).all()
    if avail:
        assert np.array_equal(ts, np.full(N, t))
    else:
        assert abs(ts.mean() - 1.0) < 4 / math.sqrt(N) and (ts > 0).all()
    # T-int: the draws imply exactly the requested genotype at their sampling time
    S = _t_sum(T, poset)
    assert np.array_equal(S <= ts[:, None], np.broadcast_to(g.astype(bool), S.shape))
    # T-fp: density of these very draws as the reference computes it
    d = oracle.otcbn_proposal_density(g, ts, lam, poset, T)
    np.testing.assert_allclose(dens, d, rtol=1e-10, atol=1e-10)
    # deterministic per seed, different across seeds
    r2 = mc.draw_hidden_vars_samples(N, g, t, lam, poset, 1.0, avail, seed=4)
    assert np.array_equal(r2["T"], T) and np.array_equal(r2["densities"], dens)
    assert not np.array_equal(mc.draw_hidden_vars_samples(N, g, t, lam, poset, 1.0, avail, seed=5)["T"], T)
    # T-stat: E_q[P(Z)/q(Z)] = P(X = genotype)
    px = exact.genotype_probs_time(poset, lam, t) if avail else exact.genotype_probs_exp(poset, lam, 1.0)
    est = []
    for s in range(1, 13):
        q = mc.draw_hidden_vars_samples(4000, g, t, lam, poset, 1.0, avail, seed=s)
        logp = (np.log(lam)[None, :] - q["T"] * lam[None, :]).sum(axis=1)
        est.append(np.exp(logp - q["densities"]).mean())
    est = np.array(est)
    se = est.std(ddof=1) / math.sqrt(est.size)
    joint_check(np.array([(est.mean() - px[x]) / se]), est.size - 1, "E_q[P/q] = P(X)")  # one comparison: the plain 3 s.e. rule


@pytest.mark.parametrize("avail", [True, False])
def test_expected_time_difference_matches_oracle(oracle, avail):
    cfg, g, _ = _problem(seed=9, p=7)
    lam, poset, t = cfg["lambdas"], cfg["poset"], 1.1
    a = np.array([mc.expected_time_difference(600, g, t, lam, poset, 1.0, avail, seed=s) for s in range(1, 25)])
    b = np.array([oracle.expected_time_difference(600, g, t, lam, poset, 1.0, avail, seed=s) for s in range(101, 125)])
    assert np.isfinite(a).all() and (a > 0).all()
    se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
    z = (a.mean(axis=0) - b.mean(axis=0)) / se
    joint_check(z, 46, "expectedTimeDifference")
    # one observation of the batched E-step is the same estimator: the per-sample draws give the same answer
    q = mc.draw_hidden_vars_samples(200000, g, t, lam, poset, 1.0, avail, seed=77)
    w = np.exp((np.log(lam)[None, :] - q["T"] * lam[None, :]).sum(axis=1) - q["densities"])
    ref = (q["T"] * w[:, None]).sum(axis=0) / w.sum()
    assert np.abs(a.mean(axis=0) / ref - 1).max() < 0.05


def test_loglike_importance_sampling_batched(oracle):
    cfg, _, _ = _problem(seed=11)
    poset, lam = cfg["poset"], cfg["lambdas"]
    ok = mc.compatible_observations(cfg["obs"], poset)["compatible"].astype(bool)
    obs, times = cfg["obs"][ok], cfg["times"][ok]
    N = obs.shape[0]
    wt = np.linspace(0.5, 2.0, N)
    # sampling times given: exact answer K2' per observation
    r = mc.loglike_importance_sampling(poset, lam, obs, times, 20000, wt, 1.0, True, seed=3)
    ex = np.array([wt[i] * math.log(exact.genotype_probs_time(poset, lam, times[i])[exact.mask_of(obs[i])])
                   for i in range(N)])
    np.testing.assert_allclose(r["probs"], ex, atol=0.05)
    assert abs(r["approx_loglike"] - r["probs"].sum()) < 1e-9
    # against the reference's per-observation loop (oracle) from an independent stream, replicate seeds
    a = np.array([mc.loglike_importance_sampling(poset, lam, obs, times, 500, wt, 1.0, True, seed=s)["approx_loglike"]
                  for s in range(1, 17)])
    b = np.array([oracle.loglike_importance_sampling(poset, lam, obs[:N], times, 500, wt, 1.0, True,
                                                     seed=1000 * s)["approx_loglike"] for s in range(1, 17)])
    se = math.sqrt(a.var(ddof=1) / a.size + b.var(ddof=1) / b.size)
    assert abs(a.mean() - b.mean()) < 4 * se, (a.mean(), b.mean(), se)
    # sampling times unknown: T_s ~ Exp(lambda_s) inside the sampler, exact answer K2
    px = exact.genotype_probs_exp(poset, lam, 1.0)
    r = mc.loglike_importance_sampling(poset, lam, obs, None, 20000, None, 1.0, False, seed=5)
    ex = np.array([math.log(px[exact.mask_of(obs[i])]) for i in range(N)])
    np.testing.assert_allclose(r["probs"], ex, atol=0.06)
    # genotype_probability_fast: the single-genotype form; 0 for a genotype the poset excludes
    gp = mc.genotype_probability_fast(poset, lam, obs[0], times[0], nrOfSamples=20000, seed=2)
    assert abs(gp / exact.genotype_probs_time(poset, lam, times[0])[exact.mask_of(obs[0])] - 1) < 0.05
    u, v = [int(a_[0]) for a_ in np.nonzero(poset)]
    bad = np.zeros(len(lam), np.int32)
    bad[v] = 1
    assert mc.genotype_probability_fast(poset, lam, bad, 1.0) == 0.0


def test_otcbn_argument_errors():
    cfg, g, _ = _problem()
    cyc = cfg["poset"].copy()
    u, v = [int(a[0]) for a in np.nonzero(cyc)]
    cyc[v, u] = 1
    with pytest.raises(mc.MccbnError) as e:
        mc.draw_hidden_vars_samples(10, g, 1.0, cfg["lambdas"], cyc)
    assert e.value.code == 1 and "acyclic" in str(e.value)
    with pytest.raises(mc.MccbnError):
        mc.expected_time_difference(0, g, 1.0, cfg["lambdas"], cfg["poset"])
    assert mc.draw_hidden_vars_samples(0, g, 1.0, cfg["lambdas"], cfg["poset"])["T"].shape == (0, 6)
