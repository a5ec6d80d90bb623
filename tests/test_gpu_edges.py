"""Edge cases of the reference-facing calls on the device: smallest and ragged sizes, p = 1, single observation,
L below a warp, duplicated observations, weights, error paths."""
import numpy as np
import pytest

import exact
import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check

pytestmark = pytest.mark.gpu


def test_single_event_single_observation():
    """p = 1, N = 1: P(Y = 1) = (1 - eps) P(T <= T_s) + eps P(T > T_s), P(T <= T_s) = lam / (lam + lam_s)"""
    poset = np.zeros((1, 1), np.int32)
    lam = np.array([2.0])
    for y in (0, 1):
        obs = np.array([[y]], np.int32)
        px = lam[0] / (lam[0] + 1.0)
        want = (0.9 * px + 0.1 * (1 - px)) if y else (0.9 * (1 - px) + 0.1 * px)
        est = np.array([mc.importance_weight(obs, 4096, poset, lam, 0.1, seed=s)["w"][0] / 4096 for s in range(1, 17)])
        joint_check(np.array([(est.mean() - want) / (est.std(ddof=1) / 4 + 1e-300)]), 15, "p = 1")


@pytest.mark.parametrize("L", [1, 5, 31, 32, 33, 100])
def test_L_below_and_around_a_warp(oracle, L):
    c = synth.make_config(7, 9, eps=0.05, seed=3, density=0.3)
    r = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=2)
    assert r["w"].shape == (9,) and np.all(r["w"] > 0) and np.all(np.isfinite(r["Tdiff"]))
    # sum of weights is at most L times the largest weight, at least L times the smallest
    assert np.all(r["w"] <= L * 0.9 ** 7 * (1 + 1e-6)) and np.all(r["w"] >= L * 0.1 ** 7 * (1 - 1e-6))
    g = np.array([mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=s)["w"] / L for s in range(1, 41)])
    o = np.array([oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=s)["w"] / L
                  for s in range(101, 141)])
    se = np.sqrt(g.var(axis=0, ddof=1) / 40 + o.var(axis=0, ddof=1) / 40)
    joint_check((g.mean(axis=0) - o.mean(axis=0)) / np.maximum(se, 1e-300), 78, f"L = {L}")


def test_ragged_N_and_duplicates_and_weights(oracle):
    c = synth.make_config(6, 37, eps=0.05, seed=9, density=0.3)   # 37: not a multiple of any tile size
    obs = np.vstack([c["obs"], c["obs"][:5]])                      # duplicated rows
    w = np.concatenate([np.ones(37), np.zeros(5)])                 # zero-weight duplicates must not change the fit
    a = mc.MCEM_hcbn(c["lambda0"], c["poset"], obs, weights=w, L=128, eps=0.1, max_iter=20, update_step_size=10, seed=3)
    b = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=128, eps=0.1, max_iter=20, update_step_size=10, seed=3)
    np.testing.assert_allclose(a["lambda_"], b["lambda_"], rtol=1e-9)
    assert abs(a["eps"] - b["eps"]) < 1e-12 and abs(a["llhood"] - b["llhood"]) < 1e-9 * abs(b["llhood"])
    # doubling every weight doubles the log-likelihood and leaves the parameters alone
    d = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], weights=2 * np.ones(37), L=128, eps=0.1, max_iter=20,
                     update_step_size=10, seed=3)
    np.testing.assert_allclose(d["lambda_"], b["lambda_"], rtol=1e-9)
    assert abs(d["llhood"] - 2 * b["llhood"]) < 1e-9 * abs(b["llhood"])


def test_argument_errors():
    c = synth.make_config(5, 10, eps=0.05, seed=1, density=0.3)
    with pytest.raises(ValueError):  # match.arg of the R wrapper
        mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.1, L=64, sampling="no-such-scheme", seed=1)
    with pytest.raises(ValueError):
        mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.1, L=64, times=np.ones(3), seed=1)
    with pytest.raises(mc.MccbnError):  # backward: L smaller than the Hamming ball -> no sample at all
        mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.1, L=3, sampling="backward", seed=1)


def test_all_zero_and_all_one_observations():
    c = synth.make_config(8, 4, eps=0.05, seed=2, density=0.3)
    obs = np.array([[0] * 8, [1] * 8], np.int32)
    px = exact.genotype_probs_exp(c["poset"], c["lambdas"], 1.0)
    want = np.array([exact.obs_prob_and_dist(px, exact.mask_of(o), 0.1, 8)[0] for o in obs])
    est = np.array([mc.importance_weight(obs, 8192, c["poset"], c["lambdas"], 0.1, seed=s)["w"] / 8192 for s in range(1, 25)])
    z = (est.mean(axis=0) - want) / (est.std(axis=0, ddof=1) / np.sqrt(24))
    joint_check(z, 23, "all-zero / all-one observations")
