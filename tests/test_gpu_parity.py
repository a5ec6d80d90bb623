"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tiers:
  T-int  bit-exact     : flattening, compatibility, counts, hidden genotypes X, Hamming distances, ball enumeration
  T-fp   <= 1e-10 rel. : weights, sum w, E[d], E[Z], log q, obs_llhood, M-step outputs from identical supplied draws
  T-stat 3 s.e.        : independent RNG streams (Philox vs mt19937): P(Y), E[d], E[Z], converged lambda/eps/llhood,
                         plus the model-derived exact answers K1/K2
"""
import math

import numpy as np
import pytest

import exact
import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check

pytestmark = pytest.mark.gpu

RTOL_FP = 1e-10  # fp64 tier tolerance stated by BASELINE.json north_star


def _cfg(p=8, N=60, eps=0.08, seed=3, density=0.3):
    return synth.make_config(p, N, eps=eps, seed=seed, density=density)


# ---------------------------------------------------------------- T-int -------------------------------------------
@pytest.mark.parametrize("p", [1, 5, 16, 30, 32])
def test_forward_from_times_bit_exact(oracle, p):
    rng = np.random.default_rng(p)
    poset = synth.random_poset(p, 0.25, rng=rng)
    lam = rng.uniform(0.3, 3.0, p)
    L = 777  # ragged: not a multiple of any block size
    Z = rng.exponential(1.0, (L, p)) / lam
    Ts = rng.exponential(1.0, L)
    y = rng.integers(0, 2, p)
    a = mc.forward_from_times(y, poset, 0.07, Z, Ts)
    b = oracle.forward_from_times(y, poset, 0.07, Z, Ts)
    assert np.array_equal(a["T_sum"], b["T_sum"])  # pure fp64 add/max: bit-exact
    assert np.array_equal(a["X"], b["X"])
    assert np.array_equal(a["dist"], b["dist"])
    assert np.array_equal(a["w"], b["w"])  # host-built pow table == reference pow


def test_compatibility_and_counts_bit_exact(oracle):
    rng = np.random.default_rng(0)
    for p, N in ((3, 1), (12, 1000), (30, 5000), (40, 300), (64, 257)):
        poset = synth.random_poset(p, 0.2, rng=rng)
        obs = rng.integers(0, 2, (N, p)).astype(np.int32)
        obs[: N // 3] = synth.sample_genotypes(max(N // 3, 1), poset, 1.0, rng.uniform(0.5, 2, p), 0.0, rng)[
            "obs_events"][: N // 3]
        r = mc.compatible_observations(obs, poset)
        assert np.array_equal(r["compatible"], oracle.compatible_observations(obs, poset))
        assert r["num_compatible"] == oracle.num_compatible_observations(obs, poset)
        assert r["num_incompatible_events"] == oracle.num_incompatible_events(obs, poset)


def test_initialize_lambda_matches_oracle(oracle):
    for seed in range(3):
        c = synth.make_config(12, 500, eps=0.05, seed=seed, density=0.25)
        lam, eps = mc.initialize_lambda(c["obs"], c["poset"], lambda_s=1.0)
        lam_o, eps_o = oracle.initialize_lambda(c["obs"], c["poset"], lambda_s=1.0)
        assert np.array_equal(lam, lam_o)  # integer counts + the same float division
        assert eps == eps_o


def test_cyclic_poset_error():
    poset = np.zeros((4, 4), np.int32)
    poset[0, 1] = poset[1, 2] = poset[2, 0] = 1
    with pytest.raises(mc.MccbnError) as e:
        mc.obs_loglikelihood(np.zeros((3, 4), np.int32), poset, np.ones(4), 0.1, L=32, seed=1)
    assert e.value.code == 1 and "acyclic" in str(e.value)


# ---------------------------------------------------------------- T-fp --------------------------------------------
@pytest.mark.parametrize("p,N,L,times", [(6, 15, 500, False), (16, 200, 1000, False), (30, 64, 333, True),
                                         (32, 7, 64, False)])
def test_estep_from_times_fp64(oracle, p, N, L, times):
    c = _cfg(p=p, N=N, seed=p + N)
    rng = np.random.default_rng(L)
    Z = rng.exponential(1.0, (L, p)) / c["lambdas"]
    Ts = rng.exponential(1.0, L)
    t = c["times"] if times else None
    wts = rng.uniform(0.5, 2.0, N)
    a = mc.estep_forward_from_times(c["obs"], c["poset"], c["lambdas"], 0.1, Z, Ts, weights=wts, times=t)
    b = oracle.estep_forward_from_times(c["obs"], c["poset"], c["lambdas"], 0.1, Z, Ts, weights=wts, times=t)
    for k in ("w", "w_sqrt", "dist", "Tdiff", "Tdiff_colsum", "lambda_new"):
        np.testing.assert_allclose(a[k], b[k], rtol=RTOL_FP, atol=1e-300, err_msg=k)
    for k in ("obs_llhood", "expected_dist", "N_eff", "eps_new"):
        assert abs(a[k] - b[k]) <= RTOL_FP * abs(b[k]), (k, a[k], b[k])


def test_conditional_from_uniforms_fp64(oracle):
    for p in (7, 25):
        c = _cfg(p=p, N=10, seed=40 + p)
        rng = np.random.default_rng(p)
        L = 600
        X = synth.sample_genotypes(L, c["poset"], 1.0, c["lambdas"], 0.0, rng)["hidden_genotypes"].copy()
        X[L // 2:] = rng.integers(0, 2, (L - L // 2, p))  # second half: arbitrary (mostly incompatible) genotypes
        u = rng.random((L, p))
        Ts = rng.exponential(1.0, L) + 0.01
        a = mc.conditional_from_uniforms(X, c["poset"], c["lambdas"], u, Ts)
        b = oracle.conditional_from_uniforms(X, c["poset"], c["lambdas"], u, Ts)
        assert np.array_equal(a["incompatible"], b["incompatible"])  # bit-exact flag
        ok = b["incompatible"] == 0
        assert ok.sum() > L // 3
        np.testing.assert_allclose(a["Tdiff"][ok], b["Tdiff"][ok], rtol=RTOL_FP, atol=1e-14)
        np.testing.assert_allclose(a["log_q"][ok], b["log_q"][ok], rtol=RTOL_FP, atol=1e-12)
        np.testing.assert_allclose(a["log_pz"][ok], b["log_pz"][ok], rtol=RTOL_FP, atol=1e-12)


# ----------------------------------------------------- production kernel vs its own per-sample dump ---------------
@pytest.mark.parametrize("p,times", [(5, False), (13, True), (20, False), (30, False), (32, True)])
def test_production_reduction_matches_dump(p, times):
    """The Philox E-step kernel and the per-sample dump kernel share the stream (seed, iteration 0, observation i,
    sample l): reducing the dump in numpy (fp64) must reproduce the kernel's per-observation statistics up to the
    fp32 accumulation error of the fast path."""
    c = _cfg(p=p, N=9, seed=p, density=0.2)
    L = 700
    eps = 0.07
    t = c["times"] if times else None
    r = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], eps, time=t, seed=11)
    for i in (0, 4, 8):
        d = mc.importance_weight(c["obs"][i], L, c["poset"], c["lambdas"], eps, time=None if t is None else t[i],
                                 seed=11, obs_offset=i)
        w = d["w"]
        assert d["Tdiff"].shape == (L, p) and (d["Tdiff"] >= 0).all()
        assert abs(r["w"][i] - w.sum()) <= 2e-5 * w.sum()
        assert abs(r["w_sqrt"][i] - (w * w).sum()) <= 1e-4 * (w * w).sum()
        assert abs(r["dist"][i] - (w * d["dist"]).sum() / w.sum()) <= 1e-4 * max(1.0, r["dist"][i])
        np.testing.assert_allclose(r["Tdiff"][i], (d["Tdiff"] * w[:, None]).sum(axis=0) / w.sum(), rtol=2e-4)


def test_dump_marginals_are_exponential():
    # raw waiting times of the Philox path: Z_j ~ Exp(lambda_j) (mean and variance), independent of the poset
    c = _cfg(p=12, N=1, seed=5)
    L = 200000
    d = mc.importance_weight(c["obs"][0], L, c["poset"], c["lambdas"], 0.05, seed=99)
    m = d["Tdiff"].mean(axis=0) * c["lambdas"]
    v = d["Tdiff"].var(axis=0) * c["lambdas"] ** 2
    assert np.abs(m - 1).max() < 5 / math.sqrt(L)
    assert np.abs(v - 1).max() < 0.03
    # different observation index / seed => different stream
    d2 = mc.importance_weight(c["obs"][0], 1000, c["poset"], c["lambdas"], 0.05, seed=99, obs_offset=1)
    assert not np.array_equal(d2["Tdiff"], d["Tdiff"][:1000])


def test_shard_invariance():
    """Philox counters use global observation indices and the record partition depends on (L, GLOBAL N) only: scoring a
    shard with obs_offset / N_global reproduces the rows of the full run BIT FOR BIT -- per-observation outputs do not
    depend on how the observations are sharded (one record per observation here, and also with several records)."""
    c = _cfg(p=10, N=300, seed=8)
    for L in (640, 5000):  # 1 record per observation; 4 records per observation (global N = 300 is small)
        full = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.05, seed=5)
        part = mc.importance_weight(c["obs"][100:220], L, c["poset"], c["lambdas"], 0.05, seed=5, obs_offset=100,
                                    N_global=300)
        assert np.array_equal(part["w"], full["w"][100:220])
        assert np.array_equal(part["Tdiff"], full["Tdiff"][100:220])
        assert np.array_equal(part["dist"], full["dist"][100:220])
    # without N_global the shard chooses its own partition: same samples, fp32 summation order may differ
    part = mc.importance_weight(c["obs"][100:220], 5000, c["poset"], c["lambdas"], 0.05, seed=5, obs_offset=100)
    np.testing.assert_allclose(part["w"], full["w"][100:220], rtol=1e-5)
    np.testing.assert_allclose(part["Tdiff"], full["Tdiff"][100:220], rtol=1e-4)


# ---------------------------------------------------------------- T-stat ------------------------------------------
def _mean_se(fn, seeds):
    v = np.array([fn(s) for s in seeds])
    return v.mean(axis=0), v.std(axis=0, ddof=1) / math.sqrt(len(seeds))


@pytest.mark.parametrize("sampling,kw,precision", [("forward", {}, 0), ("forward", {}, 1),
                                                   ("backward", dict(neighborhood_dist=6), 0),
                                                   ("backward", dict(neighborhood_dist=6), 1), ("pool", {}, 0)])
def test_k2_exact_probabilities(sampling, kw, precision):
    """IS estimate of P(Y_i) and E[d | Y_i] against the exact lattice recursion (K2)."""
    c = synth.make_config(6, 12, eps=0.1, seed=11, density=0.3)
    p = 6
    px = exact.genotype_probs_exp(c["poset"], c["lambdas"], 1.0)
    ex = np.array([exact.obs_prob_and_dist(px, exact.mask_of(c["obs"][i]), 0.1, p) for i in range(12)])
    L = 4096 if sampling != "pool" else 512  # pool: K = p L entries

    def est(seed):
        r = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, sampling=sampling, seed=seed,
                                 precision=precision, **kw)
        Leff = r["L_eff"] if sampling == "backward" else L
        # sum w d = dist * sum w: an UNBIASED estimate of L P(Y) E[d | Y]; `dist` itself is a ratio estimator with an O(1/L)
        # bias (the reference's too), visible against the exact value once enough replicates are averaged
        return np.concatenate([r["w"] / Leff, r["dist"] * r["w"] / Leff])

    mean, se = _mean_se(est, range(100, 124))
    z = (mean - np.concatenate([ex[:, 0], ex[:, 0] * ex[:, 1]])) / np.maximum(se, 1e-300)
    joint_check(z, 23, f"K2 {sampling} precision={precision}")  # one-sample: 24 replicate seeds against the exact value


def test_k1_empty_poset_given_times():
    p, N = 5, 8
    rng = np.random.default_rng(8)
    lam = rng.uniform(0.5, 2.0, p)
    poset = np.zeros((p, p), np.int32)
    obs = rng.integers(0, 2, (N, p)).astype(np.int32)
    times = rng.uniform(0.3, 2.0, N)
    ex = np.array([exact.k1_empty_poset(obs[i], lam, 0.05, times[i]) for i in range(N)])

    def est(seed):
        r = mc.importance_weight(obs, 4096, poset, lam, 0.05, time=times, seed=seed)
        return np.concatenate([r["w"] / 4096, r["dist"] * r["w"] / 4096])  # unbiased numerator, see above

    mean, se = _mean_se(est, range(300, 324))
    joint_check((mean - np.concatenate([ex[:, 0], ex[:, 0] * ex[:, 1]])) / se, 23, "K1")


@pytest.mark.parametrize("sampling,kw", [("forward", {}), ("backward", {}), ("bernoulli", {}), ("pool", {}),
                                         ("add-remove", {})])
def test_estep_statistics_match_oracle(oracle, sampling, kw):
    """Per-observation sum w / L_eff, E[d], E[Z] from independent streams: GPU (Philox) vs oracle (mt19937)."""
    c = synth.make_config(8, 10, eps=0.1, seed=17, density=0.3)
    L = {"backward": 9 * 200, "pool": 256}.get(sampling, 2048)

    def run(fn):
        def est(seed):
            r = fn(seed)
            Leff = r["L_eff"] if sampling == "backward" else L
            good = r["w"] > 0
            return np.concatenate([np.where(good, r["w"] / np.maximum(Leff, 1), 0.0), r["dist"], r["Tdiff"].ravel()])
        return est

    g = run(lambda s: mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, sampling=sampling, seed=s, **kw))
    o = run(lambda s: oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, sampling=sampling, seed=s,
                                               thrds=4, **kw))
    mg, sg = _mean_se(g, range(1, 25))
    mo, so = _mean_se(o, range(1001, 1025))
    z = (mg - mo) / np.sqrt(sg ** 2 + so ** 2 + 1e-300)
    joint_check(z, 46, f"E-step statistics {sampling}")  # 24 + 24 replicate seeds


@pytest.mark.parametrize("sampling,L,k,eps_true", [("forward", 300, 1, 0.05), ("backward", 37 * 10, 2, 0.05),
                                                   ("bernoulli", 300, 1, 0.0), ("pool", 100, 1, 0.05),
                                                   ("add-remove", 300, 1, 0.05)])
def test_mcem_fixed_point_matches_oracle(oracle, sampling, L, k, eps_true):
    """Converged lambda, eps and observed log-likelihood from independent RNG streams agree with the reference
    restatement within 3 Monte-Carlo standard errors (north_star), s.e. from replicate seeds on both sides.
    (bernoulli runs on noise-free data: with noisy data the reference itself aborts with "all samples have weight
    0" as soon as one observation has no compatible genotype among its L coin-flip proposals.)"""
    c = synth.make_config(8, 300, eps=eps_true, seed=23, density=0.3)
    kw = dict(L=L, eps=0.1, sampling=sampling, max_iter=60, update_step_size=20, tol=1e-3, neighborhood_dist=k)

    def pack(r):
        return np.concatenate([r["lambda_"], [r["eps"], r["llhood"]]])

    g = np.array([pack(mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], seed=s, **kw)) for s in range(1, 21)])
    o = np.array([pack(oracle.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], seed=s, thrds=oracle.num_threads(), **kw))
                  for s in range(101, 121)])
    se = np.sqrt(g.var(axis=0, ddof=1) / g.shape[0] + o.var(axis=0, ddof=1) / o.shape[0])
    z = (g.mean(axis=0) - o.mean(axis=0)) / np.maximum(se, 1e-12 * np.abs(o.mean(axis=0)))
    joint_check(z, 38, f"fixed point {sampling}")  # per-parameter 3 s.e. under the joint rule of tests/stat_criteria.py
    # recovery (reference validation method): estimates are close to the generating parameters
    rel = np.abs(g.mean(axis=0)[:8] - c["lambdas"]).mean() / c["lambdas"].mean()
    if sampling != "add-remove":  # its proposal only reaches genotypes one closure step from Y: biased by design
        assert rel < 0.35 and abs(g.mean(axis=0)[8] - eps_true) < 0.03


def test_mcem_trace_and_batch_average_rule():
    c = synth.make_config(6, 100, eps=0.05, seed=2, density=0.3)
    r = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=128, eps=0.1, max_iter=50, update_step_size=20, seed=4,
                     return_trace=True)
    n = r["n_iter"]
    assert n in (20, 40, 50) and r["trace"].shape == (n, 8)
    last = r["trace"][40:50] if n == 50 else r["trace"][n - 20:n]
    np.testing.assert_allclose(r["lambda_"], last[:, 2:].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(r["eps"], last[:, 1].mean(), rtol=1e-12)
    np.testing.assert_allclose(r["llhood"], last[:, 0].mean(), rtol=1e-12)
    # reproducible for a fixed seed (deterministic reductions), different for another seed
    r2 = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=128, eps=0.1, max_iter=50, update_step_size=20, seed=4)
    assert np.array_equal(r["lambda_"], r2["lambda_"]) and r["llhood"] == r2["llhood"]
    r3 = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=128, eps=0.1, max_iter=50, update_step_size=20, seed=5)
    assert r3["llhood"] != r["llhood"]


def test_backward_empty_ball_raises_like_the_reference(oracle):
    """Some noisy observations have no poset-compatible genotype within Hamming distance 1: the reference aborts the
    run with "all samples have weight 0" (src/mcem_hcbn.cpp:695-699) and so must the device path."""
    c = synth.make_config(8, 300, eps=0.05, seed=23, density=0.3)
    kw = dict(L=360, eps=0.1, sampling="backward", max_iter=5, update_step_size=5, seed=1)
    with pytest.raises(oracle.OracleError) as eo:
        oracle.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], thrds=4, **kw)
    with pytest.raises(mc.MccbnError) as eg:
        mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], **kw)
    assert eo.value.code == 2 and eg.value.code == 2 and str(eo.value) == str(eg.value)


def test_zero_weight_error():
    c = _cfg(p=6, N=5)
    with pytest.raises(mc.MccbnError) as e:  # backward with L < 1 + p: no samples (src/mcem_hcbn.cpp:362-363)
        mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=3, eps=0.1, sampling="backward", max_iter=2,
                     update_step_size=1, seed=1)
    assert e.value.code == 2 and "weight 0" in str(e.value)


def test_extreme_epsilon_values():
    c = _cfg(p=10, N=20, seed=6)
    for eps in (1e-12, 0.5, 0.9):  # tiny eps: weights span hundreds of orders of magnitude; eps > 1/2: flipped table
        ll = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], eps, L=2048, seed=3)
        assert np.isfinite(ll)
    px = exact.genotype_probs_exp(c["poset"], c["lambdas"], 1.0)
    ex = sum(math.log(exact.obs_prob_and_dist(px, exact.mask_of(c["obs"][i]), 0.9, 10)[0]) for i in range(20))
    est = np.array([mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.9, L=8192, seed=s) for s in range(16)])
    assert abs(est.mean() - ex) < 5 * est.std(ddof=1) / 4 + 0.05


# ---------------------------------------------------------------- full-size properties ----------------------------
def test_full_size_config2_properties():
    """BASELINE config 2 (p=16, N=10,000, L=1,000): size-independent properties on the full problem."""
    c = synth.make_config(16, 10000, eps=0.05, seed=10)
    r = mc.importance_weight(c["obs"], 1000, c["poset"], c["lambdas"], 0.05, seed=1)
    assert np.all(r["w"] > 0) and np.all(np.isfinite(r["Tdiff"]))
    assert np.all(r["dist"] >= 0) and np.all(r["dist"] <= 16)
    # Cauchy-Schwarz: (sum w)^2 <= L sum w^2  (ESS <= L), and ESS >= 1
    ess = r["w"] ** 2 / r["w_sqrt"]
    assert np.all(ess <= 1000 * (1 + 1e-6)) and np.all(ess >= 1 - 1e-6)
    # identical genotypes get statistically identical estimates; exact duplicates of the weight table bounds
    lo, hi = 0.05 ** 16, 0.95 ** 16
    assert np.all(r["w"] / 1000 >= lo * (1 - 1e-6)) and np.all(r["w"] / 1000 <= hi * (1 + 1e-6))
    # observed log-likelihood equals the sum of per-observation logs
    ll = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=1000, seed=1)
    assert abs(ll - np.log(r["w"] / 1000).sum()) <= 1e-9 * abs(ll)
    # one EM iteration from the truth stays near the truth (fixed point property, N = 10,000)
    pb = mc.Problem(c["obs"])
    pb.set_model(c["poset"], c["lambdas"], 0.05)
    pb.em_iterations(1000, n_iter=3, seed=2)
    out = pb.read()
    assert np.abs(out["lambda_"] / c["lambdas"] - 1).max() < 0.25 and abs(out["eps"] - 0.05) < 0.01


def test_full_size_config3_properties():
    """BASELINE config 3 (p=30, N=100,000, L=10,000: 1e9 samples per E-step) at full size, through properties that
    do not need the oracle: determinism, independence of the kernel variant and of the sharding, weight bounds."""
    c = synth.make_config(30, 100000, eps=0.05, seed=10)
    N, L = 100000, 10000
    ctx = mc.Context(0)
    try:
        mc.set_jit(1, ctx=ctx)
        ll = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx)
        assert ll == mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx)  # deterministic
        assert ll != mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, seed=2, ctx=ctx)
        # sharding: observations are independent and the Philox counters carry the global index
        parts = [mc.obs_loglikelihood(c["obs"][lo:hi], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx, obs_offset=lo)
                 for lo, hi in ((0, 12500), (12500, 50000), (50000, 100000))]
        assert abs(sum(parts) - ll) <= 1e-7 * abs(ll)   # fp32 per-lane sums depend on the chunking chosen for N
        # the generic kernel draws the same samples: same estimate up to fp32 summation order
        mc.set_jit(0, ctx=ctx)
        ll_generic = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx)
        assert abs(ll_generic - ll) <= 1e-7 * abs(ll)
        # bounds: every observation's estimate of P(Y) lies between the smallest and the largest possible weight
        mc.set_jit(1, ctx=ctx)
        r = mc.importance_weight(c["obs"][:20000], L, c["poset"], c["lambdas"], 0.05, seed=1, ctx=ctx)
        assert np.all(r["w"] / L >= 0.05 ** 30 * (1 - 1e-5)) and np.all(r["w"] / L <= 0.95 ** 30 * (1 + 1e-5))
        assert np.all(np.isfinite(r["Tdiff"])) and np.all(r["Tdiff"] > 0) and np.all((r["dist"] >= 0) & (r["dist"] <= 30))
        assert abs(np.log(r["w"] / L).sum() - parts[0] - mc.obs_loglikelihood(
            c["obs"][12500:20000], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx, obs_offset=12500)) <= 1e-6 * abs(ll)
    finally:
        mc.set_jit(-1, ctx=ctx)
        ctx.close()


def test_pool_given_times_matches_forward_exact():
    """pool with sampling times available: X_k = [T_k <= t_i]; estimate against the K2' matrix-exponential answer."""
    c = synth.make_config(6, 10, eps=0.1, seed=31, density=0.3)
    ex = []
    for i in range(10):
        px = exact.genotype_probs_time(c["poset"], c["lambdas"], c["times"][i])
        ex.append(exact.obs_prob_and_dist(px, exact.mask_of(c["obs"][i]), 0.1, 6))
    ex = np.array(ex)
    L = 512

    def est(seed):
        r = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, time=c["times"], sampling="pool", seed=seed)
        return np.concatenate([r["w"] / L, r["dist"] * r["w"] / L])  # unbiased numerator of E[d | Y]

    mean, se = _mean_se(est, range(500, 524))
    joint_check((mean - np.concatenate([ex[:, 0], ex[:, 0] * ex[:, 1]])) / se, 23, "pool, times given, K2'")


@pytest.mark.parametrize("times_given", [False, True])
def test_effective_sample_size_forward(oracle, times_given):
    """The reference's own diagnostic (utils/hcbn_test.r T4-T6): ESS = (sum w)^2 / sum w^2 per genotype from the
    `_importance_weight` outputs (w, w_sqrt), against its exact large-L limit L E[w]^2 / E[w^2] (K2 / K2') and the
    oracle's independent stream."""
    c = _cfg(p=7, N=16, eps=0.1, seed=4)
    L, eps = 100000, 0.1
    times = c["times"] if times_given else None
    r = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], eps, time=times, seed=5)
    o = oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], eps, times=times, thrds=oracle.num_threads(),
                                 seed=6)
    ess, ess_o = r["w"] ** 2 / r["w_sqrt"], o["w"] ** 2 / o["w_sqrt"]
    px = None if times_given else exact.genotype_probs_exp(c["poset"], c["lambdas"], 1.0)
    for i in range(c["obs"].shape[0]):
        if times_given:
            px = exact.genotype_probs_time(c["poset"], c["lambdas"], times[i])
        m1, m2 = exact.forward_weight_moments(px, exact.mask_of(c["obs"][i]), eps, 7)
        lim = L * m1 * m1 / m2
        assert abs(ess[i] / lim - 1) < 0.05 and abs(ess[i] / ess_o[i] - 1) < 0.07, (i, ess[i], ess_o[i], lim)
        assert abs(r["w"][i] / (L * m1) - 1) < 0.03  # sum w -> L P(Y)


_T1_CACHE = {}


def _t1_table():
    if not _T1_CACHE:
        c = _cfg(p=5, N=4, eps=0.1, seed=6)
        bf = exact.brute_force_posterior(c["poset"], c["lambdas"], 0.1, 1500000, seed=1)
        ys = sorted(bf, key=lambda y: -bf[y]["n"])[:6]
        _T1_CACHE.update(c=c, bf=bf, ys=ys)
    return _T1_CACHE["c"], _T1_CACHE["bf"], _T1_CACHE["ys"]


@pytest.mark.parametrize("sampling,L,kw", [("forward", 100000, {}), ("bernoulli", 100000, {}),
                                           ("backward", 32 * 2000, dict(neighborhood_dist=5)), ("pool", 4000, {})])
def test_t1_brute_force_posterior_means(sampling, L, kw):
    """The reference's validation method T1 (utils/hcbn_test.r:101-193): importance-sampling estimates of P(Y),
    E[Z_j | Y] and E[d | Y] against frequencies / conditional means over 1.5e6 simulated noisy genotypes (plain numpy,
    independent of the oracle and of the device code).  Every scheme with full support is consistent for them:
    forward, bernoulli, backward with the whole cube as its ball, pool."""
    c, bf, ys = _t1_table()
    p, eps = 5, 0.1
    obs = np.array([[(y >> j) & 1 for j in range(p)] for y in ys], np.int32)
    reps = [mc.importance_weight(obs, L, c["poset"], c["lambdas"], eps, sampling=sampling, seed=s, **kw)
            for s in range(1, 17)]
    zs = []
    for i, y in enumerate(ys):
        b = bf[y]
        Zs = np.array([r["Tdiff"][i] for r in reps])
        ds = np.array([r["dist"][i] for r in reps])
        Leff = np.array([(r["L_eff"][i] if sampling == "backward" else L) for r in reps], float)
        ps = np.array([r["w"][i] for r in reps]) / Leff
        se_Z = np.sqrt(Zs.var(axis=0, ddof=1) / len(reps) + b["Z_se"] ** 2)
        zs.extend((Zs.mean(axis=0) - b["Z"]) / se_Z)
        zs.append((ds.mean() - b["d"]) / math.sqrt(ds.var(ddof=1) / len(reps) + b["d_se"] ** 2))
        se_p = math.sqrt(ps.var(ddof=1) / len(reps) + b["prob"] * (1 - b["prob"]) / 1500000)
        zs.append((ps.mean() - b["prob"]) / se_p)
    # 16 replicate seeds against brute-force values with their own (large-sample) standard errors: nu = 15 is conservative
    joint_check(np.array(zs), len(reps) - 1, f"T1 {sampling}")


def test_otcbn_mcem_matches_oracle(oracle):
    """Legacy OT-CBN entry (`MCEM`, src/mcem.cpp:307-335): noise-free model, conditional proposal with X == Y."""
    c = synth.make_config(7, 300, eps=0.0, seed=41, density=0.3)
    kw = dict(max_iter=40, nrOfSamples=16)
    g = np.array([np.concatenate([r["lambda_"], [r["llhood"]]]) for r in
                  (mc.estimate_mutation_rates(c["poset"], c["obs"], ilambda=c["lambda0"], zeta=0.2, seed=s, **kw)
                   for s in range(1, 17))])
    o = np.array([np.concatenate([r["lambda_"], [r["llhood"]]]) for r in
                  (oracle.MCEM_otcbn(c["lambda0"], c["poset"], c["obs"], alpha=0.2, seed=s, **kw)
                   for s in range(101, 117))])
    se = np.sqrt(g.var(axis=0, ddof=1) / 16 + o.var(axis=0, ddof=1) / 16)
    z = (g.mean(axis=0) - o.mean(axis=0)) / se
    joint_check(z, 30, "OT-CBN MCEM fixed point")
    # with sampling times given
    g2 = mc.estimate_mutation_rates(c["poset"], c["obs"], sampling_times=c["times"], ilambda=c["lambda0"], seed=3, **kw)
    o2 = oracle.MCEM_otcbn(c["lambda0"], c["poset"], c["obs"], times=c["times"], seed=3, **kw)
    assert np.abs(g2["lambda_"] / o2["lambda_"] - 1).max() < 0.1
    # incompatible genotype with non-zero weight: the R wrapper's error (R/mcem_param_est.r:57-59)
    bad = c["obs"].copy()
    u, v = [int(a[0]) for a in np.nonzero(c["poset"])]
    bad[0, u], bad[0, v] = 0, 1
    with pytest.raises(mc.MccbnError) as e:
        mc.estimate_mutation_rates(c["poset"], bad, ilambda=c["lambda0"], seed=1, **kw)
    assert "incompatible with the poset" in str(e.value)


def test_add_remove_proposal_support_and_times_given(oracle):
    """add-remove (src/mcem_hcbn.cpp:389-459): incompatible observations, the wild type and the fully mutated genotype
    take different move sets; with sampling times given.  GPU vs oracle from independent streams."""
    c = synth.make_config(7, 6, eps=0.1, seed=31, density=0.35)
    obs = c["obs"].copy()
    obs[0, :] = 0                      # wild type: add or stand still
    obs[1, :] = 1                      # all events: remove or stand still
    topo, pm = mc.flatten_poset(c["poset"])
    child = next(j for j in range(7) if pm[j])          # an event with a parent
    obs[2, :] = 0
    obs[2, child] = 1                  # incompatible: add or remove only
    times = np.array([0.4, 2.5, 0.8, 1.0, 1.2, 0.6])
    L = 2048

    def run(fn):
        def est(seed):
            r = fn(seed)
            return np.concatenate([r["w"] / L, r["dist"], r["Tdiff"].ravel()])
        return est

    g = run(lambda s: mc.importance_weight(obs, L, c["poset"], c["lambdas"], 0.1, time=times, sampling="add-remove", seed=s))
    o = run(lambda s: oracle.importance_weight(obs, L, c["poset"], c["lambdas"], 0.1, times=times, sampling="add-remove",
                                               seed=s, thrds=4))
    mg, sg = _mean_se(g, range(1, 25))
    mo, so = _mean_se(o, range(1001, 1025))
    z = (mg - mo) / np.sqrt(sg ** 2 + so ** 2 + 1e-300)
    joint_check(z, 46, "add-remove support")


# ---------------------------------------------------------------- 32 < p <= 64 (forward scheme, 64-bit genotype words)
@pytest.mark.parametrize("p", [33, 40, 48, 57, 64])
def test_wide_forward_k1_empty_poset(p):
    """exact answer K1 (empty poset, sampling times given) for p beyond one 32-bit genotype word.  Forward sampling
    cannot resolve P(Y) for 40 free events with a few thousand samples, so all but six events are (almost) certain to
    have occurred (lambda t >> 1); the observations disagree with two of the certain events -- one of them the last
    bit of the 64-bit word path -- which multiplies P(Y) by known factors, and are random on the six free events."""
    N = 6
    rng = np.random.default_rng(p)
    lam = np.full(p, 80.0)
    free = rng.choice(p, 6, replace=False)
    lam[free] = rng.uniform(0.5, 2.0, 6)
    poset = np.zeros((p, p), np.int32)
    obs = np.ones((N, p), np.int32)
    obs[:, free] = rng.integers(0, 2, (N, 6))
    sure = np.setdiff1d(np.arange(p), free)
    obs[:, sure[-1]] = 0          # the highest certain event: observed as not mutated (one noise flip)
    obs[::2, sure[0]] = 0
    times = rng.uniform(0.3, 2.0, N)
    ex = np.array([exact.k1_empty_poset(obs[i], lam, 0.05, times[i]) for i in range(N)])
    L = 4096

    def est(seed):
        r = mc.importance_weight(obs, L, poset, lam, 0.05, time=times, seed=seed)
        return np.concatenate([r["w"] / L, r["dist"] * r["w"] / L])  # unbiased numerator of E[d | Y]

    mean, se = _mean_se(est, range(300, 324))
    ref = np.concatenate([ex[:, 0], ex[:, 0] * ex[:, 1]])
    joint_check((mean - ref) / np.maximum(se, 1e-300), 23, f"wide K1 p={p}")
    assert np.all(ex[:, 1] >= 1.0)  # at least the forced flip


@pytest.mark.parametrize("p,times", [(36, False), (50, True), (64, False)])
def test_wide_forward_matches_oracle(oracle, p, times):
    """per-observation log sum w, E[d], E[Z] and the MCEM fixed point for p > 32: GPU (Philox) vs oracle (mt19937)"""
    c = synth.make_config(p, 8, eps=0.05, seed=p, density=0.08)
    t = c["times"] if times else None
    L = 1024

    def run(fn):
        def est(seed):
            r = fn(seed)
            return np.concatenate([np.log(r["w"] / L), r["dist"], r["Tdiff"].ravel()])
        return est

    g = run(lambda s: mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, time=t, seed=s))
    o = run(lambda s: oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, times=t, seed=s, thrds=4))
    mg, sg = _mean_se(g, range(1, 17))
    mo, so = _mean_se(o, range(1001, 1017))
    z = (mg - mo) / np.sqrt(sg ** 2 + so ** 2 + 1e-300)
    joint_check(z, 30, f"wide forward p={p}")
    # one full MCEM run stays finite and moves towards the generating parameters
    c2 = synth.make_config(p, 300, eps=0.05, seed=p + 1, density=0.08)
    r = mc.MCEM_hcbn(c2["lambda0"], c2["poset"], c2["obs"], L=64, eps=0.1, max_iter=40, update_step_size=20, seed=3)
    assert np.all(np.isfinite(r["lambda_"])) and 0 < r["eps"] < 0.3 and np.isfinite(r["llhood"])


@pytest.mark.parametrize("sampling,p,eps_true,kw", [("backward", 40, 0.02, dict(neighborhood_dist=1)),
                                                    ("bernoulli", 36, 0.0, {}), ("backward", 64, 0.01, dict(neighborhood_dist=1))])
def test_wide_conditional_schemes_match_oracle(oracle, sampling, p, eps_true, kw):
    """backward / bernoulli for 32 < p <= 64 (NVRTC-specialised kernels on 64-bit genotype words) against the oracle"""
    c = synth.make_config(p, 8, eps=eps_true, seed=p + 3, density=0.06)
    L = (p + 1) * 24 if sampling == "backward" else 1024

    def run(fn):
        def est(seed):
            r = fn(seed)
            good = r["w"] > 0
            Leff = np.maximum(r["L_eff"], 1)
            return np.concatenate([np.where(good, r["w"] / Leff, 0.0), np.where(good, r["dist"], 0.0),
                                   np.where(good[:, None], r["Tdiff"], 0.0).ravel()])
        return est

    g = run(lambda s: mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.05, sampling=sampling, seed=s, **kw))
    o = run(lambda s: oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.05, sampling=sampling, seed=s,
                                               thrds=4, **kw))
    mg, sg = _mean_se(g, range(1, 25))
    mo, so = _mean_se(o, range(1001, 1025))
    z = (mg - mo) / np.sqrt(sg ** 2 + so ** 2 + 1e-300)
    joint_check(z, 46, f"wide {sampling} p={p}")
    assert (mg[:8] > 0).sum() >= 4 and (mo[:8] > 0).sum() >= 4  # the comparison is not vacuous


def test_wide_other_schemes_are_rejected():
    c = synth.make_config(40, 5, eps=0.05, seed=1, density=0.08)
    for sampling in ("pool", "add-remove"):
        with pytest.raises(mc.MccbnError) as e:
            mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.1, L=64, sampling=sampling, seed=1)
        assert "32 < p <= 64" in str(e.value)
