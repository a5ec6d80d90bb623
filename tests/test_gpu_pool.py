"""The pool scheme's own parity tier (src/mcem_hcbn.cpp:184-236, :460-491, :650-654, :668-677).

  T-int / T-fp   from a SUPPLIED pool (T_pool, Tdiff_pool, T_s): pool genotypes X_k and distances d_k bit-exact against the
                 oracle's generate_genotypes / hamming_dist_mat; sum_k q_k (=> obs_llhood_i = log(sum q / K)), E[d], E[Z]
                 <= 1e-10 relative (mccbn_pool_from_times)
  T-stat         pool_resample = 1 (the reference's L-index resampling, :477-484) against the oracle's pool branch, and
                 against the default mode (expectations over the whole pool), of which it is an unbiased, noisier estimate
"""
import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check, replicate_z

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,K,times_given", [(6, 257, False), (25, 1000, False), (25, 640, True), (32, 333, False)])
def test_pool_from_supplied_times(oracle, p, K, times_given):
    c = synth.make_config(p, 9, eps=0.07, seed=40 + p, density=0.2)
    rng = np.random.default_rng(p)
    eps = 0.07
    T_pool, Z_pool = oracle.sample_times(K, c["poset"], c["lambdas"], seed=5)  # sample_times :184-213
    Ts = rng.exponential(1.0, K)
    times = rng.uniform(0.2, 2.0, c["obs"].shape[0]) if times_given else None
    r = mc.pool_from_times(c["obs"], c["poset"], eps, T_pool, Z_pool, None if times_given else Ts, times=times)
    N = c["obs"].shape[0]
    for i in range(N):
        ts_i = np.full(K, times[i]) if times_given else Ts
        X_o, _ = oracle.generate_genotypes(T_pool, c["poset"], times=ts_i)  # :215-236
        d_o = oracle.hamming_dist_mat(X_o, c["obs"][i])                      # :326-330
        if i == 0:
            assert np.array_equal(r["X_pool"], X_o)
        assert np.array_equal(r["d_pool"][:, i], d_o)
        # the reference's expressions (:463-464, :491) evaluated in numpy fp64
        q = np.power(eps, d_o.astype(float)) * np.power(1 - eps, p - d_o.astype(float))
        assert abs(r["q_sum"][i] / q.sum() - 1) <= 1e-10
        assert abs(r["dist"][i] - (q * d_o).sum() / q.sum()) <= 1e-10 * max(1.0, r["dist"][i])
        np.testing.assert_allclose(r["Tdiff"][i], (q[:, None] * Z_pool).sum(axis=0) / q.sum(), rtol=1e-10)
        # ... and through the oracle's pool branch itself: every resampled row carries the weight sum q / K
        w = oracle.importance_weight_genotype(c["obs"][i], 8, c["poset"], c["lambdas"], eps, sampling="pool",
                                              dist_pool=d_o, Tdiff_pool=Z_pool, seed=3)["w"]
        assert np.all(np.abs(w / (r["q_sum"][i] / K) - 1) <= 1e-10)
    ll = np.sum(np.log(r["q_sum"] / K))
    assert abs(r["llhood"] / ll - 1) <= 1e-10


def test_pool_resample_against_oracle_and_default_mode(oracle):
    """pool_resample = 1: per-observation E[d] and E[Z] over replicate seeds agree with the oracle (which resamples like
    the reference); the log-likelihood part equals the default mode's for the same seed (the resampled rows all carry the
    weight sum q / K); the resampled means are unbiased for the default mode's expectations."""
    c = synth.make_config(8, 12, eps=0.1, seed=17, density=0.3)
    L, n_rep = 128, 32
    kw = dict(sampling="pool")
    g1 = [mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=1 + s, pool_resample=True, **kw)
          for s in range(n_rep)]
    g0 = [mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=1 + s, **kw) for s in range(n_rep)]
    o = [oracle.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.1, seed=901 + s, thrds=4, **kw)
         for s in range(n_rep)]

    def vec(r):
        return np.concatenate([r["w"] / L, r["dist"], r["Tdiff"].ravel(order="F")])

    # same pool, same sum q.  The two modes chunk the pool differently and each thread adds <= kPoolFlush = 512 relative
    # weights in fp32 before it flushes to its fp64 record: worst case 512 * 2^-24 = 3e-5 relative, typically ~1e-6
    for a, b in zip(g1, g0):
        np.testing.assert_allclose(a["w"], b["w"], rtol=4e-5)
    z, nu = replicate_z([vec(r) for r in g1], [vec(r) for r in o])
    joint_check(z, nu, "pool_resample vs oracle")
    # paired with the default mode on the same pools: mean difference ~ 0, and the resampling adds variance
    d = np.array([vec(a) - vec(b) for a, b in zip(g1, g0)])[:, c["obs"].shape[0]:]
    zz = d.mean(axis=0) / np.maximum(d.std(axis=0, ddof=1) / np.sqrt(n_rep), 1e-300)
    joint_check(zz, n_rep - 1, "pool_resample unbiased for the pool expectation")
    assert np.mean(np.var([vec(r) for r in g1], axis=0)[12:] >= np.var([vec(r) for r in g0], axis=0)[12:]) > 0.7


def test_pool_resample_fixed_point(oracle):
    c = synth.make_config(8, 200, eps=0.05, seed=23, density=0.3)
    kw = dict(L=60, eps=0.1, sampling="pool", max_iter=40, update_step_size=20, tol=1e-3)

    def pack(r):
        return np.concatenate([r["lambda_"], [r["eps"], r["llhood"]]])

    g = [pack(mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], seed=s, pool_resample=True, **kw)) for s in range(1, 33)]
    o = [pack(oracle.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], seed=100 + s, thrds=oracle.num_threads(), **kw))
         for s in range(1, 33)]
    z, nu = replicate_z(g, o)
    joint_check(z, nu, "pool_resample fixed point")
