"""T-stat parity on the BASELINE configurations themselves ("What's weak" 1-2).

The CUDA path -- BOTH kernels: the NVRTC poset-specialised one that bench.py times and the generic one -- against the CPU
oracle (independent streams: Philox vs the reference's mt19937) on slices of the real configurations:

  cfg2  MCEM.hcbn p=16 L=1,000 forward        cfg4  ASA-sized E-step p=20 L=1,000 forward
  cfg3  MCEM.hcbn p=30 L=10,000 forward       cfg5  p=25 L=5,000 eps=0.05: backward k=1, bernoulli, pool, add-remove

The posets and rates are those of bench.py (synth.make_config(p, ., seed=10): the poset and the rates are drawn before the
observations, so they do not depend on N); the slices are 64 observations x the configuration's L for the E-step
statistics (per-observation P^(Y), E[d | Y], E[Z_j | Y]: up to 64 x 32 simultaneous comparisons) and a few hundred
observations for the converged (lambda, eps, llhood) of MCEM.hcbn -- the quantities BASELINE.json's north_star names
("within 3 Monte Carlo standard errors").  The acceptance rule for K simultaneous comparisons is stated in
tests/stat_criteria.py (per-parameter 3 s.e. with the binomial allowance for K comparisons, Bonferroni, omnibus
chi-square; alpha = 1e-3).  Follows the reference's own validation method, utils/hcbn_test.r:101-193, and its hot loop,
src/mcem_hcbn.cpp:665-709.

test_fp32_vs_fp64_sampling_full_cfg3: the production precision (fp32 sampling / fp64 reduction) against the fp64 mode
(53-bit uniforms, log, fp64 event times -- the reference's arithmetic) at FULL cfg3 size on the same GPU.
"""
import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check, replicate_z

pytestmark = pytest.mark.gpu

# name: (p, L, eps_true) -- bench.py CONFIGS
BASE = {"cfg2": (16, 1000, 0.05), "cfg3": (30, 10000, 0.05), "cfg4": (20, 1000, 0.05), "cfg5": (25, 5000, 0.05)}
N_SLICE = 64
N_REP = 48  # replicate seeds per side, E-step statistics
N_FIT = 16  # replicate seeds per side, MCEM fits


def _config(name, N):
    p, L, eps = BASE[name]
    return synth.make_config(p, N, eps=eps, seed=10), p, L


_ORACLE_CACHE = {}


def _oracle_cached(key, fn):
    """the oracle side of a comparison does not depend on which device kernel is being checked"""
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = fn()
    return _ORACLE_CACHE[key]


class _Jit:
    """Force the specialised (1) or the generic (0) kernel for the default context, and check which one ran."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        mc.set_jit(self.mode)
        mc.jit_launches(reset=True)
        mc.kernel_launches(reset=True)
        return self

    def __exit__(self, *exc):
        n_jit, n_all = mc.jit_launches(), mc.kernel_launches()
        mc.set_jit(-1)
        if exc[0] is None:
            assert n_all > 0
            if self.mode == 1:
                assert n_jit > 0, "the specialised kernel did not run"
            else:
                assert n_jit == 0, "a specialised kernel ran although the generic one was requested"
        return False


def _estep_vector(r, L, sampling):
    """per-observation unbiased P^(Y) = sum w / L_eff, then E[d], then E[Z_j] (ratio estimators, the same on both sides)"""
    Leff = r["L_eff"] if sampling == "backward" else L
    good = r["w"] > 0
    pw = np.where(good, r["w"] / np.maximum(Leff, 1), 0.0)
    return np.concatenate([pw, r["dist"], r["Tdiff"].ravel(order="F")])


def _estep_totals(r, L, sampling):
    """the statistic vector of one E-step: obs_llhood, expected_dist, column sums of E[Z] (src/mcem_hcbn.cpp:683-694)"""
    Leff = np.asarray(r["L_eff"], float) if sampling == "backward" else np.full(r["w"].shape, float(L))
    good = r["w"] > 0
    ll = np.sum(np.log(r["w"][good] / Leff[good]))
    return np.concatenate([[ll, r["dist"][good].sum()], r["Tdiff"][good].sum(axis=0)])


# ------------------------------------------------------------------------------------------------------------------
# E-step statistics, forward scheme: cfg2 / cfg3 / cfg4, specialised and generic kernel
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,jit", [("cfg3", 1), ("cfg3", 0), ("cfg2", 1), ("cfg2", 0), ("cfg4", 1), ("cfg4", 0)])
def test_forward_estep_on_baseline_config(oracle, name, jit):
    c, p, L = _config(name, N_SLICE)
    lam, eps = c["lambda0"], 0.1  # the parameters bench.py starts (and times) from
    with _Jit(jit):
        g = [mc.importance_weight(c["obs"], L, c["poset"], lam, eps, seed=1 + s) for s in range(N_REP)]
    o = _oracle_cached(("fwd", name), lambda: [
        oracle.importance_weight(c["obs"], L, c["poset"], lam, eps, seed=5001 + s, thrds=oracle.num_threads())
        for s in range(N_REP)])
    z, nu = replicate_z([_estep_vector(r, L, "forward") for r in g], [_estep_vector(r, L, "forward") for r in o])
    info = joint_check(z, nu, f"{name} jit={jit} per-observation")
    zt, _ = replicate_z([_estep_totals(r, L, "forward") for r in g], [_estep_totals(r, L, "forward") for r in o])
    info_t = joint_check(zt, nu, f"{name} jit={jit} totals")
    print(name, "jit", jit, info, info_t)


# ------------------------------------------------------------------------------------------------------------------
# E-step statistics on cfg5 (p = 25, L = 5000, eps = 0.05): the four conditional / pool schemes
# ------------------------------------------------------------------------------------------------------------------
def _nonempty_ball(c, p):
    """backward k = 1: observations whose radius-1 Hamming ball holds at least one genotype compatible with the poset
    (the others make the reference abort with "all samples have weight 0", src/mcem_hcbn.cpp:695-699)"""
    keep = []
    for i in range(c["obs"].shape[0]):
        y = c["obs"][i]
        ball = [y] + [np.where(np.arange(p) == j, 1 - y, y) for j in range(p)]
        if mc.compatible_observations(np.array(ball, np.int32), c["poset"])["num_compatible"] > 0:
            keep.append(i)
    return np.array(keep)


@pytest.mark.parametrize("sampling,jit", [("backward", 1), ("backward", 0), ("bernoulli", 1), ("add-remove", 1),
                                          ("add-remove", 0), ("pool", 0)])
def test_cfg5_estep_schemes(oracle, sampling, jit):
    eps = 0.05
    if sampling == "bernoulli":
        # coin-flip proposals at eps = 0.05 reach a compatible genotype from a noisy observation too rarely: the reference
        # aborts on such data; the scheme is exercised on noise-free genotypes of the same poset (as in test_gpu_parity)
        c, p, L = synth.make_config(25, N_SLICE, eps=0.0, seed=10), 25, 5000
        obs = c["obs"]
    else:
        c, p, L = _config("cfg5", 3 * N_SLICE)
        obs = c["obs"]
        if sampling == "backward":
            obs = obs[_nonempty_ball(c, p)]
        obs = np.asfortranarray(obs[:N_SLICE])
    lam = c["lambdas"]
    n_rep = 24 if sampling == "pool" else N_REP  # pool: K = p L = 125,000 entries, O(N K p) in the oracle
    with _Jit(jit):
        g = [mc.importance_weight(obs, L, c["poset"], lam, eps, sampling=sampling, seed=1 + s) for s in range(n_rep)]
    o = _oracle_cached(("cfg5", sampling), lambda: [
        oracle.importance_weight(obs, L, c["poset"], lam, eps, sampling=sampling, seed=7001 + s,
                                 thrds=oracle.num_threads()) for s in range(n_rep)])
    z, nu = replicate_z([_estep_vector(r, L, sampling) for r in g], [_estep_vector(r, L, sampling) for r in o])
    info = joint_check(z, nu, f"cfg5 {sampling} jit={jit} per-observation")
    print("cfg5", sampling, "jit", jit, info)


# ------------------------------------------------------------------------------------------------------------------
# converged lambda / eps / llhood of MCEM.hcbn on the BASELINE posets
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,sampling,jit,N,L,kw", [
    ("cfg3", "forward", 1, 150, 500, {}),
    ("cfg3", "forward", 0, 150, 500, {}),
    ("cfg2", "forward", 1, 200, 500, {}),
    ("cfg4", "forward", 1, 200, 500, {}),
    ("cfg5", "backward", 1, 150, 26 * 15, dict(neighborhood_dist=1)),
    ("cfg5", "add-remove", 1, 150, 300, {}),
    ("cfg5", "pool", 0, 80, 100, {}),
])
def test_mcem_fixed_point_on_baseline_posets(oracle, name, sampling, jit, N, L, kw):
    """lambda_j (p values), eps and the observed log-likelihood returned by MCEM.hcbn -- the batch averages of
    src/mcem_hcbn.cpp:711-735 -- from N_FIT replicate seeds per side (sizes chosen so that the oracle side of the whole
    file stays around two minutes on the GPU box's host cores)."""
    c, p, _ = _config(name, 3 * N if sampling == "backward" else N)
    obs = c["obs"]
    if sampling == "backward":
        obs = np.asfortranarray(obs[_nonempty_ball(c, p)][:N])
    em = dict(L=L, eps=0.1, sampling=sampling, max_iter=60, update_step_size=20, tol=1e-3, **kw)

    def pack(r):
        return np.concatenate([r["lambda_"], [r["eps"], r["llhood"]]])

    with _Jit(jit):
        g = [pack(mc.MCEM_hcbn(c["lambda0"], c["poset"], obs, seed=1 + s, **em)) for s in range(N_FIT)]
    o = _oracle_cached(("em", name, sampling), lambda: [
        pack(oracle.MCEM_hcbn(c["lambda0"], c["poset"], obs, seed=9001 + s, thrds=oracle.num_threads(), **em))
        for s in range(N_FIT)])
    z, nu = replicate_z(g, o)
    info = joint_check(z, nu, f"{name} {sampling} jit={jit} fixed point")
    print(name, sampling, "jit", jit, info)


# ------------------------------------------------------------------------------------------------------------------
# production precision against the fp64 mode at full cfg3 size
# ------------------------------------------------------------------------------------------------------------------
def test_fp32_vs_fp64_sampling_full_cfg3():
    """BASELINE config 3 in full (p = 30, N = 100,000, L = 10,000; 1e9 samples per E-step): the observed log-likelihood,
    the expected distance and the p column sums of E[Z] of one E-step, fp32 sampling (the specialised kernel bench.py
    times) against fp64 sampling, 6 replicate seeds each.  The systematic difference must be below the replicate
    standard error (same joint criterion) -- the fp32 path's biases derived in DESIGN.md (~1e-6 relative) against a
    Monte-Carlo s.e. of ~3e-5 relative at this size."""
    p, N, L = 30, 100000, 10000
    c = synth.make_config(p, N, eps=0.05, seed=10)
    lam, eps = c["lambda0"], 0.1

    def totals(r):
        return np.concatenate([[np.log(r["w"] / L).sum(), r["dist"].sum()], r["Tdiff"].sum(axis=0)])

    with _Jit(1):
        a = [totals(mc.importance_weight(c["obs"], L, c["poset"], lam, eps, seed=11 + s)) for s in range(6)]
    mc.set_jit(0)
    try:
        b = [totals(mc.importance_weight(c["obs"], L, c["poset"], lam, eps, seed=31 + s, precision=1)) for s in range(6)]
    finally:
        mc.set_jit(-1)
    z, nu = replicate_z(a, b)
    info = joint_check(z, nu, "cfg3 full size, fp32 vs fp64 sampling")
    rel_se = np.sqrt(np.var(a, axis=0, ddof=1) / 6 + np.var(b, axis=0, ddof=1) / 6) / np.abs(np.mean(b, axis=0))
    rel_diff = np.abs(np.mean(a, axis=0) - np.mean(b, axis=0)) / np.abs(np.mean(b, axis=0))
    print("cfg3 full fp32 vs fp64:", info, "max rel diff %.2e, median rel s.e. %.2e" % (rel_diff.max(), np.median(rel_se)))
