"""ctypes binding of the CPU oracle (oracle/mccbn_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mccbn_b200) never imports this module.

All matrices are numpy arrays in R layout (column-major / Fortran order): obs is N x p int32,
poset is p x p int32 with poset[i, j] == 1 meaning "i is a parent of j".
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmccbn_oracle.so")

ERR_OK, ERR_NOT_ACYCLIC, ERR_ZERO_WEIGHT = 0, 1, 2


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "mccbn_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_log_bernoulli_process.restype = C.c_double
        _lib.ref_log_bernoulli_process.argtypes = [C.c_uint, C.c_double, C.c_uint]
        _lib.ref_n_choose_k.restype = C.c_uint
        _lib.ref_n_choose_k.argtypes = [C.c_uint, C.c_uint]
    return _lib


def _chk(rc):
    if rc != 0:
        raise OracleError(rc, lib().ref_last_error().decode())


def _i32(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int32))


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads():
    return lib().ref_num_threads()


def flatten_poset(poset):
    poset = _i32(poset)
    p = poset.shape[0]
    topo = np.zeros(p, np.int32)
    mask = np.zeros(p, np.uint64)
    cyc = C.c_int(0)
    _chk(lib().ref_flatten_poset(_p(poset), p, _p(topo), _p(mask), C.byref(cyc)))
    return topo, mask, bool(cyc.value)


def compatible_observations(obs, poset):
    obs, poset = _i32(obs), _i32(poset)
    out = np.zeros(obs.shape[0], np.int32)
    _chk(lib().ref_compatible_observations(_p(obs), obs.shape[0], _p(poset), poset.shape[0], _p(out)))
    return out


def num_compatible_observations(obs, poset):
    obs, poset = _i32(obs), _i32(poset)
    out = C.c_int(0)
    _chk(lib().ref_num_compatible_observations(_p(obs), obs.shape[0], _p(poset), poset.shape[0], C.byref(out)))
    return out.value


def num_incompatible_events(obs, poset):
    obs, poset = _i32(obs), _i32(poset)
    out = C.c_int(0)
    _chk(lib().ref_num_incompatible_events(_p(obs), obs.shape[0], _p(poset), poset.shape[0], C.byref(out)))
    return out.value


def hamming_dist_mat(x, y):
    x, y = _i32(x), _i32(y)
    out = np.zeros(x.shape[0], np.int32)
    _chk(lib().ref_hamming_dist_mat(_p(x), x.shape[0], x.shape[1], _p(y), _p(out)))
    return out


def n_choose_k(n, k):
    return lib().ref_n_choose_k(n, k)


def neighbors(p, k, genotype):
    g = _i32(genotype)
    total = sum(n_choose_k(p, i) for i in range(k + 1))
    out = np.zeros((total, p), np.int32, order="F")
    ring = np.zeros(total, np.int32)
    _chk(lib().ref_neighbors(p, k, _p(g), _p(out), _p(ring)))
    return out, ring


def coin_tossing(eps, L, genotype, seed):
    g = _i32(genotype)
    out = np.zeros((L, g.size), np.int32, order="F")
    _chk(lib().ref_coin_tossing(C.c_double(eps), L, _p(g), g.size, seed, _p(out)))
    return out


def sample_genotypes(N, poset, lambdas, lambda_s=1.0, times=None, seed=1):
    poset, lam = _i32(poset), _f64(lambdas)
    p = poset.shape[0]
    Z = np.zeros((N, p), np.float64, order="F")
    Ts = np.zeros(N) if times is None else _f64(times).copy()
    S = np.zeros((N, p), np.int32, order="F")
    _chk(lib().ref_sample_genotypes(N, _p(poset), p, _p(lam), _p(Z), _p(Ts), C.c_float(lambda_s),
                                    int(times is not None), seed, _p(S)))
    return S, Z, Ts


def sample_times(N, poset, lambdas, seed=1):
    poset, lam = _i32(poset), _f64(lambdas)
    p = poset.shape[0]
    Z = np.zeros((N, p), order="F")
    T = np.zeros((N, p), order="F")
    _chk(lib().ref_sample_times(N, _p(poset), p, _p(lam), seed, _p(Z), _p(T)))
    return T, Z


def generate_genotypes(T_sum, poset, times=None, lambda_s=1.0, seed=1):
    T_sum, poset = _f64(T_sum), _i32(poset)
    N, p = T_sum.shape
    Ts = np.zeros(N) if times is None else _f64(times).copy()
    S = np.zeros((N, p), np.int32, order="F")
    _chk(lib().ref_generate_genotypes(_p(T_sum), N, _p(poset), p, _p(Ts), C.c_float(lambda_s),
                                      int(times is not None), seed, _p(S)))
    return S, Ts


def forward_from_times(genotype, poset, eps, Z, T_sampling):
    g, poset, Z, Ts = _i32(genotype), _i32(poset), _f64(Z), _f64(T_sampling)
    L, p = Z.shape
    Tsum = np.zeros((L, p), order="F")
    X = np.zeros((L, p), np.int32, order="F")
    d = np.zeros(L, np.int32)
    w = np.zeros(L)
    _chk(lib().ref_forward_from_times(_p(g), L, _p(poset), p, C.c_double(eps), _p(Z), _p(Ts), _p(Tsum), _p(X),
                                      _p(d), _p(w)))
    return dict(T_sum=Tsum, X=X, dist=d, w=w)


def generate_mutation_times(obs, poset, lambdas, times=None, lambda_s=1.0, seed=1, uniforms=None):
    obs, poset, lam = _i32(obs), _i32(poset), _f64(lambdas)
    N, p = obs.shape
    avail = times is not None
    Ts = np.zeros(N) if times is None else _f64(times).copy()
    U = None if uniforms is None else _f64(uniforms)
    T = np.zeros((N, p), order="F")
    dens = np.zeros(N)
    _chk(lib().ref_generate_mutation_times(_p(obs), N, _p(poset), p, _p(lam), _p(Ts), C.c_float(lambda_s),
                                           int(avail), seed, _p(U), _p(T), _p(dens)))
    return dict(Tdiff=T, proposal_density=dens, sampling_time=Ts)


def conditional_from_uniforms(X, poset, lambdas, u, T_sampling):
    X, poset, lam, u, Ts = _i32(X), _i32(poset), _f64(lambdas), _f64(u), _f64(T_sampling)
    L, p = X.shape
    T = np.zeros((L, p), order="F")
    lq = np.zeros(L)
    lpz = np.zeros(L)
    inc = np.zeros(L, np.int32)
    _chk(lib().ref_conditional_from_uniforms(_p(X), L, _p(poset), p, _p(lam), _p(u), _p(Ts), _p(T), _p(lq), _p(lpz),
                                             _p(inc)))
    return dict(Tdiff=T, log_q=lq, log_pz=lpz, incompatible=inc)


def cbn_density_log(time, lambdas):
    time, lam = _f64(time), _f64(lambdas)
    out = np.zeros(time.shape[0])
    _chk(lib().ref_cbn_density_log(_p(time), time.shape[0], time.shape[1], _p(lam), _p(out)))
    return out


def scale_path_to_mutation(poset, lambdas):
    poset, lam = _i32(poset), _f64(lambdas)
    out = np.zeros(poset.shape[0])
    _chk(lib().ref_scale_path_to_mutation(_p(poset), poset.shape[0], _p(lam), _p(out)))
    return out


def complete_log_likelihood(lambdas, eps, Tdiff, dist, weights):
    lam, Tdiff, dist, weights = _f64(lambdas), _f64(Tdiff), _f64(dist), _f64(weights)
    out = C.c_double(0)
    _chk(lib().ref_complete_log_likelihood(_p(lam), lam.size, C.c_double(eps), _p(Tdiff), Tdiff.shape[0], _p(dist),
                                           _p(weights), C.byref(out)))
    return out.value


def log_bernoulli_process(dist, eps, p):
    return lib().ref_log_bernoulli_process(dist, eps, p)


def importance_weight_genotype(genotype, L, poset, lambdas, eps, time=None, sampling="forward", scale_cumulative=None,
                               dist_pool=None, Tdiff_pool=None, neighborhood_dist=1, lambda_s=1.0, seed=1):
    g, poset, lam = _i32(genotype), _i32(poset), _f64(lambdas)
    p = poset.shape[0]
    sc = None if scale_cumulative is None else _f64(scale_cumulative)
    dp = None if dist_pool is None else _i32(dist_pool)
    Tp = None if Tdiff_pool is None else _f64(Tdiff_pool)
    K = 0 if dp is None else dp.size
    w = np.zeros(L)
    d = np.zeros(L, np.int32)
    T = np.zeros(L * p)
    Lout = C.c_int(0)
    _chk(lib().ref_importance_weight_genotype(
        _p(g), L, _p(poset), p, _p(lam), C.c_double(eps), C.c_double(0.0 if time is None else time),
        sampling.encode(), _p(sc), _p(dp), _p(Tp), K, neighborhood_dist, C.c_float(lambda_s), int(time is not None),
        seed, C.byref(Lout), _p(w), _p(d), _p(T)))
    n = Lout.value
    return dict(w=w[:n], dist=d[:n], Tdiff=T[:n * p].reshape((n, p), order="F"))


def importance_weight(obs, L, poset, lambdas, eps, times=None, sampling="forward", neighborhood_dist=1, lambda_s=1.0,
                      thrds=1, seed=1):
    obs, poset, lam = _i32(obs), _i32(poset), _f64(lambdas)
    N, p = obs.shape
    t = np.zeros(N) if times is None else _f64(times)
    w = np.zeros(N)
    w2 = np.zeros(N)
    Le = np.zeros(N, np.int32)
    ed = np.zeros(N)
    eT = np.zeros((N, p), order="F")
    _chk(lib().ref_importance_weight(_p(obs), N, L, _p(poset), p, _p(lam), C.c_double(eps), _p(t), sampling.encode(),
                                     neighborhood_dist, C.c_float(lambda_s), int(times is not None), thrds, seed,
                                     _p(w), _p(w2), _p(Le), _p(ed), _p(eT)))
    return dict(w=w, w_sqrt=w2, L_eff=Le, dist=ed, Tdiff=eT)


def obs_log_likelihood(obs, poset, lambdas, eps, L, weights=None, times=None, sampling="forward", neighborhood_dist=1,
                       lambda_s=1.0, thrds=1, seed=1):
    obs, poset, lam = _i32(obs), _i32(poset), _f64(lambdas)
    N, p = obs.shape
    wt = np.ones(N) if weights is None else _f64(weights)
    t = np.zeros(N) if times is None else _f64(times)
    out = C.c_double(0)
    _chk(lib().ref_obs_log_likelihood(_p(obs), N, _p(poset), p, _p(lam), C.c_double(eps), _p(wt), _p(t), L,
                                      sampling.encode(), neighborhood_dist, C.c_float(lambda_s),
                                      int(times is not None), thrds, seed, C.byref(out)))
    return out.value


def MCEM_hcbn(ilambda, poset, obs, L, eps, lambda_s=1.0, times=None, weights=None, sampling="forward", max_iter=100,
              update_step_size=20, tol=0.001, max_lambda=1e6, neighborhood_dist=1, thrds=1, verbose=False, seed=1,
              return_trace=False):
    obs, poset, lam = _i32(obs), _i32(poset), _f64(ilambda)
    N, p = obs.shape
    wt = np.ones(N) if weights is None else _f64(weights)
    t = np.zeros(N) if times is None else _f64(times)
    out_l = np.zeros(p)
    out_e = C.c_double(0)
    out_ll = C.c_double(0)
    trace = np.zeros((max_iter, p + 2))
    nit = C.c_int(0)
    _chk(lib().ref_MCEM_hcbn(_p(lam), _p(poset), _p(obs), _p(t), C.c_float(lambda_s), C.c_double(eps), _p(wt),
                             C.c_uint(L), sampling.encode(), C.c_uint(max_iter), C.c_uint(update_step_size),
                             C.c_double(tol), C.c_float(max_lambda), C.c_uint(neighborhood_dist),
                             int(times is not None), thrds, int(verbose), seed, N, p, _p(out_l), C.byref(out_e),
                             C.byref(out_ll), _p(trace), C.byref(nit)))
    res = dict(lambda_=out_l, eps=out_e.value, llhood=out_ll.value, n_iter=nit.value)
    if return_trace:
        res["trace"] = trace[:nit.value].copy()
    return res


def estep_forward_from_times(obs, poset, lambdas, eps, Z, T_sampling, weights=None, times=None, max_lambda=1e6):
    obs, poset, lam, Z, Ts = _i32(obs), _i32(poset), _f64(lambdas), _f64(Z), _f64(T_sampling)
    N, p = obs.shape
    L = Z.shape[0]
    wt = np.ones(N) if weights is None else _f64(weights)
    t = np.zeros(N) if times is None else _f64(times)
    w = np.zeros(N)
    w2 = np.zeros(N)
    ed = np.zeros(N)
    eT = np.zeros((N, p), order="F")
    tot = np.zeros(3 + p)
    en = C.c_double(0)
    ln = np.zeros(p)
    _chk(lib().ref_estep_forward_from_times(_p(obs), N, _p(poset), p, _p(lam), C.c_double(eps), _p(wt), _p(t),
                                            int(times is not None), L, _p(Z), _p(Ts), C.c_float(max_lambda), _p(w),
                                            _p(w2), _p(ed), _p(eT), _p(tot), C.byref(en), _p(ln)))
    return dict(w=w, w_sqrt=w2, dist=ed, Tdiff=eT, obs_llhood=tot[0], expected_dist=tot[1], N_eff=tot[2],
                Tdiff_colsum=tot[3:].copy(), eps_new=en.value, lambda_new=ln)


def initialize_lambda(obs, poset, lambda_s=1.0, max_lambda=1e6):
    obs, poset = _i32(obs), _i32(poset)
    N, p = obs.shape
    lam = np.zeros(p)
    e = C.c_double(0)
    _chk(lib().ref_initialize_lambda(_p(obs), N, _p(poset), p, C.c_float(lambda_s), C.c_float(max_lambda), _p(lam),
                                     C.byref(e)))
    return lam, e.value


def mt19937_kat():
    """10000th output of a default-seeded std::mt19937 (the C++ standard says 4123659995)."""
    f = lib().ref_mt19937_kat
    f.restype = C.c_uint
    return int(f())


def rexp(n, rate, seed):
    out = np.zeros(n)
    _chk(lib().ref_rexp(n, C.c_double(rate), seed, _p(out)))
    return out


def MCEM_otcbn(ilambda, poset, obs, times=None, max_iter=100, alpha=0.2, weights=None, nrOfSamples=5, max_lambda=1e6,
               lambda_s=1.0, seed=1):
    """legacy OT-CBN `MCEM` (src/mcem.cpp:307-335) with the topological order of R's my.topological.sort"""
    obs, poset, lam = _i32(obs), _i32(poset), _f64(ilambda)
    N, p = obs.shape
    t = np.zeros(N) if times is None else _f64(times)
    wt = np.ones(N) if weights is None else _f64(weights)
    indeg = poset.sum(axis=0).astype(int)
    order, ready = [], [j for j in range(p) if indeg[j] == 0]
    while ready:
        j = ready.pop(0)
        order.append(j)
        for c in np.nonzero(poset[j])[0]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(int(c))
    topo = _i32(np.array(order))
    out_l = np.zeros(p)
    out_ll = C.c_double(0)
    _chk(lib().ref_MCEM_otcbn(_p(lam), _p(poset), _p(obs), _p(t), int(max_iter), C.c_float(alpha), _p(topo), _p(wt),
                              int(nrOfSamples), C.c_float(max_lambda), C.c_float(lambda_s), int(times is not None),
                              int(seed), N, p, _p(out_l), C.byref(out_ll)))
    return dict(lambda_=out_l, llhood=out_ll.value)


def _r_topo(poset):
    """R's my.topological.sort order (Kahn, FIFO) - 1"""
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
    return _i32(np.array(order))


def draw_hidden_vars_samples(N, genotype, time, lambda_, poset, lambda_s=1.0, sampling_time_available=True, seed=1):
    """drawHiddenVarsSamples (src/mcem.cpp:214-248): dict(T=N x p, densities=N)"""
    g, poset, lam = _i32(genotype), _i32(poset), _f64(lambda_)
    p = poset.shape[0]
    T = np.zeros((N, p), order="F")
    dens = np.zeros(N)
    _chk(lib().ref_drawHiddenVarsSamples(int(N), _p(g), C.c_double(time), _p(lam), _p(poset), _p(_r_topo(poset)),
                                         C.c_float(lambda_s), int(bool(sampling_time_available)), int(seed), p,
                                         _p(T), _p(dens)))
    return dict(T=T, densities=dens)


def expected_time_difference(nrOfSamples, genotype, time, lambda_, poset, lambda_s=1.0, sampling_time_available=True,
                             seed=1):
    """expectedTimeDifference (src/mcem.cpp:251-276)"""
    g, poset, lam = _i32(genotype), _i32(poset), _f64(lambda_)
    p = poset.shape[0]
    out = np.zeros(p)
    _chk(lib().ref_expectedTimeDifference(int(nrOfSamples), _p(g), C.c_double(time), _p(lam), _p(poset),
                                          _p(_r_topo(poset)), C.c_float(lambda_s), int(bool(sampling_time_available)),
                                          int(seed), p, _p(out)))
    return out


def otcbn_proposal_density(genotype, times, lambda_, poset, T):
    """log proposal density of drawHiddenVarsSample recomputed from its output T (N x p) and the sampling times"""
    g, poset, lam, T, t = _i32(genotype), _i32(poset), _f64(lambda_), _f64(T), _f64(times)
    N, p = T.shape
    dens = np.zeros(N)
    _chk(lib().ref_otcbn_proposal_density(N, _p(g), _p(t), _p(lam), _p(poset), _p(_r_topo(poset)), p, _p(T), _p(dens)))
    return dens


def loglike_importance_sampling(poset, lambda_, obs_events, sampling_times, nrOfSamples, weights=None, lambda_s=1.0,
                                sampling_times_available=True, seed=1):
    """loglike_importance_sampling (R/mcem_genotype_probabilities.r:3-23): one drawHiddenVarsSamples call per
    observation (seed + i stands in for the advancing global R generator), cbn_density_ (:3-6), logsumexp."""
    obs, lam = _i32(np.atleast_2d(obs_events)), _f64(lambda_)
    N = obs.shape[0]
    wt = np.ones(N) if weights is None else _f64(weights)
    probs = np.zeros(N)
    for i in range(N):
        r = draw_hidden_vars_samples(nrOfSamples, obs[i], float(sampling_times[i]) if sampling_times is not None else 0.0,
                                     lam, poset, lambda_s, sampling_times_available, seed + i)
        values = (np.log(lam)[None, :] - r["T"] * lam[None, :]).sum(axis=1) - r["densities"]  # sum dexp(x, rates, log=TRUE)
        m = values.max()
        probs[i] = wt[i] * (m + np.log(np.exp(values - m).sum()) - np.log(nrOfSamples))
    return dict(approx_loglike=float(probs.sum()), probs=probs)


def adaptive_simulated_annealing(poset, obs, L, times=None, lambda_s=1.0, weights=None, sampling="forward", max_iter=100,
                                 update_step_size=20, tol=0.001, max_lambda=1e6, T0=50.0, adap_rate=0.3,
                                 acceptance_rate=None, step_size=None, max_iter_asa=10000, neighborhood_dist=1,
                                 adaptive=True, outdir="", thrds=1, verbose=False, seed=1, score_mode=0):
    """`_adaptive_simulated_annealing` (src/asa.cpp:596-665) with the defaults of the R wrapper
    (R/adaptive_simulated_annealing.R:62-113).  score_mode=1: deterministic pseudo score (test seam)."""
    obs, poset = _i32(obs), _i32(poset)
    N, p = obs.shape
    t = np.zeros(N) if times is None else _f64(times)
    w = np.ones(N) if weights is None else _f64(weights)
    if acceptance_rate is None:
        acceptance_rate = 1.0 / p
    if step_size is None:
        step_size = max(1, max_iter_asa // 10)
    lam, eps, ll = np.zeros(p), C.c_double(0), C.c_double(0)
    out = np.zeros((p, p), np.int32, order="F")
    stats = np.zeros(6, np.int32)
    _chk(lib().ref_adaptive_simulated_annealing(
        _p(poset), _p(obs), _p(t), C.c_float(lambda_s), _p(w), C.c_uint(L), sampling.encode(), C.c_uint(max_iter),
        C.c_uint(update_step_size), C.c_double(tol), C.c_float(max_lambda), C.c_float(T0), C.c_float(adap_rate),
        C.c_float(acceptance_rate), int(step_size), C.c_uint(max_iter_asa), C.c_uint(neighborhood_dist), int(adaptive),
        outdir.encode(), int(times is not None), int(thrds), int(verbose), int(seed), N, p, int(score_mode), _p(lam),
        C.byref(eps), _p(out), C.byref(ll), _p(stats)))
    return dict(lambda_=lam, eps=eps.value, poset=out, llhood=ll.value, n_accept=int(stats[0]),
                moves=stats[1:5].tolist(), n_scored=int(stats[5]))


def draw_sample(genotype, poset, move, q_prob, seed=1):
    """`_draw_sample` (src/add_remove.cpp:244-284)."""
    g, poset, q = _i32(genotype), _i32(poset), _f64(q_prob)
    p = poset.shape[0]
    out = np.zeros(p, np.int32)
    qc = C.c_double(0)
    _chk(lib().ref_draw_sample(_p(g), _p(poset), p, C.c_uint(move), _p(q), int(seed), _p(out), C.byref(qc)))
    return out, qc.value
