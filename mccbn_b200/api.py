"""Host-side mirror of the reference's R interface for the hot path (same names, argument meaning, defaults and
error behaviour as the R wrappers), on top of the C ABI.  R is not installed in this image, so this Python layer
plays the role of R/mcem_hcbn.R, R/loglikelihood.R and R/adaptive_simulated_annealing.R: argument munging,
defaults, then one call into the shared library -- exactly what the R wrappers do before `.Call`.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import EmOptions, ErrBuf, MccbnError, f64, i32, ptr

SAMPLING = ("forward", "add-remove", "backward", "bernoulli", "pool")


def _match_sampling(s):
    if s not in SAMPLING:
        raise ValueError(f"'arg' should be one of {SAMPLING}")  # match.arg
    return s


class Context:
    """Owns a CUDA stream (and, when world > 1, an NCCL communicator) on one device."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        e = ErrBuf()
        e.check(capi.lib().mccbn_ctx_create(C.byref(self._h), int(device), *e.args))
        self.device = device
        self.rank, self.world = 0, 1

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        capi.lib().mccbn_ctx_set_stream(self._h, C.c_void_p(cuda_stream))

    def init_comm(self, rank, world, unique_id):
        e = ErrBuf()
        buf = (C.c_ubyte * capi.NCCL_ID_BYTES).from_buffer_copy(bytes(unique_id))
        e.check(capi.lib().mccbn_ctx_init_comm(self._h, int(rank), int(world), buf, *e.args))
        self.rank, self.world = rank, world

    def peer_handle(self):
        """CUDA IPC handle of this rank's statistic mailbox (allocated on first use)."""
        e = ErrBuf()
        buf = (C.c_ubyte * capi.IPC_HANDLE_BYTES)()
        e.check(capi.lib().mccbn_ctx_peer_handle(self._h, buf, *e.args))
        return bytes(buf)

    def init_peers(self, rank, world, handles):
        """`handles`: the world IPC handles in rank order (bytes of length world * 64)."""
        handles = bytes(handles)
        if len(handles) != world * capi.IPC_HANDLE_BYTES:
            raise ValueError("expected world * 64 bytes of IPC handles")
        e = ErrBuf()
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        e.check(capi.lib().mccbn_ctx_init_peers(self._h, int(rank), int(world), buf, *e.args))
        self.rank, self.world = rank, world

    def kernel_launches(self, reset=False):
        return int(capi.lib().mccbn_kernel_launches(self._h, int(reset)))

    def jit_launches(self, reset=False):
        """launches of NVRTC poset-specialised kernels among kernel_launches()"""
        return int(capi.lib().mccbn_jit_launches(self._h, int(reset)))

    def close(self):
        if self._h:
            capi.lib().mccbn_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    e = ErrBuf()
    buf = (C.c_ubyte * capi.NCCL_ID_BYTES)()
    e.check(capi.lib().mccbn_nccl_unique_id(buf, *e.args))
    return bytes(buf)


def set_gpus(n):
    """In-process multi-GPU: calls made without a context shard their observations over the first n devices
    (mccbn_set_gpus; the role of the reference's `thrds`).  n <= 1 turns it off."""
    e = ErrBuf()
    e.check(capi.lib().mccbn_set_gpus(int(n), *e.args))


def get_gpus():
    return int(capi.lib().mccbn_get_gpus())


def device_count():
    return int(capi.lib().mccbn_device_count())


def _ctx(ctx):
    return ctx.handle if ctx is not None else None


def _options(obs_offset=0, N_global=0, trace=None, n_iter=None, precision=0, pool_resample=0, stream_id=0):
    o = EmOptions()
    o.obs_offset = int(obs_offset)
    o.N_global = int(N_global)
    o.trace = trace.ctypes.data if trace is not None else None
    o.n_iter = C.addressof(n_iter) if n_iter is not None else None
    o.precision = int(precision)
    o.pool_resample = int(pool_resample)
    o.stream_id = int(stream_id)
    return o


def MCEM_hcbn(lambda_, poset, obs, lambda_s=1.0, L=None, eps=None, sampling="forward", times=None, weights=None,
              max_iter=100, update_step_size=20, tol=0.001, max_lambda=1e6, neighborhood_dist=1, thrds=1,
              verbose=False, seed=None, ctx=None, obs_offset=0, N_global=0, precision=0, stream_id=0,
              return_trace=False, pool_resample=False):
    """MCEM.hcbn (R/mcem_hcbn.R:51-90) -> `_MCEM_hcbn` (src/mcem_hcbn.cpp:804-864).

    Returns dict(lambda_=..., eps=..., llhood=...) like the R list(lambda, eps, llhood)."""
    sampling = _match_sampling(sampling)
    obs, poset, lam = i32(obs), i32(poset), f64(lambda_)
    N, p = obs.shape
    if L is None:
        raise TypeError('argument "L" is missing, with no default')
    if times is None:
        t = np.zeros(N)
        avail = False
    else:
        t = f64(times)
        avail = True
        lambda_s = 1.0 / float(np.mean(t))  # R/mcem_hcbn.R:68-71
    wt = np.ones(N) if weights is None else f64(weights)
    if update_step_size > max_iter:
        update_step_size = int(max_iter / 5)  # :75-76
    if seed is None:
        seed = int(np.random.randint(1, 30001))  # sample.int(3e4, 1)
    if eps is None:
        eps = float(np.random.default_rng(seed).uniform(0.01, 0.3))  # set.seed(seed); runif(1, 0.01, 0.3)
    out_l = np.zeros(p)
    out_e, out_ll = C.c_double(0), C.c_double(0)
    trace = np.zeros((max(max_iter, 1), p + 2)) if return_trace else None
    nit = C.c_int(0)
    opt = _options(obs_offset, N_global, trace, nit, precision, int(pool_resample), stream_id)
    e = ErrBuf()
    e.check(capi.lib().mccbn_MCEM_hcbn(
        _ctx(ctx), ptr(lam), ptr(poset), ptr(obs), ptr(t), C.c_float(lambda_s), C.c_double(eps), ptr(wt),
        C.c_uint(int(L)), sampling.encode(), C.c_uint(int(max_iter)), C.c_uint(int(update_step_size)),
        C.c_double(tol), C.c_float(max_lambda), C.c_uint(int(neighborhood_dist)), int(avail), int(thrds),
        int(bool(verbose)), int(seed), N, p, ptr(out_l), C.byref(out_e), C.byref(out_ll), C.byref(opt), *e.args))
    res = dict(lambda_=out_l, eps=out_e.value, llhood=out_ll.value, n_iter=nit.value)
    if return_trace:
        res["trace"] = trace[:nit.value].copy()
    return res


def obs_loglikelihood(obs, poset, lambda_, eps, weights=None, times=None, L=None, sampling="forward",
                      neighborhood_dist=1, lambda_s=1.0, thrds=1, seed=None, ctx=None, obs_offset=0, N_global=0,
                      precision=0, stream_id=0, pool_resample=False):
    """obs.loglikelihood (R/loglikelihood.R:25-57) -> `_obs_log_likelihood` (src/mcem_hcbn.cpp:762-797)."""
    sampling = _match_sampling(sampling)
    obs, poset, lam = i32(obs), i32(poset), f64(lambda_)
    N, p = obs.shape
    wt = np.ones(N) if weights is None else f64(weights)
    if times is None:
        t = np.zeros(N)
        avail = False
    else:
        t = f64(times)
        if t.size != N:
            raise ValueError(f"A vector of length {N} is expected")
        lambda_s = 1.0 / float(np.mean(t))  # R/loglikelihood.R:44-46
        avail = True
    if seed is None:
        seed = int(np.random.randint(1, 30001))
    out = C.c_double(0)
    opt = _options(obs_offset=obs_offset, N_global=N_global, precision=precision, stream_id=stream_id,
                   pool_resample=int(pool_resample))
    e = ErrBuf()
    e.check(capi.lib().mccbn_obs_log_likelihood(
        _ctx(ctx), ptr(obs), ptr(poset), ptr(lam), C.c_double(eps), ptr(wt), ptr(t), C.c_uint(int(L)),
        sampling.encode(), C.c_uint(int(neighborhood_dist)), C.c_float(lambda_s), int(avail), int(thrds), int(seed),
        N, p, C.byref(out), C.byref(opt), *e.args))
    return out.value


def importance_weight(genotype, L, poset, lambda_, eps, time=None, sampling="forward", neighborhood_dist=1,
                      lambda_s=1.0, thrds=1, seed=None, ctx=None, obs_offset=0, N_global=0, precision=0, stream_id=0,
                      pool_resample=False):
    """importance.weight (R/mcem_hcbn.R:138-195): matrix input -> `_importance_weight` (src/mcem_hcbn.cpp:914-1010),
    vector input -> `_importance_weight_genotype` (:866-912)."""
    sampling = _match_sampling(sampling)
    poset, lam = i32(poset), f64(lambda_)
    p = poset.shape[0]
    g = np.asarray(genotype)
    if seed is None:
        seed = int(np.random.randint(1, 30001))
    opt = _options(obs_offset=obs_offset, N_global=N_global, precision=precision, stream_id=stream_id,
                   pool_resample=int(pool_resample))
    e = ErrBuf()
    if g.ndim == 2:
        obs = i32(g)
        N = obs.shape[0]
        if time is None:
            t, avail = np.zeros(N), False
        else:
            t, avail = f64(time), True
            if t.size != N:
                raise ValueError(f"A vector of length {N} is expected")
        w, w2, ed = np.zeros(N), np.zeros(N), np.zeros(N)
        Le = np.zeros(N, np.int32)
        eT = np.zeros((N, p), order="F")
        e.check(capi.lib().mccbn_importance_weight(
            _ctx(ctx), ptr(obs), C.c_uint(int(L)), ptr(poset), ptr(lam), C.c_double(eps), ptr(t), sampling.encode(),
            C.c_uint(int(neighborhood_dist)), C.c_float(lambda_s), int(avail), int(thrds), int(seed), N, p, ptr(w),
            ptr(w2), ptr(Le), ptr(ed), ptr(eT), C.byref(opt), *e.args))
        res = dict(w=w, w_sqrt=w2, dist=ed, Tdiff=eT)
        if sampling in ("backward", "bernoulli"):
            res["L_eff"] = Le
        return res
    gv = i32(g.ravel())
    t, avail = (0.0, False) if time is None else (float(time), True)
    w = np.zeros(int(L))
    d = np.zeros(int(L), np.int32)
    T = np.zeros(int(L) * p)
    Lout = C.c_int(0)
    e.check(capi.lib().mccbn_importance_weight_genotype(
        _ctx(ctx), ptr(gv), C.c_uint(int(L)), ptr(poset), ptr(lam), C.c_double(eps), C.c_double(t), sampling.encode(),
        C.c_uint(int(neighborhood_dist)), C.c_float(lambda_s), int(avail), int(seed), p, C.byref(Lout), ptr(w), ptr(d),
        ptr(T), C.byref(opt), *e.args))
    n = Lout.value
    return dict(w=w[:n], dist=d[:n], Tdiff=T[:n * p].reshape((n, p), order="F"))


def adaptive_simulated_annealing(poset, obs, times=None, lambda_s=1.0, weights=None, L=None, sampling="forward",
                                 max_iter=100, update_step_size=20, tol=0.001, max_lambda_val=1e6, T0=50.0,
                                 adap_rate=0.3, acceptance_rate=None, step_size=None, max_iter_asa=10000,
                                 neighborhood_dist=1, adaptive=True, outdir=None, thrds=1, verbose=False, seed=None,
                                 ctx=None, _test_structural_score=False):
    """adaptive.simulated.annealing (R/adaptive_simulated_annealing.R:62-113) -> `_adaptive_simulated_annealing`
    (src/asa.cpp:596-665).  Returns dict(lambda_, eps, poset, llhood)."""
    import os
    sampling = _match_sampling(sampling)
    obs, poset = i32(obs), i32(poset)
    N, p = obs.shape
    if times is None:
        t, avail = np.zeros(N), False
    else:
        t, avail = f64(times), True  # NB: the ASA wrapper does NOT override lambda.s (:79-84)
    wt = np.ones(N) if weights is None else f64(weights)
    if update_step_size > max_iter:
        update_step_size = int(max_iter / 5)
    if acceptance_rate is None:
        acceptance_rate = 1.0 / p
    if step_size is None:
        step_size = max_iter_asa // 10
    if outdir is None:
        outdir = os.path.join(os.getcwd(), "")
    if seed is None:
        seed = int(np.random.randint(1, 30001))
    out_l = np.zeros(p)
    out_e, out_ll = C.c_double(0), C.c_double(0)
    out_p = np.zeros((p, p), np.int32, order="F")
    e = ErrBuf()
    entry = (capi.lib().mccbn_test_asa_structural_score if _test_structural_score  # test-only seam, see the header
             else capi.lib().mccbn_adaptive_simulated_annealing)
    e.check(entry(
        _ctx(ctx), ptr(poset), ptr(obs), ptr(t), C.c_float(lambda_s), ptr(wt), C.c_uint(int(L)), sampling.encode(),
        C.c_uint(int(max_iter)), C.c_uint(int(update_step_size)), C.c_double(tol), C.c_float(max_lambda_val),
        C.c_float(T0), C.c_float(adap_rate), C.c_float(acceptance_rate), int(step_size), C.c_uint(int(max_iter_asa)),
        C.c_uint(int(neighborhood_dist)), int(bool(adaptive)), str(outdir).encode(), int(avail), int(thrds),
        int(bool(verbose)), int(seed), N, p, ptr(out_l), C.byref(out_e), ptr(out_p), C.byref(out_ll), *e.args))
    return dict(lambda_=out_l, eps=out_e.value, poset=out_p, llhood=out_ll.value)


def _initialize_lambda_otcbn(obs, average_sampling_times, poset):
    """initialize_lambda of the OT-CBN wrapper (R/mcem_param_est.r:2-36): direct parents only."""
    N, p = obs.shape
    if p == 1:
        return np.array([1.0 / average_sampling_times])
    theta = np.zeros(p)
    for i in range(p):
        parents = np.nonzero(poset[:, i] == 1)[0]
        if parents.size == 0:
            theta[i] = obs[:, i].sum() / N
        else:
            idx = np.all(obs[:, parents] == 1, axis=1)
            theta[i] = 0.000001 if idx.sum() == 0 else obs[idx, i].sum() / idx.sum()
    return -np.log(np.maximum(1 - theta, 0.00001)) / average_sampling_times


def estimate_mutation_rates(poset, genotypes, sampling_times=None, weights=None, max_iter=100, zeta=0.2, ilambda=None,
                            nrOfSamples=5, verbose=False, maxLambdaValue=1e6, lambda_s=1.0, seed=None, ctx=None):
    """estimate_mutation_rates / MCEM (R/mcem_param_est.r:40-72) -> `MCEM` (src/mcem.cpp:307-335), the legacy
    noise-free OT-CBN fit.  Returns dict(lambda_, llhood)."""
    obs, poset = i32(genotypes), i32(poset)
    N, p = obs.shape
    avail = sampling_times is not None
    if not avail:
        average_sampling_times = lambda_s  # sic (R/mcem_param_est.r:50-52)
        t = np.zeros(N)
    else:
        t = f64(sampling_times)
        average_sampling_times = float(np.mean(t))
    wt = np.ones(N) if weights is None else f64(weights)
    if ilambda is None:
        ilambda = _initialize_lambda_otcbn(obs, average_sampling_times, poset)
        ilambda[ilambda == 0] = 1e-7
    lam = f64(ilambda)
    from .synth import topological_sort
    topo = i32(np.array(topological_sort(poset)))
    if seed is None:
        seed = int(np.random.randint(1, 30001))
    out_l = np.zeros(p)
    out_ll = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_MCEM(_ctx(ctx), ptr(lam), ptr(poset), ptr(obs), ptr(t), int(max_iter), C.c_float(zeta),
                                  ptr(topo), ptr(wt), int(nrOfSamples), int(bool(verbose)), C.c_float(maxLambdaValue),
                                  C.c_float(lambda_s), int(avail), int(seed), N, p, ptr(out_l), C.byref(out_ll),
                                  *e.args))
    return dict(lambda_=out_l, llhood=out_ll.value)


# ---- OT-CBN callers of the conditional proposal (R/mcem_genotype_probabilities.r) ---------------------------------
def _topo(poset):
    """my.topological.sort(poset) - 1; the library derives its own order and reports cycles itself"""
    from .synth import topological_sort
    try:
        return i32(np.array(topological_sort(poset)))
    except ValueError:
        return i32(np.arange(poset.shape[0]))


def draw_hidden_vars_samples(N, genotype, time, lambda_, poset, lambda_s=1.0, sampling_time_available=True, seed=1,
                             ctx=None):
    """.Call("drawHiddenVarsSamples", ...) (src/mcem.cpp:214-248): dict(T=N x p, densities=N, time=N)."""
    g, poset, lam = i32(genotype), i32(poset), f64(lambda_)
    p = poset.shape[0]
    T = np.zeros((int(N), p), order="F")
    dens = np.zeros(int(N))
    ts = np.zeros(int(N))
    e = ErrBuf()
    e.check(capi.lib().mccbn_drawHiddenVarsSamples(_ctx(ctx), int(N), ptr(g), C.c_double(float(time)), ptr(lam),
                                                   ptr(poset), ptr(_topo(poset)), C.c_float(lambda_s),
                                                   int(bool(sampling_time_available)), int(seed), p, ptr(T), ptr(dens),
                                                   ptr(ts), *e.args))
    return dict(T=T, densities=dens, time=ts)


def expected_time_difference(nrOfSamples, genotype, time, lambda_, poset, lambda_s=1.0, sampling_time_available=True,
                             seed=1, ctx=None):
    """.Call("expectedTimeDifference", ...) (src/mcem.cpp:251-276)."""
    g, poset, lam = i32(genotype), i32(poset), f64(lambda_)
    p = poset.shape[0]
    out = np.zeros(p)
    e = ErrBuf()
    e.check(capi.lib().mccbn_expectedTimeDifference(_ctx(ctx), int(nrOfSamples), ptr(g), C.c_double(float(time)),
                                                    ptr(lam), ptr(poset), ptr(_topo(poset)), C.c_float(lambda_s),
                                                    int(bool(sampling_time_available)), int(seed), p, ptr(out),
                                                    *e.args))
    return out


def loglike_importance_sampling(poset, lambda_, obs_events, sampling_times, nrOfSamples, weights=None, lambda_s=1.0,
                                sampling_times_available=True, seed=1, ctx=None):
    """loglike_importance_sampling (R/mcem_genotype_probabilities.r:10-23), all observations in one launch:
    dict(approx_loglike, probs)."""
    obs, poset, lam = i32(np.atleast_2d(obs_events)), i32(poset), f64(lambda_)
    N, p = obs.shape
    t = np.zeros(N) if sampling_times is None else f64(sampling_times)
    wt = np.ones(N) if weights is None else f64(weights)
    probs = np.zeros(N)
    ll = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_loglike_importance_sampling(_ctx(ctx), ptr(poset), ptr(lam), ptr(obs), ptr(t),
                                                         int(nrOfSamples), ptr(wt), C.c_float(lambda_s),
                                                         int(bool(sampling_times_available)), int(seed), N, p,
                                                         C.byref(ll), ptr(probs), *e.args))
    return dict(approx_loglike=ll.value, probs=probs)


def genotype_probability_fast(poset, lambda_, genotype, time, nrOfSamples=100, lambda_s=1.0,
                              sampling_time_available=True, seed=1, ctx=None):
    """genotype_probability_fast (R/mcem_genotype_probabilities.r:54-70): 0 for a genotype the poset excludes."""
    g = i32(genotype)
    if not compatible_observations(g[None, :], poset, ctx=ctx)["compatible"][0]:
        return 0.0
    r = loglike_importance_sampling(poset, lambda_, g[None, :], [time], nrOfSamples, None, lambda_s,
                                    sampling_time_available, seed, ctx)
    return float(np.exp(r["approx_loglike"]))


# ---- indexing and supplied-randomness seams -----------------------------------------------------------------------
def flatten_poset(poset):
    poset = i32(poset)
    p = poset.shape[0]
    topo = np.zeros(p, np.int32)
    mask = np.zeros(p, np.uint64)
    e = ErrBuf()
    e.check(capi.lib().mccbn_flatten_poset(ptr(poset), p, ptr(topo), ptr(mask), *e.args))
    return topo, mask


def compatible_observations(obs, poset, ctx=None):
    obs, poset = i32(obs), i32(poset)
    N, p = obs.shape
    out = np.zeros(N, np.int32)
    nc, ni = C.c_int(0), C.c_int(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_compatible_observations(_ctx(ctx), ptr(obs), ptr(poset), N, p, ptr(out), C.byref(nc),
                                                     C.byref(ni), *e.args))
    return dict(compatible=out, num_compatible=nc.value, num_incompatible_events=ni.value)


def neighbors(p, k, genotype):
    g = i32(genotype)
    n = C.c_int(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_neighbors(p, k, ptr(g), None, None, C.byref(n), *e.args))
    out = np.zeros((n.value, p), np.int32, order="F")
    ring = np.zeros(n.value, np.int32)
    e.check(capi.lib().mccbn_neighbors(p, k, ptr(g), ptr(out), ptr(ring), C.byref(n), *e.args))
    return out, ring


def jit_launches(ctx=None, reset=False):
    """Launches of NVRTC poset-specialised kernels by `ctx` (None: the default context) since the last reset."""
    return int(capi.lib().mccbn_jit_launches(_ctx(ctx), int(reset)))


def kernel_launches(ctx=None, reset=False):
    return int(capi.lib().mccbn_kernel_launches(_ctx(ctx), int(reset)))


def set_jit(mode, ctx=None):
    """-1 auto, 0 never, 1 always use the NVRTC poset-specialised forward kernel (csrc/jit.cuh)."""
    capi.lib().mccbn_set_jit(_ctx(ctx), int(mode))


def jit_compile_forward(poset, times_given=False):
    """Host-only: generate + compile the specialised kernel of a poset; returns dict(cubin_bytes, ms, log, source)."""
    poset = i32(poset)
    n, ms = C.c_size_t(0), C.c_double(0)
    log = C.create_string_buffer(1 << 16)
    src = C.create_string_buffer(1 << 18)
    e = ErrBuf()
    e.check(capi.lib().mccbn_jit_compile_forward(ptr(poset), poset.shape[0], int(times_given), C.byref(n), C.byref(ms),
                                                 log, C.c_size_t(len(log)), src, C.c_size_t(len(src)), *e.args))
    return dict(cubin_bytes=n.value, ms=ms.value, log=log.value.decode(), source=src.value.decode())


def sample_genotypes(N, poset, lambda_, lambda_s=1.0, times=None, seed=1, ctx=None):
    """`_sample_genotypes` (src/mcem_hcbn.cpp:1012-1049): N draws from the generative model on the device.
    Returns dict(samples N x p int32, Tdiff N x p, sampling_time N)."""
    poset, lam = i32(poset), f64(lambda_)
    p = poset.shape[0]
    ts = np.zeros(N) if times is None else f64(times).copy()
    X = np.zeros((N, p), np.int32, order="F")
    Z = np.zeros((N, p), order="F")
    e = ErrBuf()
    e.check(capi.lib().mccbn_sample_genotypes(_ctx(ctx), C.c_uint(int(N)), ptr(poset), ptr(lam), ptr(ts),
                                              C.c_float(lambda_s), int(times is not None), int(seed), p, ptr(X), ptr(Z),
                                              *e.args))
    return dict(samples=X, Tdiff=Z, sampling_time=ts)


def sample_times(N, poset, lambda_, seed=1, ctx=None):
    """`_sample_times` (src/mcem_hcbn.cpp:1052-1083): returns dict(mutation_times N x p, Tdiff N x p)."""
    poset, lam = i32(poset), f64(lambda_)
    p = poset.shape[0]
    T = np.zeros((N, p), order="F")
    Z = np.zeros((N, p), order="F")
    e = ErrBuf()
    e.check(capi.lib().mccbn_sample_times(_ctx(ctx), C.c_uint(int(N)), ptr(poset), ptr(lam), int(seed), p, ptr(T), ptr(Z),
                                          *e.args))
    return dict(mutation_times=T, Tdiff=Z)


def generate_genotypes(T_events_sum, poset, times=None, lambda_s=1.0, seed=1, ctx=None):
    """`_generate_genotypes` (src/mcem_hcbn.cpp:1086-1120): returns dict(samples N x p int32, sampling_time N)."""
    T = f64(T_events_sum)
    poset = i32(poset)
    N, p = T.shape
    ts = np.zeros(N) if times is None else f64(times).copy()
    X = np.zeros((N, p), np.int32, order="F")
    e = ErrBuf()
    e.check(capi.lib().mccbn_generate_genotypes(_ctx(ctx), ptr(T), ptr(poset), ptr(ts), C.c_float(lambda_s),
                                                int(times is not None), int(seed), N, p, ptr(X), *e.args))
    return dict(samples=X, sampling_time=ts)


def scale_path_to_mutation(poset, lambda_):
    """`_scale_path_to_mutation` (src/add_remove.cpp:218-242)."""
    poset, lam = i32(poset), f64(lambda_)
    out = np.zeros(poset.shape[0])
    e = ErrBuf()
    e.check(capi.lib().mccbn_scale_path_to_mutation(ptr(poset), poset.shape[0], ptr(lam), ptr(out), *e.args))
    return out


def cbn_density_log(time, lambda_):
    """`_cbn_density_log` (src/add_remove.cpp:323-335)."""
    t, lam = f64(time), f64(lambda_)
    N, p = t.shape
    out = np.zeros(N)
    e = ErrBuf()
    e.check(capi.lib().mccbn_cbn_density_log(ptr(t), N, p, ptr(lam), ptr(out), *e.args))
    return out


def complete_log_likelihood(lambda_, eps, Tdiff, dist, weights):
    """`_complete_log_likelihood` (src/mcem_hcbn.cpp:738-760)."""
    lam, T, d, w = f64(lambda_), f64(Tdiff), f64(dist), f64(weights)
    N, p = T.shape
    out = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_complete_log_likelihood(ptr(lam), p, C.c_double(eps), ptr(T), N, ptr(d), ptr(w),
                                                     C.byref(out), *e.args))
    return out.value


def draw_sample(genotype, poset, move, q_prob, seed=1):
    """`_draw_sample` (src/add_remove.cpp:244-284): returns (sample, q_choice)."""
    g, poset, q = i32(genotype), i32(poset), f64(q_prob)
    p = poset.shape[0]
    out = np.zeros(p, np.int32)
    qc = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_draw_sample(ptr(g), ptr(poset), p, C.c_uint(int(move)), ptr(q), int(seed), ptr(out),
                                         C.byref(qc), *e.args))
    return out, qc.value


def coin_tossing(eps, L, genotype, seed=1, ctx=None):
    """`_coin_tossing` (src/backward_sampling.cpp:105-123): L x p noisy copies of `genotype`."""
    g = i32(genotype).ravel()
    out = np.zeros((int(L), g.size), np.int32, order="F")
    e = ErrBuf()
    e.check(capi.lib().mccbn_coin_tossing(_ctx(ctx), C.c_double(eps), C.c_uint(int(L)), ptr(g), g.size, int(seed),
                                          ptr(out), *e.args))
    return out


def generate_mutation_times(obs, poset, lambda_, times=None, lambda_s=1.0, seed=1, ctx=None):
    """`_generate_mutation_times` (src/add_remove.cpp:286-321): returns dict(Tdiff, proposal_density, sampling_time)."""
    obs, poset, lam = i32(obs), i32(poset), f64(lambda_)
    N, p = obs.shape
    ts = np.zeros(N) if times is None else f64(times).copy()
    T = np.zeros((N, p), order="F")
    dens = np.zeros(N)
    e = ErrBuf()
    e.check(capi.lib().mccbn_generate_mutation_times(_ctx(ctx), ptr(obs), ptr(poset), ptr(lam), ptr(ts),
                                                     C.c_float(lambda_s), int(times is not None), int(seed), N, p,
                                                     ptr(T), ptr(dens), *e.args))
    return dict(Tdiff=T, proposal_density=dens, sampling_time=ts)


def jit_compile_conditional(poset, mode=0, times_given=False):
    """Host-only: generate + compile the specialised conditional-proposal kernel of a poset (mode 0 backward,
    1 bernoulli, 2 OT-CBN, 3 add-remove); returns dict(cubin_bytes, ms, log, source)."""
    poset = i32(poset)
    n, ms = C.c_size_t(0), C.c_double(0)
    log = C.create_string_buffer(1 << 16)
    src = C.create_string_buffer(1 << 18)
    e = ErrBuf()
    e.check(capi.lib().mccbn_jit_compile_conditional(ptr(poset), poset.shape[0], int(times_given), int(mode), C.byref(n),
                                                     C.byref(ms), log, C.c_size_t(len(log)), src, C.c_size_t(len(src)),
                                                     *e.args))
    return dict(cubin_bytes=n.value, ms=ms.value, log=log.value.decode(), source=src.value.decode())


POSET_OPS = dict(add_edge=0, remove_edge=1, add_relation=2, remove_relation=3, transitive_reduction=4, swap_node=5)


def poset_edit(poset, op, u=0, v=0):
    """Host-only seam for the annealer's poset edit operations (src/model_impl.cpp:145-316)."""
    poset = i32(poset)
    p = poset.shape[0]
    out = np.zeros((p, p), np.int32, order="F")
    flags = C.c_int(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_poset_edit(ptr(poset), p, POSET_OPS[op], int(u), int(v), ptr(out), C.byref(flags),
                                        *e.args))
    return dict(poset=out, added=bool(flags.value & 1), cycle=bool(flags.value & 2),
                reduction_flag=bool(flags.value & 4))


def initialize_lambda(obs, poset, lambda_s=1.0, max_lambda=1e6, ctx=None):
    obs, poset = i32(obs), i32(poset)
    N, p = obs.shape
    lam = np.zeros(p)
    eps = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_initialize_lambda(_ctx(ctx), ptr(obs), ptr(poset), N, p, C.c_float(lambda_s),
                                               C.c_float(max_lambda), ptr(lam), C.byref(eps), *e.args))
    return lam, eps.value


def forward_from_times(genotype, poset, eps, Z, T_sampling, ctx=None):
    g, poset, Z, Ts = i32(genotype), i32(poset), f64(Z), f64(T_sampling)
    L, p = Z.shape
    Tsum = np.zeros((L, p), order="F")
    X = np.zeros((L, p), np.int32, order="F")
    d = np.zeros(L, np.int32)
    w = np.zeros(L)
    e = ErrBuf()
    e.check(capi.lib().mccbn_forward_from_times(_ctx(ctx), ptr(g), L, ptr(poset), p, C.c_double(eps), ptr(Z), ptr(Ts),
                                                ptr(Tsum), ptr(X), ptr(d), ptr(w), *e.args))
    return dict(T_sum=Tsum, X=X, dist=d, w=w)


def estep_forward_from_times(obs, poset, lambda_, eps, Z, T_sampling, weights=None, times=None, max_lambda=1e6,
                             ctx=None):
    obs, poset, lam, Z, Ts = i32(obs), i32(poset), f64(lambda_), f64(Z), f64(T_sampling)
    N, p = obs.shape
    L = Z.shape[0]
    wt = np.ones(N) if weights is None else f64(weights)
    t = np.zeros(N) if times is None else f64(times)
    w, w2, ed = np.zeros(N), np.zeros(N), np.zeros(N)
    eT = np.zeros((N, p), order="F")
    tot = np.zeros(3 + p)
    en = C.c_double(0)
    ln = np.zeros(p)
    e = ErrBuf()
    e.check(capi.lib().mccbn_estep_forward_from_times(
        _ctx(ctx), ptr(obs), N, ptr(poset), p, ptr(lam), C.c_double(eps), ptr(wt), ptr(t), int(times is not None), L,
        ptr(Z), ptr(Ts), C.c_float(max_lambda), ptr(w), ptr(w2), ptr(ed), ptr(eT), ptr(tot), C.byref(en), ptr(ln),
        *e.args))
    return dict(w=w, w_sqrt=w2, dist=ed, Tdiff=eT, obs_llhood=tot[0], expected_dist=tot[1], N_eff=tot[2],
                Tdiff_colsum=tot[3:].copy(), eps_new=en.value, lambda_new=ln)


def pool_from_times(obs, poset, eps, T_pool, Tdiff_pool, T_sampling=None, times=None, weights=None, ctx=None):
    """Pool scheme from a supplied pool (the supplied-randomness seam of the `pool` branch, src/mcem_hcbn.cpp:460-491)."""
    obs, poset = i32(obs), i32(poset)
    N, p = obs.shape
    Tp, Zp = f64(T_pool), f64(Tdiff_pool)
    K = Tp.shape[0]
    ts = None if T_sampling is None else f64(T_sampling)
    t = np.zeros(N) if times is None else f64(times)
    wt = np.ones(N) if weights is None else f64(weights)
    X = np.zeros((K, p), np.int32, order="F")
    d = np.zeros((K, N), np.int32, order="F")
    q, ed = np.zeros(N), np.zeros(N)
    eT = np.zeros((N, p), order="F")
    ll = C.c_double(0)
    e = ErrBuf()
    e.check(capi.lib().mccbn_pool_from_times(
        _ctx(ctx), ptr(obs), N, ptr(poset), p, C.c_double(eps), ptr(wt), ptr(t), int(times is not None), K, ptr(Tp),
        ptr(Zp), ptr(ts), ptr(X), ptr(d), ptr(q), ptr(ed), ptr(eT), C.byref(ll), *e.args))
    return dict(X_pool=X, d_pool=d, q_sum=q, dist=ed, Tdiff=eT, llhood=ll.value)


def conditional_from_uniforms(X, poset, lambda_, u, T_sampling, ctx=None):
    X, poset, lam, u, Ts = i32(X), i32(poset), f64(lambda_), f64(u), f64(T_sampling)
    L, p = X.shape
    T = np.zeros((L, p), order="F")
    lq, lpz = np.zeros(L), np.zeros(L)
    inc = np.zeros(L, np.int32)
    e = ErrBuf()
    e.check(capi.lib().mccbn_conditional_from_uniforms(_ctx(ctx), ptr(X), L, ptr(poset), p, ptr(lam), ptr(u), ptr(Ts),
                                                       ptr(T), ptr(lq), ptr(lpz), ptr(inc), *e.args))
    return dict(Tdiff=T, log_q=lq, log_pz=lpz, incompatible=inc)


class Problem:
    """Device-resident observation shard + model: what bench.py times with inputs already in HBM."""

    def __init__(self, obs, times=None, weights=None, ctx=None, obs_offset=0, N_global=0):
        obs = i32(obs)
        self.N, self.p = obs.shape
        self._ctx = ctx
        t = None if times is None else f64(times)
        w = None if weights is None else f64(weights)
        self._h = C.c_void_p()
        e = ErrBuf()
        e.check(capi.lib().mccbn_problem_create(_ctx(ctx), ptr(obs), ptr(t), ptr(w), self.N, self.p,
                                                C.c_int64(obs_offset), C.c_int64(N_global), C.byref(self._h),
                                                *e.args))

    def set_model(self, poset, lambda_, eps, lambda_s=1.0, max_lambda=1e6):
        poset, lam = i32(poset), f64(lambda_)
        e = ErrBuf()
        e.check(capi.lib().mccbn_problem_set_model(self._h, ptr(poset), ptr(lam), C.c_double(eps),
                                                   C.c_float(lambda_s), C.c_float(max_lambda), *e.args))

    def em_iterations(self, L, n_iter=1, sampling="forward", neighborhood_dist=1, times_available=False, seed=1,
                      update_params=True):
        e = ErrBuf()
        e.check(capi.lib().mccbn_problem_em_iterations(self._h, C.c_uint(int(L)), sampling.encode(),
                                                       C.c_uint(int(neighborhood_dist)), int(times_available),
                                                       int(seed), int(n_iter), int(update_params), *e.args))

    def read(self):
        lam = np.zeros(self.p)
        eps, ll = C.c_double(0), C.c_double(0)
        e = ErrBuf()
        e.check(capi.lib().mccbn_problem_read(self._h, ptr(lam), C.byref(eps), C.byref(ll), *e.args))
        return dict(lambda_=lam, eps=eps.value, llhood=ll.value)

    @property
    def stream(self):
        return capi.lib().mccbn_problem_stream(self._h)

    def close(self):
        if self._h:
            capi.lib().mccbn_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
