#!/usr/bin/env python
"""bench.py -- headline benchmark of the MCEM E-step hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3|cfg2|cfg1|cfg5]

metric   importance samples / second of one EM iteration (E-step + reduction + [allreduce] + M-step), whole job
workload BASELINE config 3: MCEM.hcbn forward scheme, p=30, N=100,000, L=10,000 (1e9 samples per E-step); the
         N observations are sharded over the ranks (strong scaling: total work fixed)
value    inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
e2e      the same metric through the reference-facing C ABI call (`mccbn_obs_log_likelihood`: one E-step) with HOST
         buffers: H2D of the observation matrix and D2H of the result inside the timed region
roofline the path is instruction-issue bound (Philox INT32 + FP32/SFU), not HBM or tensor bound: see DESIGN.md.
         peak  = 32-bit random words/s of a pure Philox4x32-10 generator micro-kernel measured live on this GPU
         achieved = random words/s consumed by the E-step kernel (PMAX/4 Philox blocks per sample)
cpu_baseline / --impl reference: the CPU oracle (oracle/, a restatement of the reference's OpenMP loop -- the
         reference itself cannot be compiled in this image) on all host cores over a bounded slice of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (p, N, L, eps_true, sampling, description)
    "cfg1": (10, 100, 100, 0.01, "forward", "README H-CBN2 example p=10 N=100 L=100"),
    "cfg2": (16, 10000, 1000, 0.05, "forward", "MCEM.hcbn p=16 N=10,000 L=1,000"),
    "cfg3": (30, 100000, 10000, 0.05, "forward", "MCEM.hcbn p=30 N=100,000 L=10,000"),
    "cfg4": (20, 10000, 1000, 0.05, "forward", "ASA-sized E-step p=20 N=10,000 L=1,000"),
    "cfg5": (25, 50000, 5000, 0.05, "backward", "MCEM.hcbn backward k=1 p=25 N=50,000 L=5,000 eps=0.05"),
    "cfg5pool": (25, 50000, 5000, 0.05, "pool", "MCEM.hcbn pool K=125,000 p=25 N=50,000 L=5,000 eps=0.05"),
}


def make_problem(name):
    from mccbn_b200 import synth
    p, N, L, eps, sampling, desc = CONFIGS[name]
    cfg = synth.make_config(p, N, eps=eps, seed=10)
    return cfg, p, N, L, sampling, desc


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def philox_peak_words_per_s(torch):
    """Pure Philox4x32-10 generator micro-kernel (same round function as the E-step), words/s."""
    import ctypes as C
    from mccbn_b200 import _capi
    lib = _capi.lib()
    if not hasattr(lib, "mccbn_microbench_philox"):
        return None
    lib.mccbn_microbench_philox.restype = C.c_double
    lib.mccbn_microbench_philox.argtypes = [C.c_void_p, C.c_int]
    return float(lib.mccbn_microbench_philox(None, 5))


def host_cores(orc):
    """Threads for the CPU arms: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers, so
    omp_get_max_threads() would make the reference arm single-threaded in the N > 1 runs; the oracle takes the count as
    an argument (`thrds`, like the reference) and calls omp_set_num_threads itself."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(int(n), int(orc.num_threads()), 1)


def run_reference(args):
    """Reference arm: the CPU oracle (the reference's OpenMP loop restated) on all host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    cfg, p, N, L, sampling, desc = make_problem(args.config)
    cores = host_cores(orc)
    kw = dict(L=L, sampling=sampling, thrds=cores)
    # calibrate a bounded slice: ~args.ref_seconds of CPU work per step
    n0 = min(N, 4 * cores)
    t0 = time.perf_counter()
    orc.obs_log_likelihood(cfg["obs"][:n0], cfg["poset"], cfg["lambdas"], cfg["eps"], seed=1, **kw)
    dt = time.perf_counter() - t0
    n_s = int(max(cores, min(N, n0 * args.ref_seconds / max(dt, 1e-6))))
    n_s -= n_s % cores or 0
    n_s = max(n_s, cores)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.obs_log_likelihood(cfg["obs"][:n_s], cfg["poset"], cfg["lambdas"], cfg["eps"], seed=2 + it, **kw)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    L_eff = L if sampling != "backward" else (L // (p + 1)) * (p + 1)
    ms = 1e3 * float(np.mean(times))
    val = n_s * L_eff / (ms * 1e-3)
    sample = f"{n_s} of {N} observations x L={L} per step (observations are independent in the reference loop)"
    line = {
        "impl": "reference", "metric": "importance samples/sec per E-step", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "sampling": sampling, "p": p, "N": N, "L": L},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import mccbn_b200 as mc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg, p, N, L, sampling, desc = make_problem(args.config)
    # shard observations contiguously over ranks (strong scaling)
    lo = (N * rank) // world
    hi = (N * (rank + 1)) // world
    obs = np.ascontiguousarray(cfg["obs"][lo:hi])
    obs = np.asfortranarray(obs)

    ctx = mc.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    # NVRTC poset-specialised kernels (csrc/jit.cuh) under the library's DEFAULT policy: the device-resident problem
    # announces its work, the kernel is compiled during warm-up and the one-call e2e path finds it in the context's cache
    if args.no_jit:
        mc.set_jit(0, ctx=ctx)
    if world > 1:
        from mccbn_b200 import dist as mdist
        if args.exchange == "peer":  # statistic exchange fused into the reduction / M-step kernels over NVLink peer memory
            ctx.init_peers(rank, world, mdist.gather_peer_handles(ctx, dist, torch.device("cuda", local)))
        else:                        # one ncclAllReduce per EM iteration
            ctx.init_comm(rank, world, mdist.broadcast_unique_id(dist, torch.device("cuda", local)))

    pb = mc.Problem(obs, ctx=ctx, obs_offset=lo, N_global=N)
    pb.set_model(cfg["poset"], cfg["lambda0"], 0.1 if cfg["eps"] != 0.01 else 0.01)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(seed):
        pb.em_iterations(L, n_iter=1, sampling=sampling, seed=seed, update_params=True)

    with torch.cuda.stream(stream):
        for w in range(args.warmup):
            one_step(100 + w)
        # steady state: under the default JIT policy a mid-sized per-rank workload (>= 5e7 samples per E-step, e.g. cfg 5 on
        # 4 or 8 ranks) compiles its specialised kernel on a background thread while the generic kernel runs; a real fit of
        # >= 100 iterations switches over after ~0.4 s.  Let that finish before the timed steps, then take it up.
        barrier()
        time.sleep(0.8)
        one_step(100 + args.warmup)
        barrier()
        ctx.kernel_launches(reset=True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        evs = []
        barrier()
        for k in range(args.steps):
            flush.zero_()  # evict L2 between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one_step(1000 + k)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        launches = ctx.kernel_launches()
        jit_launches = ctx.jit_launches(reset=True)
        clocks = sampler.stop() if rank == 0 else None
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    run_note = None
    check = None
    try:
        out = pb.read()
        assert np.all(np.isfinite(out["lambda_"])), "E-step produced non-finite parameters"
        # result checksum: the parameters after the W + K EM iterations of this run.  The Philox counters carry global
        # observation indices, so every --gpus N must print the same numbers (up to fp64 summation order, ~1e-12)
        check = {"llhood": float(out["llhood"]), "eps": float(out["eps"]), "sum_lambda": float(np.sum(out["lambda_"])),
                 "em_iterations": args.warmup + 1 + args.steps}
    except mc.MccbnError as e:
        if e.code != 2:
            raise
        # "all samples have weight 0": with the backward scheme at k = 1 some noisy observations have no compatible
        # genotype in their Hamming ball; the reference aborts the run at this point as well (mcem_hcbn.cpp:695-699).
        # The timing of the E-step is unaffected.
        run_note = "run ends with the reference's error: " + str(e)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    L_eff = L if sampling != "backward" else (L // (p + 1)) * (p + 1)
    value = N * L_eff / (ms * 1e-3)

    # ---- end-to-end through the C ABI with host buffers (one E-step per call) ----
    e2e_steps = max(2, min(args.steps, 5))
    # the observations of the timed calls live in pinned host memory (column-major N x p int32, R's layout)
    host_kind, pinned_keepalive = "pageable", None
    try:
        pinned_keepalive = torch.empty((obs.shape[1], obs.shape[0]), dtype=torch.int32).pin_memory()
        view = pinned_keepalive.numpy().T  # F-contiguous (N, p) view of the pinned block: passed through without a copy
        view[...] = obs
        if view.flags["F_CONTIGUOUS"] and view.dtype == np.int32:
            obs, host_kind = view, "pinned"
    except Exception:
        pinned_keepalive = None

    def e2e_call(seed):
        try:
            mc.obs_loglikelihood(obs, cfg["poset"], cfg["lambda0"], 0.1, L=L, sampling=sampling, seed=seed, ctx=ctx,
                                 obs_offset=lo, N_global=N)
        except mc.MccbnError as e:  # the reference's "all samples have weight 0" abort (see run_note above)
            if e.code != 2:
                raise

    with torch.cuda.stream(stream):
        e2e_call(7)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_call(8 + k)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e_jit = ctx.jit_launches()
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    n_loc = hi - lo
    h2d = n_loc * p * 4 + 2 * n_loc * 8
    d2h = 1600

    line = {
        "metric": "importance samples/sec per E-step", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 sampling / f64 reduction", "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "sampling": sampling, "p": p, "N": N, "L": L,
                   "sec_per_mcem_iteration": ms * 1e-3, "l2": "256 MiB buffer written between timed iterations",
                   "parallelism": (f"observations sharded over {world} rank(s); per iteration one exchange of p+5 doubles, " +
                                   ("fused into the reduction/M-step kernels over NVLink peer memory"
                                    if args.exchange == "peer" else "ncclAllReduce")),
                   "kernel": ({"forward": "estep_forward", "pool": "pool_generate + estep_pool_mma"}.get(
                       sampling, "estep_conditional") + (" (NVRTC poset-specialised, default JIT policy)"
                                                         if jit_launches else " (generic)"))},
        "e2e": {"value": N * L_eff / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world,
                "call": f"mccbn_obs_log_likelihood ({host_kind} host buffers, one E-step, default JIT policy: "
                        f"{'specialised kernel from the context cache' if e2e_jit else 'generic kernel'})"},
        "gpu_launches": int(launches),
        "gpu_launches_specialised": int(jit_launches),
        "check": check,
        "clocks": clocks,
    }
    if run_note:
        line["config"]["note"] = run_note
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = None
        try:
            with torch.cuda.stream(stream):
                peak = philox_peak_words_per_s(torch)
        except Exception:
            peak = None
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        line["roofline"] = workload_roofline(args.config, cfg, p, N // world, L, sampling, ms, peak, sm_hz, n_sm, peaks)
        if world == 1 and not args.no_other:
            other = {}
            for name in ("cfg2", "cfg4", "cfg5", "cfg5pool"):
                if name == args.config:
                    continue
                try:
                    other[name] = time_other_workload(torch, mc, ctx, stream, flush, name, peak, sm_hz, n_sm, peaks)
                    if CONFIGS[name][4] == "forward" and not other[name]["specialised_kernel"]:
                        # 1e7 samples per E-step is below the default policy's compile threshold: the number above is the
                        # generic kernel's; the specialised kernel's (mccbn_set_jit(1), what a long fit would reach) beside it
                        mc.set_jit(1, ctx=ctx)
                        try:
                            other[name + "_specialised"] = time_other_workload(torch, mc, ctx, stream, flush, name, peak,
                                                                               sm_hz, n_sm, peaks)
                        finally:
                            mc.set_jit(-1, ctx=ctx)
                except Exception as e:  # a side measurement must never take the headline down
                    other[name] = {"error": str(e)[:200]}
            try:
                other["cfg4_asa"] = time_asa(mc, ctx)
            except Exception as e:
                other["cfg4_asa"] = {"error": str(e)[:200]}
            line["config"]["other_workloads"] = other
            if sampling == "forward":
                try:
                    line["config"]["fp64_mode"] = time_fp64_mode(torch, mc, ctx, stream, cfg, obs, p, N, L)
                except Exception as e:
                    line["config"]["fp64_mode"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, p, N, L, sampling, args)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def fwd_words_per_sample(p):
    """p + 1 uniforms of 24 bits packed back to back in the Philox stream (kernels_forward.cuh): 32-bit words drawn"""
    return 4 * (((24 * p + 23) // 32) // 4 + 1)


def executed_backward_samples(cfg, p, L):
    """backward scheme, k = 1: the kernel draws `reps` samples for every genotype of the Hamming ball that is compatible
    with the poset and skips the others (their weight is 0 in the reference, src/mcem_hcbn.cpp:524-532)."""
    obs = np.asarray(cfg["obs"]) != 0
    poset = np.asarray(cfg["poset"]) != 0
    reps = L // (p + 1)
    par = [np.nonzero(poset[:, j])[0] for j in range(p)]

    def n_incompatible(X):  # rows of X that violate some cover relation
        bad = np.zeros(X.shape[0], bool)
        for j in range(p):
            if par[j].size:
                bad |= X[:, j] & ~X[:, par[j]].all(axis=1)
        return bad

    total = (~n_incompatible(obs)).sum()
    for j in range(p):
        X = obs.copy()
        X[:, j] = ~X[:, j]
        total += (~n_incompatible(X)).sum()
    return int(total) * reps


def workload_roofline(name, cfg, p, N_local, L, sampling, ms, philox_peak, sm_hz, n_sm, peaks):
    """Roofline of the dominant kernel of one workload.  None of these kernels is HBM- or tensor-bound (no contraction;
    the observation stream is ~1e-3 B per sample): the binding resource is instruction issue.  Ceilings, all per GPU:
      philox  32-bit random words/s of a pure Philox4x32-10 generator kernel measured live (IMAD.WIDE-bound)
      mufu    16 MUFU results / clk / SM (XU pipe, tools/pipe_microbench.cu: MUFU.LG2 = 8 cycles per warp per scheduler)
      fp32    128 FFMA / clk / SM
    `achieved` counts ALGORITHMIC work per launch / the E-step time."""
    t = ms * 1e-3
    mufu_peak = 16.0 * n_sm * sm_hz
    fp32_peak = 128.0 * n_sm * sm_hz
    hbm_bytes = N_local * (8 + (5 + p) * 8 * 2)  # observation word read + partial record write/read per observation
    out = {"bound": "instruction issue (Philox INT32 + FP32/SFU); not hbm/tensor", "traffic": ncu_traffic_bytes(name),
           "hbm": {"achieved_GBs": hbm_bytes / t / 1e9, "peak_GBs": peaks.get("hbm_gbs"),
                   "note": "algorithmic HBM bytes per E-step / step time: memory is not the limiter"}}
    if sampling == "forward":
        wps = fwd_words_per_sample(p)
        ach = N_local * L * wps / t
        out.update({"achieved": ach / 1e9, "peak": philox_peak / 1e9 if philox_peak else None,
                    "unit": "Gword/s (32-bit Philox output)", "frac": ach / philox_peak if philox_peak else None,
                    "words_per_sample": wps,
                    "also": {"mufu_frac": N_local * L * (p + 1) / t / mufu_peak,
                             "fp32_frac": N_local * L * (2 * p + 3) / t / fp32_peak},
                    "peak_source": "pure Philox4x32-10 generator micro-kernel (IMAD.WIDE-bound) measured live on this GPU; "
                                   "frac = share of the step the multiplier pipe alone would need (see profiles/README.md "
                                   "for the measured issue-cost model that accounts for the rest)"})
    elif sampling == "pool":
        PM = 8 if p <= 8 else (16 if p <= 16 else 32)
        pairs = float(N_local) * p * L  # (observation, pool entry) pairs, K = p L entries
        # per pair: one 32-bit Philox word + MUFU.LG2 (the pair's sampling time), a 6-step rank search in the entry's sorted
        # times, one popc, one table look-up -- and PMAX + 3 weighted sums, which run as TF32 mma with split operands
        # (estep_pool_mma_kernel).  No single pipe bounds it (ncu r2l: issue slots 74 %, shared-memory wavefronts 65 %); the
        # line reports the same ceiling as the forward kernels, the multiplier pipe's share, and the others beside it.
        ach = pairs / t  # Philox words/s: one per pair
        mma_flops = pairs * PM * 2 * 3  # three TF32 products per (pair, event)
        out.update({"achieved": ach / 1e9, "peak": philox_peak / 1e9 if philox_peak else None,
                    "unit": "Gword/s (32-bit Philox output, one word per pair)",
                    "frac": ach / philox_peak if philox_peak else None, "pairs_per_s": pairs / t,
                    "also": {"mufu_frac": pairs / t / mufu_peak, "ffma_equivalent_frac": pairs * (PM + 3) / t / fp32_peak,
                             "mma_tf32_TFLOPs": mma_flops / t / 1e12},
                    "peak_source": "pure Philox4x32-10 generator micro-kernel measured live on this GPU; the kernel is bound by "
                                   "the issue of its per-pair scalar work, not by one pipe: without any of the weighted-sum "
                                   "work it takes 21.8 of its 25.8 ms (profiles/README.md r2n)"})
    else:
        nblk = ((24 * p + 23) // 32) // 4 + 1
        executed = executed_backward_samples(cfg, p, L) * (N_local / cfg["obs"].shape[0]) if sampling == "backward" \
            else float(N_local) * L
        mufu = executed * (3 * p + 2)  # per node EX2 (F), LG2 (z), LG2 (log F); + T_s + the final EX2 of the weight
        ach = mufu / t
        out.update({"achieved": ach / 1e12, "peak": mufu_peak / 1e12, "unit": "TMUFU/s (3p + 2 per executed sample)",
                    "frac": ach / mufu_peak, "executed_samples": executed,
                    "also": {"philox_frac": (executed * 4 * nblk / t / philox_peak) if philox_peak else None},
                    "peak_source": "16 MUFU/clk/SM x SMs x SM clock (XU pipe); executed samples = compatible genotypes of "
                                   "the Hamming balls x reps (incompatible ones are skipped: their weight is 0)"})
    return out


def time_other_workload(torch, mc, ctx, stream, flush, name, philox_peak, sm_hz, n_sm, peaks, steps=5, warmup=3):
    """One EM iteration of another BASELINE configuration, inputs resident, CUDA events on the launching stream."""
    cfg, p, N, L, sampling, desc = make_problem(name)
    pb = mc.Problem(np.asfortranarray(cfg["obs"]), ctx=ctx)
    pb.set_model(cfg["poset"], cfg["lambda0"], 0.1)
    ctx.jit_launches(reset=True)
    with torch.cuda.stream(stream):
        for w in range(warmup):
            pb.em_iterations(L, n_iter=1, sampling=sampling, seed=50 + w, update_params=False)
        torch.cuda.synchronize()
        evs = []
        for k in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            pb.em_iterations(L, n_iter=1, sampling=sampling, seed=60 + k, update_params=False)
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    L_eff = L if sampling != "backward" else (L // (p + 1)) * (p + 1)
    r = {"workload": f"{name}: {desc}", "ms_per_step": ms, "value": N * L_eff / (ms * 1e-3), "unit": "samples/s",
         "specialised_kernel": bool(ctx.jit_launches()),
         "roofline": workload_roofline(name, cfg, p, N, L, sampling, ms, philox_peak, sm_hz, n_sm, peaks)}
    pb.close()
    return r


def time_asa(mc, ctx, steps=100):
    """BASELINE config 4 end to end through the reference-facing call: adaptive.simulated.annealing, p = 20, N = 10,000,
    L = 1,000, max.iter.asa = 100 (host buffers in, files and the fitted model out; every step is one candidate MCEM fit of
    up to 100 EM iterations, scored speculatively in batches).  Wall clock of the whole call."""
    import tempfile
    cfg, p, N, L, sampling, desc = make_problem("cfg4")
    outdir = tempfile.mkdtemp() + "/"
    ctx.kernel_launches(reset=True)
    ctx.jit_launches(reset=True)
    t0 = time.perf_counter()
    r = mc.adaptive_simulated_annealing(cfg["poset"], cfg["obs"], L=L, max_iter=100, update_step_size=20,
                                        max_iter_asa=steps, outdir=outdir, seed=10, ctx=ctx)
    dt = time.perf_counter() - t0
    return {"workload": "cfg4: adaptive.simulated.annealing p=20 N=10,000 L=1,000 max.iter.asa=100", "seconds": dt,
            "candidate_fits_per_s": steps / dt, "llhood": float(r["llhood"]), "gpu_launches": int(ctx.kernel_launches()),
            "gpu_launches_specialised": int(ctx.jit_launches()),
            "call": "mccbn_adaptive_simulated_annealing (host buffers, default JIT policy and batch size)"}


def time_fp64_mode(torch, mc, ctx, stream, cfg, obs, p, N, L, steps=2):
    """The headline workload in the library's fp64 mode (precision = 1: 53-bit uniforms, log, fp64 event times and sums --
    the reference's arithmetic), through the one-call entry with host buffers."""
    with torch.cuda.stream(stream):
        mc.obs_loglikelihood(obs, cfg["poset"], cfg["lambda0"], 0.1, L=L, seed=3, ctx=ctx, precision=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            mc.obs_loglikelihood(obs, cfg["poset"], cfg["lambda0"], 0.1, L=L, seed=4 + k, ctx=ctx, precision=1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    return {"value": N * L / dt, "unit": "samples/s", "ms_per_step": dt * 1e3,
            "call": "mccbn_obs_log_likelihood, precision = 1 (fp64 sampling), host buffers"}


def ncu_traffic_bytes(config):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json); None if there is no capture for it."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(config, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_baseline(cfg, p, N, L, sampling, args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    cores = host_cores(orc)
    kw = dict(L=L, sampling=sampling, thrds=cores)
    n0 = min(N, 4 * cores)
    t0 = time.perf_counter()
    orc.obs_log_likelihood(cfg["obs"][:n0], cfg["poset"], cfg["lambdas"], cfg["eps"], seed=1, **kw)
    dt = time.perf_counter() - t0
    n_s = int(max(cores, min(N, n0 * args.cpu_seconds / max(dt, 1e-6))))
    n_s = max(cores, n_s - n_s % cores)
    t0 = time.perf_counter()
    orc.obs_log_likelihood(cfg["obs"][:n_s], cfg["poset"], cfg["lambdas"], cfg["eps"], seed=2, **kw)
    dt = time.perf_counter() - t0
    L_eff = L if sampling != "backward" else (L // (p + 1)) * (p + 1)
    return {"value": n_s * L_eff / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{n_s} of {N} observations x L={L}, one E-step, {dt:.1f} s (oracle = restated reference loop)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the side measurements (other configs, fp64 mode)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU statistic exchange: fused peer-memory kernels (default) or ncclAllReduce")
    ap.add_argument("--no-jit", action="store_true", help="time the generic (poset-interpreting) forward kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
