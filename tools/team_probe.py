#!/usr/bin/env python
"""In-process device team (mccbn_set_gpus): one process, G devices, host buffers in, result out.
Times the reference-facing one-call E-step (mccbn_obs_log_likelihood) and a short MCEM.hcbn fit for G = 1, 2, 4, 8
(as many as the box has).  Usage: team_probe.py [cfg3] [max_gpus]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mccbn_b200 as mc
from bench import make_problem

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg, p, N, L, sampling, desc = make_problem(name)
n_dev = min(torch.cuda.device_count(), int(sys.argv[2]) if len(sys.argv) > 2 else 8)
obs = np.asfortranarray(cfg["obs"])
mc.set_jit(1)  # poset-specialised kernels on every device (the automatic policy compiles only for announced work >= 5e9 samples)
out = {"workload": desc, "N": N, "L": L, "p": p, "sampling": sampling, "rows": []}
for G in [g for g in (1, 2, 4, 8) if g <= n_dev]:
    mc.set_gpus(G)
    ts, lls = [], []
    for rep in range(6):
        t0 = time.perf_counter()
        ll = mc.obs_loglikelihood(obs, cfg["poset"], cfg["lambda0"], 0.1, L=L, sampling=sampling, seed=7)
        ts.append(time.perf_counter() - t0)
        lls.append(ll)
    t0 = time.perf_counter()
    fit = mc.MCEM_hcbn(cfg["lambda0"], cfg["poset"], obs, L=L, eps=0.1, sampling=sampling, seed=3, max_iter=10,
                       update_step_size=5)
    t_fit = time.perf_counter() - t0
    best = min(ts[2:])
    row = {"gpus": G, "estep_call_ms": round(1e3 * best, 3), "samples_per_s": N * L / best,
           "first_call_ms": round(1e3 * ts[0], 1), "mcem_10_iterations_ms": round(1e3 * t_fit, 2), "llhood": lls[-1],
           "fit_eps": fit["eps"]}
    out["rows"].append(row)
    print(row, flush=True)
print("TEAM_PROBE " + json.dumps(out))
