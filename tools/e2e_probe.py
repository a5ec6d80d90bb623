#!/usr/bin/env python
"""Dev helper: fixed overhead of the reference-facing one-call E-step (host buffers) vs the resident-problem path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mccbn_b200 as mc
from bench import make_problem

cfg, p, N, L, sampling, desc = make_problem(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
ctx = mc.Context(0)
mc.set_jit(1, ctx=ctx)
obs = np.asfortranarray(cfg["obs"])
for LL in (32, L):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mc.obs_loglikelihood(obs, cfg["poset"], cfg["lambda0"], 0.1, L=LL, sampling=sampling, seed=7 + rep, ctx=ctx)
        torch.cuda.synchronize()
        print(f"L={LL:6d} call {rep}: {1e3 * (time.perf_counter() - t0):8.2f} ms", flush=True)
