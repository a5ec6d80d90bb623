#!/usr/bin/env python
"""Dev helper: sizes beyond the BASELINE configs (index arithmetic, memory): N = 1e6 x L = 1e4 forward (1e10 samples per
E-step), the shard-sum property on it, and the 64-event wide path at N = 1e5."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mccbn_b200 as mc
from mccbn_b200 import synth

t0 = time.time()
c = synth.make_config(30, 1000000, eps=0.05, seed=10)
print("synthetic data: %.1f s" % (time.time() - t0), flush=True)
ctx = mc.Context(0)
mc.set_jit(1, ctx=ctx)
L = 10000
for rep in range(2):
    t0 = time.time()
    ll = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx)
    dt = time.time() - t0
    print(f"N=1e6 L=1e4: llhood {ll:.3f} in {dt*1e3:.1f} ms -> {1e10/dt:.3e} samples/s (host buffers)", flush=True)
parts = [mc.obs_loglikelihood(c["obs"][lo:hi], c["poset"], c["lambdas"], 0.05, L=L, seed=1, ctx=ctx, obs_offset=lo)
         for lo, hi in ((0, 333333), (333333, 1000000))]
print("shard sum rel. diff %.2e" % abs(sum(parts) / ll - 1), flush=True)
pb = mc.Problem(c["obs"], ctx=ctx)
pb.set_model(c["poset"], c["lambda0"], 0.1)
pb.em_iterations(L, n_iter=2, seed=3)
out = pb.read()
print("2 EM iterations: eps %.4f, max |lambda/true - 1| %.3f" % (out["eps"], np.abs(out["lambda_"] / c["lambdas"] - 1).max()), flush=True)
pb.close()
cw = synth.make_config(64, 100000, eps=0.05, seed=7, density=0.05)
t0 = time.time()
llw = mc.obs_loglikelihood(cw["obs"], cw["poset"], cw["lambdas"], 0.05, L=1000, seed=1, ctx=ctx)
print(f"p=64 N=1e5 L=1e3: llhood {llw:.3f} in {(time.time()-t0)*1e3:.1f} ms", flush=True)
print("peak device memory %.2f GB" % (torch.cuda.max_memory_allocated() / 1e9 if torch.cuda.is_available() else 0), flush=True)
