#!/usr/bin/env python
"""Dev helper: one small call of every kernel family, generic and specialised (compute-sanitizer is closed on this
pool; the kernel variants are cross-checked against each other and the oracle by the tests instead)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mccbn_b200 as mc
from mccbn_b200 import synth

c = synth.make_config(9, 70, eps=0.05, seed=3, density=0.3)
c0 = synth.make_config(9, 70, eps=0.0, seed=3, density=0.3)
for jit in (0, 1):
    mc.set_jit(jit)
    for sampling, kw, cfg in (("forward", {}, c), ("backward", dict(neighborhood_dist=2), c), ("bernoulli", {}, c0),
                              ("add-remove", {}, c), ("pool", {}, c)):
        L = 46 * 3 if sampling == "backward" else 96
        r = mc.importance_weight(cfg["obs"], L, cfg["poset"], cfg["lambdas"], 0.07, sampling=sampling, seed=1, **kw)
        print(jit, sampling, float(np.sum(r["w"])), flush=True)
    r = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=64, eps=0.1, max_iter=12, update_step_size=4, seed=2)
    print(jit, "mcem", r["eps"], flush=True)
mc.set_jit(-1)
cw = synth.make_config(40, 20, eps=0.05, seed=4, density=0.08)
print("wide", mc.obs_loglikelihood(cw["obs"], cw["poset"], cw["lambdas"], 0.1, L=64, seed=1), flush=True)
print("otcbn", mc.estimate_mutation_rates(c0["poset"], c0["obs"], max_iter=5, seed=1)["llhood"], flush=True)
print("sim", mc.sample_genotypes(100, c["poset"], c["lambdas"], seed=1)["samples"].sum(), flush=True)
import tempfile
r = mc.adaptive_simulated_annealing(np.zeros((9, 9), np.int32), c["obs"], L=32, max_iter=8, update_step_size=4, max_iter_asa=6,
                                    outdir=tempfile.mkdtemp() + "/", seed=1)
print("asa", r["llhood"], flush=True)
