#!/usr/bin/env python
"""Dev helper: BASELINE config 4 end to end -- adaptive.simulated.annealing p=20, N=10,000, L=1,000, max.iter.asa=100 on the
GPU (speculative batches of candidate posets), and the CPU oracle's cost of ONE candidate score for comparison."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import mccbn_b200 as mc
from bench import make_problem

cfg, p, N, L, sampling, desc = make_problem("cfg4")
steps = int(os.environ.get("ASA_STEPS", "100"))
for batch in [int(b) for b in os.environ.get("ASA_BATCHES", "1,8,0").split(",")]:
    if batch:
        os.environ["MCCBN_ASA_BATCH"] = str(batch)
    else:
        os.environ.pop("MCCBN_ASA_BATCH", None)  # automatic
    d = tempfile.mkdtemp() + "/"
    t0 = time.perf_counter()
    r = mc.adaptive_simulated_annealing(cfg["poset"], cfg["obs"], L=L, max_iter=100, update_step_size=20, max_iter_asa=steps,
                                        outdir=d, seed=10)
    dt = time.perf_counter() - t0
    n_all, n_jit = mc.kernel_launches(reset=True), mc.jit_launches(reset=True)
    print(f"GPU ASA cfg4: batch={batch:2d}  {steps} steps in {dt:7.2f} s  -> {steps/dt:6.2f} candidate MCEM fits/s, llhood {r['llhood']:.2f}"
          f"  ({n_all} launches, {n_jit} specialised)", flush=True)
if os.environ.get("PROBE_CPU", "1") == "1":
    import oracle as orc
    orc.build()
    t0 = time.perf_counter()
    orc.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], L=L, eps=0.1, max_iter=20, update_step_size=20, tol=0.0,
                  seed=1, thrds=orc.num_threads())
    dt = (time.perf_counter() - t0) * 5
    print(f"CPU oracle ({orc.num_threads()} threads): one candidate fit (100 EM iterations, extrapolated from 20) = {dt:.1f} s "
          f"-> {1/dt:.4f} fits/s", flush=True)
