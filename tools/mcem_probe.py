#!/usr/bin/env python
"""Dev helper: wall time of whole MCEM.hcbn runs (host buffers, 100 EM iterations) on the small configs, GPU vs the
CPU oracle -- the "seconds per MCEM iteration" half of the BASELINE metric."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import mccbn_b200 as mc
from bench import make_problem

names = sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]
with_cpu = os.environ.get("PROBE_CPU", "1") == "1"
if with_cpu:
    import oracle as orc
    orc.build()
for name in names:
    cfg, p, N, L, sampling, desc = make_problem(name)
    kw = dict(L=L, eps=0.1, max_iter=100, update_step_size=20, tol=0.0, seed=10)   # tol = 0: run all 100 iterations
    mc.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], **kw)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        r = mc.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], **kw)
        ts.append(time.perf_counter() - t0)
    g = min(ts)
    line = f"{name}: GPU {g*1e3:8.2f} ms / 100 iterations = {g*10:7.3f} ms per iteration ({N*L*100/g:.3e} samples/s)"
    if with_cpu and N * L <= 1e7:
        t0 = time.perf_counter()
        o = orc.MCEM_hcbn(cfg["lambda0"], cfg["poset"], cfg["obs"], thrds=orc.num_threads(), **kw)
        c = time.perf_counter() - t0
        line += f" | CPU oracle ({orc.num_threads()} threads) {c*1e3:9.1f} ms = {c*10:8.2f} ms per iteration -> x{c/g:.0f}"
    print(line, flush=True)
