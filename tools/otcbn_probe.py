#!/usr/bin/env python
"""Dev helper: the legacy OT-CBN entry (estimate_mutation_rates -> MCEM, src/mcem.cpp) GPU vs oracle wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import mccbn_b200 as mc
from mccbn_b200 import synth
import oracle as orc
orc.build()
for p, N, S in ((16, 10000, 5), (16, 10000, 100), (30, 100000, 5)):
    c = synth.make_config(p, N, eps=0.0, seed=10)
    kw = dict(max_iter=100, nrOfSamples=S, seed=3)
    mc.estimate_mutation_rates(c["poset"], c["obs"], **kw)
    t0 = time.perf_counter(); g = mc.estimate_mutation_rates(c["poset"], c["obs"], **kw); tg = time.perf_counter() - t0
    line = f"OT-CBN p={p} N={N} nrOfSamples={S}: GPU {tg*1e3:8.1f} ms"
    if N * S <= 1e6:
        t0 = time.perf_counter(); o = orc.MCEM_otcbn(c["lambda0"] * 0 + g["lambda_"].mean(), c["poset"], c["obs"], max_iter=100, nrOfSamples=S, seed=3); to = time.perf_counter() - t0
        line += f" | oracle {to*1e3:8.1f} ms -> x{to/tg:.0f}"
    print(line, flush=True)
