#!/usr/bin/env python
"""Dev helper: time one EM iteration (E-step + reduction + M-step) of a config with inputs resident on the GPU.
usage: python tools/kbench.py cfg3 [cfg2 ...]   (MCCBN_B200_LIB selects a kernel-variant build)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mccbn_b200 as mc
from bench import CONFIGS, make_problem

def run(name, steps=5):
    cfg, p, N, L, sampling, desc = make_problem(name)
    if name.endswith("pool"):
        sampling = "pool"
    ctx = mc.Context(0)
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    pb = mc.Problem(cfg["obs"], ctx=ctx)
    pb.set_model(cfg["poset"], cfg["lambda0"], 0.1)
    with torch.cuda.stream(st):
        for w in range(2):
            pb.em_iterations(L, n_iter=1, sampling=sampling, seed=w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        pb.em_iterations(L, n_iter=steps, sampling=sampling, seed=9)
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    Leff = L if sampling != "backward" else (L // (p + 1)) * (p + 1)
    try:
        eps = "%.4f" % pb.read()["eps"]
    except mc.MccbnError as e:
        eps = "ERR(" + str(e)[:30] + ")"
    print(f"{os.environ.get('MCCBN_B200_LIB','default')[-20:]:20s} jit={os.environ.get('MCCBN_JIT','auto'):4s} {name:9s} {sampling:9s} {ms:9.3f} ms/iter  {N*Leff/ms/1e6:9.2f} Gsamples/s  eps={eps}", flush=True)

if __name__ == "__main__":
    for n in sys.argv[1:]:
        run(n)
