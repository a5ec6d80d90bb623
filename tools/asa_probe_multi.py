#!/usr/bin/env python
"""Dev helper (torchrun, one rank per GPU): BASELINE config 4 -- adaptive.simulated.annealing p=20, N=10,000, L=1,000,
max.iter.asa=100 -- candidate-parallel over the GPUs of one box.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/asa_probe_multi.py"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import mccbn_b200 as mc
from mccbn_b200 import dist as mdist
from bench import make_problem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, p, N, L, sampling, desc = make_problem("cfg4")
steps = int(os.environ.get("ASA_STEPS", "100"))
ctx = mdist.init_context(local, dist, exchange="peer")
for rep in range(2):
    d = tempfile.mkdtemp() + "/"
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    r = mc.adaptive_simulated_annealing(cfg["poset"], cfg["obs"], L=L, max_iter=100, update_step_size=20, max_iter_asa=steps,
                                        outdir=d, seed=10, ctx=ctx)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"ASA cfg4 on {world} GPU(s), run {rep}: {steps} steps in {dt:7.2f} s -> {steps/dt:6.2f} steps/s, llhood {r['llhood']:.2f}",
              flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
