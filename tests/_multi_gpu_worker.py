"""Worker of tests/test_gpu_multi.py (launched through torch.distributed.run, one rank per GPU).

Each rank holds a contiguous shard of the observations.  For both statistic exchanges (peer memory, NCCL) it runs a few EM
iterations and checks that (a) all ranks end with bit-identical parameters, (b) they equal the single-GPU run over all
observations up to the fp64 summation order (the Philox stream is indexed by the GLOBAL observation index)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import mccbn_b200 as mc
    from mccbn_b200 import dist as mdist, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # MCCBN_TEST_ONE_DEVICE=1: every rank on device 0 (a one-GPU box).  NCCL refuses two ranks on one device, so the
    # plumbing (handle all-gather, result comparison) goes over gloo and only the peer-memory exchange is exercised.
    one_device = os.environ.get("MCCBN_TEST_ONE_DEVICE") == "1"
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    if one_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    exchanges = ("peer",) if one_device else ("peer", "nccl")
    tdev = "cpu" if one_device else "cuda"
    res = {}
    verbose = bool(os.environ.get("MCCBN_TEST_VERBOSE"))

    def say(*a):
        if verbose:
            print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)

    try:
        # (scheme, L, neighbourhood, config): odd N = ragged split; the backward config is one in which every observation
        # has a compatible genotype within distance 2 (otherwise the run aborts like the reference, mcem_hcbn.cpp:695-699)
        cases = (("forward", 512, 1, synth.make_config(14, 4001, eps=0.05, seed=5, density=0.2)),
                 ("backward", 560, 2, synth.make_config(10, 2001, eps=0.02, seed=5, density=0.2)))
        for sampling, L, kdist, c in cases:
            N = c["obs"].shape[0]
            lo, hi = mdist.shard_bounds(N, rank, world)
            obs = np.asfortranarray(c["obs"][lo:hi])
            ref = None
            if rank == 0:  # unsharded run on this rank's GPU
                ctx1 = mc.Context(local)
                pb = mc.Problem(c["obs"], ctx=ctx1)
                pb.set_model(c["poset"], c["lambda0"], 0.1)
                pb.em_iterations(L, n_iter=4, sampling=sampling, neighborhood_dist=kdist, seed=3)
                ref = pb.read()
                pb.close()
                ctx1.close()
                say(sampling, "single-GPU reference done")
            for exchange in exchanges:
                ctx = mdist.init_context(local, dist, exchange=exchange)
                say(sampling, exchange, "context ready")
                pb = mc.Problem(obs, ctx=ctx, obs_offset=lo, N_global=N)
                pb.set_model(c["poset"], c["lambda0"], 0.1)
                pb.em_iterations(L, n_iter=4, sampling=sampling, neighborhood_dist=kdist, seed=3)
                out = pb.read()
                say(sampling, exchange, "EM done", out["eps"])
                vec = torch.tensor(np.concatenate([out["lambda_"], [out["eps"], out["llhood"]]]), device=tdev)
                allv = [torch.zeros_like(vec) for _ in range(world)]
                dist.all_gather(allv, vec)
                same = all(bool(torch.equal(allv[0], v)) for v in allv)
                key = f"{sampling}/{exchange}"
                res[key] = {"identical_across_ranks": same}
                if rank == 0:
                    res[key]["max_rel_vs_single_gpu"] = float(np.max(np.abs(out["lambda_"] / ref["lambda_"] - 1)))
                    res[key]["eps_abs"] = abs(out["eps"] - ref["eps"])
                    res[key]["ll_rel"] = abs(out["llhood"] / ref["llhood"] - 1)
                pb.close()
                dist.barrier()
                ctx.close()
                say(sampling, exchange, "context closed")
        # candidate-parallel annealing: every rank holds all observations and fits its share of each speculative batch;
        # all ranks walk the same trajectory and end with the single-GPU result, bit for bit
        import tempfile
        ca = synth.make_config(8, 400, eps=0.05, seed=5, density=0.3)
        start = np.zeros((8, 8), np.int32)
        kw = dict(L=64, max_iter=40, update_step_size=20, max_iter_asa=30, seed=7)
        os.environ["MCCBN_ASA_BATCH"] = "8"
        ctx1 = mc.Context(local)
        ref = mc.adaptive_simulated_annealing(start, ca["obs"], outdir=tempfile.mkdtemp() + "/", ctx=ctx1, **kw)
        ctx1.close()
        for exchange in exchanges:
            ctx = mdist.init_context(local, dist, exchange=exchange)
            out = tempfile.mkdtemp() + "/"
            r = mc.adaptive_simulated_annealing(start, ca["obs"], outdir=out, ctx=ctx, **kw)
            same = (np.array_equal(r["poset"], ref["poset"]) and np.array_equal(r["lambda_"], ref["lambda_"]) and
                    r["eps"] == ref["eps"] and r["llhood"] == ref["llhood"])
            flag = torch.tensor([1.0 if same else 0.0], device=tdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            res[f"asa/{exchange}"] = {"identical_to_single_gpu_on_all_ranks": bool(flag.item() == 1.0),
                                      "rank0_wrote_files": (rank != 0) or os.path.exists(out + "params.txt")}
            say("asa", exchange, same)
            dist.barrier()
            ctx.close()
    except BaseException:
        # a rank that fails must not leave its peers waiting in a collective: take the whole job down
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
