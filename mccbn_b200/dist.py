"""One-process-per-GPU plumbing on top of torch.distributed (NCCL over NVLink): observation sharding and the
bootstrap of the library's own NCCL communicator.

The path shards by observation (the reference's OpenMP loop over i, src/mcem_hcbn.cpp:665-700, has only `+`
reductions): rank r holds rows [lo, hi) of obs / times / weights, runs the E-step on them with GLOBAL row indices in
the Philox counters, and one exchange of the p+5 statistic vector [LL, N_eff, D, zero_count, 0, S_0..S_{p-1}] per EM
iteration precedes the replicated on-device M-step: either fused into the library's reduction and M-step kernels over
NVLink peer memory (default on one box) or one ncclAllReduce issued by the C library on its own stream.
"""
import numpy as np


def shard_bounds(N, rank, world):
    """Contiguous block partition of N observations (the OpenMP schedule(static) partition of the reference)."""
    return (N * rank) // world, (N * (rank + 1)) // world


def combine_statistics(local_stats, dist=None):
    """Sum-allreduce of the statistic vector (what ncclAllReduce does on the device); numpy/torch CPU version used by
    the gloo tests and by callers that drive the E-step themselves."""
    import torch
    t = torch.as_tensor(np.asarray(local_stats, dtype=np.float64)).clone()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def m_step(stats, p, max_lambda=1e6):
    """The replicated M-step on the combined statistics (src/mcem_hcbn.cpp:707-709, model_impl.cpp:21-34)."""
    LL, N_eff, D = stats[0], stats[1], stats[2]
    S = np.asarray(stats[5:5 + p])
    eps = max(min(D / (N_eff * p), 1 - np.finfo(float).eps), np.finfo(float).eps)
    lam = 1.0 / (S / N_eff)
    lam = np.where((lam > np.float32(max_lambda)) | ~np.isfinite(lam), float(np.float32(max_lambda)), lam)
    return lam, eps, LL


def broadcast_unique_id(dist, device=None):
    """Rank 0 creates the NCCL unique id of the library's communicator; everyone receives it through the already
    initialised torch.distributed group (gloo or nccl)."""
    import torch
    from . import api
    rank = dist.get_rank()
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf = torch.tensor(list(api.nccl_unique_id()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def gather_peer_handles(ctx, dist, device=None):
    """All-gather of the 64-byte CUDA IPC handles of the ranks' statistic mailboxes, in rank order."""
    import torch
    mine = torch.tensor(list(ctx.peer_handle()), dtype=torch.uint8, device=device)
    parts = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return b"".join(bytes(t.cpu().tolist()) for t in parts)


def init_context(local_rank, dist=None, exchange="peer"):
    """Context on cuda:local_rank for one rank of the torch.distributed world.

    exchange="peer": the statistic exchange runs inside the reduction / M-step kernels over NVLink peer memory (one box);
    exchange="nccl": one ncclAllReduce per EM iteration on the library's own communicator (also across boxes)."""
    from . import api
    ctx = api.Context(local_rank)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        dev = torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else None
        if exchange == "peer":
            ctx.init_peers(dist.get_rank(), dist.get_world_size(), gather_peer_handles(ctx, dist, dev))
        else:
            ctx.init_comm(dist.get_rank(), dist.get_world_size(), broadcast_unique_id(dist, dev))
    return ctx
