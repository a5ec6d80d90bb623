"""CPU, world_size = 2, gloo: the N>1 host logic -- observation sharding, the (p+5)-double statistic exchange and the
replicated M-step -- checked against the unsharded oracle E-step; plus the unique-id bootstrap plumbing."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle as orc
    from mccbn_b200 import dist as mdist, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = synth.make_config(9, 101, eps=0.05, seed=4, density=0.3)  # 101: ragged split
        rng = np.random.default_rng(1)
        L, p = 200, 9
        Z = rng.exponential(1.0, (L, p)) / c["lambdas"]
        Ts = rng.exponential(1.0, L)
        wts = rng.uniform(0.5, 2.0, 101)
        lo, hi = mdist.shard_bounds(101, rank, world)
        r = orc.estep_forward_from_times(c["obs"][lo:hi], c["poset"], c["lambdas"], 0.1, Z, Ts, weights=wts[lo:hi])
        local = np.concatenate([[r["obs_llhood"], r["N_eff"], r["expected_dist"], 0.0, 0.0], r["Tdiff_colsum"]])
        tot = mdist.combine_statistics(local, dist)
        lam, eps, ll = mdist.m_step(tot, p)
        full = orc.estep_forward_from_times(c["obs"], c["poset"], c["lambdas"], 0.1, Z, Ts, weights=wts)
        ok = (np.allclose(lam, full["lambda_new"], rtol=1e-12) and abs(eps - full["eps_new"]) < 1e-14 and
              abs(ll - full["obs_llhood"]) < 1e-10 * abs(ll))
        # bootstrap plumbing: a 128-byte id created on rank 0 arrives unchanged everywhere
        ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        ok = ok and ident.tolist() == list(range(128))
        q.put((rank, lo, hi, bool(ok), lam.tolist()))
    finally:
        dist.destroy_process_group()


def test_sharded_statistics_match_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 101  # contiguous cover, no overlap
    assert all(r[3] for r in res)
    assert res[0][4] == res[1][4]  # bit-identical parameters on both ranks (replicated M-step)


def test_shard_bounds_cover():
    from mccbn_b200.dist import shard_bounds
    for N in (1, 7, 100, 100000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(N, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
