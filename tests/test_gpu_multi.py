"""N > 1: the peer-memory statistic exchange fused into the tail kernel of an EM iteration and the NCCL allreduce path,
against the unsharded single-GPU run.  The world-2/4/8 tests need that many devices (skipped on a one-GPU box); the
`test_sharded_em_two_ranks_one_gpu` runs the same exchange kernels with TWO ranks on device 0 -- two processes whose
mailboxes are mapped through CUDA IPC -- so that the exchange path is exercised on a one-GPU box as well."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_em_matches_single_gpu(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() + world) % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_RESULT ")][-1]
    res = json.loads(line[len("MULTI_GPU_RESULT "):])
    assert set(res) == {"forward/peer", "forward/nccl", "backward/peer", "backward/nccl", "asa/peer", "asa/nccl"}
    for key in ("asa/peer", "asa/nccl"):  # candidate-parallel annealing: the single-GPU result on every rank
        v = res.pop(key)
        assert v["identical_to_single_gpu_on_all_ranks"] and v["rank0_wrote_files"], (key, v)
    for key, v in res.items():
        assert v["identical_across_ranks"], key
        # same samples (Philox counters carry the global observation index); only the fp64 summation order differs
        assert v["max_rel_vs_single_gpu"] < 1e-9 and v["eps_abs"] < 1e-11 and v["ll_rel"] < 1e-10, (key, v)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_in_process_team(world):
    """mccbn_set_gpus: ONE process (the reference's callers are an R session) shards over `world` devices itself."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MCCBN_PEER_TIMEOUT_S="60")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_team_worker.py"), str(world)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("TEAM_RESULT ")][-1]
    res = json.loads(line[len("TEAM_RESULT "):])
    for sampling in ("forward", "backward"):
        v = res[sampling]
        assert v["lambda_rel"] < 1e-9 and v["eps_abs"] < 1e-11 and v["ll_rel"] < 1e-10 and v["obs_ll_rel"] < 1e-10, v
        assert v["iw_w_rel"] < 1e-10 and v["iw_dist_abs"] < 1e-10 and v["iw_Tdiff_rel"] < 1e-9, v
    assert res["weights_times_ll_rel"] < 1e-10 and res["after_error_ll_rel"] < 1e-10, res
    assert res["zero_weight_error"][0] == 2 and "weight 0" in res["zero_weight_error"][1], res
    assert res["asa_identical"], res


def test_sharded_em_two_ranks_one_gpu():
    """Two processes (torch.distributed over gloo), both on device 0: observations sharded, statistic vectors exchanged
    through CUDA-IPC mapped mailboxes by the fused tail kernels, replicated M-step -- against the unsharded run."""
    if _n_gpus() < 1:
        pytest.skip("needs a GPU")
    port = 29900 + os.getpid() % 90
    env = dict(os.environ, MCCBN_TEST_ONE_DEVICE="1", MCCBN_PEER_TIMEOUT_S="120")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_RESULT ")][-1]
    res = json.loads(line[len("MULTI_GPU_RESULT "):])
    assert set(res) == {"forward/peer", "backward/peer", "asa/peer"}
    v = res.pop("asa/peer")
    assert v["identical_to_single_gpu_on_all_ranks"] and v["rank0_wrote_files"], v
    for key, v in res.items():
        assert v["identical_across_ranks"], key
        assert v["max_rel_vs_single_gpu"] < 1e-9 and v["eps_abs"] < 1e-11 and v["ll_rel"] < 1e-10, (key, v)
