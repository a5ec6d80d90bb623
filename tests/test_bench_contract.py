"""CPU-only: the reference arm of bench.py (`--impl reference`) -- the restated reference loop on the host cores.

The driver launches it like our arm (torchrun for N > 1), and torchrun exports OMP_NUM_THREADS=1 to its workers: the arm
must still use every core it may run on, rank 0 alone runs and prints, the other ranks exit 0 without output."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1", "--ref-seconds", "0.5", *flags], capture_output=True, text=True, env=env,
                          timeout=300)


def test_reference_arm_line_and_thread_count():
    r = _run({"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["config"]["workload"].startswith("cfg3") and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "observations" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))  # not the 1 thread OMP_NUM_THREADS asked for
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
