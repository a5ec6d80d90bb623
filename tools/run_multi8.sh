#!/bin/bash
# Dev helper (GPU box with 8 GPUs, `gpurun --gpus 8 -- tools/run_multi8.sh <tag>`): the multi-GPU evidence of a round in one
# call -- tests/test_gpu_multi.py (worlds 2/4/8, both exchanges, in-process team), the headline bench at N = 8/4/2/1, the
# other sharded workloads at N = 8, the in-process team probe and candidate-parallel annealing.
tag=${1:-r2g}
port=29650
tr() { n=$1; shift; port=$((port + 1)); timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/${tag}_multi8.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_multi8.log; tail -4 gpurun_out/${tag}_multi8.log
for n in 8 4 2; do
  tr $n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  echo "bench N=$n rc=$?"; grep -o '"value": [0-9.e+]*, "unit": "samples/s", "n_gpus": [0-9]*' gpurun_out/${tag}_scale_$n.json | head -1
done
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/${tag}_scale_1.json 2> gpurun_out/${tag}_scale_1.err
echo "bench N=1 rc=$?"
for c in cfg5 cfg5pool; do
  tr 8 bench.py --gpus 8 --steps 5 --warmup 3 --config $c > gpurun_out/${tag}_scale8_$c.json 2> gpurun_out/${tag}_scale8_$c.err
  echo "bench $c N=8 rc=$?"
done
tr 4 bench.py --gpus 4 --steps 5 --warmup 3 --config cfg5 > gpurun_out/${tag}_scale4_cfg5.json 2> gpurun_out/${tag}_scale4_cfg5.err
echo "bench cfg5 N=4 rc=$?"
timeout 200 python tools/team_probe.py cfg3 8 > gpurun_out/${tag}_team_probe_8gpu.log 2>&1; tail -5 gpurun_out/${tag}_team_probe_8gpu.log | cut -c1-400
tr 8 tools/asa_probe_multi.py > gpurun_out/${tag}_asa_8gpu.log 2>&1; tail -3 gpurun_out/${tag}_asa_8gpu.log
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2g"
for f in sorted(glob.glob("gpurun_out/%s_scale*.json" % tag)):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], "%.3f ms" % d["ms_per_step"], "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d.get("gpu_launches"), d.get("check"))
PY
