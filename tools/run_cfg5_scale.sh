#!/bin/bash
# Dev helper (`gpurun --gpus 4 -- tools/run_cfg5_scale.sh <tag>`): BASELINE config 5 (backward k = 1, and the pool scheme) at
# N = 4 / 2 / 1 GPUs; N = 8 comes from tools/run_multi8.sh.
tag=${1:-r2p}
port=29700
for c in cfg5 cfg5pool; do
  for n in 4 2; do
    port=$((port + 1))
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 5 --warmup 3 --config $c > gpurun_out/${tag}_scale${n}_$c.json 2> gpurun_out/${tag}_scale${n}_$c.err
    echo "$c N=$n rc=$?"
  done
  timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --config $c --no-cpu-baseline --no-other > gpurun_out/${tag}_scale1_$c.json 2> gpurun_out/${tag}_scale1_$c.err
  echo "$c N=1 rc=$?"
done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/%s_scale*_cfg5*.json" % sys.argv[1])):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], "%.3f ms" % d["ms_per_step"], "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d.get("check"))
PY
