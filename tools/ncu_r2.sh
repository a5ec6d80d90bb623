#!/bin/bash
# Dev helper (GPU box), round 2: launch list of the bench command and one `ncu --set full` capture per E-step kernel on its
# BASELINE-sized workload.  Run AFTER the same commands have exited 0 without ncu.   usage: tools/ncu_r2.sh <tag>
tag=${1:-r2}
# 1. launch list of the bench command (per-launch times under ncu are cold-cache and serialised: the SHARE must agree)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/${tag}_ncu_launches.log 2>&1
run() {  # name, kernel regex, env, kbench config, launches to skip
  env $3 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${5:-1} -c 1 -f -o gpurun_out/prof_$1_$tag \
    python tools/kbench.py $4 > gpurun_out/ncu_$1.log 2>&1 || tail -5 gpurun_out/ncu_$1.log
  # the reports are ~30 MB each with sources: keep the raw-metric page as CSV, drop the report
  ncu -i gpurun_out/prof_$1_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw_$1.csv 2>/dev/null
  rm -f gpurun_out/prof_$1_$tag.ncu-rep
}
run fwd_spec "mccbn_fwd_spec" "MCCBN_JIT=1" cfg3
run fwd_generic "estep_forward_kernel" "MCCBN_JIT=0" cfg3
run cond_spec "mccbn_cond_spec" "MCCBN_JIT=1" cfg5
run pool "estep_pool_kernel" "MCCBN_JIT=0" cfg5pool
run fwd_spec_cfg4 "mccbn_fwd_spec" "MCCBN_JIT=1" cfg4
run tail "finalize_mstep_kernel" "MCCBN_JIT=1" cfg3
ls -la gpurun_out/${tag}_*
