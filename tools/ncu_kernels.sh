#!/bin/bash
# Dev helper (GPU box): one `ncu --set full` capture of each remaining E-step kernel on its BASELINE-sized workload.
# usage: tools/ncu_kernels.sh <tag>     -> gpurun_out/prof_<kernel>_<tag>.ncu-rep
set -e
tag=${1:-r1}
run() {  # name, kernel regex, env, kbench config
  env $3 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -f -o gpurun_out/prof_$1_$tag \
    python tools/kbench.py $4 > gpurun_out/ncu_$1.log 2>&1 || tail -5 gpurun_out/ncu_$1.log
  # the reports are ~30 MB each with sources: keep the raw-metric page as CSV, drop the report
  ncu -i gpurun_out/prof_$1_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$1_$tag.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$1_$tag.ncu-rep
}
run fwd_generic "estep_forward_kernel" "MCCBN_JIT=0" cfg3
run backward "estep_conditional_kernel" "MCCBN_JIT=0" cfg5
run pool "estep_pool_kernel" "MCCBN_JIT=0" cfg5pool
run fwd_cfg4 "estep_forward_kernel" "MCCBN_JIT=0" cfg4
ls -la gpurun_out/*.raw.csv
