MCCBN_JIT=1 python tools/kbench.py cfg5 > /dev/null 2>&1
MCCBN_JIT=1 ncu --set full --clock-control none -k "regex:mccbn_cond_spec" -c 1 -f -o gpurun_out/prof_cond_spec_r1f python tools/kbench.py cfg5 > gpurun_out/ncu_cond.log 2>&1
ncu -i gpurun_out/prof_cond_spec_r1f.ncu-rep --page raw --csv > gpurun_out/prof_cond_spec_r1f.raw.csv 2>/dev/null
rm -f gpurun_out/prof_cond_spec_r1f.ncu-rep
ls -la gpurun_out/prof_cond_spec_r1f.raw.csv
