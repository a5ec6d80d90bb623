#!/bin/bash
V=tools/_variants
for rep in 1 2; do
for lib in lib_base.so NEW; do
  if [ $lib = NEW ]; then unset MCCBN_B200_LIB; else export MCCBN_B200_LIB=$PWD/$V/$lib; fi
  MCCBN_JIT=1 timeout 120 python tools/kbench.py cfg3 cfg2 cfg4 cfg5 2>&1 | grep ms/iter
  MCCBN_JIT=0 timeout 120 python tools/kbench.py cfg3 cfg4 cfg2 2>&1 | grep ms/iter
done
done
