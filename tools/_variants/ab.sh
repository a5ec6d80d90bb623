#!/bin/bash
V=tools/_variants
for lib in lib_base.so NEW; do
  if [ $lib = NEW ]; then unset MCCBN_B200_LIB; else export MCCBN_B200_LIB=$PWD/$V/$lib; fi
  MCCBN_JIT=1 timeout 120 python tools/kbench.py cfg3 cfg2 cfg4 cfg5 2>&1 | grep ms/iter
  MCCBN_JIT=0 timeout 120 python tools/kbench.py cfg3 cfg4 2>&1 | grep ms/iter
done
export MCCBN_B200_LIB=$PWD/$V/lib_diag.so
for c in 1 2 3 5 8 12; do echo "chunks=$c"; MCCBN_DIAG_CHUNKS=$c MCCBN_JIT=1 timeout 120 python tools/kbench.py cfg3 cfg4 cfg2 2>&1 | grep ms/iter; done
for c in 1 2 3; do echo "generic chunks=$c"; MCCBN_DIAG_CHUNKS=$c MCCBN_JIT=0 timeout 120 python tools/kbench.py cfg4 cfg2 2>&1 | grep ms/iter; done
