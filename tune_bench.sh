#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --profile 2>&1 | tail -1; }
for lib in tune/lib_q_mb4.so tune/lib_q_mb5.so; do
 for lut in 2049 1025; do
  run FPYV_B200_LIB=$PWD/$lib FPV_BENCH_LUT=$lut
 done
done
run FPYV_B200_LIB=$PWD/tune/lib_q_mb5.so FPV_BENCH_LUT=1025 FPV_TUNE_STAGGER_NS=1500
