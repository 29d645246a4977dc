#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --profile 2>&1 | tail -1 | cut -c1-120; }
run FPV_TUNE_STATIC=1
run FPV_TUNE_STATIC=0
run FPV_TUNE_STATIC=0 FPV_TUNE_STAGES=3
timeout 200 python tune_trace.py 2>&1 | grep -E "^K=|end us"
