#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --profile 2>&1 | tail -1; }
run FPV_TUNE_NOTMA=1
run FPV_TUNE_STAGES=2
run FPV_TUNE_STAGES=3
