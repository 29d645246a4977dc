#!/bin/bash
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --profile 2>&1 | tail -1 | cut -c1-200; }
for lib in tune/*.so; do run FPYV_B200_LIB=$PWD/$lib; done
