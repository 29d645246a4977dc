for o in 1 2 3 4; do echo "occ=$o"; FPV_TUNE_OCC=$o FPV_BENCH_CHAINED=0 python bench.py --profile --steps 20 --warmup 5 --envs 4194304 2>&1 | tail -1 | cut -c1-150; done
