for o in 4 3 2 1; do echo "occ=$o chained"; FPV_TUNE_OCC=$o python bench.py --profile --steps 100 --warmup 10 2>&1 | tail -1 | cut -c1-110; done
