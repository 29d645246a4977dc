#!/bin/bash
# dev helper: A/B builds + tuning switches on the GPU box
python - <<'PY'
import torch
a=torch.empty(1<<29,dtype=torch.bfloat16,device='cuda'); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
best=1e9
for _ in range(10):
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
print(f"copy bandwidth {2*a.numel()*2/best/1e6:.0f} GB/s")
PY
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --profile 2>&1 | tail -1; }
for lib in tune/lib_ctaring.so tune/lib_warpring.so; do
  run FPYV_B200_LIB=$PWD/$lib FPV_TUNE_NOTMA=1
  run FPYV_B200_LIB=$PWD/$lib FPV_TUNE_STAGES=2
done
