"""dev helper: host-side cost of one step() call vs device time"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone
dev='cuda:0'
for n, K in ((1<<20, 8), (1<<20, 1), (4096, 8)):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    d.reset()
    a = torch.zeros(n, 4, device=dev)
    for _ in range(20): d.step(a, return_obs=False)
    torch.cuda.synchronize()
    N = 2000
    t0 = time.perf_counter()
    for _ in range(N): d.step(a, return_obs=False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n} K={K}: host issue {1e6*(t1-t0)/N:.1f} us/step, total {1e6*(t2-t0)/N:.1f} us/step")
