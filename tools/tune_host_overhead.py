import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from fpyv_b200 import BatchedDrone
dev, n = "cuda:0", 1 << 20
ds = []
for j in range(4):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=2)
    g = torch.Generator(device=dev).manual_seed(j)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    ds.append(d)
a = torch.rand(n, 4, device=dev) * 2 - 1
for i in range(20): ds[i % 4].step(a, return_obs=False, chained=True)
torch.cuda.synchronize()
for chained in (True, False):
    torch.cuda._sleep(int(4e7))     # GPU busy: measure pure enqueue cost
    t0 = time.perf_counter()
    for i in range(200): ds[i % 4].step(a, return_obs=False, chained=chained)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"chained={chained}: host enqueue {1e6*(t1-t0)/200:.1f} us per step()")
