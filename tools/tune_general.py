import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from fpyv_b200 import BatchedDrone, Cylinder, Ground, Target
dev, n = "cuda:0", 1 << 20
def run(label, K, objs, **kw):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, thrust_lut=2049, **kw)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 8; pos[:, 2] = 20 + torch.rand(n, device=dev, generator=g) * 8
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    act = (torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous(); act[:, 3] = -0.6
    for _ in range(3): d.step(act, None, objs, return_obs=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.step(act, None, objs, return_obs=False)
    e1.record(); torch.cuda.synchronize()
    print(f"{label:40s} K={K}: {e0.elapsed_time(e1)/10*1e3:8.1f} us")
far = [Target(np.array([500.0, 0, 3]), 1.0), Ground()]
rng = np.random.default_rng(5)
six = [Target(np.array([0.0, 0.0, 3.0]), 1.0)] + [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0) for _ in range(5)] + [Ground()]
for K in (1, 8):
    run("hot path (ground only)", K, None)
    run("general, no ground no objects", K, None, ground=False)
    run("general, one far sphere + ground", K, far)
    run("general, 6 objects, drones high above", K, six)
