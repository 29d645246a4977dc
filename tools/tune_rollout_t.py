import os, sys, torch
sys.path.insert(0, os.getcwd())
from fpyv_b200 import BatchedDrone
dev, n = "cuda:0", 1 << 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); fr = torch.ones(64 << 20, device=dev)
for K in (8, 1):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    for T in (1, 2, 4):
        acts = (torch.rand(T, n, 4, device=dev, generator=g) * 2 - 1).contiguous()
        for name, fn in (("step x T", lambda: [d.step(acts[t], return_obs=False) for t in range(T)]), ("fused", lambda: d.rollout(acts, fused=True))):
            for _ in range(3): fn()
            ts = []
            for _ in range(15):
                flush.zero_(); fr.sum()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print(f"K={K} T={T} {name:9s}: {sorted(ts)[7]*1e3/T:6.1f} us per control step (isolated call, flushed L2)")
