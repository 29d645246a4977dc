import os, sys, torch
sys.path.insert(0, os.getcwd())
from fpyv_b200 import BatchedDrone
dev, n = "cuda:0", 1 << 20
d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
g = torch.Generator(device=dev).manual_seed(1)
pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
ha = [torch.empty(n, 4, dtype=torch.float32, pin_memory=True).uniform_(-1, 1) for _ in range(2)]
hd = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for s in (1, 2, 4, 8, 16):
    for i in range(5): d.step_host(ha[i % 2], hd, slices=s); torch.cuda.current_stream().synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        d.step_host(ha[i % 2], hd, slices=s); torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    print(f"slices={s:2d}: {e0.elapsed_time(e1)/50*1e3:7.1f} us per step")
