import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from fpyv_b200 import BatchedDrone
dev, n = "cuda:0", 1 << 20
ds = []
for j in range(4):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=2)
    g = torch.Generator(device=dev).manual_seed(j)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    d._epoch = 0xFFFFFFF0 if j == 0 else 0     # drone 0 crosses the uint32 wrap of the epoch counter
    d._chunk_epoch.fill_(-16 if j == 0 else 0)
    ds.append(d)
acts = [torch.rand(n, 4, device=dev) * 2 - 1 for _ in range(4)]
t0 = time.time()
for i in range(20000):
    ds[i % 4].step(acts[i % 4], return_obs=False, chained=True)
torch.cuda.synchronize()
print(f"20,000 chained steps in {time.time()-t0:.2f} s; epochs", [int(d._chunk_epoch.view(torch.int64 if False else torch.int32)[0]) & 0xFFFFFFFF for d in ds],
      "; nonfinite", [d.episode_stats()["nonfinite"] for d in ds], "; env_steps", [d.episode_stats()["env_steps"] for d in ds])
# same-drone chain across the wrap as well
d = ds[0]
for i in range(64): d.step(acts[i % 4], return_obs=False, chained=True)
torch.cuda.synchronize(); print("same-drone chain ok, epoch", int(d._chunk_epoch[0]) & 0xFFFFFFFF)
