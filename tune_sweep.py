"""dev helper: T(K) sweeps on the GPU box"""
import sys, torch, json
sys.path.insert(0, '.')
from fpyv_b200 import BatchedDrone
dev = 'cuda:0'
def timeit(n, K, packed=True, lut=2049, flush=True, steps=30, auto_reset=True, stats=True):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=auto_reset, thrust_lut=lut, packed=packed)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    acts = [torch.rand(n, 4, device=dev, generator=g) * 2 - 1 for _ in range(4)]
    fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(5): d.step(acts[i % 4], return_obs=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        if flush: fl.zero_()
        ev[i][0].record(); d.step(acts[i % 4], return_obs=False); ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3
n = 1 << 20
for packed in (True, False):
    for K in (1, 2, 4, 8, 16, 32):
        med, mn = timeit(n, K, packed)
        print(f"n=1M packed={packed} K={K:2d}: median {med:7.1f} us  min {mn:7.1f} us  -> {n*K/med*1e6:.3e} substeps/s")
for K in (1, 8):
    med, mn = timeit(n, K, True, flush=False)
    print(f"n=1M packed noflush K={K}: median {med:7.1f} us min {mn:7.1f}")
    med, mn = timeit(n, K, True, lut=0)
    print(f"n=1M packed nolut K={K}: median {med:7.1f} us min {mn:7.1f}")
    med, mn = timeit(4 * n, K, True)
    print(f"n=4M packed K={K}: median {med:7.1f} us min {mn:7.1f} -> {4*n*K/med*1e6:.3e} substeps/s")
