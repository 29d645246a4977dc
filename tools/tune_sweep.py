"""dev helper: flush-mode and size sweeps on the GPU box"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from fpyv_b200 import BatchedDrone
dev = 'cuda:0'
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fr = torch.ones(64 << 20, dtype=torch.float32, device=dev)
def flush(mode):
    if mode in ('write', 'write+read'): fl.zero_()
    if mode in ('read', 'write+read'): fr.sum()
def timeit(n, K, mode, steps=30, lut=2049):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=lut)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    acts = [torch.rand(n, 4, device=dev, generator=g) * 2 - 1 for _ in range(4)]
    for i in range(5): d.step(acts[i % 4], return_obs=False)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        flush(mode)
        ev[i][0].record(); d.step(acts[i % 4], return_obs=False); ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3
def copyit(nbytes, mode, steps=30):
    a = torch.empty(nbytes // 4, dtype=torch.float32, device=dev); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for i in range(steps):
        flush(mode)
        ev[i][0].record(); b.copy_(a); ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in ev)
    return ts[len(ts) // 2] * 1e3
n = 1 << 20
for mode in ('none', 'write', 'read', 'write+read'):
    c = copyit(76 << 20, mode)
    print(f"flush={mode:10s} copy 76MB->76MB: {c:6.1f} us ({2*76*1.048576/c*1e3:.0f} GB/s)   K1: {timeit(n,1,mode):6.1f} us   K8: {timeit(n,8,mode):6.1f} us")
for K in (1, 4, 8, 16, 32):
    print(f"K={K}: read-flush {timeit(n, K, 'read'):.1f} us")
