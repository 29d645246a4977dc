"""dev helper: per-warp timelines of three back-to-back CHAINED launches (three independent drones), to see how early
launch i+1's CTAs get onto the SMs that launch i has left"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone
dev = 'cuda:0'; n = 1 << 20; K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
chained = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
ds = []
for j in range(3):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    g = torch.Generator(device=dev).manual_seed(1 + j)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    ds.append(d)
a = torch.rand(n, 4, device=dev) * 2 - 1
for _ in range(4):
    for d in ds: d.step(a, return_obs=False, chained=chained)
for d in ds:
    d._trace = torch.zeros(3 * 4 * 1024, dtype=torch.int64, device=dev); d._fast_ok = False
torch.cuda.synchronize()
torch.cuda._sleep(int(2e6))     # ~1 ms: all three launches are queued before the first one starts
for d in ds: d.step(a, return_obs=False, chained=chained)
torch.cuda.synchronize()
T = [d._trace.cpu().numpy().reshape(-1, 3) for d in ds]
T = [t[t[:, 0] > 0] for t in T]
t0 = min(t[:, 0].min() for t in T)
print(f"K={K} chained={chained}")
for j, t in enumerate(T):
    st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    # CTA = 4 consecutive warps
    cta_end = en.reshape(-1, 4).max(1); cta_start = st.reshape(-1, 4).min(1)
    pc = lambda x: " ".join(f"{np.percentile(x, p):6.1f}" for p in (0, 10, 50, 90, 100))
    print(f" launch {j}: warp start [{pc(st)}]  warp end [{pc(en)}]  CTA end [{pc(cta_end)}]  (us; percentiles 0 10 50 90 100)")
print(f" span of the three launches: {max(t[:,1].max() for t in T)/1e3 - t0/1e3:.1f} us -> {(max(t[:,1].max() for t in T) - t0)/3e3:.1f} us per launch")
