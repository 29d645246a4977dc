"""dev helper: per-warp cycles by phase (needs a -DFPV_TRACE_PHASES build via FPYV_B200_LIB)"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone
dev='cuda:0'; n=1<<20
for K in (8, 1, 32):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    a = torch.rand(n, 4, device=dev, generator=g) * 2 - 1
    fl = torch.ones(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(5): d.step(a, return_obs=False)
    W = 592 * 4
    d._trace = torch.zeros(4 * W, dtype=torch.int64, device=dev); d._fast_ok = False
    fl.sum(); torch.cuda.synchronize()
    d.step(a, return_obs=False); torch.cuda.synchronize()
    t = d._trace.cpu().numpy()
    ph = t[:3 * W].reshape(W, 3).astype(np.float64); tot = t[3 * W:].astype(np.float64)
    print(f"K={K}: per-warp cycles (mean over {W} warps): total loop {tot.mean():.0f}  wait {ph[:,0].mean():.0f}  read+unpack-issue {ph[:,1].mean():.0f}  tile(substeps+epilogue) {ph[:,2].mean():.0f};  chunks/warp {16384/W:.2f}")
    print(f"      per chunk: wait {ph[:,0].mean()/6.92:.0f}  read {ph[:,1].mean()/6.92:.0f}  tile {ph[:,2].mean()/6.92:.0f}   (max total {tot.max():.0f}, min {tot.min():.0f})")
