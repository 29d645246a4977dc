"""dev helper: per-warp start/end timeline of the hot kernel"""
import os, sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from fpyv_b200 import BatchedDrone
dev='cuda:0'; n=1<<20
for K in (8,):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    g = torch.Generator(device=dev).manual_seed(1)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5; pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    a = torch.rand(n, 4, device=dev, generator=g) * 2 - 1
    fl = torch.ones(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(5): d.step(a, return_obs=False)
    d._trace = torch.zeros(3 * 4 * 1024, dtype=torch.int64, device=dev); d._fast_ok = False
    fl.sum(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.step(a, return_obs=False); e1.record(); torch.cuda.synchronize()
    t = d._trace.cpu().numpy().reshape(-1, 3); t = t[t[:, 0] > 0]
    np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"trace_k{K}.npy"), t)
    t0 = t[:, 0].min()
    st, en, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
    print(f"K={K}: event time {e0.elapsed_time(e1)*1e3:.1f} us; warps {len(t)}; start us: min {st.min():.1f} p50 {np.median(st):.1f} p90 {np.percentile(st,90):.1f} max {st.max():.1f}")
    print(f"   end us: min {en.min():.1f} p10 {np.percentile(en,10):.1f} p50 {np.median(en):.1f} p90 {np.percentile(en,90):.1f} max {en.max():.1f}; dur us: min {(en-st).min():.1f} p50 {np.median(en-st):.1f} max {(en-st).max():.1f}")
    # per SM: number of warps, span
    per = {}
    for s_, e_, m_ in zip(st, en, sm): per.setdefault(int(m_), []).append((s_, e_))
    cnt = np.array([len(v) for v in per.values()])
    print(f"   SMs {len(per)}; warps/SM min {cnt.min()} max {cnt.max()}; per-SM last end: min {min(max(e for _,e in v) for v in per.values()):.1f} max {max(max(e for _,e in v) for v in per.values()):.1f}")
    print(f"   warp-residency utilisation: {(en-st).sum()/(en.max()*len(t)):.3f} (sum of warp durations / (last end x warps))")
    hist, edges = np.histogram(en, bins=12)
    print("   end histogram:", list(zip(np.round(edges[:-1],1), hist)))
    hist, edges = np.histogram(st, bins=8)
    print("   start histogram:", list(zip(np.round(edges[:-1],1), hist)))
