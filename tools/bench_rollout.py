#!/usr/bin/env python
"""Open-loop rollout throughput on one B200: T control steps per call on ONE 1,048,576-env batch, three ways --
plain steps (every launch waits for the previous grid), chained steps (FPV_F_CHAINED), and the fused rollout kernel
(fpv_drone_rollout: state in registers across the T steps).  Same workload as bench.py otherwise (drag + LUT + ground
contact + auto-reset, fresh random sticks every control step).  The action block of a call is T x 16 MiB, so the
working set exceeds the L2 for T >= 8.  CUDA events around whole calls; prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpyv_b200 import BatchedDrone  # noqa: E402


def main():
    dev, n = "cuda:0", 1 << 20
    out = {}
    for K, T in ((8, 16), (1, 32)):
        d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
        g = torch.Generator(device=dev).manual_seed(1)
        pos = torch.randn(n, 3, device=dev, generator=g) * 5
        pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
        d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
        acts = (torch.rand(T, n, 4, device=dev, generator=g) * 2 - 1).contiguous()
        dones = torch.empty((T, n), dtype=torch.uint8, device=dev)
        res = {}
        for name, fn in (("plain_steps", lambda: [d.step(acts[t], return_obs=False) for t in range(T)]),
                         ("chained_steps", lambda: d.rollout(acts, done_out=dones, fused=False)),
                         ("fused_rollout", lambda: d.rollout(acts, done_out=dones, fused=True))):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            res[name] = {"ms_per_control_step": ms / T, "env_steps_per_sec": n * T / (ms * 1e-3)}
        # algorithmic traffic of the fused kernel: state once each way + (action + done) per step
        b = n * (128 + 17 * T)
        res["fused_rollout"]["hbm_GBps_algorithmic"] = b / (res["fused_rollout"]["ms_per_control_step"] * T * 1e-3) / 1e9
        res["fused_rollout"]["fp32_TFLOPs_algorithmic"] = 252.0 * K * n * T / (res["fused_rollout"]["ms_per_control_step"] * T * 1e-3) / 1e12
        out[f"K{K}_T{T}"] = res
    print(json.dumps({"metric": "rollout_env_steps_per_sec", "envs": n, "results": out}))


if __name__ == "__main__":
    main()
