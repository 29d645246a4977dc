#!/usr/bin/env python
"""Measurement of the two other dynamics modes on one B200 at 1,048,576 envs: mode B (`Racer`, tests/racer_drone_test.py)
and mode C (acro: rate PID -> mixer -> per-motor LUT thrust -> rigid body; parity unpinned).  CUDA events, L2 flushed
before every timed step; prints one JSON line.  Both kernels are plain one-env-per-thread grids with 7 float4 planes:
algorithmic bytes per env per control step = 7*16*2 (state) + 16 (action) [+ 16 torque / motor-thrust output + 1 done]."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpyv_b200 import BatchedAcroDrone, BatchedDrone, BatchedRacer, Cylinder, Ground, Target  # noqa: E402


def timeit(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    dev, n = "cuda:0", 1 << 20
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    g = torch.Generator(device=dev).manual_seed(2)
    out = {}
    for K in (1, 8):
        r = BatchedRacer(5, {"roll": [2, 0.1, 1e-4], "pitch": [2, 0.1, 1e-4], "yaw": [0.1, 0, 0]}, num_envs=n, device=dev, dt=1e-3, substeps=K)
        r.reset()
        a = torch.cat([torch.rand(n, 3, device=dev, generator=g) * 6 - 3, torch.rand(n, 1, device=dev, generator=g) * 10], dim=1).contiguous()
        ms = timeit(lambda: r.step(a), flush)
        b = n * (7 * 16 * 2 + 16 + 16)
        out[f"racer_K{K}"] = {"ms_per_step": ms, "env_steps_per_sec": n / (ms * 1e-3), "hbm_GBps": b / (ms * 1e-3) / 1e9,
                              "hbm_frac": b / (ms * 1e-3) / 1e9 / peak}
        d = BatchedAcroDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True)
        pos = torch.randn(n, 3, device=dev, generator=g) * 5
        pos[:, 2] = 0.3 + torch.rand(n, device=dev, generator=g) * 5
        d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 20)
        act = (torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous()
        ms = timeit(lambda: d.step(act), flush)
        b = n * (7 * 16 * 2 + 16 + 16 + 1)
        out[f"acro_K{K}"] = {"ms_per_step": ms, "env_steps_per_sec": n / (ms * 1e-3), "env_substeps_per_sec": n * K / (ms * 1e-3),
                             "hbm_GBps": b / (ms * 1e-3) / 1e9, "hbm_frac": b / (ms * 1e-3) / 1e9 / peak}
    # mode A on the GENERAL path: the stock world of params.yaml (5 cylinders, one spherical target, ground plane)
    rng = np.random.default_rng(5)
    objs = [Target(np.array([0.0, 0.0, 3.0]), 1.0)] + [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0) for _ in range(5)] + [Ground()]
    for K in (1, 8):
        d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
        pos = torch.randn(n, 3, device=dev, generator=g) * 8
        pos[:, 2] = 0.3 + torch.rand(n, device=dev, generator=g) * 8
        d.reset(pos, torch.randn(n, 3, device=dev, generator=g) * 2, (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
        act = (torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous()
        ms = timeit(lambda: d.step(act, None, objs, return_obs=False), flush)
        b = n * 145
        out[f"drone_obstacles_K{K}"] = {"ms_per_step": ms, "env_steps_per_sec": n / (ms * 1e-3), "env_substeps_per_sec": n * K / (ms * 1e-3),
                                        "hbm_GBps": b / (ms * 1e-3) / 1e9, "hbm_frac": b / (ms * 1e-3) / 1e9 / peak, "objects": 6,
                                        "crashes": d.episode_stats()["crashes"]}
    print(json.dumps({"metric": "mode_B_C_env_steps_per_sec", "envs": n, "hbm_peak_GBps": peak, "results": out}))


if __name__ == "__main__":
    main()
