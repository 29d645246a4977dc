#!/usr/bin/env python
"""Run a few launches of ONE ring-kernel mode at its bench size (for `ncu -k regex:ring_step_kernel -s 3 -c 1 --set full`).
usage: profile_modes.py {hot|general_clear|general_contact|gate_race|racer|acro|acro_k1} [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedAcroDrone, BatchedDrone, BatchedRacer, Cylinder, Ground, Target  # noqa: E402
from fpyv_b200.env import GateRaceEnv  # noqa: E402

mode = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n = 1 << 20
if mode in ("hot", "general_clear", "general_contact"):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    if mode == "general_clear":
        ang = torch.rand(n, device=dev, generator=g) * 6.2831853
        rad = 60.0 + 20.0 * torch.rand(n, device=dev, generator=g)
        pos = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang), 0.3 + torch.rand(n, device=dev, generator=g) * 8], 1)
    else:
        pos = torch.randn(n, 3, device=dev, generator=g) * (8.0 if mode == "general_contact" else 5.0)
        pos[:, 2] = 0.3 + torch.rand(n, device=dev, generator=g) * (8.0 if mode == "general_contact" else 2.7)
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    rng = np.random.default_rng(5)
    objs = None if mode == "hot" else [Target(np.array([0.0, 0.0, 3.0]), 1.0)] + \
        [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0) for _ in range(5)] + [Ground()]
    act = (torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous()
    fn = lambda: d.step(act, None, objs, return_obs=False)
elif mode == "gate_race":
    env = GateRaceEnv(None, num_envs=8192, agents_per_env=32, device=dev, substeps=8, dt=1e-3, thrust_lut=2049)
    env.reset()
    a = torch.rand(8192, 32, 4, device=dev, generator=g) * 2 - 1
    a[..., 3] = a[..., 3] * 0.3 - 0.3
    env.step(a)
    fn = lambda: env.step(a, fused=True)
elif mode == "racer":
    r = BatchedRacer(5, {"roll": [2, 0.1, 1e-4], "pitch": [2, 0.1, 1e-4], "yaw": [0.1, 0, 0]}, num_envs=n, device=dev, dt=1e-3, substeps=8)
    r.reset()
    a = torch.cat([torch.rand(n, 3, device=dev, generator=g) * 6 - 3, torch.rand(n, 1, device=dev, generator=g) * 10], 1).contiguous()
    fn = lambda: r.step(a)
else:
    K = 1 if mode == "acro_k1" else 8
    d = BatchedAcroDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5
    pos[:, 2] = 0.3 + torch.rand(n, device=dev, generator=g) * 5
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 20)
    a = (torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous()
    fn = lambda: d.step(a)
for _ in range(steps):
    fn()
torch.cuda.synchronize()
print("ok", mode)
