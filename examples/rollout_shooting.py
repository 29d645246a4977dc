#!/usr/bin/env python
"""Random-shooting MPC with the fused rollout: 262,144 candidate stick sequences of 25 control steps (8 substeps of 1 ms
each) from one start state, ONE launch per evaluation (state in registers across the 25 steps), pick the sequence that
ends closest to a goal without crashing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone  # noqa: E402

n, T, dev = 1 << 18, 25, "cuda:0"
drone = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, thrust_lut=2049)
start = torch.tensor([0.0, 0.0, 2.0], device=dev)
goal = torch.tensor([0.6, 0.3, 2.3], device=dev)
g = torch.Generator(device=dev).manual_seed(0)
best = None
for it in range(4):
    drone.reset(start.expand(n, 3), torch.zeros(n, 3, device=dev), torch.zeros(n, 3, device=dev))
    sticks = torch.rand(T, n, 4, device=dev, generator=g) * 2 - 1
    sticks[..., 3] = sticks[..., 3] * 0.5 - 0.4                       # throttle around hover
    if best is not None:                                              # keep the incumbent in slot 0
        sticks[:, 0] = best
    flags = torch.zeros(T, n, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    drone.rollout(sticks.contiguous(), done_out=flags)
    e1.record()
    cost = (drone.position - goal).norm(dim=1) + 100.0 * flags.any(dim=0)
    k = int(cost.argmin())
    best = sticks[:, k].clone()
    torch.cuda.synchronize()
    print(f"iteration {it}: best end distance {cost[k].item():.3f} m; {n * T / (e0.elapsed_time(e1) * 1e-3):.3g} env-steps/s "
          f"({e0.elapsed_time(e1):.2f} ms for {n} x {T} steps)")
