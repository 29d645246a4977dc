#!/usr/bin/env python
"""Closed-loop stepping at the open-loop rate: the population is split into two independent halves on two CUDA streams.
Each stream is in plain order -- policy(state) -> step -> policy(state) -> ... -- so nothing is known ahead of time and
nothing is chained; every step launch takes 2 of the 4 CTA slots per SM (cta_slots=2), so the two streams' launches run side
by side and the start-up and tail of one are covered by the bulk of the other.
The "policy" here is a proportional attitude-hold: sticks from the drone's own body rates and height error."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone  # noqa: E402

n, steps, dev = 1 << 19, 60, "cuda:0"
halves, streams, actions = [], [torch.cuda.Stream(), torch.cuda.Stream()], []
g = torch.Generator(device=dev).manual_seed(0)
for h in range(2):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, cta_slots=2)
    pos = torch.randn(n, 3, device=dev, generator=g) * 3
    pos[:, 2] = 2.0 + torch.rand(n, device=dev, generator=g)
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g) * 0.3, (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 10)
    halves.append(d)
    actions.append(torch.zeros(n, 4, device=dev))
hover = -0.16                                   # throttle stick that roughly carries the weight


def policy(d, out):
    """sticks from the CURRENT state (this is what closes the loop): damp the body rates, hold 2.5 m."""
    out[:, :3] = (d.prev_rates / d.max_rates).clamp(-1, 1) * 0.5
    out[:, 3] = (hover + 0.4 * (2.5 - d.position[:, 2]) - 0.2 * d.velocity[:, 2]).clamp(-1, 1)


torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_stream(torch.cuda.current_stream())
for t in range(steps):
    for h in range(2):
        with torch.cuda.stream(streams[h]):
            policy(halves[h], actions[h])
            halves[h].step(actions[h], return_obs=False)
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
z = torch.cat([d.position[:, 2] for d in halves])
crashes = sum(d.episode_stats()["crashes"] for d in halves)
print(f"{2 * n} drones in two halves on two streams, {steps} closed-loop control steps (policy + 8 substeps each): "
      f"{ms / steps * 1e3:.1f} us per step of the whole population = {2 * n * steps / (ms * 1e-3):.3g} env-steps/s; "
      f"mean height {z.mean().item():.2f} m, crashes {int(crashes)}")
