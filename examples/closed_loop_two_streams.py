#!/usr/bin/env python
"""Closed-loop stepping at the open-loop rate: `TwoStreamDrones` splits the population into two independent halves on two
CUDA streams.  Each stream is in plain order -- policy(state) -> step -> policy(state) -> ... -- so nothing is known ahead of
time and nothing is chained; every step launch takes 2 of the 4 CTA slots per SM, so the two streams' launches run side by
side and the start-up and tail of one are covered by the bulk of the other.
The "policy" here is a proportional attitude-hold: sticks from the drone's own body rates and height error."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import TwoStreamDrones  # noqa: E402

n, steps, dev = 1 << 20, 60, "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
pop = TwoStreamDrones(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
pos = torch.randn(n, 3, device=dev, generator=g) * 3
pos[:, 2] = 2.0 + torch.rand(n, device=dev, generator=g)
pop.reset(pos, torch.randn(n, 3, device=dev, generator=g) * 0.3, (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 10)
actions = [torch.zeros(d.num_envs, 4, device=dev) for d in pop.parts]
hover = -0.16                                   # throttle stick that roughly carries the weight


def policy(d, i):
    """sticks from the CURRENT state of part i (this is what closes the loop): damp the body rates, hold 2.5 m."""
    out = actions[i]
    out[:, :3] = (d.prev_rates / d.max_rates).clamp(-1, 1) * 0.5
    out[:, 3] = (hover + 0.4 * (2.5 - d.position[:, 2]) - 0.2 * d.velocity[:, 2]).clamp(-1, 1)
    return out


for t in range(5):        # warm-up (allocator, first launches)
    pop.step(policy)
pop.join()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(steps):
    pop.step(policy)
pop.join()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{n} drones in two halves on two streams, {steps} closed-loop control steps (policy + 8 substeps each): "
      f"{ms / steps * 1e3:.1f} us per step of the whole population = {n * steps / (ms * 1e-3):.3g} env-steps/s "
      f"(this toy policy is a dozen eager PyTorch element-wise launches per part and dominates; the dynamics step itself "
      f"takes ~37 us in this form, bench.py ms_per_step_two_streams); "
      f"mean height {pop.position[:, 2].mean().item():.2f} m, crashes {int(pop.episode_stats()['crashes'])}")
