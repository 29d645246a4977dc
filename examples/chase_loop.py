#!/usr/bin/env python
"""The reference's chase loop (src/core/simulator.py:98-110) for 4,096 drones at once: render the target, average its
pixels, point-and-shoot autopilot, step with the per-env rotation / thrust override; then one full depth frame each."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import Autopilot, BatchedCamera, BatchedDrone, Cylinder, Ground, Target, World, config  # noqa: E402

n, dev = 4096, "cuda:0"
params = config.load_params(None)
drone = BatchedDrone(params, num_envs=n, device=dev)
cam = BatchedCamera.from_params(params, n, dev)
ap = Autopilot(drone, cam)
rng = np.random.default_rng(0)
target = Target(np.array([12.0, 0.0, 4.0]), 1.0, nu=5)
ground = Ground(60, 50, random=True, rng=rng)
cylinders = [Cylinder(np.array([rng.normal(6, 4), rng.normal(0, 8), 0.0]), 1.0, 8.0, 10, 25, random=True, rng=rng) for _ in range(5)]
g = torch.Generator(device=dev).manual_seed(0)
pos = torch.randn(n, 3, device=dev, generator=g) * torch.tensor([2.0, 4.0, 0.0], device=dev)
pos[:, 2] = 3 + torch.rand(n, device=dev, generator=g) * 3
drone.reset(pos, torch.tensor([1.0, 0.0, 0.0], device=dev).expand(n, 3), torch.zeros(n, 3, device=dev))
ap.reset()
tworld = World([target], dev)
action = np.tile([-0.1, 0.0, 0.0, 0.0], (n, 1))                     # simulator.py:88
d0 = (drone.position - torch.as_tensor(target.position, dtype=torch.float32, device=dev)).norm(dim=1).mean().item()
for _ in range(180):                                                  # 3 s at 60 fps
    cam.update_from(drone)
    pixel, seen = cam.target_pixel(tworld, max_depth=15)
    q, force = ap.calculate_needed_force_orientation(pixel, target.position, target.radius, seen=seen, as_quaternion=True)
    drone.step(action, np.zeros(3), [target, *cylinders, ground], rotation_matrix=q, thrust_force=force, return_obs=False)
d1 = (drone.position - torch.as_tensor(target.position, dtype=torch.float32, device=dev)).norm(dim=1)
frames = cam.render_depth_image([target, *cylinders, ground], max_depth=25)
print(f"mean distance to the target {d0:.1f} m -> {d1[~drone.done].mean().item():.1f} m (keep_distance = "
      f"{params['drone']['keep_distance']} m); crashed {int(drone.done.sum())} of {n}; target in view {int(seen.sum())}; "
      f"frames {tuple(frames.shape)} {frames.dtype}")
