#!/usr/bin/env python
"""The reference's driver loop (src/core/simulator.py:53-59, :83-93, :156) with the drop-in `Drone`:
one drone, NumPy in / NumPy out, dt = 1/fps, ground in the object list, stop on crash."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import Drone, Ground, config  # noqa: E402

params = config.load_params(None)                       # params.yaml (same keys as the reference's)
drone = Drone(params)
ground = Ground(**{k: params["simulator"]["ground"][k] for k in ("size", "resolution")}, random=False)
drone.reset(position=np.array(params["drone"]["initial_position"]), velocity=np.array(params["drone"]["initial_velocity"]),
            ypr=np.array(params["drone"]["initial_orientation"]))
wind = np.zeros(3)
for i in range(600):
    action = np.array([-0.1, 0.0, 0.0, -0.62])          # roll a little, throttle just under hover
    Rt, gyro, acc = drone.step(action=action, wind_velocity_vector=wind, object_list=[ground])
    if drone.done:
        print(f"crashed at step {i}, position {drone.position.round(2)}")
        break
else:
    print("flew 10 s, position", drone.position.round(2), "velocity", drone.velocity.round(2))
