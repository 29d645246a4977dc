"""small, fast exercise of every kernel for compute-sanitizer memcheck (ragged sizes hit the partial-chunk paths)"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone, BatchedRacer, Joystick, Cylinder, Target, Ground
from fpyv_b200.env import GateRaceEnv
dev = "cuda:0"
rng = np.random.default_rng(0)
for n in (1, 63, 64, 65, 129, 1000):
    for kw in (dict(), dict(packed=False), dict(thrust_lut=257, auto_reset=True, substeps=8, dt=1e-3), dict(freeze_done=True)):
        d = BatchedDrone(None, num_envs=n, device=dev, **kw)
        d.reset(np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 2, n)], 1), rng.normal(size=(n, 3)), rng.uniform(-20, 20, (n, 3)))
        for _ in range(4):
            d.step(rng.uniform(-1, 1, (n, 4)))
        d.step(rng.uniform(-1, 1, (n, 4)), wind_velocity_vector=torch.randn(n, 3, device=dev), return_obs=False)
        d.step(rng.uniform(-1, 1, (n, 4)), object_list=[Target([0, 0, 3], 1.0), Cylinder([2, 0, 0], 0.5, 4.0), Ground()], return_obs=False)
        if d.substeps == 1:
            d.step(rng.uniform(-1, 1, (n, 4)), rotation_matrix=np.tile(np.eye(3), (n, 1, 1)), thrust_force=np.full(n, 7.0), return_obs=False)
        _ = d.rotation_matrix; d.set_rotation_matrix(np.tile(np.eye(3), (n, 1, 1)))
    r = BatchedRacer(5, {"roll": [2, 0.1, 1e-4], "pitch": [2, 0, 0], "yaw": [0.1, 0, 0]}, num_envs=n, device=dev)
    r.reset(); r.step(rng.uniform(-3, 3, (n, 4))); r.step(rng.uniform(-3, 3, (n, 4)))
    rc = Joystick(device=dev); rc.calibrate(os.path.join(os.path.dirname(__file__), "..", "fpyv_b200", "config", "frsky.json"))
    rc.feed(rng.integers(0, 65536, (n, 6))); rc.calib_read(); rc.read_actions()
env = GateRaceEnv(None, num_envs=24, agents_per_env=4, device=dev, substeps=4)
env.reset()
for _ in range(3):
    env.step(torch.rand(24, 4, 4) * 2 - 1)
torch.cuda.synchronize()
print("sanitize_small: ok")
