"""Small, fast exercise of every kernel for compute-sanitizer (ragged sizes hit the partial-chunk paths; a 4,096-env chained
run exercises the per-chunk epoch protocol).  Run ONE tool per process / per gpurun call (B200_PROFILING.md):
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
tests/test_gpu_sanitizer.py does exactly that when FPV_RUN_SANITIZER names the tool."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedAcroDrone, BatchedDrone, BatchedRacer, Joystick, Cylinder, Target, Ground, hostmem
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
    a = BatchedAcroDrone(None, num_envs=n, device=dev, substeps=4, dt=1e-3, auto_reset=True)
    a.reset(np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.3, 2, n)], 1), rng.normal(size=(n, 3)), rng.uniform(-20, 20, (n, 3)))
    a.step(rng.uniform(-1, 1, (n, 4))); a.step(rng.uniform(-1, 1, (n, 4)))
# chained launches (per-chunk epochs, batched publication), done bitmask, fused rollout, zero-copy host inputs
n = 4096
d = BatchedDrone(None, num_envs=n, device=dev, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049, done_bits=True)
d.reset(np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 2, n)], 1), rng.normal(size=(n, 3)), rng.uniform(-20, 20, (n, 3)))
acts = torch.as_tensor(rng.uniform(-1, 1, (12, n, 4)), dtype=torch.float32, device=dev)
for t in range(12):
    d.step(acts[t], return_obs=False, chained=True)
d.rollout(acts, fused=True)
ha, hb = hostmem.pinned((n, 4), torch.float32), hostmem.pinned(((n + 31) // 32,), torch.int32)
ha.copy_(acts[0].cpu())
d.step_host(ha, hb, zero_copy=True); d.step_host(ha, hb, slices=2)
hs = hostmem.pinned((n, 6), torch.uint8); hs.zero_()
d.step_host_sticks(hs, hb, zero_copy=True); d.step_host_sticks(hs, hb)
env = GateRaceEnv(None, num_envs=24, agents_per_env=4, device=dev, substeps=4)
env.reset()
for _ in range(3):
    env.step(torch.rand(24, 4, 4) * 2 - 1)
    env.step(torch.rand(24, 4, 4) * 2 - 1, fused=False)
torch.cuda.synchronize()
print("sanitize_small: ok")
