#!/usr/bin/env python
"""Measurement of the chase pipeline on one B200: per-stage time of
  camera pose -> depth frame of the whole world -> target pixel -> point-and-shoot autopilot -> override step
for N drones with the stock 640x480 camera and the stock-sized world (params.yaml: 50x50 ground points, 5 cylinders of
25x10 points, one nu=5 target).  CUDA events, L2 flushed before every timed call; prints one JSON line.

Roofline of the depth frame (the dominant stage): HBM.  Algorithmic bytes per env = W*H (the frame is written once)
+ 32 B per world point (its double4 is read once per env, from L2).  The reference renders the same frame with a Python
loop over the points (components.py:620-625)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpyv_b200 import Autopilot, BatchedCamera, BatchedDrone, Cylinder, Ground, Target, World, config  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = "cuda:0"
    params = config.load_params(None)
    rng = np.random.default_rng(3)
    ground = Ground(60, 50, random=True, rng=rng)
    cyls = [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0, 10, 25, random=True, rng=rng) for _ in range(5)]
    tgt = Target(np.array([0.0, 0.0, 3.0]), 1.0, nu=5)
    world = World([tgt, *cyls, ground], dev)
    tworld = World([tgt], dev)
    g = torch.Generator(device=dev).manual_seed(9)
    pos = torch.randn(n, 3, device=dev, generator=g) * torch.tensor([8.0, 8.0, 0.0], device=dev)
    pos[:, 2] = 1.0 + torch.rand(n, device=dev, generator=g) * 9
    d = BatchedDrone(params, num_envs=n, device=dev)
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 40)
    cam = BatchedCamera.from_params(params, n, dev)
    ap = Autopilot(d, cam)
    act = torch.rand(n, 4, device=dev, generator=g) * 2 - 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    state = {}

    def stage_pose():
        cam.update_from(d)

    def stage_frame():
        state["img"] = cam.render_depth_image(world, 25)

    def stage_pixel():
        state["px"], state["seen"] = cam.target_pixel(tworld, 15)

    def stage_autopilot():
        state["q"], state["f"] = ap.calculate_needed_force_orientation(state["px"], tgt.position, tgt.radius, seen=state["seen"],
                                                                       as_quaternion=True)

    def stage_step():
        d.step(act, np.zeros(3), [Ground()], rotation_matrix=state["q"], thrust_force=state["f"], return_obs=False)

    stages = [("camera_pose", stage_pose), ("depth_frame", stage_frame), ("target_pixel", stage_pixel),
              ("autopilot", stage_autopilot), ("override_step", stage_step)]
    for _, fn in stages:
        fn()
    # pre-allocate once: the frame tensor is re-created by every render call (torch caching allocator, no cudaMalloc)
    out = {}
    for name, fn in stages:
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name] = sorted(ts)[len(ts) // 2]
    W, H = int(cam.resolution[0]), int(cam.resolution[1])
    frame_bytes = n * (W * H + 32 * world.n_points)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    ach = frame_bytes / (out["depth_frame"] * 1e-3) / 1e9
    # CPU beside it: the oracle's restatement of the reference's per-point Python loop, one frame
    from oracle import chase_oracle as co
    c = co.camera_consts(params)
    cp, cR = cam.position.cpu().numpy(), cam.rotation_matrix.cpu().numpy()
    objs = [np.asarray(o.points) for o in (tgt, *cyls, ground)]
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 5.0:
        co.render_depth_image(c, cp[k % n], cR[k % n], objs, 25)
        k += 1
    cpu_fps = k / (time.perf_counter() - t0)
    total = sum(out.values())
    print(json.dumps({"metric": "chase_pipeline_env_steps_per_sec", "value": n / (total * 1e-3), "unit": "env-steps/s",
                      "n_envs": n, "resolution": [W, H], "world_points": world.n_points, "ms": out, "ms_total": total,
                      "frames_seen": int(state["seen"].sum()), "nonzero_pixels_per_frame": float((state["img"] > 0).sum()) / n,
                      "roofline": {"kernel": "cudaMemsetAsync + fpv::camera_prune_kernel + fpv::camera_splat_kernel",
                                   "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "peak_source": src, "algorithmic": f"{W*H} B frame + 32 B x {world.n_points} points per env x {n} envs"},
                      "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                       "sample": f"{k} frames of oracle/chase_oracle.render_depth_image (float64 NumPy restatement of "
                                                 "components.py:614-629 incl. its per-point Python loop), 5 s"},
                      "gpu_frames_per_sec": n / (out["depth_frame"] * 1e-3)}))


if __name__ == "__main__":
    main()
