#!/usr/bin/env python
"""Small batches (BASELINE.json configs[1]: 4,096 drones, K = 8, one GPU): the step is launch-latency bound, so what
matters is how the launches are issued.  Per control step, one B200:
  step()          one Python call + one launch per step (host-bound: ~6 us of enqueue per call)
  cuda graph      16 steps captured once, replayed (device-side launch-to-launch latency)
  rollout(T=64)   one launch per 64 control steps, state in registers (fpv_drone_rollout)
Prints one JSON line; profiles/r1_small_batch_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpyv_b200 import BatchedDrone  # noqa: E402

dev = "cuda:0"
K, REPS = 8, 50


def make(n):
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    g = torch.Generator(device=dev).manual_seed(7)
    pos = torch.randn(n, 3, device=dev, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=dev, generator=g) * 2.95
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
    return d, g


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us per call


out = {"substeps": K, "unit": "us per control step", "sizes": {}}
for n in (4096, 65536):
    d, g = make(n)
    acts = torch.rand(64, n, 4, device=dev, generator=g) * 2 - 1
    d.step(acts[0], return_obs=False)
    i = [0]

    def one_step():
        d.step(acts[i[0] & 63], return_obs=False)
        i[0] += 1

    t_step = timed(one_step, 64 * REPS // 8)
    # 16 steps in one CUDA graph (static action buffers: the caller refreshes them between replays)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for t in range(16):
            d.step(acts[t], return_obs=False)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for t in range(16):
            d.step(acts[t], return_obs=False)
    t_graph = timed(graph.replay, REPS) / 16
    t_roll = timed(lambda: d.rollout(acts), REPS) / 64
    out["sizes"][str(n)] = {"step_call": round(t_step, 2), "cuda_graph_16": round(t_graph, 2), "rollout_64": round(t_roll, 2),
                            "env_steps_per_sec": {"step_call": n / t_step * 1e6, "cuda_graph_16": n / t_graph * 1e6,
                                                  "rollout_64": n / t_roll * 1e6}}
print(json.dumps(out))
