#!/usr/bin/env python
"""Measurement of BASELINE.json configs[4] on one B200: the multi-agent gate-race env, 262,144 drones = 8,192 envs x 32
agents (one warp per env), 8 substeps of 1 ms per control step.  One env step = the dynamics launch + the env kernel
(gate-plane crossing, per-agent reward, warp-reduced team reward / termination, 16-float observations).
CUDA events, L2 flushed before every timed step; prints one JSON line.

Roofline of the env kernel: HBM.  Algorithmic bytes per agent = 64 (state planes read) + 1 (done) + 8 + 4 (race
bookkeeping read) + 8 + 4 (written back) + 4 (agent reward) + 64 (observation) = 157 B, + 5 B per env (team reward,
done).  The reward rules are ours (parity unpinned, include/fpv_api.h); only the gate geometry is the reference's."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fpyv_b200.env import GateRaceEnv  # noqa: E402


def main():
    envs, agents = 8192, 32
    env = GateRaceEnv(None, num_envs=envs, agents_per_env=agents, device="cuda:0", substeps=8, dt=1e-3, thrust_lut=2049)
    env.reset()
    dev = env.device
    n = env.n_agents
    g = torch.Generator(device=dev).manual_seed(3)
    acts = [torch.rand(envs, agents, 4, device=dev, generator=g) * 2 - 1 for _ in range(4)]
    for a in acts:
        a[..., 3] = a[..., 3] * 0.3 - 0.3           # near-hover throttle: agents fly instead of falling
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(5):
        env.step(acts[i % 4])
    t_all, t_env = [], []
    for i in range(30):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        act = acts[i % 4].reshape(n, 4)
        e0.record()
        env.drone.step(act, return_obs=False)
        e1.record()
        env._run_env_kernel(env.drone._done)
        e2.record()
        torch.cuda.synchronize()
        t_all.append(e0.elapsed_time(e2))
        t_env.append(e1.elapsed_time(e2))
    med = lambda v: sorted(v)[len(v) // 2]
    ms_all, ms_env = med(t_all), med(t_env)
    # the same env step in ONE launch (fpv_gate_race_step: dynamics + env logic, one agent per thread)
    t_fused = []
    for i in range(30):
        flush.zero_()
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record()
        env.step(acts[i % 4], fused=True)
        e1.record()
        torch.cuda.synchronize()
        t_fused.append(e0.elapsed_time(e1))
    ms_fused = med(t_fused)
    env_bytes = n * 157 + envs * 5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    ach = env_bytes / (ms_env * 1e-3) / 1e9
    st = env.episode_stats()
    print(json.dumps({"metric": "gate_race_agent_steps_per_sec", "value": n / (ms_all * 1e-3), "unit": "agent-steps/s",
                      "config": {"workload": "BASELINE.json configs[4]: 262,144 drones = 8,192 envs x 32 agents, 8 substeps x 1 ms, "
                                             "8-gate track (generate_track), team reward / termination by warp reduction"},
                      "ms_per_env_step": ms_all, "ms_dynamics": ms_all - ms_env, "ms_env_kernel": ms_env,
                      "fused_single_launch": {"ms_per_env_step": ms_fused, "agent_steps_per_sec": n / (ms_fused * 1e-3),
                                              "hbm_GBps_algorithmic": n * (64 + 64 + 16 + 1 + 12 + 12 + 4 + 64) / (ms_fused * 1e-3) / 1e9,
                                              "api": "GateRaceEnv.step(fused=True) = fpv_gate_race_step; bit-identical to the "
                                                     "scalar dynamics kernel + fpv_gate_env_step"},
                      "env_steps_per_sec": envs / (ms_all * 1e-3),
                      "roofline": {"kernel": "fpv::gate_env_step_kernel", "bound": "hbm", "achieved": ach, "peak": peak,
                                   "unit": "GB/s", "frac": ach / peak, "peak_source": src,
                                   "algorithmic": "157 B per agent + 5 B per env",
                                   "note": "40 MB per launch: the kernel lasts a few microseconds, launch latency and the "
                                           "ramp of 1,024 CTAs are a large share of it"},
                      "episode_stats": st}))


if __name__ == "__main__":
    main()
