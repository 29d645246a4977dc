"""Short measurement legs for the BASELINE.json configs other than the headline one, run by bench.py after the headline
legs so that the driver's BENCH/SCALE records carry them (`extra` in the JSON line):

  config0_racer      configs[0]  single `Racer` (tests/racer_drone_test.py), batch 1, dt 1 ms, 10,000 steps -- and the same
                                 kernel at 1,048,576 racers (where a roofline means something)
  config1_small      configs[1]  4,096 drones, K = 8: step() per launch / CUDA graph / fused rollout
  config3_sharded    configs[3]  16,777,216 drones sharded over the ranks (N > 1 only) + all-reduced episode statistics
  config4_gate_race  configs[4]  262,144 agents = 8,192 envs x 32, dynamics + env step in one launch
  mode_c             north_star's acro inner loop (rate PID -> mixer -> per-motor LUT thrust), 1,048,576 envs
  general_path       mode A with the stock obstacle world (sphere target + 5 cylinders + ground)
  chase              SURVEY 8f rows 1+3: depth frame (HBM-bound) and the closed chase loop

Every leg is timed on the device with CUDA events on the launching stream.  Two forms:
  ms_isolated   ONE launch after an explicit L2 flush (256 MiB write), median of `reps` -- pays the whole launch start-up/tail
  ms_stream     back-to-back launches in plain stream order over 4 independent batches stepped round-robin (cold L2 by
                working-set size), one event bracket -- the sustained rate; used for `roofline` when the batch is large
`roofline.frac` is algorithmic work / time / peak for the BINDING roofline of that leg (the other one is reported too).
Algorithmic counts are derived in DESIGN.md section 4.3."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))

# algorithmic work per env
FLOP_DRONE_SUBSTEP = 252           # SURVEY 8(d), reference formulation of Drone.step
FLOP_RACER_SUBSTEP = 126           # ~120 flop + 6 sin/cos
FLOP_ACRO_SUBSTEP = 355            # DESIGN 4.3 (our count of oracle/acro_oracle.py's step)
FLOP_GATE_ENV_STEP = 120           # gate metrics x2, pass test, reward, 16-float observation (per control step)
BYTES_DRONE_STEP = 64 + 64 + 16 + 1
BYTES_RACER_STEP = None            # filled from the library's plane count below
BYTES_GATE_ENV = 12 + 12 + 4 + 64  # race bookkeeping read + written, agent reward, observation


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


class Timer:
    def __init__(self, dev):
        self.dev = dev
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def isolated(self, fn, reps=15, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return _median(ts)

    def stream(self, fns, steps=40, warm=8):
        """`fns`: one callable per independent batch; step i runs fns[i % len(fns)].  The `steps` launches are captured into
        ONE CUDA graph and the replay is timed, so that the number is device time (the Python side of some of these calls --
        dict building, object-list lowering -- costs more than the kernel; a trainer would capture its loop the same way).
        Falls back to eager launches if a call cannot be captured."""
        for i in range(warm):
            fns[i % len(fns)]()
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(len(fns)):
                    fns[i]()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(steps):
                    fns[i % len(fns)]()
            graph = g
        except Exception:       # noqa: BLE001
            graph = None
            torch.cuda.synchronize()
        self.last_stream_form = "cuda graph" if graph is not None else "eager"
        if graph is not None:
            graph.replay()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for i in range(steps):
                fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps


def two_streams(fns, steps=40, warm=8):
    """Plain stream order on TWO streams: fns[0], fns[2] (independent batches) alternate on one stream, fns[1], fns[3] on the
    other; every launch waits for its predecessor on its stream, the two streams are not synchronised with each other, so the
    start-up and tail of one stream's launch are covered by the other's bulk.  Eager launches, one event bracket behind a
    spin kernel; device time per step."""
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()

    def run(count):
        for i in range(count):
            with torch.cuda.stream(sa if i % 2 == 0 else sb):
                fns[i % len(fns)]()
    sa.wait_stream(cur)
    sb.wait_stream(cur)
    run(warm)
    cur.wait_stream(sa)
    cur.wait_stream(sb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(1_000_000)
    e0.record()
    sa.wait_stream(cur)
    sb.wait_stream(cur)
    run(steps)
    cur.wait_stream(sa)
    cur.wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _roof(n, ms, flop_per_env, bytes_per_env, pk, kernel, timing):
    s = ms * 1e-3
    tf, gb = flop_per_env * n / s / 1e12, bytes_per_env * n / s / 1e9
    f32, hbm = tf / pk["fp32_tflops"], gb / pk["hbm_gbs"]
    bound = "fp32" if f32 >= hbm else "hbm"
    return {"kernel": kernel, "bound": bound, "achieved": tf if bound == "fp32" else gb,
            "peak": pk["fp32_tflops"] if bound == "fp32" else pk["hbm_gbs"], "unit": "TFLOP/s" if bound == "fp32" else "GB/s",
            "frac": max(f32, hbm), "fp32_frac": f32, "hbm_frac": hbm, "timing": timing,
            "algorithmic": f"{flop_per_env} flop + {bytes_per_env} B per env per launch x {n} envs"}


def _rand_init(n, dev, g, zlo=0.05, zhi=3.0, spread=5.0):
    pos = torch.randn(n, 3, device=dev, generator=g) * spread
    pos[:, 2] = zlo + torch.rand(n, device=dev, generator=g) * (zhi - zlo)
    return pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30


# ------------------------------------------------------------------------------------------------------------------
def leg_config0_racer(dev, pk, tm, K=8):
    from fpyv_b200 import BatchedRacer, _lib
    pid = {"roll": [2, 0.1, 1e-4], "pitch": [2, 0.1, 1e-4], "yaw": [0.1, 0, 0]}
    out = {"workload": "BASELINE.json configs[0]: Racer (tests/racer_drone_test.py:68-103), dt 1 ms"}
    # (i) the config as stated: ONE racer, 10,000 steps of 1 ms.  The demo holds its action (racer_drone_test.py:113-122),
    #     so the 10,000 steps are also one launch with substeps = 10,000.
    r1 = BatchedRacer(5, pid, num_envs=1, device=dev, dt=1e-3, substeps=1)
    r1.reset()
    a1 = torch.tensor([[3.0, -2.0, 1.0, 6.0]], device=dev)
    for _ in range(20):
        r1.step(a1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2000):
        r1.step(a1)
    e1.record()
    torch.cuda.synchronize()
    us_call = e0.elapsed_time(e1) / 2000 * 1e3
    rK = BatchedRacer(5, pid, num_envs=1, device=dev, dt=1e-3, substeps=10000)
    rK.reset()
    rK.step(a1)
    ms_10k = tm.isolated(lambda: rK.step(a1), reps=5, warm=1)
    out["batch1"] = {"us_per_step_one_launch_per_step": us_call, "steps_per_sec_one_launch_per_step": 1e6 / us_call,
                     "ms_10000_steps_one_launch": ms_10k, "steps_per_sec_one_launch": 1e4 / (ms_10k * 1e-3),
                     "note": "one env = one thread: launch-latency / dependent-chain bound, no roofline applies"}
    # (ii) the same kernel where a roofline applies: 1,048,576 racers, K substeps per launch
    n = 1 << 20
    g = torch.Generator(device=dev).manual_seed(2)
    planes = getattr(_lib, "RACER_PLANES", 7)
    rs, acts = [], []
    for j in range(4):
        r = BatchedRacer(5, pid, num_envs=n, device=dev, dt=1e-3, substeps=K)
        r.reset()
        rs.append(r)
        acts.append(torch.cat([torch.rand(n, 3, device=dev, generator=g) * 6 - 3, torch.rand(n, 1, device=dev, generator=g) * 10], 1).contiguous())
    ms_iso = tm.isolated(lambda: rs[0].step(acts[0]))
    ms_str = tm.stream([(lambda r=r, a=a: r.step(a)) for r, a in zip(rs, acts)])
    ms_2s = two_streams([(lambda r=r, a=a: r.step(a)) for r, a in zip(rs, acts)])
    b = planes * 16 * 2 + 16 + 16
    out["batch_1M"] = {"envs": n, "substeps": K, "ms_isolated": ms_iso, "ms_stream": ms_str, "ms_two_streams": ms_2s,
                       "env_steps_per_sec": n / (ms_2s * 1e-3), "env_substeps_per_sec": n * K / (ms_2s * 1e-3)}
    out["roofline"] = _roof(n, min(ms_str, ms_2s), FLOP_RACER_SUBSTEP * K, b, pk, "racer step kernel (mode B)",
                            "the faster of ms_stream (one stream) and ms_two_streams, both plain stream order")
    return out


def leg_config1_small(dev, pk, tm, n=4096, K=8):
    from fpyv_b200 import BatchedDrone
    g = torch.Generator(device=dev).manual_seed(7)
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    d.reset(*_rand_init(n, dev, g))
    acts = (torch.rand(64, n, 4, device=dev, generator=g) * 2 - 1).contiguous()
    d.step(acts[0], return_obs=False)
    i = [0]

    def one():
        d.step(acts[i[0] & 63], return_obs=False)
        i[0] += 1

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
        return e0.elapsed_time(e1) / reps * 1e3

    us_step = timed(one, 400)
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
    us_graph = timed(graph.replay, 50) / 16
    us_roll = timed(lambda: d.rollout(acts), 50) / 64
    best = min(us_step, us_graph, us_roll)
    return {"workload": f"BASELINE.json configs[1]: {n} drones, {K} substeps x 1 ms, LUT, ground, auto-reset, random sticks",
            "us_per_control_step": {"step_call": us_step, "cuda_graph_16_steps": us_graph, "rollout_64_steps_per_launch": us_roll},
            "env_steps_per_sec": {"step_call": n / us_step * 1e6, "cuda_graph_16_steps": n / us_graph * 1e6,
                                  "rollout_64_steps_per_launch": n / us_roll * 1e6},
            "roofline": dict(_roof(n, best * 1e-3, FLOP_DRONE_SUBSTEP * K, 17 if best == us_roll else BYTES_DRONE_STEP, pk,
                                   "drone_rollout_kernel" if best == us_roll else "ring_step_kernel<DroneMode hot path>", "best of the three forms"),
                             note=f"{n} envs = {n // 64} warp-chunks on 2,368 resident warps: latency-bound by construction "
                                  "(one chunk's 8 substeps are a dependent chain); the fraction is reported for completeness")}


def leg_config4_gate_race(dev, pk, tm, envs=8192, agents=32, K=8):
    from fpyv_b200.env import GateRaceEnv
    n = envs * agents
    es, acts = [], []
    g = torch.Generator(device=dev).manual_seed(3)
    for j in range(4):
        env = GateRaceEnv(None, num_envs=envs, agents_per_env=agents, device=dev, substeps=K, dt=1e-3, thrust_lut=2049, seed=j)
        env.reset()
        a = torch.rand(envs, agents, 4, device=dev, generator=g) * 2 - 1
        a[..., 3] = a[..., 3] * 0.3 - 0.3
        env.step(a)
        es.append(env)
        acts.append(a)
    ms_iso = tm.isolated(lambda: es[0].step(acts[0], fused=True))
    ms_str = tm.stream([(lambda e=e, a=a: e.step(a, fused=True)) for e, a in zip(es, acts)])
    ms_two = tm.isolated(lambda: es[0].step(acts[0], fused=False))
    ms_2s = two_streams([(lambda e=e, a=a: e.step(a, fused=True)) for e, a in zip(es, acts)])
    # the headline's form: the sticks of all steps exist up front, so launches of the 4 independent envs are chained and
    # take 2 of the 4 CTA slots per SM each (side by side); eager launches in one event bracket behind a spin kernel
    for e in es:
        e.drone.cta_slots = 2
    for i in range(8):
        es[i % 4].step(acts[i % 4], fused=True, chained=True)
    torch.cuda.synchronize()
    torch.cuda._sleep(1_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        es[i % 4].step(acts[i % 4], fused=True, chained=True)
    e1.record()
    torch.cuda.synchronize()
    ms_chained = e0.elapsed_time(e1) / 40
    import time as _time
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    for i in range(40):
        es[i % 4].step(acts[i % 4], fused=True, chained=True)
    host_us = (_time.perf_counter() - t0) / 40 * 1e6       # enqueue cost of one call (the device runs behind)
    torch.cuda.synchronize()
    for e in es:
        e.drone.cta_slots = 0
    flop = FLOP_DRONE_SUBSTEP * K + FLOP_GATE_ENV_STEP
    byt = BYTES_DRONE_STEP + BYTES_GATE_ENV
    return {"workload": f"BASELINE.json configs[4]: {n} drones = {envs} envs x {agents} agents, {K} substeps x 1 ms, 8-gate track, "
                        "per-env team reward / termination by warp reduction (reward rules: ours, parity unpinned)",
            "ms_isolated": ms_iso, "ms_stream": ms_str, "ms_two_streams": ms_2s, "ms_chained": ms_chained,
            "ms_two_launches_isolated": ms_two, "host_us_per_call": host_us,
            "agent_steps_per_sec": n / (ms_chained * 1e-3), "env_steps_per_sec": envs / (ms_chained * 1e-3),
            "agent_steps_per_sec_plain_stream_order": n / (ms_str * 1e-3),
            "roofline": _roof(n, ms_chained, flop, byt, pk, "fused gate-race step (fpv_gate_race_step)",
                              "ms_chained (open-loop form, like the headline); frac_plain_stream_order beside it"),
            "frac_plain_stream_order": _roof(n, ms_str, flop, byt, pk, "", "")["frac"],
            "episode_stats": es[0].episode_stats()}


def leg_mode_c(dev, pk, tm, n=1 << 20):
    from fpyv_b200 import BatchedAcroDrone, _lib
    g = torch.Generator(device=dev).manual_seed(4)
    out = {"workload": f"mode C (north_star): stick -> rate PID -> mixer -> 4 LUT lookups -> rigid body, {n} envs, auto-reset "
                       "(model: ours, parity unpinned; oracle/acro_oracle.py)"}
    planes = _lib.ACRO_PLANES
    for K in (8, 1):
        ds, acts = [], []
        for j in range(4):
            d = BatchedAcroDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True)
            pos, vel, rpy = _rand_init(n, dev, g, 0.3, 5.3)
            d.reset(pos, vel, rpy * (2.0 / 3.0))
            ds.append(d)
            acts.append((torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous())
        ms_iso = tm.isolated(lambda: ds[0].step(acts[0]))
        ms_str = tm.stream([(lambda d=d, a=a: d.step(a)) for d, a in zip(ds, acts)])
        ms_2s = two_streams([(lambda d=d, a=a: d.step(a)) for d, a in zip(ds, acts)])
        b = planes * 16 * 2 + 16 + 16 + 1
        best = min(ms_str, ms_2s)
        out[f"K{K}"] = {"ms_isolated": ms_iso, "ms_stream": ms_str, "ms_two_streams": ms_2s, "env_steps_per_sec": n / (best * 1e-3),
                        "env_substeps_per_sec": n * K / (best * 1e-3),
                        "roofline": _roof(n, best, FLOP_ACRO_SUBSTEP * K, b, pk, "acro step kernel (mode C)",
                                          "the faster of ms_stream (one stream) and ms_two_streams, both plain stream order")}
        del ds, acts
    out["roofline"] = out["K8"]["roofline"]
    return out


def leg_general_path(dev, pk, tm, n=1 << 20, K=8):
    from fpyv_b200 import BatchedDrone, Cylinder, Ground, Target
    rng = np.random.default_rng(5)
    objs = [Target(np.array([0.0, 0.0, 3.0]), 1.0)] + \
           [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0) for _ in range(5)] + [Ground()]
    g = torch.Generator(device=dev).manual_seed(6)
    out = {"workload": f"mode A general path: {n} drones, {K} substeps x 1 ms, LUT, auto-reset, object_list = 1 sphere target + "
                       "5 cylinders (r 2 m, h 10 m) + ground (the stock world of params.yaml)"}

    def run(spread, zhi, tag, note, clear_of=0.0):
        ds, acts = [], []
        for j in range(4):
            d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
            pos = torch.randn(n, 3, device=dev, generator=g) * spread
            pos[:, 2] = 0.3 + torch.rand(n, device=dev, generator=g) * zhi
            if clear_of > 0:      # push every spawn point radially out of the obstacles, `clear_of` metres off their surfaces
                for o in objs[:-1]:
                    c = torch.as_tensor(np.asarray(o.position, dtype=np.float32), device=dev)
                    horiz = hasattr(o, "height")
                    dvec = pos - c
                    if horiz:
                        dvec[:, 2] = 0.0
                    dist = dvec.norm(dim=1).clamp_min(1e-3)
                    need = float(o.radius) + clear_of
                    inside = dist < need
                    pos[inside] = (pos + dvec / dist[:, None] * (need - dist)[:, None])[inside]
            d.reset(pos, torch.randn(n, 3, device=dev, generator=g) * 2, (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
            ds.append(d)
            acts.append((torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous())
        ms_iso = tm.isolated(lambda: ds[0].step(acts[0], None, objs, return_obs=False))
        ms_str = tm.stream([(lambda d=d, a=a: d.step(a, None, objs, return_obs=False)) for d, a in zip(ds, acts)], steps=24, warm=4)
        out[tag] = {"ms_isolated": ms_iso, "ms_stream": ms_str, "env_steps_per_sec": n / (ms_str * 1e-3), "note": note,
                    "crashes_per_step": ds[0].episode_stats()["crashes"] / max(1.0, ds[0].episode_stats()["env_steps"] / n),
                    "roofline": _roof(n, ms_str, FLOP_DRONE_SUBSTEP * K, BYTES_DRONE_STEP, pk, "drone general-path kernel", "ms_stream")}

    run(8.0, 8.0, "contact_heavy", "spawn sigma 8 m around the obstacles: ~20 % of the drones start INSIDE a cylinder, every warp "
                                  "takes the contact path every substep")
    run(10.0, 8.0, "among_obstacles", "spawn sigma 10 m around the obstacles, every spawn point at least 0.6 m off their surfaces: "
                                     "drones fly AMONG the obstacles, some brush or hit them during the steps", clear_of=0.6)
    # clear: the same world, drones spawned in a ring 60-80 m from the origin (no obstacle in reach of any drone)
    def run_clear():
        ds, acts = [], []
        for j in range(4):
            d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
            ang = torch.rand(n, device=dev, generator=g) * 6.2831853
            rad = 60.0 + 20.0 * torch.rand(n, device=dev, generator=g)
            pos = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang), 0.3 + torch.rand(n, device=dev, generator=g) * 8], 1)
            d.reset(pos, torch.randn(n, 3, device=dev, generator=g) * 2, (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 30)
            ds.append(d)
            acts.append((torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous())
        ms_iso = tm.isolated(lambda: ds[0].step(acts[0], None, objs, return_obs=False))
        ms_str = tm.stream([(lambda d=d, a=a: d.step(a, None, objs, return_obs=False)) for d, a in zip(ds, acts)], steps=24, warm=4)
        # the headline's form: the world registered once (set_static_objects: fast path), sticks known up front -> chained
        # launches of the 4 independent batches, 2 of 4 CTA slots each
        for d in ds:
            d.set_static_objects(objs)
            d.cta_slots = 2
        ms_2s = two_streams([(lambda d=d, a=a: d.step(a, return_obs=False)) for d, a in zip(ds, acts)])
        for i in range(8):
            ds[i % 4].step(acts[i % 4], return_obs=False, chained=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(1_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            ds[i % 4].step(acts[i % 4], return_obs=False, chained=True)
        e1.record()
        torch.cuda.synchronize()
        ms_ch = e0.elapsed_time(e1) / 40
        out["clear"] = {"ms_isolated": ms_iso, "ms_stream": ms_str, "ms_two_streams": ms_2s, "ms_chained": ms_ch,
                        "env_steps_per_sec": n / (ms_ch * 1e-3),
                        "env_steps_per_sec_plain_stream_order": n / (ms_str * 1e-3),
                        "note": "same object list, every drone 60-80 m away from the obstacles",
                        "roofline": _roof(n, ms_ch, FLOP_DRONE_SUBSTEP * K, BYTES_DRONE_STEP, pk, "drone general-path kernel",
                                          "ms_chained (open-loop form, like the headline); frac_plain_stream_order beside it"),
                        "frac_plain_stream_order": _roof(n, ms_str, FLOP_DRONE_SUBSTEP * K, BYTES_DRONE_STEP, pk, "", "")["frac"]}
    run_clear()
    out["roofline"] = out["clear"]["roofline"]
    return out


def leg_chase(dev, pk, tm, n=4096):
    from fpyv_b200 import Autopilot, BatchedCamera, BatchedDrone, Cylinder, Ground, Target, World, config
    params = config.load_params(None)
    rng = np.random.default_rng(3)
    ground = Ground(60, 50, random=True, rng=rng)
    cyls = [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0, 10, 25, random=True, rng=rng) for _ in range(5)]
    tgt = Target(np.array([0.0, 0.0, 3.0]), 1.0, nu=5)
    world, tworld = World([tgt, *cyls, ground], dev), World([tgt], dev)
    g = torch.Generator(device=dev).manual_seed(9)
    pos = torch.randn(n, 3, device=dev, generator=g) * torch.tensor([8.0, 8.0, 0.0], device=dev)
    pos[:, 2] = 1.0 + torch.rand(n, device=dev, generator=g) * 9
    d = BatchedDrone(params, num_envs=n, device=dev)
    d.reset(pos, torch.randn(n, 3, device=dev, generator=g), (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 40)
    cam = BatchedCamera.from_params(params, n, dev)
    ap = Autopilot(d, cam)
    act = torch.rand(n, 4, device=dev, generator=g) * 2 - 1
    st = {}
    gnd = [Ground()]
    zero_wind = np.zeros(3)

    def s_pose():
        cam.update_from(d)

    def s_frame():
        st["img"] = cam.render_depth_image(world, 25)

    def s_pixel():
        st["px"], st["seen"] = cam.target_pixel(tworld, 15)

    def s_auto():
        st["q"], st["f"] = ap.calculate_needed_force_orientation(st["px"], tgt.position, tgt.radius, seen=st["seen"], as_quaternion=True)

    def s_step():
        d.step(act, zero_wind, gnd, rotation_matrix=st["q"], thrust_force=st["f"], return_obs=False)

    stages = [("camera_pose", s_pose), ("depth_frame", s_frame), ("target_pixel", s_pixel), ("autopilot", s_auto), ("override_step", s_step)]
    for _, fn in stages:
        fn()
    ms = {name: tm.isolated(fn, reps=7, warm=2) for name, fn in stages}
    W, H = int(cam.resolution[0]), int(cam.resolution[1])
    fb = n * (W * H + 32 * world.n_points)
    ach = fb / (ms["depth_frame"] * 1e-3) / 1e9
    loop = ms["camera_pose"] + ms["target_pixel"] + ms["autopilot"] + ms["override_step"]
    return {"workload": f"SURVEY 8f rows 1+3: {n} cameras {W}x{H}, world of {world.n_points} points (stock sizes), float64 geometry",
            "ms": ms, "frames_per_sec": n / (ms["depth_frame"] * 1e-3), "closed_loop_env_steps_per_sec": n / (loop * 1e-3),
            "roofline": {"kernel": "cudaMemsetAsync + camera_prune_kernel + camera_splat_kernel", "bound": "hbm", "achieved": ach,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "timing": "ms_isolated",
                         "algorithmic": f"{W * H} B frame + 32 B x {world.n_points} points per camera x {n} cameras"}}


def leg_plain_order_vs_batch(dev, pk, tm, K=8):
    """The headline kernel in PLAIN STREAM ORDER (what a closed-loop policy can use: every launch waits for the previous grid)
    as a function of the batch size.  A launch has a fixed cost that does not scale with the batch (~4 us start-up, ~6 us of
    end-of-kernel imbalance: the ring commits every warp to two 64-env chunks); at 1,048,576 envs that is ~25 % of the step,
    from 2,097,152 envs per GPU the plain-order launch is above 0.70 of the FP32 roofline."""
    from fpyv_b200 import BatchedDrone
    out = {"workload": f"configs[2] kernel and inputs, {K} substeps, plain stream order, 3 independent batches round-robin (cold L2)"}
    g = torch.Generator(device=dev).manual_seed(12)
    for n in (1 << 20, 1 << 21, 1 << 22, 1 << 23):
        ds, acts = [], []
        for j in range(3):
            d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
            d.reset(*_rand_init(n, dev, g))
            ds.append(d)
            acts.append((torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous())
        ms = tm.stream([(lambda d=d, a=a: d.step(a, return_obs=False)) for d, a in zip(ds, acts)], steps=18, warm=6)
        r = _roof(n, ms, FLOP_DRONE_SUBSTEP * K, BYTES_DRONE_STEP, pk, "ring_step_kernel<DroneMode<F2, hot>>", "ms_stream")
        out[str(n)] = {"ms_per_step": ms, "env_steps_per_sec": n / (ms * 1e-3), "fp32_frac": r["fp32_frac"], "hbm_frac": r["hbm_frac"]}
        del ds, acts
        torch.cuda.empty_cache()
    out["roofline"] = {"bound": "fp32", "unit": "TFLOP/s", "frac_by_envs": {k: v["fp32_frac"] for k, v in out.items() if k.isdigit()}}
    return out


def leg_config3_sharded(dev, pk, tm, world, rank, K=8, total=1 << 24, steps=12):
    """configs[3]: 16,777,216 drones sharded over `world` ranks (contiguous env slices, fpyv_b200.shard.env_shard), stepped K = 8
    with no data-path collective, then the engine's only collective: the all-reduce of the episode statistics (NCCL)."""
    import torch.distributed as dist
    from fpyv_b200 import BatchedDrone
    from fpyv_b200.shard import env_shard
    start, n = env_shard(total, rank, world)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    d = BatchedDrone(None, num_envs=n, device=dev, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    # initial state generated in slices (bounded temporaries)
    pos = torch.empty(n, 3, device=dev)
    vel = torch.empty(n, 3, device=dev)
    rpy = torch.empty(n, 3, device=dev)
    for a in range(0, n, 1 << 21):
        b = min(n, a + (1 << 21))
        pos[a:b], vel[a:b], rpy[a:b] = _rand_init(b - a, dev, g)
    d.reset(pos, vel, rpy)
    del pos, vel, rpy
    acts = [(torch.rand(n, 4, device=dev, generator=g) * 2 - 1).contiguous() for _ in range(2)]
    for i in range(4):
        d.step(acts[i % 2], return_obs=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        d.step(acts[i % 2], return_obs=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # the collective itself, timed on the device (latency-bound: 64 bytes)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = d.episode_stats(all_reduce=world > 1)
    e2.record()
    for _ in range(10):
        d.episode_stats(all_reduce=world > 1)
    e3.record()
    torch.cuda.synchronize()
    out = {"workload": f"BASELINE.json configs[3]: {total} drones env-sharded over {world} GPU(s) ({n} on this rank), {K} substeps x 1 ms, "
                       "LUT, ground, auto-reset; plain stream order (state >> L2); statistics all-reduced over NCCL after the loop",
           "envs_per_gpu": n, "total_envs": total, "ms_per_step": ms, "env_steps_per_sec": total / (ms * 1e-3),
           "env_substeps_per_sec": total * K / (ms * 1e-3), "stats_allreduce_ms_incl_d2h": e2.elapsed_time(e3) / 10,
           "roofline": _roof(n, ms, FLOP_DRONE_SUBSTEP * K, BYTES_DRONE_STEP, pk, "ring_step_kernel<DroneMode hot path>", "plain stream order, max over ranks"),
           "episode_stats": stats}
    del d, acts
    torch.cuda.empty_cache()
    return out


def run_extra(dev, pk, world=1, rank=0, which=None):
    """All legs (rank 0 runs the single-GPU ones; config3 runs on every rank).  A failing leg reports its error instead of
    taking the headline line down with it."""
    tm = Timer(dev)
    out = {}
    legs = [("config1_small", leg_config1_small), ("config4_gate_race", leg_config4_gate_race), ("config0_racer", leg_config0_racer),
            ("mode_c", leg_mode_c), ("general_path", leg_general_path), ("chase", leg_chase),
            ("plain_order_vs_batch", leg_plain_order_vs_batch)]
    if rank == 0:
        for name, fn in legs:
            if which and name not in which:
                continue
            try:
                out[name] = fn(dev, pk, tm)
            except Exception as e:   # noqa: BLE001
                out[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ROOT)
    dev = torch.device("cuda", 0)
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        pk = {"hbm_gbs": float(m["hbm_gbs"]), "fp32_tflops": sm * 128 * 2 * float(m.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12}
    except Exception:
        pk = {"hbm_gbs": 6650.0, "fp32_tflops": sm * 128 * 2 * 1965e6 / 1e12}
    print(json.dumps(run_extra(dev, pk, which=set(sys.argv[1:]) or None)))
