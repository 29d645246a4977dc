#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched FPV-drone dynamics step on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the config the metric is quoted on): 1,048,576 drones per GPU,
8 physics substeps of 1 ms per control step, body drag + motor-curve shared-memory LUT + ground contact
(initial heights 0.05..3 m so a fraction bounce and crash), auto-reset on crash, fresh random stick
actions every control step.  A "step" is one control step of every env.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
SUBSTEPS = 8
DT = 1e-3
LUT_N = int(os.environ.get("FPV_BENCH_LUT", "2049"))   # env override: developer tuning only
CHAINED = os.environ.get("FPV_BENCH_CHAINED", "1") != "0"   # developer A/B: 0 = every launch waits for the previous grid
E2E_SLICES = int(os.environ.get("FPV_BENCH_E2E_SLICES", "4"))
CTA_SLOTS = int(os.environ.get("FPV_BENCH_CTA_SLOTS", "2"))  # CTA slots per SM one launch takes in the primary loop (0 = all)
# algorithmic work per env (quaternion state, 64 B each way)
BYTES_PER_ENV_STEP = 64 + 64 + 16 + 1          # state read + state write + action + done flag
FLOP_PER_ENV_SUBSTEP = 252                     # 245 arithmetic + 6 sin/cos + 1 sqrt
FP32_LANES_PER_SM = 128
SPIN_CYCLES = 500_000                          # ~0.25 ms at 1.965 GHz: lets the host run ahead of the device before e0


def ncu_traffic_bytes():
    """DRAM bytes per K=8 launch from the committed ncu capture (None if the profile is absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_drone_step.json")) as f:
            return float(json.load(f)[0]["traffic_bytes"])
    except Exception:
        return None


def peaks():
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)),
                 source="measured (MEASURED_PEAKS.json)")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


def synthetic_init(n, device, seed):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    pos = torch.randn(n, 3, device=device, generator=g) * torch.tensor([5.0, 5.0, 0.0], device=device)
    pos[:, 2] = 0.05 + torch.rand(n, device=device, generator=g) * 2.95       # ground-contact config
    vel = torch.randn(n, 3, device=device, generator=g)
    rpy = (torch.rand(n, 3, device=device, generator=g) * 2 - 1) * 30
    return pos, vel, rpy, g


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------------
NUMPY_REFERENCE = {"value": 2.8e3, "unit": "env-steps/s", "cores": 1, "kind": "reference",
                   "sample": "the reference's own per-object NumPy Drone.step (src/utils/components.py:220-248) driven through "
                             "oracle/ref_shim.py, 2,000 steps, ground-only object list, stdout swallowed; measured in the build "
                             "container (SURVEY.md section 6) -- a CONSTANT quoted for scale, not re-measured on this box: the Python "
                             "reference does not travel to the GPU box"}


def _cpu_workload(n, seed=1234):
    import numpy as np
    import yaml
    from oracle import fpv_oracle as fo
    cfg = os.path.join(ROOT, "fpyv_b200", "config")
    with open(os.path.join(cfg, "params.yaml")) as f:
        params = yaml.safe_load(f)
    c = fo.derive_consts(params, os.path.join(cfg, "t_motos_f80_motor_test.csv"), dt=DT)
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.05, 3.0, n)], 1)
    vel = rng.normal(0, 1, (n, 3))
    rpy = rng.uniform(-30, 30, (n, 3))
    s = fo.drone_reset(c, pos, vel, rpy)
    acts = [np.ascontiguousarray(rng.uniform(-1, 1, (n, 4))) for _ in range(4)]
    return c, s.pos.copy(), s.vel.copy(), np.ascontiguousarray(s.R), np.zeros((n, 3)), np.zeros(n), acts


def numba_arm(n=ENVS_PER_GPU, target_seconds=6.0):
    """env-steps/s of oracle/numba_oracle.py: a Numba @njit(parallel=True) float64 restatement of Drone.step.  The reference
    ships NO Numba path (src/utils/kinematics.py:6,14 are commented out) -- this is what north_star's "Numba CPU path" can mean."""
    import numpy as np
    from oracle import numba_oracle as no
    c, P, V, R, pr, pt, acts = _cpu_workload(n)
    k, mrel = no.make_consts(c)
    wind, done = np.zeros(3), np.zeros(n, dtype=np.bool_)
    no.drone_step(k, mrel, P[:64], V[:64], R[:64], pr[:64], pt[:64], acts[0][:64], wind, 1, done[:64])   # JIT compile
    no.drone_step(k, mrel, P, V, R, pr, pt, acts[0], wind, SUBSTEPS, done)
    t0, steps = time.perf_counter(), 0
    while True:
        no.drone_step(k, mrel, P, V, R, pr, pt, acts[steps % 4], wind, SUBSTEPS, done)
        steps += 1
        if time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": no.threads(), "kind": "restatement",
            "sample": f"{n} envs x {SUBSTEPS} substeps x {steps} control steps, Numba @njit(parallel=True) float64 restatement of "
                      f"Drone.step (oracle/numba_oracle.py; the reference ships no Numba path: its @jit lines are commented out, "
                      f"src/utils/kinematics.py:6,14), {no.threads()} threads, {dt:.1f} s"}


def cpu_arm(steps, warmup, n_sample=ENVS_PER_GPU, target_seconds=None):
    """env-steps/s of oracle/fpv_oracle.c (float64, all host cores) on the workload's full batch."""
    from oracle import c_oracle
    c_oracle.build()
    n = n_sample
    c, P, V, R, pr, pt, acts = _cpu_workload(n)
    k = c_oracle.make_consts(c)
    cores = c_oracle.max_threads()
    for i in range(warmup):
        c_oracle.drone_step(k, P, V, R, pr, pt, acts[i % 4], substeps=SUBSTEPS, threads=cores)
    t0 = time.perf_counter()
    done_steps = 0
    for i in range(steps):
        c_oracle.drone_step(k, P, V, R, pr, pt, acts[i % 4], substeps=SUBSTEPS, threads=cores)
        done_steps += 1
        if target_seconds and time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": n * done_steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {SUBSTEPS} substeps x {done_steps} control steps, float64 C restatement of "
                      f"Drone.step (oracle/fpv_oracle.c), {cores} threads, {dt:.1f} s",
            "ms_per_step": 1e3 * dt / done_steps, "steps": done_steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": r["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus, **extra):
    c = {"workload": "BASELINE.json configs[2]: 1,048,576 drones per GPU, 8 substeps x 1 ms per control step, "
                     "drag + motor-curve LUT + ground contact, auto-reset, random sticks",
         "envs_per_gpu": ENVS_PER_GPU, "total_envs": ENVS_PER_GPU * n_gpus, "substeps_per_step": SUBSTEPS,
         "dt_substep_s": DT, "thrust_lut_entries": LUT_N, "parallelism": f"env-sharded x{n_gpus}, no data-path collective",
         "l2": "cold by working-set size: 4 independent batches of envs_per_gpu drones are stepped round-robin (each step "
               "= one control step of ONE batch; 4 x (64 MiB state + 16 MiB actions) read + 4 x 64 MiB written per "
               "rotation > 126 MB L2), steps back to back in one CUDA-event bracket; ms_per_step_flushed is the "
               "cross-check with ONE batch and an explicit L2 flush (256 MiB write + 256 MiB read) before every step, "
               "per-step event intervals summed",
         "launches": f"one launch per control step; the stick commands of all steps exist before the loop, so launches are "
                     f"chained (FPV_F_CHAINED: no grid-wide wait, state ordered per 64-env chunk) and each takes {CTA_SLOTS or 'all'} "
                     f"of the 4 CTA slots per SM, so launches of consecutive (independent) batches overlap; "
                     f"ms_per_step_chained_full_grid = chained with all slots, ms_per_step_unchained = every launch waits "
                     f"for the previous grid, ms_per_step_two_streams = plain stream order on TWO streams (independent halves of "
                     f"the population, 2 CTA slots per launch: the closed-loop form; ..._closed_loop: with a stand-in policy kernel "
                     f"rewriting the batch's actions in front of every step, its time included), ms_per_step_flushed = one isolated launch"}
    c.update(extra)
    return c


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from fpyv_b200 import BatchedDrone

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1 when NCCL_DEBUG is
    # set by the environment or an nccl.conf) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left as the environment set it: fd 1 already points at stderr (above), so NCCL's communicator lines
        # ("... nranks N ...") reach stderr and stdout still carries exactly the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    n, K, W = args.envs, args.steps, args.warmup

    NB = 4   # independent 1,048,576-env batches stepped round-robin: the working set (4 x 80 MiB read per step set)
             # exceeds the 126 MB L2, so every step finds its state cold WITHOUT flush kernels between the steps

    def make(substeps, lut=LUT_N, seed_off=0):
        d = BatchedDrone(None, num_envs=n, device=dev, substeps=substeps, dt=DT, auto_reset=True, thrust_lut=lut)
        pos, vel, rpy, g = synthetic_init(n, dev, 1234 + rank + 1000 * seed_off)
        d.reset(pos, vel, rpy)
        return d, g

    drones, gen = [], None
    for j in range(NB):
        d, g = make(SUBSTEPS, seed_off=j)
        drones.append(d)
        gen = gen or g
    drone = drones[0]
    ring = [torch.rand(n, 4, device=dev, generator=gen) * 2 - 1 for _ in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_r = torch.ones(64 << 20, dtype=torch.float32, device=dev)

    def timed_rotation(ds, steps, warm, chained=CHAINED, cta_slots=CTA_SLOTS):
        """The contract's timed region: `steps` control steps back to back inside ONE CUDA-event bracket (barrier +
        synchronize on both sides).  Step i advances batch i % NB, so its state was evicted from L2 by the other
        batches' traffic (inputs larger than L2).  The NB batches are independent, so with `cta_slots` = 2 every launch
        takes two of the four CTA slots of each SM and two consecutive launches are resident side by side: the idle
        tail and the start-up of one launch are covered by the bulk of its neighbour."""
        for d_ in ds:
            d_.cta_slots = cta_slots if chained else 0
        for i in range(warm):
            ds[i % len(ds)].step(ring[i % 4], return_obs=False, chained=chained)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # The bracket holds DEVICE time of exactly `steps` steps: a ~0.25 ms spin kernel is queued first, so that the host
        # has enqueued the event and the first launches before the device reaches them (otherwise the first launch's host
        # latency and the idle-to-busy ramp of the slowest rank -- a fixed ~70 us per bracket, MAX-reduced over the ranks --
        # are charged to a 20-step run and read as a scaling loss although the step has no collective).
        torch.cuda._sleep(SPIN_CYCLES)
        e0.record()
        for i in range(steps):
            # the stick commands of every step exist before the loop starts (open-loop rollout), so consecutive
            # launches may be chained: launch i+1 starts on the SMs launch i has left (per-chunk ordering of the state)
            ds[i % len(ds)].step(ring[i % 4], return_obs=False, chained=chained)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for d_ in ds:
            d_.cta_slots = 0
        return e0.elapsed_time(e1)

    def timed_loop(d, steps, warm):
        """Cross-check: ONE batch, L2 flushed explicitly before every step, per-step CUDA-event intervals summed
        (each interval then also contains the launch latency that back-to-back steps hide)."""
        for i in range(warm):   # warm-up mirrors the timed iteration exactly (flush kernels included)
            flush.zero_()
            flush_r.sum()
            d.step(ring[i % 4], return_obs=False)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for i in range(steps):
            flush.zero_()        # evict the state from L2 (write > L2 size) ...
            flush_r.sum()        # ... then a 256 MiB read pass so no dirty flush lines are written back inside the step
            ev[i][0].record()
            d.step(ring[i % 4], return_obs=False)
            ev[i][1].record()
        torch.cuda.synchronize()
        per = [a.elapsed_time(b) for a, b in ev]
        timed_loop.last = sorted(per)
        return sum(per)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    ms = timed_rotation(drones, K, W)
    ms_full_grid = timed_rotation(drones, K, W, chained=True, cta_slots=0)     # chained, every launch takes all CTA slots
    ms_unchained = timed_rotation(drones, K, W, chained=False)                 # every launch waits for the previous grid
    ms_flushed = timed_loop(drone, min(K, 200), W)
    K_fl = min(K, 200)

    def timed_two_streams(ds, steps, warm, policy=False):
        """Plain stream order in a form a CLOSED-LOOP trainer can use: the population is split into independent halves
        (here: batches 0, 2 on one stream and 1, 3 on the other), every launch waits for its predecessor ON ITS STREAM
        (no chaining, no knowledge of future sticks) and takes 2 of the 4 CTA slots per SM, so the two streams' launches run
        side by side and the start-up / tail of one stream's launch is covered by the other stream's bulk.
        policy=True puts a stand-in policy kernel in front of every step ON THE SAME STREAM (it rewrites that batch's 16 MiB
        action buffer after the batch's previous step has finished -- a real policy would read the new state there), so the
        loop is closed: no launch can start before the previous step of its batch AND the policy that followed it are done."""
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        acts_b = [torch.empty_like(ring[0]) for _ in ds] if policy else None
        for d_ in ds:
            d_.cta_slots = 2
        cur = torch.cuda.current_stream()

        def run(count):
            for i in range(count):
                with torch.cuda.stream(sa if i % 2 == 0 else sb):
                    if policy:
                        j = i % len(ds)
                        acts_b[j].copy_(ring[(i // len(ds) + j) % 4])
                        ds[j].step(acts_b[j], return_obs=False)
                    else:
                        ds[i % len(ds)].step(ring[i % 4], return_obs=False)
        sa.wait_stream(cur)
        sb.wait_stream(cur)
        run(warm + (warm % 2))
        cur.wait_stream(sa)
        cur.wait_stream(sb)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(SPIN_CYCLES)
        e0.record()
        sa.wait_stream(cur)
        sb.wait_stream(cur)
        run(steps)
        cur.wait_stream(sa)
        cur.wait_stream(sb)
        e1.record()
        torch.cuda.synchronize()
        for d_ in ds:
            d_.cta_slots = 0
        return e0.elapsed_time(e1)

    ms_two_streams = timed_two_streams(drones, K, W)
    ms_two_streams_policy = timed_two_streams(drones, K, W, policy=True)

    if args.profile:      # ncu / launch-list runs: only the kernel loops (K=8 then K=1)
        if sampler:
            sampler.stop(t0, time.time())
        k8_sorted = [round(x, 4) for x in timed_loop.last]
        d1s = [make(1, seed_off=j)[0] for j in range(NB)]
        ms_k1 = timed_rotation(d1s, K, W)
        ms_k1_fl = timed_loop(d1s[0], K_fl, W)
        if rank == 0:
            p = timed_loop.last
            os.dup2(real_stdout, 1)
            print(json.dumps({"profile_run": True, "ms_per_step_k8": ms / K, "ms_per_step_k1": ms_k1 / K,
                              "flushed_k8": ms_flushed / K_fl, "flushed_k1": ms_k1_fl / K_fl,
                              "k1_fl_min_med_max": [p[0], p[len(p) // 2], p[-1]], "k8_fl_sorted": k8_sorted[:3] + k8_sorted[-3:]}))
        return

    # ---- end to end through the public API with HOST buffers (page-locked): actions in, step, done flags out.
    # The step's result travels back as the done BITMASK (fpv_drone_io_t.done_bits: one ballot per 32 envs in the step's
    # epilogue): n / 8 bytes instead of n.  Two forms of the same call are timed and the faster one is `e2e`:
    #   zero copy   ONE launch; the step's TMA engine reads the host buffer over PCIe chunk by chunk while earlier chunks
    #               compute, the warps write the flag words straight to host memory
    #   sliced      fpv_drone_step_host: 4 env slices pipelined over H2D / step / D2H streams (round 1's form)
    # Static inputs are generated on the device and written into the pinned buffers by a D2H copy, so no line of them sits
    # dirty in a CPU cache (a buffer the CPU has just written is read ~1.7x slower: profiles/r1_h2d_cpu_cache_effect.txt);
    # `e2e_producer` below is the honest case of a CPU that REWRITES its input every step.
    from fpyv_b200 import hostmem
    from fpyv_b200.sticks import pack_crsf
    drone_bytes = drone
    drone = BatchedDrone(None, num_envs=n, device=dev, substeps=SUBSTEPS, dt=DT, auto_reset=True, thrust_lut=LUT_N, done_bits=True)
    pos_, vel_, rpy_, _ = synthetic_init(n, dev, 1234 + rank)
    drone.reset(pos_, vel_, rpy_)
    del pos_, vel_, rpy_
    drone.step(ring[0], return_obs=False)
    host_done = hostmem.pinned(((n + 31) // 32,), torch.int32)
    host_actions = [hostmem.pinned((n, 4), torch.float32) for _ in range(2)]
    for h in host_actions:
        h.copy_(torch.rand(n, 4, device=dev, generator=gen) * 2 - 1)
    host_u16 = [hostmem.pinned((n, 4), torch.uint16) for _ in range(2)]
    host_crsf = [hostmem.pinned((n, 6), torch.uint8) for _ in range(2)]
    for h16, h11 in zip(host_u16, host_crsf):
        v11 = torch.randint(0, 2048, (n, 4), dtype=torch.int32, device=dev, generator=gen)
        h16.copy_(((v11 << 5) | (v11 >> 6)).to(torch.uint16))
        bits = (v11[:, 0].long() | (v11[:, 1].long() << 11) | (v11[:, 2].long() << 22) | (v11[:, 3].long() << 33))
        h11.copy_(torch.stack([(bits >> (8 * b)) & 0xFF for b in range(6)], 1).to(torch.uint8))
    torch.cuda.synchronize()
    flagged = [0]

    def e2e_loop(call, bufs):
        """K calls, the host waits for and reads the flags every step; device time of the whole loop (events)."""
        for i in range(max(3, W)):      # the timed loop's body exactly (the first CPU-side read of the flags costs milliseconds)
            call(bufs[i % 2])
            torch.cuda.current_stream().synchronize()
            flagged[0] += int(host_done[:64].count_nonzero())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            call(bufs[i % 2])
            torch.cuda.current_stream().synchronize()      # the caller consumes the done flags every step
            flagged[0] += int(host_done[:64].count_nonzero())
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    ms_e2e_zero = e2e_loop(lambda h: drone.step_host(h, host_done, zero_copy=True), host_actions)
    ms_e2e_sliced = e2e_loop(lambda h: drone.step_host(h, host_done, slices=E2E_SLICES), host_actions)
    # sliced H2D copies (the copy engine's rate) + flag words written straight to host memory by the step kernels (no D2H)
    ms_e2e_direct = {sl: e2e_loop(lambda h, sl=sl: drone.step_host(h, host_done, slices=sl, flags_direct=True), host_actions)
                     for sl in (2, 3, 4, 6)}
    best_sl = min(ms_e2e_direct, key=ms_e2e_direct.get)
    ms_e2e_direct_all = dict(ms_e2e_direct)
    if ms_e2e_direct[best_sl] < ms_e2e_sliced:
        ms_e2e_sliced, sliced_form = ms_e2e_direct[best_sl], f"sliced copies ({best_sl} slices), flags written to host memory by the kernels"
    else:
        sliced_form = f"sliced copies ({E2E_SLICES} slices), flags copied back"
    # ---- the same end-to-end step fed the way the reference's simulator feeds it (step(action=None): raw joystick axes,
    #      components.py:227-228, :250-253) in compact transport form, calibrated on the device: uint16 x 4 (8 B/env) and the
    #      6-byte CRSF packing of four 11-bit channels (what an RC link carries).  Extra metrics; `e2e` stays on float32 actions.
    ms_st = {"u16_zero_copy": e2e_loop(lambda h: drone.step_host_sticks(h, host_done, zero_copy=True), host_u16),
             "u16_sliced": e2e_loop(lambda h: drone.step_host_sticks(h, host_done, slices=E2E_SLICES), host_u16),
             "crsf_zero_copy": e2e_loop(lambda h: drone.step_host_sticks(h, host_done, zero_copy=True), host_crsf),
             "crsf_sliced": e2e_loop(lambda h: drone.step_host_sticks(h, host_done, slices=E2E_SLICES), host_crsf)}

    # ---- a producer that REWRITES its action buffer every step (double-buffered: while the device works on buffer A the
    #      CPU fills buffer B from pageable memory), ordinary pinned memory vs write-combined memory; host wall clock
    def producer_loop(bufs, zero_copy):
        src = [torch.rand(n, 4) * 2 - 1 for _ in range(2)]
        bufs[0].copy_(src[0])
        call = (lambda h: drone.step_host(h, host_done, zero_copy=True)) if zero_copy else (lambda h: drone.step_host(h, host_done, slices=E2E_SLICES))
        for i in range(3):
            call(bufs[i % 2])
            bufs[(i + 1) % 2].copy_(src[(i + 1) % 2])
            torch.cuda.current_stream().synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            call(bufs[i % 2])
            bufs[(i + 1) % 2].copy_(src[(i + 1) % 2])      # the next step's actions are produced while this step runs
            torch.cuda.current_stream().synchronize()
            flagged[0] += int(host_done[:64].count_nonzero())
        return (time.perf_counter() - t0) * 1e3

    # the link itself: one plain pinned -> device copy of the same 16 MiB, best of 7 (the roofline `e2e` sits on)
    link_ms = None
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        drone._actions.copy_(host_actions[0], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        link_ms = e0.elapsed_time(e1) if link_ms is None else min(link_ms, e0.elapsed_time(e1))
    link_gbs = 16 * n / (link_ms * 1e-3) / 1e9
    best_zero = ms_e2e_zero <= ms_e2e_sliced
    wc = [hostmem.pinned((n, 4), torch.float32, write_combined=True) for _ in range(2)]
    ms_prod = {"pinned": producer_loop(host_actions, best_zero), "write_combined": producer_loop(wc, best_zero)}
    del wc
    for h in host_actions:      # leave the static buffers as they were (not dirty in the CPU caches)
        h.copy_(torch.rand(n, 4, device=dev, generator=gen) * 2 - 1)
    ms_e2e = min(ms_e2e_zero, ms_e2e_sliced)
    ms_e2e_sticks = min(ms_st.values())
    drone = drone_bytes

    # ---- the same workload as an open-loop rollout in ONE launch per 16 control steps (fpv_drone_rollout: state in
    #      registers across the steps).  Reported next to the headline, never as it: the headline keeps one launch and
    #      one HBM round trip of the state per control step.  4 batches x 16 steps x 16 MiB of actions >> L2.
    T_ro = 16
    ro_actions = (torch.rand(T_ro, n, 4, device=dev, generator=gen) * 2 - 1).contiguous()
    reps = max(1, K // T_ro)
    for j in range(len(drones)):
        drones[j].rollout(ro_actions, fused=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        drones[i % len(drones)].rollout(ro_actions, fused=True)
    e1.record()
    torch.cuda.synchronize()
    ms_rollout = e0.elapsed_time(e1) / (reps * T_ro)
    del ro_actions

    # ---- HBM-bound variant of the same kernel (K = 1) for the memory-roofline placement
    d1s = [make(1, seed_off=j)[0] for j in range(NB)]
    ms_k1 = timed_rotation(d1s, K, W)

    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None      # covers the K=8 loop, the e2e loop and the K=1 loop

    # ---- measured FP32 peak: a pure-FMA probe (fpv_probe_fp32) timed with events, scalar and packed forms
    def fp32_probe(packed):
        import ctypes as C
        from fpyv_b200 import _lib
        lib = _lib.load()
        sink = torch.zeros(torch.cuda.get_device_properties(dev).multi_processor_count * 8 * 256, device=dev)
        flop = C.c_double(0.0)
        best = None
        for r in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.fpv_probe_fp32(int(packed), 4096, _lib.ptr(sink), sink.numel(), C.byref(flop), _lib.current_stream(dev)))
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1)
            best = ms_ if best is None or ms_ < best else best
        return flop.value / (best * 1e-3) / 1e12
    peak_meas = {"ffma_scalar_tflops": fp32_probe(0), "ffma2_packed_tflops": fp32_probe(1)} if rank == 0 else None

    # ---- the other BASELINE configs, short legs (bench_legs.py); configs[3] runs on every rank when N > 1
    del d1s
    del drones[1:]
    torch.cuda.empty_cache()
    import bench_legs
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    pk_legs = {"hbm_gbs": peaks()["hbm_gbs"], "fp32_tflops": sm_count * FP32_LANES_PER_SM * 2 * peaks()["sm_max_mhz"] * 1e6 / 1e12}
    extra = {}
    if not args.no_extra:
        if world > 1:
            try:
                extra["config3_sharded"] = bench_legs.leg_config3_sharded(dev, pk_legs, None, world, rank)
            except Exception as e:      # noqa: BLE001
                extra["config3_sharded"] = {"error": f"{type(e).__name__}: {e}"}
            dist.barrier()
        extra.update(bench_legs.run_extra(dev, pk_legs, world, rank))
    t = torch.tensor([ms, ms_e2e, ms_k1, ms_rollout, ms_full_grid, ms_unchained, ms_e2e_sticks, ms_two_streams, ms_two_streams_policy], dtype=torch.float64, device=dev)
    t_forms = torch.tensor([ms_e2e_zero, ms_e2e_sliced] + [ms_st[k_] for k_ in sorted(ms_st)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_forms, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_k1, ms_rollout, ms_full_grid, ms_unchained, ms_e2e_sticks, ms_two_streams, ms_two_streams_policy = t.tolist()
    ms_e2e_zero, ms_e2e_sliced = t_forms.tolist()[:2]
    ms_st = dict(zip(sorted(ms_st), t_forms.tolist()[2:]))
    ms_e2e, ms_e2e_sticks = min(ms_e2e_zero, ms_e2e_sliced), min(ms_st.values())
    stats = drone.episode_stats(all_reduce=world > 1)      # the engine's only collective (NCCL), outside the timed loop
    ms_flushed_per_step = ms_flushed / K_fl
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    total_envs = n * world
    value = total_envs * K / (ms * 1e-3)
    per_gpu_launch_s = ms * 1e-3 / K
    fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    fp32_ach = FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / per_gpu_launch_s / 1e12
    hbm_ach = BYTES_PER_ENV_STEP * n / per_gpu_launch_s / 1e9
    hbm_k1 = BYTES_PER_ENV_STEP * n / (ms_k1 * 1e-3 / K) / 1e9
    roof = {"bound": "fp32", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak,
            "traffic": ncu_traffic_bytes(), "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one K=8 launch of this kernel on this workload, "
            "ncu --set full (profiles/r2_ncu_drone_step.json, tools/profile_modes.py hot); part of the state stores is still in L2 at kernel end",
            "kernel": "fpv::ring_step_kernel<DroneMode<F2, ANG=4, hot path>, 128 threads, 4 CTAs/SM, 2-slot TMA ring per warp> (K=8)",
            "peak_source": f"{sm_count} SMs x {FP32_LANES_PER_SM} FP32 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz (clocks.max.sm); "
                           "tensor cores unused by design (no dense contraction on this path)",
            "algorithmic": f"{FLOP_PER_ENV_SUBSTEP} flop/env/substep x {SUBSTEPS} substeps x {n} envs per launch",
            "frac_chained_full_grid": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_full_grid * 1e-3 / K) / 1e12 / fp32_peak,
            "frac_unchained": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_unchained * 1e-3 / K) / 1e12 / fp32_peak,
            "frac_two_streams": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_two_streams * 1e-3 / K) / 1e12 / fp32_peak,
            "frac_two_streams_closed_loop": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_two_streams_policy * 1e-3 / K) / 1e12 / fp32_peak,
            "frac_isolated_launch": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_flushed_per_step * 1e-3) / 1e12 / fp32_peak,
            "peak_measured": dict(peak_meas, unit="TFLOP/s", frac_of_measured_scalar=fp32_ach / peak_meas["ffma_scalar_tflops"],
                                  how="fpv_probe_fp32: SMs x 8 CTAs x 256 threads x 4096 iterations x 16 independent FMA chains with "
                                      "shared multiplicand/addend (operand-reuse hits), best of 6, CUDA events"),
            "hbm": {"achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"],
                    "algorithmic": f"{BYTES_PER_ENV_STEP} B/env/control-step x {n} envs per launch", "peak_source": pk["source"]},
            "hbm_bound_variant_k1": {"bound": "hbm", "achieved": hbm_k1, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                     "frac": hbm_k1 / pk["hbm_gbs"], "ms_per_step": ms_k1 / K,
                                     "env_steps_per_sec": total_envs * K / (ms_k1 * 1e-3)}}
    cpu = cpu_arm(steps=1000, warmup=1, target_seconds=10.0)
    try:
        cpu_numba = numba_arm()
    except Exception as e:      # noqa: BLE001
        cpu_numba = {"error": f"{type(e).__name__}: {e}"}
    line = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world),
            "env_substeps_per_sec": value * SUBSTEPS,
            "e2e": {"value": total_envs * K / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": 16 * n * world,
                    "d2h_bytes_per_step": 4 * ((n + 31) // 32) * world, "ms_per_step": ms_e2e / K,
                    "form": "zero copy" if ms_e2e_zero <= ms_e2e_sliced else sliced_form,
                    "ms_per_step_sliced_flags_direct_by_slices": {str(k_): v_ / K for k_, v_ in ms_e2e_direct_all.items()},
                    "link": {"bound": "pcie", "h2d_copy_gbs_measured": link_gbs, "achieved_gbs": (16 * n + 4 * ((n + 31) // 32)) / (ms_e2e / K * 1e-3) / 1e9,
                             "frac": (16 * n + 4 * ((n + 31) // 32)) / (ms_e2e / K * 1e-3) / 1e9 / link_gbs,
                             "note": "e2e moves 16 B/env in and 1 bit/env out; a bare cudaMemcpyAsync of the same 16 MiB (best of 7, "
                                     "this run) is the link rate it is compared with"},
                    "ms_per_step_zero_copy": ms_e2e_zero / K, "ms_per_step_sliced_copies": ms_e2e_sliced / K,
                    "api": "BatchedDrone(done_bits=True).step_host(pinned actions [n,4] float32, pinned done bitmask): zero copy = ONE "
                           "fpv_drone_step launch whose TMA loads read the host buffer over PCIe and whose warps write the flag "
                           "words to host memory; sliced copies = fpv_drone_step_host (4 env slices over H2D / step / D2H streams). "
                           "The host waits for and reads the flags every step.",
                    "host_buffers": "2 static page-locked action buffers alternated (not rewritten inside the loop), filled by a D2H copy "
                                    "of device-generated sticks so no line is dirty in a CPU cache; see e2e_producer for a CPU that "
                                    "rewrites its input every step"},
            "e2e_producer": {"unit": "env-steps/s", "form": "zero copy" if best_zero else "sliced copies",
                             "pinned": {"value": total_envs * K / (ms_prod["pinned"] * 1e-3), "ms_per_step": ms_prod["pinned"] / K},
                             "write_combined": {"value": total_envs * K / (ms_prod["write_combined"] * 1e-3), "ms_per_step": ms_prod["write_combined"] / K},
                             "note": "the CPU copies 16 MiB of fresh actions from pageable memory into the other input buffer while the "
                                     "device steps (torch CPU copy, all host threads); host wall clock, rank 0's own loop"},
            "e2e_raw_sticks": {"value": total_envs * K / (ms_e2e_sticks * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_e2e_sticks / K,
                               "form": min(ms_st, key=ms_st.get), "ms_per_step_by_form": {k_: v_ / K for k_, v_ in ms_st.items()},
                               "h2d_bytes_per_step": (6 if "crsf" in min(ms_st, key=ms_st.get) else 8) * n * world,
                               "d2h_bytes_per_step": 4 * ((n + 31) // 32) * world,
                               "api": "BatchedDrone.step_host_sticks(pinned raw sticks): the reference's step(action=None) joystick path, "
                                      "calibrated on the device; uint16 x 4 (8 B/env) or CRSF 4 x 11 bit (6 B/env); zero copy = the step "
                                      "itself reads the sticks from host memory and calibrates them in registers; extra metric"},
            "ms_per_step_flushed": ms_flushed_per_step,
            "ms_per_step_chained_full_grid": ms_full_grid / K, "ms_per_step_unchained": ms_unchained / K,
            "ms_per_step_two_streams": ms_two_streams / K, "ms_per_step_two_streams_closed_loop": ms_two_streams_policy / K,
            "rollout_fused": {"ms_per_step": ms_rollout, "env_steps_per_sec": total_envs / (ms_rollout * 1e-3),
                              "steps_per_launch": T_ro, "fp32_frac": FLOP_PER_ENV_SUBSTEP * SUBSTEPS * n / (ms_rollout * 1e-3) / 1e12 / (sm_count * FP32_LANES_PER_SM * 2 * pk["sm_max_mhz"] * 1e6 / 1e12),
                              "api": "BatchedDrone.rollout(actions[16, n, 4]): fpv_drone_rollout, bit-identical to 16 step() calls; "
                                     "extra metric, not the headline (the state stays in registers between control steps)"},
            "gpu_launches": K, "roofline": roof,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cpu_baseline_numba": cpu_numba, "cpu_baseline_numpy_reference": NUMPY_REFERENCE,
            "extra": extra, "clocks": clocks, "episode_stats": stats}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: the BASELINE config)")
    ap.add_argument("--profile", action="store_true", help="kernel loops only (for ncu): no e2e / CPU legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the legs for the other BASELINE configs (bench_legs.py)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
