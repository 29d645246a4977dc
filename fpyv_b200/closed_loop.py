"""Closed-loop stepping at the open-loop rate: `TwoStreamDrones`.

A trainer whose policy needs the NEW state before it can produce the next sticks cannot chain launches (FPV_F_CHAINED wants
the sticks of the next step to exist up front), and one launch alone on the machine pays its start-up and its tail in full
(0.60-0.64 of the FP32 roofline at 1,048,576 drones against 0.76 chained).  The envs are independent, though:
split into `parts` populations, each on its own CUDA stream in plain order -- policy(state) -> step -> policy(state) -> ... --
and every step launch limited to 4 / parts of the CTA slots of an SM, the parts' launches run side by side and the start-up
and tail of one are covered by the bulk of the others.  Nothing is known ahead of time, nothing is chained; measured on B200
(bench.py `ms_per_step_two_streams`): 37.0 us per 1,048,576-drone step = the chained rate.

The result is bit-identical to stepping one `BatchedDrone` with the same per-env inputs (tests/test_gpu_ring_modes.py)."""
from __future__ import annotations

import torch

from .drone import BatchedDrone


class TwoStreamDrones:
    def __init__(self, params=None, num_envs: int = 1 << 20, device="cuda:0", parts: int = 2, **drone_kw):
        if parts not in (1, 2, 4):
            raise ValueError("parts must be 1, 2 or 4 (a step launch takes 4 / parts of an SM's four CTA slots)")
        self.num_envs, self.device = int(num_envs), torch.device(device)
        # part sizes: multiples of 64 envs (the kernel's chunk), the last part takes the remainder
        per = (self.num_envs // parts + 63) // 64 * 64
        self.bounds = [min(i * per, self.num_envs) for i in range(parts)] + [self.num_envs]
        if any(b1 <= b0 for b0, b1 in zip(self.bounds, self.bounds[1:])):
            raise ValueError(f"{num_envs} envs are too few for {parts} parts")
        drone_kw.setdefault("cta_slots", 4 // parts if parts > 1 else 0)
        self.parts = [BatchedDrone(params, num_envs=b1 - b0, device=self.device, **drone_kw)
                      for b0, b1 in zip(self.bounds, self.bounds[1:])]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.parts]
        self._forked = False

    def _slices(self, t):
        per_env = t is not None and hasattr(t, "shape") and len(t.shape) >= 2 and t.shape[0] == self.num_envs
        return [t[b0:b1] if per_env else t for b0, b1 in zip(self.bounds, self.bounds[1:])]   # (else: broadcast over envs)

    def reset(self, position=None, velocity=None, ypr=None):
        self.join()
        for d, p, v, r in zip(self.parts, self._slices(position), self._slices(velocity), self._slices(ypr)):
            d.reset(p, v, r)

    def _fork(self):
        if not self._forked:   # the parts' streams start behind whatever the caller has queued so far
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                s.wait_stream(cur)
            self._forked = True

    def step(self, policy, **step_kw):
        """One control step of every part: on part i's stream, `actions = policy(part_i, i)` (a float32 [n_i, 4] CUDA tensor
        computed from the part's current state -- anything enqueued inside the call lands on that stream) followed by
        `part_i.step(actions)`.  Returns at once; the parts run unsynchronised with each other until join()."""
        self._fork()
        step_kw.setdefault("return_obs", False)
        for i, (d, s) in enumerate(zip(self.parts, self.streams)):
            with torch.cuda.stream(s):
                d.step(policy(d, i), **step_kw)

    def step_actions(self, actions, **step_kw):
        """The same with the sticks of the whole population given as one [num_envs, 4] tensor (sliced per part)."""
        sl = self._slices(actions)
        self.step(lambda d, i: sl[i], **step_kw)

    def join(self):
        """The caller's current stream waits for every part (call before reading state on it)."""
        if self._forked:
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                cur.wait_stream(s)
            self._forked = False

    def _cat(self, name):
        self.join()
        return torch.cat([getattr(d, name) for d in self.parts])

    position = property(lambda self: self._cat("position"))
    velocity = property(lambda self: self._cat("velocity"))
    quaternion = property(lambda self: self._cat("quaternion"))
    prev_rates = property(lambda self: self._cat("prev_rates"))
    prev_thrust = property(lambda self: self._cat("prev_thrust"))
    done = property(lambda self: self._cat("done"))

    def episode_stats(self) -> dict:
        self.join()
        out: dict = {}
        for d in self.parts:
            for k, v in d.episode_stats().items():
                out[k] = out.get(k, 0) + v
        if out.get("episodes"):
            out["mean_episode_len"] = out["episode_len_sum"] / out["episodes"]
        return out
