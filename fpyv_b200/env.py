"""`GateRaceEnv`: multi-agent gate-race environment on top of `BatchedDrone` (BASELINE.json configs[4]).

API style of the reference's tests/ma_com_simple_env.py:17-57 (old-gym 4-tuple, one observation per agent, one
reward / done per env): `obs = env.reset()`, `obs, reward, done, info = env.step(action)`.
Gate geometry is the reference's (`Gate` plane components.py:811-822, `generate_track` generators.py:7-18); the
reward and termination rules are OUR definition (include/fpv_api.h) -- the reference has no gate-passing reward, so
their parity is unpinned and checked only against oracle/gate_env_oracle.py."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .drone import BatchedDrone
from .objects import generate_track


class GateRaceEnv:
    def __init__(self, params=None, num_envs=8192, agents_per_env=32, device="cuda:0", substeps=8, dt=1e-3, gates=None,
                 track=None, w_gate=10.0, w_progress=1.0, w_crash=5.0, laps_to_finish=0, spawn_height=(1.0, 3.0),
                 seed=0, **drone_kw):
        self.num_envs, self.agents_per_env = int(num_envs), int(agents_per_env)
        self.n_agents = self.num_envs * self.agents_per_env
        self.drone = BatchedDrone(params, num_envs=self.n_agents, device=device, substeps=substeps, dt=dt,
                                  auto_reset=True, **drone_kw)
        self.device = self.drone.device
        self._lib = self.drone._lib
        if gates is None:
            tr = dict(count=8, radius=12, gate_size=5, gate_resolution=17)
            tr.update(track or {})
            gates = generate_track(**tr)
        if not 1 <= len(gates) <= _lib.MAX_GATES:
            raise ValueError(f"need 1..{_lib.MAX_GATES} gates")
        self.gates = gates
        p = self._p = _lib.GateEnvParams()
        p.n_gates, p.agents_per_env, p.laps_to_finish = len(gates), self.agents_per_env, int(laps_to_finish)
        p.w_gate, p.w_progress, p.w_crash = w_gate, w_progress, w_crash
        for i, g in enumerate(gates):
            c, nrm = np.asarray(g.position, dtype=np.float64), np.asarray(g.normal, dtype=np.float64)
            p.gates[i] = _lib.Gate(c[0], c[1], c[2], nrm[0], nrm[1], nrm[2], float(g.size) / 2, 0.0)
        n, dev = self.n_agents, self.device
        self._prev = torch.zeros((n, 2), dtype=torch.float32, device=dev)
        self._progress = torch.zeros(n, dtype=torch.int32, device=dev)
        self._agent_reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._env_reward = torch.zeros(self.num_envs, dtype=torch.float32, device=dev)
        self._env_done = torch.zeros(self.num_envs, dtype=torch.uint8, device=dev)
        self._obs = torch.zeros((self.num_envs, self.agents_per_env, _lib.ENV_OBS_FLOATS), dtype=torch.float32, device=dev)
        self._gen = torch.Generator(device=dev).manual_seed(seed)
        self.spawn_height = spawn_height
        self.agent_names = [f"agent_{a}" for a in range(self.agents_per_env)]
        self._obs_views = None
        self._chain_env = False
        self._fused_args = None
        self._done_view = None

    # -- helpers
    def _obs_dict(self):
        # views into the persistent observation buffer: built once (32 slicing calls per step would cost more host time
        # than the env step takes on the device)
        if self._obs_views is None:
            self._obs_views = {name: self._obs[:, a] for a, name in enumerate(self.agent_names)}
        return self._obs_views

    def _spawn(self):
        """Start grid behind gate 0 (on its negative side), agents of an env spread laterally."""
        n, dev = self.n_agents, self.device
        g0 = self.gates[0]
        c = torch.as_tensor(np.asarray(g0.position), dtype=torch.float32, device=dev)
        nrm = torch.as_tensor(np.asarray(g0.normal), dtype=torch.float32, device=dev)
        side = torch.as_tensor(np.asarray(g0.rotation_matrix)[:, 1], dtype=torch.float32, device=dev)
        u = torch.rand((n, 3), device=dev, generator=self._gen)
        pos = c - nrm * (3.0 + 3.0 * u[:, :1]) + side * ((u[:, 1:2] - 0.5) * 2.0 * min(float(g0.size) / 2, 4.0))
        pos[:, 2] = self.spawn_height[0] + (self.spawn_height[1] - self.spawn_height[0]) * u[:, 2]
        yaw = float(np.degrees(np.arctan2(np.asarray(g0.normal)[1], np.asarray(g0.normal)[0])))
        rpy = torch.zeros((n, 3), device=dev)
        rpy[:, 2] = yaw
        return pos, torch.zeros((n, 3), device=dev), rpy

    # -- gym-style surface
    def reset(self):
        pos, vel, rpy = self._spawn()
        self.drone.reset(pos, vel, rpy)
        d = self.drone
        _lib.check(self._lib.fpv_gate_env_reset(C.byref(self._p), _lib.ptr(d._state), self.n_agents, d._stride, None,
                                                _lib.ptr(self._prev), _lib.ptr(self._progress),
                                                _lib.current_stream(self.device)))
        self._run_env_kernel(torch.zeros(self.n_agents, dtype=torch.uint8, device=self.device), stats=False)
        return self._obs_dict()

    def _run_env_kernel(self, agent_done, stats=True):
        d = self.drone
        _lib.check(self._lib.fpv_gate_env_step(
            C.byref(self._p), _lib.ptr(d._state), self.n_agents, d._stride, _lib.ptr(agent_done), _lib.ptr(self._prev),
            _lib.ptr(self._progress), _lib.ptr(self._agent_reward), _lib.ptr(self._env_reward), _lib.ptr(self._env_done),
            _lib.ptr(self._obs), C.c_void_p(d._stats.data_ptr()) if stats else None, _lib.current_stream(self.device)))

    def step(self, action, fused=True, chained=False):
        """action: dict {agent name: [num_envs,4]} or tensor [num_envs, agents_per_env, 4] (roll, pitch, yaw, throttle).
        Returns (obs dict, reward [num_envs], done [num_envs] bool, {}).
        fused=True: dynamics and env step in ONE launch (fpv_gate_race_step: the env step is the per-chunk epilogue of the
        packed TMA-ring kernel, every agent's state is read once and written once); fused=False (or a scalar-kernel
        drone): fpv_drone_step followed by fpv_gate_env_step.  Both give the same bits.
        chained=True (fused only; open-loop use: the actions of several steps exist up front): the launch may start on the
        SMs the previous step's launch has already left (FPV_F_CHAINED, see BatchedDrone.step) -- the env's per-agent arrays
        are ordered chunk by chunk together with the state."""
        if type(action) is torch.Tensor and action.is_cuda and action.dtype is torch.float32 and action.is_contiguous() \
                and action.numel() == 4 * self.n_agents:
            act = action.view(self.n_agents, 4)          # the trainer's own device tensor: no conversion, no copy
        else:
            if isinstance(action, dict):
                action = torch.stack([torch.as_tensor(action[k]).to(self.device, torch.float32) for k in self.agent_names], dim=1)
            act = torch.as_tensor(action).to(self.device, torch.float32).reshape(self.n_agents, 4).contiguous()
        d = self.drone
        if fused and d._fast_ok and d._static is None and not (d._flags & (_lib.F_FREEZE_DONE | _lib.F_SCALAR)):
            d._last_action = act
            if chained and torch.cuda.is_current_stream_capturing():
                chained = False
            chain_now = chained and d._chain_ready and self._chain_env
            d._chain_ready = False
            d._p.flags = (d._flags | _lib.F_CHAINED) if chain_now else d._flags
            d._io.actions = act.data_ptr()
            if chained:
                d._io.epoch = d._epoch & 0xFFFFFFFF
                d._io.chunk_epoch = d._chunk_epoch_ptr
                d._chain_armed = True
            else:
                d._io.chunk_epoch = None
                d._chain_armed = False
            if self._fused_args is None:       # the env's own buffers never move: their pointers are built once
                self._fused_args = (d._p_ref, d._io_ref, C.byref(self._p), _lib.ptr(self._prev), _lib.ptr(self._progress),
                                    _lib.ptr(self._agent_reward), _lib.ptr(self._env_reward), _lib.ptr(self._env_done),
                                    _lib.ptr(self._obs))
            rc = self._lib.fpv_gate_race_step(*self._fused_args, _lib.raw_stream(d._dev_index))
            if rc:
                _lib.check(rc)
            if chained:
                d._epoch += 1
                d._chain_ready = True
            self._chain_env = chained      # the previous writer of the env arrays was a (chain-published) fused step
        else:
            self._chain_env = False
            d.step(act, return_obs=False)      # (also the first call: it configures the drone's io block)
            self._run_env_kernel(d._done)
        if self._done_view is None:
            self._done_view = self._env_done.view(torch.bool)      # a view of the persistent flag buffer, built once
        return self._obs_dict(), self._env_reward, self._done_view, {}

    @property
    def next_gate(self):
        return (self._progress & 0xffff).view(self.num_envs, self.agents_per_env)

    @property
    def laps(self):
        return (self._progress >> 16).view(self.num_envs, self.agents_per_env)

    @property
    def agent_reward(self):
        return self._agent_reward.view(self.num_envs, self.agents_per_env)

    def episode_stats(self, **kw):
        return self.drone.episode_stats(**kw)
