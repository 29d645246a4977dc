"""`BatchedRacer`: the acro rate-PID drone of the reference's tests/racer_drone_test.py (`PID` :11-32,
`Racer` :68-103) for N envs on the GPU (C ABI: fpv_racer_reset / fpv_racer_step)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class BatchedRacer:
    def __init__(self, prop_size_inch, pid_values, num_envs=1, device="cuda:0", dt=1e-3, substeps=1):
        """pid_values: {"roll": [P,I,D], "pitch": [...], "yaw": [...]} as in racer_drone_test.py:113."""
        self._lib = _lib.load()
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        self.r = (prop_size_inch / 2) * 2.54 / 100          # :70
        self.mass = 0.5                                      # :82
        self.I = self.mass * self.r ** 2 * np.ones(3)        # :83
        self.dt = float(dt)
        self.substeps = int(substeps)
        self.pid_values = {k: np.asarray(pid_values[k], dtype=np.float64) for k in ("roll", "pitch", "yaw")}
        n = self.num_envs
        self._stride = (n + 3) // 4 * 4
        self._state = torch.zeros((_lib.RACER_PLANES, self._stride, 4), dtype=torch.float32, device=self.device)
        self._torque = torch.zeros((n, 4), dtype=torch.float32, device=self.device)
        self._work = torch.zeros(32, dtype=torch.int32, device=self.device)     # chunk counters of the ring kernel
        p = self._p = _lib.RacerParams()
        p.dt, p.substeps, p.mass, p.vel_decay, p.flags = self.dt, self.substeps, self.mass, 0.9, 0
        for i, k in enumerate(("roll", "pitch", "yaw")):
            p.inertia[i] = float(self.I[i])
            for j in range(3):
                p.gains[i][j] = float(self.pid_values[k][j])

    position = property(lambda self: self._state[0, :self.num_envs, :3])
    linear_velocity = property(lambda self: self._state[1, :self.num_envs, :3])
    quaternion = property(lambda self: self._state[2, :self.num_envs])
    torque = property(lambda self: self._torque[:, :3])

    def _observe(self, want_R, want_w):
        n = self.num_envs
        R = torch.empty((n, 3, 3), dtype=torch.float32, device=self.device) if want_R else None
        w = torch.empty((n, 3), dtype=torch.float32, device=self.device) if want_w else None
        _lib.check(self._lib.fpv_racer_observe(_lib.ptr(self._state), n, self._stride, _lib.ptr(R), _lib.ptr(w),
                                               _lib.current_stream(self.device)))
        return R, w

    @property
    def orientation(self):
        """[n,3,3]: the matrix the reference keeps (racer_drone_test.py:73), converted from the quaternion plane."""
        return self._observe(True, False)[0]

    @property
    def angular_velocity(self):
        return self._observe(False, True)[1]

    def reset(self, mask=None):
        m = None if mask is None else torch.as_tensor(mask).to(self.device, torch.uint8).contiguous()
        _lib.check(self._lib.fpv_racer_reset(_lib.ptr(self._state), self.num_envs, self._stride, _lib.ptr(m),
                                             _lib.current_stream(self.device)))

    def step(self, action):
        """action [n,4] = [roll, pitch, yaw rate set-points (rad/s), thrust (N)]  (racer_drone_test.py:95-103)."""
        a = action if isinstance(action, torch.Tensor) else torch.as_tensor(np.asarray(action), dtype=torch.float32)
        a = a.to(self.device, torch.float32)
        if tuple(a.shape) != (self.num_envs, 4):
            a = torch.broadcast_to(a, (self.num_envs, 4))
        a = a.contiguous()
        _lib.check(self._lib.fpv_racer_step(C.byref(self._p), _lib.ptr(self._state), self.num_envs, self._stride,
                                            _lib.ptr(a), _lib.ptr(self._torque), _lib.ptr(self._work),
                                            _lib.current_stream(self.device)))


class Racer:
    """Single-env NumPy-returning stand-in with the reference's attribute names."""

    def __init__(self, prop_size_inch, pid_values, device="cuda:0", dt=1e-3):
        self._b = BatchedRacer(prop_size_inch, pid_values, 1, device, dt)
        self.r, self.mass, self.I = self._b.r, self._b.mass, self._b.I

    def reset(self):
        self._b.reset()

    def step(self, action):
        self._b.step(np.asarray(action, dtype=np.float64)[None])

    _np = staticmethod(lambda t: t.double().cpu().numpy())
    position = property(lambda self: self._np(self._b.position[0]))
    linear_velocity = property(lambda self: self._np(self._b.linear_velocity[0]))
    angular_velocity = property(lambda self: self._np(self._b.angular_velocity[0]))
    torque = property(lambda self: self._np(self._b.torque[0]))
    orientation_matrix = property(lambda self: self._np(self._b.orientation[0]))
