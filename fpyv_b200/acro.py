"""`BatchedAcroDrone` -- mode C: an acro flight-controller inner loop in front of the reference's translational model.

    stick -> rate set-point (Drone.action2force's map and low-pass, components.py:185-189)
          -> rate PID per axis (the PID of tests/racer_drone_test.py:22-32)
          -> X-quad mixer in throttle units with per-motor saturation
          -> per-motor thrust from the T-Motor F80 bench curve (components.py:133-136; shared-memory LUT on the device)
          -> body torques, Euler's rigid-body equation, body-rate quaternion kinematics
          -> thrust + drag + gravity + ground spring / crash, semi-explicit Euler (components.py:233-243)

PARITY UNPINNED: the reference itself has no motor mixer, per-motor thrust, torque or inertia;
this model is ours, defined by `oracle/acro_oracle.py`, which is also its only oracle.  Same construction inputs as
`BatchedDrone` (params.yaml dict, motor CSV, calibration JSON), same reset/step call shapes."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, config
from .drone import _as_dev


class BatchedAcroDrone:
    def __init__(self, params=None, num_envs: int = 1, device="cuda:0", substeps: int = 1, dt: float | None = None,
                 gains=None, inertia=None, kappa: float = 0.016, spin=(1.0, -1.0, 1.0, -1.0), u_min: float = -0.9,
                 u_max: float = 1.0, integral_limit: float = 0.5, thrust_lut: int = 2049, auto_reset: bool = False,
                 ground: bool = True, packed: bool = True, rate_curve=None):
        self._lib = _lib.load()
        if isinstance(params, str) or params is None:
            params = config.load_params(params)
        self.params = params
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("fpyv_b200 runs on CUDA devices only (no CPU fallback)")
        self.num_envs, self.substeps = int(num_envs), int(substeps)
        c = self.constants = config.derive_constants(params, dt=dt)
        self.dt, self.mass, self.max_rates, self.gravity = c.dt, c.mass, c.max_rates, c.gravity
        ixx = 0.3 * c.mass * config.ARM_RADIUS ** 2
        self.gains = np.asarray(gains if gains is not None else [[0.06, 0.25, 0.0006], [0.06, 0.25, 0.0006], [0.08, 0.3, 0.0]],
                                dtype=np.float64)
        self.inertia = np.asarray(inertia if inertia is not None else [ixx, ixx, 1.8 * ixx], dtype=np.float64)
        self.kappa, self.spin = float(kappa), np.asarray(spin, dtype=np.float64)
        self.u_min, self.u_max, self.integral_limit = float(u_min), float(u_max), float(integral_limit)
        n, dev = self.num_envs, self.device
        self._stride = (n + 3) // 4 * 4
        self._state = torch.zeros((_lib.ACRO_PLANES, self._stride, 4), dtype=torch.float32, device=dev)
        self._reset_state = torch.zeros_like(self._state) if auto_reset else None
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._motor = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self._stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self._work = torch.zeros(32, dtype=torch.int32, device=dev)       # chunk counters of the ring kernel
        self._lut = torch.from_numpy(config.thrust_table(c, int(thrust_lut), "poly")).to(dev) if thrust_lut else None
        self._flags = ((_lib.F_GROUND if ground else 0) | (_lib.F_AUTO_RESET if auto_reset else 0) |
                       (_lib.F_THRUST_LUT if thrust_lut else 0) | (0 if packed else _lib.F_SCALAR))
        p = self._p = _lib.AcroParams()
        p.dt, p.substeps, p.gravity, p.mass, p.max_rates = c.dt, self.substeps, c.gravity, c.mass, c.max_rates
        p.rates_transition_rate, p.thrust_transition_rate = c.rates_transition_rate, c.thrust_transition_rate
        for i in range(3):
            p.k_drag[i], p.inertia[i] = float(c.k_drag[i]), float(self.inertia[i])
            for j in range(3):
                p.gains[i][j] = float(self.gains[i, j])
        for m in range(4):
            p.motor_xy[m][0], p.motor_xy[m][1] = float(c.motors_relative_position[m, 0]), float(c.motors_relative_position[m, 1])
            p.spin[m] = float(self.spin[m])
        p.motor_radius, p.spring_k = config.MOTOR_RADIUS, config.COLLISION_SPRING
        p.integral_limit, p.kappa, p.u_min, p.u_max = self.integral_limit, self.kappa, self.u_min, self.u_max
        for i in range(4):
            p.thrust_poly[i] = float(c.thrust_poly.coeffs[i])
        # stick -> rate curve: None = the reference's linear map (components.py:185); else (centre sensitivity, max rate,
        # expo) in deg/s, one triple for all axes or one per axis -- the flight-controller "actual rates" curve
        self.rate_curve = None
        if rate_curve is not None:
            rc = np.broadcast_to(np.asarray(rate_curve, dtype=np.float64), (3, 3))
            self.rate_curve = rc.copy()
            for i in range(3):
                for j in range(3):
                    p.rate_curve[i][j] = float(rc[i, j])
            self._flags |= _lib.F_RATE_CURVE
        p.flags = self._flags

    # ------------------------------------------------------------------ state views (leading env axis)
    position = property(lambda self: self._state[0, :self.num_envs, :3])
    velocity = property(lambda self: self._state[1, :self.num_envs, :3])
    quaternion = property(lambda self: self._state[2, :self.num_envs])
    throttle = property(lambda self: self._state[0, :self.num_envs, 3])
    rate_setpoint = property(lambda self: self._state[3, :self.num_envs, :3])          # deg/s
    angular_velocity = property(lambda self: self._state[4, :self.num_envs, :3])       # rad/s, body frame
    pid_integral = property(lambda self: self._state[5, :self.num_envs, :3])
    motor_thrust = property(lambda self: self._motor)                                  # N per motor, last substep
    done = property(lambda self: self._done)

    def reset(self, position, velocity, ypr, mask=None):
        n, dev = self.num_envs, self.device
        pos, vel, rpy = _as_dev(position, dev, (n, 3)), _as_dev(velocity, dev, (n, 3)), _as_dev(ypr, dev, (n, 3))
        m = None if mask is None else _as_dev(mask, dev, (n,), torch.uint8)
        _lib.check(self._lib.fpv_acro_reset(_lib.ptr(self._state), n, self._stride, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(rpy),
                                            _lib.ptr(m), _lib.current_stream(dev)))
        if self._reset_state is not None:
            if m is None:
                self._reset_state.copy_(self._state)
            else:
                sel = m.bool()
                self._reset_state[:, :n][:, sel] = self._state[:, :n][:, sel]

    def step(self, action, wind_velocity_vector=None):
        """action [n,4] = (roll, pitch, yaw, throttle) sticks in [-1,1], the reference's action layout
        (components.py:222-226); `substeps` model steps with the sticks held.  Returns the done flags."""
        n, dev = self.num_envs, self.device
        act = _as_dev(action, dev, (n, 4))
        w = np.zeros(3) if wind_velocity_vector is None else np.asarray(wind_velocity_vector, dtype=np.float64).reshape(3)
        for i in range(3):
            self._p.wind[i] = float(w[i])
        _lib.check(self._lib.fpv_acro_step(C.byref(self._p), _lib.ptr(self._state), n, self._stride, _lib.ptr(act),
                                           _lib.ptr(self._lut), 0 if self._lut is None else self._lut.numel(),
                                           _lib.ptr(self._done), _lib.ptr(self._motor), _lib.ptr(self._reset_state),
                                           _lib.ptr(self._stats), _lib.ptr(self._work), _lib.current_stream(dev)))
        return self._done

    def rollout(self, actions, done_out=None):
        """Open-loop rollout: actions [T,n,4] float32 on the device, all T control steps in ONE launch with the state in
        registers (fpv_acro_rollout); bit-identical to T step() calls.  done_out: optional uint8 [T,n]."""
        n, dev = self.num_envs, self.device
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype is torch.float32 and actions.dim() == 3
                and tuple(actions.shape[1:]) == (n, 4) and actions.is_contiguous()):
            raise ValueError("rollout expects a contiguous float32 CUDA tensor [T, num_envs, 4]")
        if done_out is not None and not (done_out.is_cuda and done_out.dtype is torch.uint8 and done_out.is_contiguous()
                                         and tuple(done_out.shape) == (actions.shape[0], n)):
            raise ValueError("done_out must be a contiguous uint8 CUDA tensor [T, num_envs]")
        for i in range(3):
            self._p.wind[i] = 0.0
        _lib.check(self._lib.fpv_acro_rollout(C.byref(self._p), _lib.ptr(self._state), n, self._stride, _lib.ptr(actions), n,
                                              int(actions.shape[0]), _lib.ptr(self._lut), 0 if self._lut is None else self._lut.numel(),
                                              _lib.ptr(done_out), n, _lib.ptr(self._done), _lib.ptr(self._motor),
                                              _lib.ptr(self._reset_state), _lib.ptr(self._stats), _lib.current_stream(dev)))
        return done_out if done_out is not None else self._done

    def hover_throttle(self):
        """Stick throttle in [-1,1] at which the four motors carry the weight (bench cubic, components.py:136)."""
        lo, hi = -1.0, 1.0
        f = lambda u: float(self.constants.thrust_poly(100 * (u + 1) / 2)) - self.mass * self.gravity
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if f(mid) < 0 else (lo, mid)
        return 0.5 * (lo + hi)
