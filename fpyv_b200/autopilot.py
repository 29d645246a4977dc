"""The reference's point-and-shoot autopilot for N drones: `Drone.calculate_needed_force_orientation`
(src/utils/components.py:258-304) driven by its `components.PID` (:15-54), one launch through the C ABI.

    ap = Autopilot(drone)                                   # reads the same params dict as the reference
    rot, force = ap.calculate_needed_force_orientation(pixel, target_position, target_radius)
    drone.step(action, wind, objects, rotation_matrix=rot, thrust_force=force)

`seen` (uint8 [n]) restricts the call to the envs whose target is in view, like the `if target_pixels.shape[1] == 0`
branch of simulator.py:103-110: the others keep their PID state and get force = NaN, which `step` reads as "no
override for this env"."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .camera import BatchedCamera

_FRAMES = {"world": 0, "drone": 1}
_MODES = {"level": 0, "frontarget": 1}


class PID:
    """State of components.PID (components.py:15-54) for n envs: columns integral, prev_derivative, previous_error,
    is_first.  The update itself runs inside the autopilot kernel."""

    def __init__(self, kP, kI, kD, dt, integral_clip=1, min_output=0.3, max_output=1, derivative_transition_rate=0.5,
                 num_envs=1, device="cuda:0"):
        self.kP, self.kI, self.kD, self.dt = kP, kI, kD, dt
        self.integral_clip, self.min_output, self.max_output = integral_clip, min_output, max_output
        self.derivative_transition_rate = derivative_transition_rate
        self.state = torch.zeros((num_envs, 4), dtype=torch.float64, device=device)
        self.reset()

    def reset(self, mask=None):
        """components.py:35-41."""
        fresh = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64, device=self.state.device)
        if mask is None:
            self.state[:] = fresh
        else:
            self.state[torch.as_tensor(mask, device=self.state.device).bool()] = fresh

    integral = property(lambda self: self.state[:, 0])
    prev_derivative = property(lambda self: self.state[:, 1])
    previous_error = property(lambda self: self.state[:, 2])
    is_first = property(lambda self: self.state[:, 3] != 0)


class Autopilot:
    def __init__(self, drone, camera: BatchedCamera | None = None):
        self._lib = _lib.load()
        self.drone = drone
        params = drone.params
        n, dev = drone.num_envs, drone.device
        self.camera = camera if camera is not None else BatchedCamera.from_params(params, n, dev)
        pns, dr = params["point_and_shoot"], params["drone"]
        self.virtual_drag_coef = pns["virtual_drag_coefficient"]          # components.py:114-118
        self.virtual_lift_coef = pns["virtual_lift_coefficient"]
        self.tof_effective_dist = pns["tof_effective_distance"]
        self.keep_distance = dr["keep_distance"]
        self.UWB_sensor_max_range = dr["UWB_sensor_max_range"]
        # Drone.__init__ overwrites the PID's output limits with the thrust limits (components.py:143-145)
        self.force_multiplier_pid = PID(**dr["force_multiplier_pid"], dt=drone.dt, num_envs=n, device=dev)
        self.prev_pixel = None                                            # components.py:146-147
        self.pixel_velocity = torch.zeros((n, 2), dtype=torch.float64, device=dev)

    def reset(self, mask=None):
        """The autopilot part of Drone.reset, components.py:166-168."""
        self.force_multiplier_pid.reset(mask)
        if mask is None or self.prev_pixel is None:
            self.prev_pixel = None
            self.pixel_velocity = torch.zeros((self.drone.num_envs, 2), dtype=torch.float64, device=self.drone.device)

    def _params(self, ref_frame, mode) -> _lib.AutopilotParams:
        if ref_frame not in _FRAMES:
            raise ValueError("Unknown reference frame")
        if mode not in _MODES:
            raise ValueError("Unknown mode")
        pid = self.force_multiplier_pid
        p = _lib.AutopilotParams()
        p.mass, p.dt = self.drone.mass, pid.dt
        p.virtual_drag_coef, p.virtual_lift_coef = self.virtual_drag_coef, self.virtual_lift_coef
        p.tof_effective_dist, p.keep_distance, p.uwb_max_range = self.tof_effective_dist, self.keep_distance, self.UWB_sensor_max_range
        p.kP, p.kI, p.kD, p.integral_clip = pid.kP, pid.kI, pid.kD, pid.integral_clip
        p.min_output, p.max_output = pid.min_output, pid.max_output
        p.derivative_transition_rate = pid.derivative_transition_rate
        p.ref_frame, p.mode = _FRAMES[ref_frame], _MODES[mode]
        p.max_throttle_force = float(self.drone.max_throttle_in_force)
        p.max_limit_iterations = 64
        return p

    def calculate_needed_force_orientation(self, pixel, target_position, target_radius=0.0, ref_frame="world",
                                           mode="level", seen=None, as_quaternion=False):
        """components.py:258-304.  pixel [n,2]; target_position [n,3] (or [3]); target_radius [n] or scalar
        (Target.calculate_distance = |p - c| - radius, :770-771).
        Returns (rotation_to_apply_force float32 [n,3,3] -- or its quaternion float32 [n,4] -- , force float32 [n])."""
        d = self.drone
        n, dev = d.num_envs, d.device
        f64 = lambda x, shape: torch.broadcast_to(torch.as_tensor(x, dtype=torch.float64, device=dev), shape).contiguous()
        px = f64(pixel, (n, 2))
        tp = f64(target_position, (n, 3))
        tr = f64(target_radius, (n,))
        sn = None if seen is None else torch.as_tensor(seen, device=dev).to(torch.uint8).contiguous()
        rot = None if as_quaternion else torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        quat = torch.empty((n, 4), dtype=torch.float32, device=dev) if as_quaternion else None
        force = torch.empty(n, dtype=torch.float32, device=dev)
        _lib.check(self._lib.fpv_autopilot(self._params(ref_frame, mode), self.camera._params(), _lib.ptr(d._state), n,
                                           d._stride, _lib.ptr(px), _lib.ptr(sn), _lib.ptr(tp), _lib.ptr(tr),
                                           _lib.ptr(self.force_multiplier_pid.state), _lib.ptr(rot), _lib.ptr(quat),
                                           _lib.ptr(force), _lib.current_stream(dev)))
        return (quat if as_quaternion else rot), force

    def convert_action2position(self, action):
        """components.py:383-387: on-screen target position in pixels, truncated to int."""
        a = torch.as_tensor(action, dtype=torch.float64, device=self.drone.device).reshape(-1, 4)
        res = torch.as_tensor(np.asarray(self.camera.resolution, dtype=np.float64), device=a.device)
        return (res / 2 * (1 + a[:, :2])).to(torch.int64)

    def point_and_shoot(self, pixel, action, ref_frame="world", mode="level", seen=None, as_quaternion=False):
        """components.py:312-381.  pixel [n,2]; action [n,4] = (target column, target row on the screen, virtual-target
        x / y offset) in [-1,1].  Updates `prev_pixel` / `pixel_velocity` like :325-330.
        Returns (rotation_to_apply_force float32 [n,3,3] or its quaternion [n,4], force float32 [n])."""
        d = self.drone
        n, dev = d.num_envs, d.device
        f64 = lambda x, shape: torch.broadcast_to(torch.as_tensor(x, dtype=torch.float64, device=dev), shape).contiguous()
        px, act = f64(pixel, (n, 2)), f64(action, (n, 4))
        sn = None if seen is None else torch.as_tensor(seen, device=dev).to(torch.uint8).contiguous()
        rot = None if as_quaternion else torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        quat = torch.empty((n, 4), dtype=torch.float32, device=dev) if as_quaternion else None
        force = torch.empty(n, dtype=torch.float32, device=dev)
        shifted = torch.empty((n, 2), dtype=torch.float64, device=dev)
        _lib.check(self._lib.fpv_point_and_shoot(self._params(ref_frame, mode), self.camera._params(), _lib.ptr(d._state), n,
                                                 d._stride, _lib.ptr(px), _lib.ptr(act), _lib.ptr(sn),
                                                 _lib.ptr(self.force_multiplier_pid.state), _lib.ptr(rot), _lib.ptr(quat),
                                                 _lib.ptr(force), _lib.ptr(shifted), _lib.current_stream(dev)))
        if self.prev_pixel is None:                                       # :325-330
            self.pixel_velocity = torch.zeros_like(shifted)
        else:
            self.pixel_velocity = (shifted - self.prev_pixel) / self.force_multiplier_pid.dt
        self.prev_pixel = shifted
        return (quat if as_quaternion else rot), force
