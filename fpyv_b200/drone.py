"""`BatchedDrone`: the reference's `utils.components.Drone` protocol (src/utils/components.py:72-253)
for N independent drones on one B200.  PyTorch owns the device memory; every `reset`/`step` is one
launch of the hand-written sm_100a kernels in libfpyv_b200.so through the C ABI (include/fpv_api.h).
There is no CPU path: without the CUDA library or a device this module raises.

Name-for-name mirror of the reference object:
  Drone(params)                         -> BatchedDrone(params, num_envs, device=...)
  reset(position, velocity, ypr)        -> same (ypr in DEGREES, consumed as roll, pitch, yaw: components.py:154)
  step(action, wind_velocity_vector, object_list, rotation_matrix=None, thrust_force=None)
                                        -> same; returns (R^T, gyro matrix, R @ acc) batched
  .position .velocity .state .rotation_matrix .prev_rates .prev_thrust .rates .thrust .acceleration
  .done .dt .mass .max_rates ...        -> same names, leading env axis
Attitude lives on the device as the unit quaternion of the rotation (`.quaternion`, reference convention
helper_functions.py:65-117); `.rotation_matrix` converts on read and `set_rotation_matrix` on write.
`Drone(params)` (bottom of this file) is the num_envs=1 NumPy-returning stand-in for simulator.py-style loops.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, config
from .objects import lower_object_list
from .sticks import Joystick


def _as_dev(x, device, shape=None, dtype=torch.float32):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=dtype)
    t = t.to(device=device, dtype=dtype, non_blocking=True)
    if shape is not None and tuple(t.shape) != tuple(shape):
        t = torch.broadcast_to(t, shape)
    return t.contiguous()


class BatchedDrone:
    def __init__(self, params=None, num_envs: int = 1, device="cuda:0", substeps: int = 1, dt: float | None = None,
                 auto_reset: bool = False, freeze_done: bool = False, thrust_lut: int = 0, lut_source: str = "poly",
                 packed: bool = True, ground: bool = True, joystick=None, cta_slots: int = 0, done_bits: bool = False):
        self._lib = _lib.load()
        if isinstance(params, str) or params is None:
            params = config.load_params(params)
        self.params = params
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("fpyv_b200 runs on CUDA devices only (no CPU fallback)")
        if num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.num_envs = int(num_envs)
        self.substeps = int(substeps)
        c = self.constants = config.derive_constants(params, dt=dt)

        # --- same attribute names as the reference (components.py:84-142)
        self.dim = 3
        self.max_rates = c.max_rates
        self.gravity = c.gravity
        self.dt = c.dt
        self.mass = c.mass
        self.drag_coef = c.drag_coef
        self.dimensions = c.dimensions
        self.cross_section_areas = c.cross_section_areas
        self.rates_transition_rate = c.rates_transition_rate
        self.thrust_transition_rate = c.thrust_transition_rate
        self.n_motors = config.N_MOTORS
        self.motor_radius = config.MOTOR_RADIUS
        self.radius = config.ARM_RADIUS
        self.motors_relative_position = c.motors_relative_position
        self.throttle2thrust = c.throttle2thrust
        self.thrust2throttle = c.thrust2throttle
        self.min_throttle_in_force = c.min_throttle_in_force
        self.max_throttle_in_force = c.max_throttle_in_force
        # the reference writes the thrust limits back into the caller's dict (components.py:143-144)
        if "force_multiplier_pid" in params.get("drone", {}):
            params["drone"]["force_multiplier_pid"]["min_output"] = self.min_throttle_in_force
            params["drone"]["force_multiplier_pid"]["max_output"] = self.max_throttle_in_force

        # --- stick front-end (components.py:75-81); file problems raise exactly like the reference
        self.rc = joystick if joystick is not None else Joystick(device=self.device)
        if "joystick_calib_path" in params["drone"]:
            search = (params.get("__config_dir__"),) if params.get("__config_dir__") else ()
            try:
                path = config.resolve_path(params["drone"]["joystick_calib_path"], search)
            except FileNotFoundError:
                path = params["drone"]["joystick_calib_path"]
            self.rc.calibrate(path, load_calibration_file=True)

        # --- device state: 4 float4 planes (fpv_api.h "Drone state layout")
        n = self.num_envs
        self._stride = (n + 3) // 4 * 4
        dev = self.device
        self._state = torch.zeros((_lib.DRONE_PLANES, self._stride, 4), dtype=torch.float32, device=dev)
        self._reset_state = torch.zeros_like(self._state) if auto_reset else None
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        # done_bits=True: the step also writes the flags as a bitmask (bit e % 32 of word e // 32; fpv_drone_io_t.done_bits)
        self._done_bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev) if done_bits else None
        self._acc = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self._actions = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self._stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self._work = torch.zeros(32, dtype=torch.int32, device=dev)      # chunk counters [0..15] + error words (fpv_drone_io_t.work)
        # chained launches (fpv_drone_io_t.chunk_epoch): one step count per 64-env chunk, and the host's copy of it
        self._chunk_epoch = torch.zeros((n + 63) // 64, dtype=torch.int32, device=dev)
        self._chunk_epoch_ptr = self._chunk_epoch.data_ptr()
        self._epoch = 0
        self._chain_ready = False    # True while the last writer of the state was step() itself
        self._chain_armed = False    # True while _io.chunk_epoch is set
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._lut = None
        if thrust_lut:
            self._lut = torch.from_numpy(config.thrust_table(c, int(thrust_lut), lut_source)).to(dev)
        self._flags = ((_lib.F_GROUND if ground else 0) | (_lib.F_AUTO_RESET if auto_reset else 0) |
                       (_lib.F_FREEZE_DONE if freeze_done else 0) | (_lib.F_THRUST_LUT if thrust_lut else 0) |
                       (0 if packed else _lib.F_SCALAR))
        self._p = self._make_params()
        self._io = _lib.DroneIO()
        # cta_slots > 0: the step kernel takes at most that many CTA slots per SM, so that chained launches of
        # INDEPENDENT batches stepped round-robin run side by side (fpv_drone_io_t.max_ctas_per_sm)
        self._io.max_ctas_per_sm = int(cta_slots)
        self._static = None          # (has_ground, ctypes array, count) of set_static_objects()
        self._host_done = None
        self._sticks_dev = None
        self._last_action = None
        self._is_reset = False
        self._fast_ok = False
        self._fast_flags = self._flags
        self._act_shape = torch.Size((n, 4))
        self._step_fn = self._lib.fpv_drone_step
        self._p_ref, self._io_ref = C.byref(self._p), C.byref(self._io)

    # ------------------------------------------------------------------ parameters
    def _make_params(self) -> _lib.DroneParams:
        c = self.constants
        p = _lib.DroneParams()
        p.dt, p.substeps, p.gravity, p.mass = c.dt, self.substeps, c.gravity, c.mass
        p.max_rates = c.max_rates
        p.rates_transition_rate, p.thrust_transition_rate = c.rates_transition_rate, c.thrust_transition_rate
        for i in range(3):
            p.k_drag[i] = float(c.k_drag[i])
        for m in range(4):
            p.motor_xy[m][0] = float(c.motors_relative_position[m, 0])
            p.motor_xy[m][1] = float(c.motors_relative_position[m, 1])
        p.motor_radius, p.spring_k, p.spring_c = config.MOTOR_RADIUS, config.COLLISION_SPRING, config.COLLISION_DAMPING
        for i in range(4):
            p.thrust_poly[i] = float(c.thrust_poly.coeffs[i])
        p.flags = self._flags
        p.n_objects = 0
        return p

    # ------------------------------------------------------------------ views of the state
    @property
    def position(self):
        return self._state[0, :self.num_envs, :3]

    @property
    def velocity(self):
        return self._state[1, :self.num_envs, :3]

    @property
    def state(self):
        """[n,6] = [x,y,z,vx,vy,vz] like Drone.state (a copy; write through .position/.velocity)."""
        return torch.cat([self.position, self.velocity], dim=1)

    @property
    def quaternion(self):
        """[n,4] view: attitude as (w, x, y, z)."""
        return self._state[2, :self.num_envs]

    @property
    def rotation_matrix(self):
        """[n,3,3] body->world rotation (Drone.rotation_matrix); converted from the quaternion plane on every read.
        Write with set_rotation_matrix()."""
        n = self.num_envs
        R = torch.empty((n, 3, 3), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.fpv_drone_get_rotation(_lib.ptr(self._state), n, self._stride, _lib.ptr(R),
                                                    _lib.current_stream(self.device)))
        return R

    def set_rotation_matrix(self, R, mask=None):
        n, dev = self.num_envs, self.device
        R = _as_dev(R, dev, (n, 3, 3))
        m = None if mask is None else _as_dev(mask, dev, (n,), torch.uint8)
        _lib.check(self._lib.fpv_drone_set_rotation(_lib.ptr(self._state), n, self._stride, _lib.ptr(R), _lib.ptr(m),
                                                    _lib.current_stream(dev)))
        self._chain_ready = False

    @property
    def prev_rates(self):
        return self._state[3, :self.num_envs, :3]

    rates = prev_rates          # components.py:188-189: step stores the filtered rates as both

    @property
    def prev_thrust(self):
        return self._state[0, :self.num_envs, 3]

    @property
    def thrust(self):
        """World-frame thrust vector of the state (kinematics.thrust_vector, kinematics.py:48-49)."""
        return self.rotation_matrix[:, :, 2] * self.prev_thrust[:, None]

    @property
    def acceleration(self):
        return self._acc[:, :3]

    @property
    def throttle(self):
        """Drone.throttle (components.py:186): the throttle stick of the last action."""
        return None if self._last_action is None else self._last_action[:, 3]

    @property
    def episode_steps(self):
        return self._state[1, :self.num_envs, 3].view(torch.int32)

    @property
    def done(self):
        return self._done.bool()

    @property
    def done_bits(self):
        """int32[ceil(n/32)]: the done flags of the last step as a bitmask (only with done_bits=True)."""
        return self._done_bits

    @property
    def motors_orientation(self):
        m = torch.as_tensor(self.motors_relative_position, dtype=torch.float32, device=self.device)
        return torch.einsum("mj,nij->nmi", m, self.rotation_matrix)

    # ------------------------------------------------------------------ reset / step
    def reset(self, position=None, velocity=None, ypr=None, mask=None):
        """Drone.reset (components.py:150-169).  Arguments broadcast over envs; `mask` (bool[n]) restricts
        the reset to a subset.  Defaults are params.yaml's initial_* (as simulator.py:59 passes them)."""
        n, dev, dr = self.num_envs, self.device, self.params["drone"]
        pos = _as_dev(dr["initial_position"] if position is None else position, dev, (n, 3))
        vel = _as_dev(dr["initial_velocity"] if velocity is None else velocity, dev, (n, 3))
        rpy = _as_dev(dr["initial_orientation"] if ypr is None else ypr, dev, (n, 3))
        m = None if mask is None else _as_dev(mask, dev, (n,), torch.uint8)
        _lib.check(self._lib.fpv_drone_reset(_lib.ptr(self._state), n, self._stride, _lib.ptr(pos), _lib.ptr(vel),
                                             _lib.ptr(rpy), _lib.ptr(m), _lib.current_stream(dev)))
        if m is None:
            self._done.zero_()
            self._acc.zero_()
        else:
            self._done.masked_fill_(m.bool(), 0)
        if self._reset_state is not None:
            if m is None:
                self._reset_state.copy_(self._state)
            else:
                sel = m.bool()
                self._reset_state[:, :n][:, sel] = self._state[:, :n][:, sel]
        self._is_reset = True
        self._chain_ready = False

    def set_static_objects(self, object_list):
        """Register a world that does not change between steps (the reference passes the same `object_list` to every
        `step`, simulator.py:84-85, :156): it is lowered once, `step(action)` without an object_list then steps against it,
        and such steps ride the allocation-free fast path (and may be chained).  Objects are read NOW: call again after
        moving one (e.g. `Target.update()`); `None` clears it.  An explicit `object_list=` in a call still wins."""
        if object_list is None:
            self._static = None
        else:
            has_ground, lowered = lower_object_list(object_list)
            arr = (_lib.Object * len(lowered))(*lowered) if lowered else None
            self._static = (has_ground, arr, len(lowered))
        self._fast_ok = False
        self._chain_ready = False

    def read_sticks(self):
        """components.py:250-253 on the batched joystick source."""
        return self.rc.read_actions()

    def step(self, action, wind_velocity_vector=None, object_list=None, rotation_matrix=None, thrust_force=None,
             return_obs: bool = True, chained: bool = False):
        """Drone.step (components.py:220-248) for every env: `substeps` reference steps with `action` held.
        action: [n,4] (roll, pitch, yaw, throttle) in [-1,1], or None to poll the joystick source.
        Returns (R^T [n,3,3], euler_matrix(*rates) [n,3,3], R @ acc [n,3]) when return_obs.

        chained=True (open-loop rollouts: the actions of several steps exist up front) lets this launch start on the
        SMs the previous launch on the stream has already left instead of waiting for its last chunk
        (FPV_F_CHAINED, fpv_api.h).  The caller promises that `action` was complete before the PREVIOUS launch on
        this stream was enqueued and that nothing but step() touched the state since the last step; it is ignored
        after reset()/set_*() and on the general (obstacle / override) path."""
        if not self._is_reset:
            raise RuntimeError("call reset() before step() (the reference's state is None until reset)")
        n, dev = self.num_envs, self.device
        # chunk epochs are published only by launches that ask for chaining (the first one of a run is still plain
        # stream order: it has nothing published to wait on)
        if chained and torch.cuda.is_current_stream_capturing():
            chained = False      # a captured launch would replay a stale epoch: graphs use plain stream order
        chain_now = chained and self._chain_ready
        # The chained-launch bookkeeping (epoch counter, "the last writer was step()") is committed only AFTER the launch
        # has been accepted: an exception on the way (bad action shape, non-zero return code) must not leave a published
        # epoch that no kernel will ever write -- the next chained launch would wait for it on the device.
        self._chain_ready = False
        if chained:
            self._io.epoch = self._epoch & 0xFFFFFFFF
            self._io.chunk_epoch = self._chunk_epoch_ptr
            self._chain_armed = True
        elif self._chain_armed:
            self._io.chunk_epoch = None
            self._chain_armed = False
        # fast path (the RL inner loop): device float32 actions, nothing else changed since the last full call --
        # one pointer store and one C call, no allocation (so it can be captured in a CUDA graph)
        if (self._fast_ok and wind_velocity_vector is None and object_list is None and rotation_matrix is None
                and type(action) is torch.Tensor and action.dtype is torch.float32 and action.is_cuda
                and action.shape == self._act_shape and action.is_contiguous()):
            self._last_action = action
            self._io.actions = action.data_ptr()
            self._p.flags = (self._fast_flags | _lib.F_CHAINED) if chain_now else self._fast_flags
            rc = self._step_fn(self._p_ref, self._io_ref, _lib.raw_stream(self._dev_index))
            if rc:
                _lib.check(rc)
            if chained:
                self._epoch += 1
                self._chain_ready = True
            return self.observe() if return_obs else None
        self._fast_ok = False        # re-established below only if this call completes in the plain configuration
        if action is None:
            action = self.read_sticks()
        act = _as_dev(action, dev, (n, 4))
        self._last_action = act
        p, io = self._p, self._io
        p.flags = (self._flags | _lib.F_CHAINED) if chain_now else self._flags
        wind_env = None
        if wind_velocity_vector is None:
            p.wind[0] = p.wind[1] = p.wind[2] = 0.0
        else:
            w = wind_velocity_vector
            if isinstance(w, torch.Tensor) and w.dim() == 2:
                wind_env = torch.zeros((n, 4), dtype=torch.float32, device=dev)
                wind_env[:, :3] = w.to(dev, torch.float32)
            else:
                w = np.asarray(w.cpu() if isinstance(w, torch.Tensor) else w, dtype=np.float64).reshape(3)
                p.wind[0], p.wind[1], p.wind[2] = float(w[0]), float(w[1]), float(w[2])
        objs = None
        p.n_objects = 0
        if object_list is not None:
            has_ground, lowered = lower_object_list(object_list)
            p.flags = (p.flags & ~_lib.F_GROUND) | (_lib.F_GROUND if has_ground else 0)
            if lowered:
                objs = (_lib.Object * len(lowered))(*lowered)
                p.n_objects = len(lowered)
        elif self._static is not None:        # the world registered with set_static_objects()
            has_ground, objs, p.n_objects = self._static
            p.flags = (p.flags & ~_lib.F_GROUND) | (_lib.F_GROUND if has_ground else 0)
        ovr_q = ovr_t = None
        if rotation_matrix is not None:
            if thrust_force is None:
                raise ValueError("rotation_matrix override needs thrust_force (components.py:230-232)")
            if isinstance(rotation_matrix, torch.Tensor) and rotation_matrix.shape == (n, 4):
                ovr_q = _as_dev(rotation_matrix, dev, (n, 4))      # already a quaternion (Autopilot(as_quaternion=True))
            else:
                R = _as_dev(rotation_matrix, dev, (n, 3, 3))
                ovr_q = torch.empty((n, 4), dtype=torch.float32, device=dev)
                _lib.check(self._lib.fpv_matrix_to_quat(_lib.ptr(R), n, _lib.ptr(ovr_q), _lib.current_stream(dev)))
            ovr_t = _as_dev(thrust_force, dev, (n,))
        io.state, io.n, io.plane_stride = self._state.data_ptr(), n, self._stride
        io.actions = act.data_ptr()
        io.wind_env = None if wind_env is None else wind_env.data_ptr()
        io.lut = None if self._lut is None else self._lut.data_ptr()
        io.lut_n = 0 if self._lut is None else self._lut.numel()
        io.done = self._done.data_ptr()
        io.done_bits = None if self._done_bits is None else self._done_bits.data_ptr()
        io.acc_out = self._acc.data_ptr()
        io.reset_state = None if self._reset_state is None else self._reset_state.data_ptr()
        io.override_q = None if ovr_q is None else ovr_q.data_ptr()
        io.override_thrust = None if ovr_t is None else ovr_t.data_ptr()
        io.objects = objs if objs is not None else C.POINTER(_lib.Object)()
        io.stats = self._stats.data_ptr()
        io.work = self._work.data_ptr()
        io.trace = None if getattr(self, "_trace", None) is None else self._trace.data_ptr()
        _lib.check(self._lib.fpv_drone_step(C.byref(p), C.byref(io), _lib.current_stream(dev)))
        if chained:
            self._epoch += 1
            self._chain_ready = True
        # the fast path may reuse p / io as they are only if this call left them in the plain configuration (incl. the
        # static world, whose lowered objects and ground flag stay in p / io)
        self._fast_ok = (wind_velocity_vector is None and object_list is None and rotation_matrix is None
                         and getattr(self, "_trace", None) is None)
        self._fast_flags = p.flags & ~_lib.F_CHAINED
        if return_obs:
            return self.observe()
        return None

    def _configure_plain_io(self):
        """Point the io block at this drone's own buffers in the plain configuration (no wind, objects or overrides):
        what the slow path of step() leaves behind after such a call, so the allocation-free fast path may follow."""
        p, io = self._p, self._io
        p.flags, p.n_objects = self._flags, 0
        if self._static is not None:
            self._fast_ok = False        # the static world is configured by the slow path of step()
            return
        p.wind[0] = p.wind[1] = p.wind[2] = 0.0
        io.state, io.n, io.plane_stride = self._state.data_ptr(), self.num_envs, self._stride
        io.actions = self._actions.data_ptr()
        io.wind_env = None
        io.lut = None if self._lut is None else self._lut.data_ptr()
        io.lut_n = 0 if self._lut is None else self._lut.numel()
        io.done, io.acc_out = self._done.data_ptr(), self._acc.data_ptr()
        io.done_bits = None if self._done_bits is None else self._done_bits.data_ptr()
        io.reset_state = None if self._reset_state is None else self._reset_state.data_ptr()
        io.override_q = io.override_thrust = None
        io.objects = C.POINTER(_lib.Object)()
        io.stats, io.work = self._stats.data_ptr(), self._work.data_ptr()
        io.chunk_epoch = io.trace = None
        self._chain_armed = False
        self._fast_flags = self._flags
        self._fast_ok = getattr(self, "_trace", None) is None

    @property
    def cta_slots(self) -> int:
        """CTA slots per SM one step launch takes (0 = all); see fpv_drone_io_t.max_ctas_per_sm."""
        return int(self._io.max_ctas_per_sm)

    @cta_slots.setter
    def cta_slots(self, k: int):
        self._io.max_ctas_per_sm = int(k)

    def rollout(self, actions, done_out=None, fused=True):
        """Open-loop rollout: T control steps with the stick commands of all steps given up front
        (actions [T,n,4] float32 on the device -- random exploration, MPC shooting, or recorded sticks replayed through
        `Joystick.replay`, the sim-to-real check of SURVEY section 8f row 4).
        fused=True (default): ONE launch for all T steps (fpv_drone_rollout) -- every env's state stays in registers from
        the first step to the last, so per control step only its action is read and its done flag written.
        fused=False (or a configuration the fused kernel does not cover: scalar kernel, freeze_done): one launch per
        control step, chained (FPV_F_CHAINED), each starting on the SMs the previous one has already left.
        Both are bit-identical to calling step() T times.
        done_out: optional uint8 [T,n] receiving every step's done flags.  Returns done_out (or the last flags)."""
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype is torch.float32
                and actions.dim() == 3 and tuple(actions.shape[1:]) == (self.num_envs, 4) and actions.is_contiguous()):
            raise ValueError("rollout expects a contiguous float32 CUDA tensor [T, num_envs, 4]")
        if done_out is not None and tuple(done_out.shape) != (actions.shape[0], self.num_envs):
            raise ValueError("done_out must be uint8 [T, num_envs]")
        if done_out is not None and not (done_out.is_cuda and done_out.dtype is torch.uint8 and done_out.is_contiguous()):
            raise ValueError("done_out must be a contiguous uint8 CUDA tensor")
        T = int(actions.shape[0])
        if T == 0:
            return done_out if done_out is not None else self._done
        if not self._fast_ok:       # the first call configures the io block; the rest ride the allocation-free path
            self.step(actions[0], return_obs=False, chained=True)
            if done_out is not None:
                done_out[0].copy_(self._done, non_blocking=True)
            first = 1
        else:
            first = 0
        if fused and first < T and self._static is None and not (self._flags & (_lib.F_SCALAR | _lib.F_FREEZE_DONE)):
            if not self._is_reset:
                raise RuntimeError("call reset() before step() (the reference's state is None until reset)")
            self._p.flags = self._flags
            self._io.chunk_epoch = None
            self._chain_ready = False
            self._last_action = actions[T - 1]
            n = self.num_envs
            _lib.check(self._lib.fpv_drone_rollout(self._p_ref, self._io_ref, actions[first].data_ptr(), n, T - first,
                                                   None if done_out is None else done_out[first].data_ptr(), n,
                                                   torch.cuda.current_stream(self.device).cuda_stream))
            return done_out if done_out is not None else self._done
        try:
            for t in range(first, T):
                if done_out is not None:     # every step's flags land directly in their row: nothing between the launches
                    self._io.done = done_out[t].data_ptr()
                self.step(actions[t], return_obs=False, chained=True)
        finally:
            self._io.done = self._done.data_ptr()
        if done_out is not None and T > first:
            self._done.copy_(done_out[T - 1], non_blocking=True)
        return done_out if done_out is not None else self._done

    def observe(self):
        """The tuple Drone.step returns (components.py:247-248)."""
        n, dev = self.num_envs, self.device
        Rt = torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        gyro = torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        accel = torch.empty((n, 3), dtype=torch.float32, device=dev)
        _lib.check(self._lib.fpv_drone_observe(_lib.ptr(self._state), n, self._stride, _lib.ptr(self._acc), _lib.ptr(Rt),
                                               _lib.ptr(gyro), _lib.ptr(accel), _lib.current_stream(dev)))
        return Rt, gyro, accel

    # ------------------------------------------------------------------ host-buffer entry (end-to-end path)
    def step_host(self, actions_host: torch.Tensor, done_host: torch.Tensor | None = None, slices: int = 4,
                  zero_copy: bool = False, flags_direct: bool = False):
        """One control step with HOST buffers: pinned actions [n,4] -> device, step, done flags -> pinned host.
        This is the call timed as `e2e` in bench.py.

        zero_copy=False (fpv_drone_step_host): the library cuts the batch into `slices` env ranges and pipelines them over
        three streams -- the H2D copy of slice c+1 runs while slice c is stepped and slice c-1's flags travel back.
        zero_copy=True: ONE launch of the step itself with the pinned host buffers as its `actions` input and `done_bits`
        output (fpv_drone_io_t: "may be a pinned host pointer"): the kernel's TMA engine fetches every 64-env chunk of
        actions straight over PCIe while earlier chunks compute, and the flags go back as 32-bit words written by the warps.
        No staging copy, no copy-engine hand-offs; needs a drone built with done_bits=True and page-locked buffers.
        flags_direct=True (sliced form, done_bits drones, pinned done_host): the slices' step kernels write the flag words
        straight into done_host instead of a device buffer + D2H copies (no third stream, no copy hand-offs at the end).
        Either way the caller's current stream is ordered after the flags: synchronising it means they are on the host."""
        n, dev = self.num_envs, self.device
        done_host = self._check_done_host(done_host)
        if self._static is not None:
            raise RuntimeError("step_host serves the hot-path configuration; a drone with set_static_objects() steps with step()")
        if not self._is_reset:
            raise RuntimeError("call reset() before step() (the reference's state is None until reset)")
        if not isinstance(actions_host, torch.Tensor) or tuple(actions_host.shape) != (n, 4):
            raise ValueError("step_host expects a [num_envs, 4] tensor of actions")
        if zero_copy:
            self._require_zero_copy(actions_host, done_host)
        if (not self._fast_ok or actions_host.is_cuda or actions_host.dtype is not torch.float32
                or not actions_host.is_contiguous()):
            self._actions.copy_(actions_host, non_blocking=True)      # first call / odd inputs: the plain path
            self.step(self._actions, return_obs=False)
            done_host.copy_(self._done if self._done_bits is None else self._done_bits, non_blocking=True)
            return done_host
        self._last_action = self._actions
        self._chain_ready = False
        self._p.flags = self._flags
        self._io.chunk_epoch = None
        if zero_copy:
            self._io.actions = actions_host.data_ptr()
            self._io.done_bits = done_host.data_ptr()
            try:
                _lib.check(self._step_fn(self._p_ref, self._io_ref, _lib.raw_stream(self._dev_index)))
            finally:
                self._io.actions = self._actions.data_ptr()
                self._io.done_bits = self._done_bits.data_ptr()
            self._last_action = None          # the actions never existed on the device
            return done_host
        self._io.actions = self._actions.data_ptr()
        if flags_direct:
            self._require_zero_copy(actions_host, done_host)
            self._io.done_bits = done_host.data_ptr()
            try:
                _lib.check(self._lib.fpv_drone_step_host(self._p_ref, self._io_ref, actions_host.data_ptr(), None, int(slices),
                                                         torch.cuda.current_stream(dev).cuda_stream))
            finally:
                self._io.done_bits = self._done_bits.data_ptr()
            return done_host
        _lib.check(self._lib.fpv_drone_step_host(self._p_ref, self._io_ref, actions_host.data_ptr(), done_host.data_ptr(),
                                                 int(slices), torch.cuda.current_stream(dev).cuda_stream))
        return done_host

    def _require_zero_copy(self, in_host, done_host):
        if self._done_bits is None:
            raise ValueError("zero_copy needs a drone built with done_bits=True (the flags go back as a bitmask)")
        if not (in_host.is_pinned() or hasattr(in_host, "_fpv_block")) or not (done_host.is_pinned() or hasattr(done_host, "_fpv_block")):
            raise ValueError("zero_copy needs page-locked host buffers (pin_memory=True or fpyv_b200.hostmem.pinned)")
        if in_host.data_ptr() % 16 or done_host.data_ptr() % 4:
            raise ValueError("zero_copy: the input buffer must be 16-byte aligned")

    def step_host_sticks(self, sticks_host: torch.Tensor, done_host: torch.Tensor | None = None, slices: int = 4,
                         zero_copy: bool = False):
        """`step(action=None)` -- the reference's joystick path (components.py:227-228, :250-253) -- with HOST buffers in a
        compact transport form, calibrated on the device by this drone's `rc` calibration (bit-identical to
        `rc.feed(raw); step(None)`), stepped, flags back to pinned host memory.  sticks_host (pinned):
          uint16 [n,4]  raw readings 0..65535 of axes 0, 1, 2, 5 (throttle, roll, pitch, yaw): 8 B/env   (FPV_STICKS_U16)
          uint8  [n,6]  four 11-bit channels as an RC link carries them (`sticks.pack_crsf`): 6 B/env    (FPV_STICKS_CRSF)
        zero_copy=True: one launch, the step reads the raw sticks straight from the pinned host buffer and calibrates them
        in registers (fpv_drone_io_t.sticks); else the sliced copy pipeline of fpv_drone_step_host_sticks."""
        n, dev = self.num_envs, self.device
        if self.rc._c is None:
            raise RuntimeError("Joystick is not calibrated: call calibrate(path) first")
        if self._static is not None:
            raise RuntimeError("step_host_sticks serves the hot-path configuration; a drone with set_static_objects() steps with step()")
        done_host = self._check_done_host(done_host)
        ok16 = isinstance(sticks_host, torch.Tensor) and sticks_host.dtype is torch.uint16 and tuple(sticks_host.shape) == (n, 4)
        ok11 = isinstance(sticks_host, torch.Tensor) and sticks_host.dtype is torch.uint8 and tuple(sticks_host.shape) == (n, 6)
        if not (ok16 or ok11) or sticks_host.is_cuda or not sticks_host.is_contiguous():
            raise ValueError("step_host_sticks expects a contiguous host tensor uint16 [num_envs, 4] or uint8 [num_envs, 6]")
        fmt = _lib.STICKS_U16 if ok16 else _lib.STICKS_CRSF
        if zero_copy:
            self._require_zero_copy(sticks_host, done_host)
            if sticks_host.numel() * sticks_host.element_size() % 16 and not hasattr(sticks_host, "_fpv_block"):
                raise ValueError("zero_copy: the stick buffer must be padded to a multiple of 16 bytes (fpyv_b200.hostmem.pinned)")
        if not self._is_reset:
            raise RuntimeError("call reset() before step() (the reference's state is None until reset)")
        if not self._fast_ok:      # configure the io block once through the plain path
            raw6 = torch.zeros((n, 6), dtype=torch.int32)
            if ok16:
                raw6[:, [0, 1, 2, 5]] = sticks_host.to(torch.int32)
            else:
                from .sticks import crsf_to_raw16
                b = sticks_host.numpy().astype(np.uint64)
                bits = sum(b[:, i] << np.uint64(8 * i) for i in range(6))
                v11 = np.stack([(bits >> np.uint64(11 * c)) & np.uint64(0x7FF) for c in range(4)], 1)
                raw6[:, [0, 1, 2, 5]] = torch.from_numpy(crsf_to_raw16(v11).astype(np.int32))
            self.rc.feed(raw6)
            self.step(None, return_obs=False)
            done_host.copy_(self._done if self._done_bits is None else self._done_bits, non_blocking=True)
            return done_host
        self._last_action = self._actions
        self._chain_ready = False
        self._p.flags = self._flags
        self._io.actions = self._actions.data_ptr()
        self._io.chunk_epoch = None
        if zero_copy:
            io = self._io
            io.sticks, io.stick_calib, io.stick_format = sticks_host.data_ptr(), C.pointer(self.rc._c), fmt
            io.done_bits = done_host.data_ptr()
            try:
                _lib.check(self._step_fn(self._p_ref, self._io_ref, _lib.raw_stream(self._dev_index)))
            finally:
                io.sticks, io.stick_format = None, 0
                io.done_bits = self._done_bits.data_ptr()
            self._last_action = None
            return done_host
        if self._sticks_dev is None or self._sticks_dev.numel() < 8 * n:
            self._sticks_dev = torch.empty(8 * n, dtype=torch.uint8, device=dev)
        _lib.check(self._lib.fpv_drone_step_host_sticks(self._p_ref, self._io_ref, C.byref(self.rc._c), sticks_host.data_ptr(),
                                                        self._sticks_dev.data_ptr(), fmt, done_host.data_ptr(), int(slices),
                                                        torch.cuda.current_stream(dev).cuda_stream))
        return done_host

    def _check_done_host(self, done_host):
        """The library DMAs the flags into done_host: num_envs bytes (uint8 [num_envs]) -- or, for a drone built with
        done_bits=True, the bitmask (int32 [ceil(num_envs / 32)], 1/8 of the bytes).  It must be a contiguous CPU tensor of
        exactly that shape (page-locked for full speed).  None -> this drone's own pinned buffer."""
        bits = self._done_bits is not None
        want_n = (self.num_envs + 31) // 32 if bits else self.num_envs
        want_t = torch.int32 if bits else torch.uint8
        if done_host is None:
            if self._host_done is None:
                self._host_done = torch.empty(want_n, dtype=want_t, pin_memory=True)
            return self._host_done
        if (not isinstance(done_host, torch.Tensor) or done_host.is_cuda or done_host.dtype is not want_t
                or not done_host.is_contiguous() or done_host.numel() != want_n):
            raise ValueError(f"done_host must be a contiguous CPU {want_t} tensor with {want_n} elements (ideally pinned)"
                             + (" -- this drone returns the done flags as a bitmask (done_bits=True)" if bits else ""))
        return done_host

    def _slice_bounds(self, slices):
        return host_slice_bounds(self.num_envs, slices)

    # ------------------------------------------------------------------ episode statistics
    def episode_stats(self, all_reduce: bool = False, reset: bool = False) -> dict:
        """Device-accumulated counters (fpv_stats_t).  With all_reduce the 8 doubles are summed over the
        torch.distributed world (NCCL on GPUs) -- the only collective of the engine, never inside step."""
        from .shard import reduce_stats, stats_dict
        s = reduce_stats(self._stats) if all_reduce else self._stats.clone()
        if reset:
            self._stats.zero_()
        out = stats_dict(s)
        # chained launches whose per-chunk wait timed out and fell back to a grid-wide wait (fpv_api.h, FPV_F_CHAINED):
        # non-zero means the chaining contract was broken or the device was too slow for it -- never a hang or a trap
        out["chain_timeouts"] = int(self._work[16].item())
        return out


class Drone:
    """num_envs = 1 stand-in for `utils.components.Drone` (NumPy in, NumPy out), for simulator.py-style
    loops: `drone = Drone(params); drone.reset(p, v, ypr); drone.step(action, wind, [ground])`."""

    def __init__(self, params=None, device="cuda:0", **kw):
        self._b = BatchedDrone(params, num_envs=1, device=device, **kw)
        for k in ("dim", "max_rates", "gravity", "dt", "mass", "drag_coef", "dimensions", "cross_section_areas",
                  "rates_transition_rate", "thrust_transition_rate", "n_motors", "motor_radius", "radius",
                  "motors_relative_position", "throttle2thrust", "thrust2throttle", "min_throttle_in_force",
                  "max_throttle_in_force", "rc", "params"):
            setattr(self, k, getattr(self._b, k))
        self.done = None
        self.throttle = None

    def reset(self, position, velocity, ypr):
        self._b.reset(np.asarray(position, dtype=np.float64)[None], np.asarray(velocity, dtype=np.float64)[None],
                      np.asarray(ypr, dtype=np.float64)[None])
        self.done = False

    def step(self, action, wind_velocity_vector=None, object_list=None, rotation_matrix=None, thrust_force=None):
        if action is not None:
            action = np.asarray(action, dtype=np.float64)[None]
        if rotation_matrix is not None:
            rotation_matrix = np.asarray(rotation_matrix, dtype=np.float64)[None]
            thrust_force = np.asarray([thrust_force], dtype=np.float64)
        Rt, gyro, acc = self._b.step(action, wind_velocity_vector, object_list, rotation_matrix, thrust_force)
        self.done = bool(self._b.done[0].item())
        self.throttle = float(self._b.throttle[0].item())
        return Rt[0].double().cpu().numpy(), gyro[0].double().cpu().numpy(), acc[0].double().cpu().numpy()

    def read_sticks(self):
        return self._b.read_sticks()[0].double().cpu().numpy()

    def _np(self, t):
        return t.double().cpu().numpy()

    position = property(lambda self: self._np(self._b.position[0]))
    velocity = property(lambda self: self._np(self._b.velocity[0]))
    state = property(lambda self: self._np(self._b.state[0]))
    rotation_matrix = property(lambda self: self._np(self._b.rotation_matrix[0]))
    prev_rates = property(lambda self: self._np(self._b.prev_rates[0]))
    rates = prev_rates
    prev_thrust = property(lambda self: float(self._b.prev_thrust[0].item()))
    thrust = property(lambda self: self._np(self._b.thrust[0]))
    acceleration = property(lambda self: self._np(self._b.acceleration[0]))
    motors_orientation = property(lambda self: self._np(self._b.motors_orientation[0]))


def host_slice_bounds(n: int, slices: int):
    """Env ranges of a sliced host step as the library cuts them (host_slice_bounds, fpv_api.cu): up to slices-1 equal
    ranges of at least ~65,536 envs and one short tail range (n/16, at least 65,536 envs), all starting on 64-env
    boundaries."""
    k_min = 65536
    slices = min(16, int(slices)) if int(slices) > 0 else 4
    body = n
    if slices >= 2 and n >= 4 * k_min:
        body = (n - max(n // 16, k_min)) // 64 * 64
        slices -= 1
    count = max(1, min(slices, body // k_min))
    per = -(-(-(-body // count)) // 64) * 64
    out = [(a, min(body, a + per)) for a in range(0, body, per)]
    if body < n:
        out.append((body, n))
    return out
