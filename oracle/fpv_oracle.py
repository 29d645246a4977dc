"""TEST INFRASTRUCTURE ONLY -- CPU restatement (float64 NumPy, batched over envs) of the
reference's dynamics hot path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
package `fpyv_b200/` never does (it fails loudly when the CUDA library is missing).

Parity status: PINNED.  The reference's own tests hold no golden vectors,
so this restatement is pinned against the reference ITSELF, executed here through
`oracle/ref_shim.py`; the resulting trajectories are frozen in `tests/golden/*.npz` by
`oracle/make_golden.py` and re-checked by `tests/test_oracle_golden.py` on every run.

Every function cites the reference lines (relative to /root/reference) it follows.
Conventions: all arrays carry a leading env axis n; R[n,3,3] is "how the world sees the body"
(src/utils/kinematics.py:9-12); rates are in deg/s as in the reference.
"""
from __future__ import annotations

import csv
import json
from dataclasses import dataclass, field

import numpy as np

AIR_DENSITY = 1.2225          # src/utils/kinematics.py:33
N_MOTORS = 4                  # src/utils/components.py:120
MOTOR_RADIUS = 0.1            # src/utils/components.py:121
ARM_RADIUS = 5 * 2.54 / 100   # src/utils/components.py:122
SPRING_K = 100.0              # src/utils/components.py:198 (handle_collisions default)
SPRING_C = 0.0                # src/utils/components.py:198


# --------------------------------------------------------------------------------------
# config formats (params.yaml dict, calibration JSON, motor-test CSV)
# --------------------------------------------------------------------------------------
def read_motor_blocks(path):
    """Motor bench CSV -> list of (throttle_percent[], thrust_g[]) blocks.

    Follows src/utils/flight_time_calculator.py:16-40: drop the header row, strip '%' from the
    throttle column, decimal comma -> point in the thrust column, and cut a new block after
    every row whose throttle is 100 %."""
    thr, thrust = [], []
    with open(path, newline="", encoding="utf-8") as f:
        for row in csv.reader(f):
            if not row or row[0] == "Type":
                continue
            thr.append(float(row[2].replace("%", "")))
            thrust.append(float(row[3].replace(",", ".")))
    return _split_like_reference(np.array(thr), np.array(thrust))


def _split_like_reference(thr, thrust):
    """flight_time_calculator.py:33-38.  After dropping the header (iloc[1:]) the frame keeps
    its ORIGINAL integer labels 1..N, `idx` holds the LABELS of the 100 % rows, and the slice
    `frame[b:n]` is POSITIONAL; label = position + 1, so each block ends just after its
    100 % row: block k = positions [idx_k, idx_{k+1}) with idx_0 = 0."""
    labels = np.nonzero(thr == 100)[0] + 1
    idx = np.append(0, np.append(labels, len(thr)))
    blocks = []
    for b, n in zip(idx[:-1], idx[1:]):
        blocks.append((thr[b:n], thrust[b:n]))
    if len(blocks[-1][0]) == 0:
        del blocks[-1]
    return blocks


def fit_cubic_with_origin(x, y, degree=3):
    """flight_time_calculator.py:43-52 (model_xy): prepend (0,0), least-squares polyfit.
    Returns coefficients high -> low (np.poly1d order)."""
    x = np.append(0.0, x)
    y = np.append(0.0, y)
    return np.polyfit(x, y, degree)


@dataclass
class DroneConsts:
    """Everything Drone.__init__ derives (src/utils/components.py:73-147) that the step uses."""
    dt: float
    gravity: float
    mass: float
    max_rates: float
    rtr: float
    ttr: float
    k_drag: np.ndarray            # -0.5 * Cd * rho * A           kinematics.py:36
    motor_rel: np.ndarray         # [4,3] body-frame motor offsets  components.py:123-125
    poly: np.ndarray              # throttle% -> total thrust [N], high -> low
    inv_poly: np.ndarray          # thrust [N] -> throttle %        components.py:137
    min_force: float
    max_force: float
    motor_radius: float = MOTOR_RADIUS
    spring_k: float = SPRING_K
    spring_c: float = SPRING_C
    ground: bool = True
    initial_position: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 10.0]))
    initial_velocity: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.0, 0.0]))
    initial_orientation: np.ndarray = field(default_factory=lambda: np.zeros(3))


def derive_consts(params: dict, motor_csv_path: str, dt: float | None = None) -> DroneConsts:
    """components.py:84-142."""
    dr, sim = params["drone"], params["simulator"]
    g = float(sim["gravity"])
    dims = np.array(dr["dimensions"], dtype=np.float64) / 100.0
    area = np.array([dims[1] * dims[2], dims[0] * dims[2], dims[0] * dims[1]])
    cd = np.array(dr["drag_coefficients"], dtype=np.float64)
    t = np.linspace(0, 2 * np.pi, N_MOTORS + 1)[:-1]
    t = t + (t[1] - t[0]) / 2
    motor_rel = ARM_RADIUS * np.array([np.cos(t), np.sin(t), np.zeros(N_MOTORS)]).T
    thr_pct, thrust_g = read_motor_blocks(motor_csv_path)[int(dr["motor_test_report_idx"])]
    thrust_n = N_MOTORS * thrust_g / 1000 * g
    poly = fit_cubic_with_origin(thr_pct, thrust_n)
    inv_poly = fit_cubic_with_origin(thrust_n, thr_pct)
    f = lambda x: np.polyval(poly, 100 * (x + 1) / 2)          # components.py:136
    min_force = float(f(-1 + 5 / 100 * 2))                       # components.py:139-140
    if not min_force > 0:
        raise AssertionError(f"fitted thrust at the 5 % idle throttle is {min_force} N, must be positive (components.py:141)")
    return DroneConsts(
        dt=float(1 / sim["fps"]) if dt is None else float(dt), gravity=g,
        mass=dr["mass"] / 1000, max_rates=float(dr["max_rates"]),
        rtr=float(dr["rates_transition_rate"]), ttr=float(dr["thrust_transition_rate"]),
        k_drag=-0.5 * cd * AIR_DENSITY * area, motor_rel=motor_rel, poly=poly, inv_poly=inv_poly,
        min_force=min_force, max_force=float(f(1.0)),
        initial_position=np.array(dr["initial_position"], dtype=np.float64),
        initial_velocity=np.array(dr["initial_velocity"], dtype=np.float64),
        initial_orientation=np.array(dr["initial_orientation"], dtype=np.float64))


def throttle2thrust(c: DroneConsts, x):
    """components.py:136 -- no clipping of throttle or thrust."""
    return np.polyval(c.poly, 100 * (np.asarray(x, dtype=np.float64) + 1) / 2)


def thrust2throttle(c: DroneConsts, x):
    """components.py:137."""
    return np.clip(np.polyval(c.inv_poly, np.asarray(x, dtype=np.float64)) / 100 * 2 - 1, -1, 1)


# --------------------------------------------------------------------------------------
# rotations
# --------------------------------------------------------------------------------------
def euler_matrix(roll, pitch, yaw):
    """Rz(yaw) @ Ry(pitch) @ Rx(roll), src/utils/helper_functions.py:19-44.  Batched: [n]->[n,3,3]."""
    roll, pitch, yaw = (np.asarray(a, dtype=np.float64) for a in (roll, pitch, yaw))
    sr, cr = np.sin(roll), np.cos(roll)
    sp, cp = np.sin(pitch), np.cos(pitch)
    sy, cy = np.sin(yaw), np.cos(yaw)
    m = np.empty(roll.shape + (3, 3))
    m[..., 0, 0] = cy * cp
    m[..., 0, 1] = cy * sp * sr - sy * cr
    m[..., 0, 2] = cy * sp * cr + sy * sr
    m[..., 1, 0] = sy * cp
    m[..., 1, 1] = sy * sp * sr + cy * cr
    m[..., 1, 2] = sy * sp * cr - cy * sr
    m[..., 2, 0] = -sp
    m[..., 2, 1] = cp * sr
    m[..., 2, 2] = cp * cr
    return m


def intrinsic_xyz_matrix(a, b, c):
    """scipy Rotation.from_euler("XYZ", [a,b,c]).as_matrix() == Rx(a) @ Ry(b) @ Rz(c)
    (tests/racer_drone_test.py:99).  Batched."""
    a, b, c = (np.asarray(v, dtype=np.float64) for v in (a, b, c))
    sa, ca = np.sin(a), np.cos(a)
    sb, cb = np.sin(b), np.cos(b)
    sc, cc = np.sin(c), np.cos(c)
    m = np.empty(a.shape + (3, 3))
    m[..., 0, 0] = cb * cc
    m[..., 0, 1] = -cb * sc
    m[..., 0, 2] = sb
    m[..., 1, 0] = sa * sb * cc + ca * sc
    m[..., 1, 1] = -sa * sb * sc + ca * cc
    m[..., 1, 2] = -sa * cb
    m[..., 2, 0] = -ca * sb * cc + sa * sc
    m[..., 2, 1] = ca * sb * sc + sa * cc
    m[..., 2, 2] = ca * cb
    return m


def matrix_to_quaternion(R):
    """helper_functions.py:65-80, q = [w,x,y,z] (valid while 1+trace > 0)."""
    qw = np.sqrt(1 + R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]) / 2
    qx = (R[..., 2, 1] - R[..., 1, 2]) / (4 * qw)
    qy = (R[..., 0, 2] - R[..., 2, 0]) / (4 * qw)
    qz = (R[..., 1, 0] - R[..., 0, 1]) / (4 * qw)
    return np.stack([qw, qx, qy, qz], axis=-1)


def quaternion_to_matrix(q):
    """helper_functions.py:100-117."""
    qw, qx, qy, qz = (q[..., i] for i in range(4))
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = 1 - 2 * qy ** 2 - 2 * qz ** 2
    m[..., 0, 1] = 2 * qx * qy - 2 * qz * qw
    m[..., 0, 2] = 2 * qx * qz + 2 * qy * qw
    m[..., 1, 0] = 2 * qx * qy + 2 * qz * qw
    m[..., 1, 1] = 1 - 2 * qx ** 2 - 2 * qz ** 2
    m[..., 1, 2] = 2 * qy * qz - 2 * qx * qw
    m[..., 2, 0] = 2 * qx * qz - 2 * qy * qw
    m[..., 2, 1] = 2 * qy * qz + 2 * qx * qw
    m[..., 2, 2] = 1 - 2 * qx ** 2 - 2 * qy ** 2
    return m


# --------------------------------------------------------------------------------------
# stick front-end
# --------------------------------------------------------------------------------------
@dataclass
class StickCalib:
    """Joystick calibration JSON (config/frsky.json, config/calibration.json)."""
    min_vals: np.ndarray
    max_vals: np.ndarray
    sign_reverse: np.ndarray
    stick_idx: np.ndarray        # in JSON key order: Throttle, Roll, Pitch, Yaw
    stick_center: np.ndarray
    names: tuple = ("Throttle", "Roll", "Pitch", "Yaw")

    @classmethod
    def from_json(cls, path):
        with open(path) as f:
            d = json.load(f)
        keys = list(d["sticks"].keys())
        return cls(np.array(d["min_vals"], dtype=np.float64), np.array(d["max_vals"], dtype=np.float64),
                   np.array(d["sign_reverse"], dtype=np.float64),
                   np.array([d["sticks"][k]["idx"] for k in keys]),
                   np.array([d["sticks"][k]["center"] for k in keys], dtype=np.float64), tuple(keys))


def calib_read(cal: StickCalib, raw):
    """raw[n,6] -> calibrated[n,6].  src/utils/get_sticks.py:245-265: linear map of every axis
    to [-1,1], times sign_reverse, then the four sticks are re-centred piecewise-linearly."""
    raw = np.asarray(raw, dtype=np.float64)
    v = (raw - cal.min_vals) / (cal.max_vals - cal.min_vals) * 2.0 + (-1.0)
    v = v * cal.sign_reverse
    for i, c in zip(cal.stick_idx, cal.stick_center):
        x = v[..., i]
        lo = (x - (-1.0)) / (c - (-1.0)) * (0.0 - (-1.0)) + (-1.0)
        hi = (x - c) / (1.0 - c) * (1.0 - 0.0) + 0.0
        v[..., i] = np.where(x <= c, lo, hi)
    return v


def sticks_to_action(cal: StickCalib, raw):
    """Drone.read_sticks, src/utils/components.py:250-253: positional unpack
    (throttle, roll, pitch, arm, _, yaw) = calib[0..5]; action = [-roll, pitch, yaw, throttle]."""
    v = calib_read(cal, raw)
    return np.stack([-v[..., 1], v[..., 2], v[..., 5], v[..., 0]], axis=-1)


# --------------------------------------------------------------------------------------
# mode A: reference `Drone`
# --------------------------------------------------------------------------------------
class DroneState:
    """Batched mirror of the mutable fields of `Drone` (components.py:150-169)."""

    def __init__(self, n):
        self.pos = np.zeros((n, 3))
        self.vel = np.zeros((n, 3))
        self.R = np.tile(np.eye(3), (n, 1, 1))
        self.prev_rates = np.zeros((n, 3))
        self.prev_thrust = np.zeros(n)
        self.done = np.zeros(n, dtype=bool)
        self.acc = np.zeros((n, 3))
        self.rates = np.zeros((n, 3))

    def copy(self):
        s = DroneState(len(self.pos))
        for k, v in self.__dict__.items():
            setattr(s, k, v.copy())
        return s


def drone_reset(c: DroneConsts, position, velocity, rpy_deg) -> DroneState:
    """components.py:150-169 (the `ypr` argument is consumed positionally as roll, pitch, yaw)."""
    position = np.atleast_2d(np.asarray(position, dtype=np.float64))
    n = len(position)
    s = DroneState(n)
    s.pos[:] = position
    s.vel[:] = np.asarray(velocity, dtype=np.float64)
    a = np.deg2rad(np.broadcast_to(np.asarray(rpy_deg, dtype=np.float64), (n, 3)))
    s.R = euler_matrix(a[:, 0], a[:, 1], a[:, 2])
    return s


def drone_substep(c: DroneConsts, s: DroneState, action, wind=None, dt=None,
                  R_override=None, thrust_override=None, extra_objects=()):
    """One call of Drone.step (components.py:220-248), in place on `s`; returns the step's
    (R^T, gyro matrix, R @ acc) tuple.  `done` is this step's crash flag (not sticky: :236 rebinds it)."""
    dt = c.dt if dt is None else dt
    n = len(s.pos)
    action = np.broadcast_to(np.asarray(action, dtype=np.float64), (n, 4))
    wind = np.zeros(3) if wind is None else np.asarray(wind, dtype=np.float64)
    # --- action2force, components.py:179-196
    cmd = np.clip(-action[:, :3] * c.max_rates, -c.max_rates, c.max_rates)
    rates = cmd * c.rtr + s.prev_rates * (1 - c.rtr)
    s.prev_rates = rates
    thr = throttle2thrust(c, action[:, 3]) * c.ttr + s.prev_thrust * (1 - c.ttr)
    s.prev_thrust = thr
    thrust_vec = s.R[:, :, 2] * thr[:, None]                       # kinematics.py:48-49
    # --- optional override, components.py:230-232
    if R_override is not None:
        s.R = np.array(np.broadcast_to(R_override, (n, 3, 3)), dtype=np.float64)
        thrust_vec = s.R[:, :, 2] * np.broadcast_to(thrust_override, (n,))[:, None]
    # --- drag, kinematics.py:33-38 (velocity PLUS wind)
    vs = s.vel + wind
    v_body = np.einsum("nji,nj->ni", s.R, vs)
    f_body = c.k_drag * v_body * np.linalg.norm(vs, axis=1)[:, None]
    drag = np.einsum("nij,nj->ni", s.R, f_body)
    grav = np.array([0.0, 0.0, -c.gravity * c.mass])               # kinematics.py:41-45
    # --- motors + collisions, components.py:235-239 / :198-214
    motors = s.pos[:, None, :] + np.einsum("mj,nij->nmi", c.motor_rel, s.R)
    coll = np.zeros((n, 3))
    crashed = np.zeros(n, dtype=bool)
    objs = ([_GroundPlane()] if c.ground else []) + list(extra_objects)
    for obj in objs:
        d = obj.distance(motors)                                   # [n,4]
        nrm = obj.normal(motors)                                   # [n,4,3]
        hit = (d < 0).any(axis=1) & ~crashed
        crashed |= hit                                             # early return: later objects skipped
        live = ~crashed
        pen = d - c.motor_radius
        vn = np.einsum("nj,nmj->nm", s.vel, nrm)
        f = (-c.spring_k * pen - c.spring_c * vn)[:, :, None] * nrm        # kinematics.py:56-59
        f = np.where((pen < 0)[:, :, None], f, 0.0)
        coll += np.where(live[:, None], f.sum(axis=1), 0.0)
    done = crashed | (motors[:, :, 2] < 0.0).any(axis=1)          # components.py:239
    s.done = done
    # --- forces -> acceleration, components.py:242-243
    total = thrust_vec + grav + drag + coll
    acc = total / c.mass
    s.acc = acc
    s.rates = rates
    # --- integration, components.py:216-218 + kinematics.py:15-30 (attitude increment applied TWICE)
    s.pos = s.pos + s.vel * dt
    s.vel = s.vel + acc * dt
    ang = np.deg2rad(rates) * dt
    E = euler_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    Et = np.swapaxes(E, 1, 2)
    s.R = s.R @ Et @ Et
    # --- return value, components.py:247-248 (rates in deg/s fed as radians -- reference quirk)
    gyro = euler_matrix(rates[:, 0], rates[:, 1], rates[:, 2])
    return np.swapaxes(s.R, 1, 2), gyro, np.einsum("nij,nj->ni", s.R, acc)


def drone_step(c: DroneConsts, s: DroneState, action, wind=None, substeps=1, dt=None, **kw):
    """K calls of Drone.step with the same action; `done` = OR over the substeps."""
    done = np.zeros(len(s.pos), dtype=bool)
    ret = None
    for _ in range(substeps):
        ret = drone_substep(c, s, action, wind, dt, **kw)
        done |= s.done
    s.done = done
    return ret


class _GroundPlane:
    """components.py:674-680: plane z = 0 with normal +z."""

    def distance(self, p):
        return p[..., 2]

    def normal(self, p):
        nrm = np.zeros_like(p)
        nrm[..., 2] = 1.0
        return nrm


class SphereObj:
    """Target, components.py:773-777."""

    def __init__(self, center, radius):
        self.c = np.asarray(center, dtype=np.float64)
        self.r = float(radius)

    def distance(self, p):
        return np.linalg.norm(p - self.c, axis=-1) - self.r

    def normal(self, p):
        d = p - self.c
        return d / np.linalg.norm(d, axis=-1, keepdims=True)


class CylinderObj:
    """Cylinder, components.py:710-729, including its frame quirk in calculate_normal
    (the point is made RELATIVE to the base at :719 and then compared with ABSOLUTE heights)."""

    def __init__(self, position, radius, height):
        self.p = np.asarray(position, dtype=np.float64)
        self.r = float(radius)
        self.h = float(height)

    def distance(self, p):
        d2 = np.linalg.norm(p[..., :2] - self.p[:2], axis=-1) - self.r
        z = p[..., 2]
        inside = (self.p[2] < z) & (z < self.p[2] + self.h)
        dh = np.minimum(np.abs(z - self.p[2]), np.abs(z - (self.p[2] + self.h)))
        return np.where(inside, d2, np.sqrt(d2 ** 2 + dh ** 2))

    def normal(self, p):
        q = p - self.p
        z = q[..., 2]
        inside = (self.p[2] < z) & (z < self.p[2] + self.h)
        radial = np.stack([q[..., 0], q[..., 1], np.zeros_like(z)], axis=-1)
        with np.errstate(invalid="ignore", divide="ignore"):
            radial = radial / np.linalg.norm(radial, axis=-1, keepdims=True)
        below = np.abs(z - self.p[2]) < np.abs(z - (self.p[2] + self.h))
        cap = np.zeros_like(q)
        cap[..., 2] = np.where(below, -1.0, 1.0)
        return np.where(inside[..., None], radial, cap)


# --------------------------------------------------------------------------------------
# mode B: `Racer` (tests/racer_drone_test.py)
# --------------------------------------------------------------------------------------
@dataclass
class RacerConsts:
    gains: np.ndarray             # [3 axes, 3 (P,I,D)]    racer_drone_test.py:113
    mass: float = 0.5             # :82
    radius: float = (5 / 2) * 2.54 / 100   # :70
    dt: float = 1e-3              # :8
    vel_decay: float = 0.9        # :102

    @property
    def inertia(self):
        return self.mass * self.radius ** 2 * np.ones(3)     # :83


class RacerState:
    def __init__(self, n):
        self.pos = np.zeros((n, 3))
        self.vel = np.zeros((n, 3))
        self.R = np.tile(np.eye(3), (n, 1, 1))
        self.omega = np.zeros((n, 3))
        self.i_err = np.zeros((n, 3))
        self.last_err = np.zeros((n, 3))
        self.first = np.ones(n, dtype=bool)
        self.torque = np.zeros((n, 3))


def racer_step(c: RacerConsts, s: RacerState, action, orthonormalise=True):
    """Racer.step, tests/racer_drone_test.py:95-103 with PID.step :22-32."""
    n = len(s.pos)
    action = np.broadcast_to(np.asarray(action, dtype=np.float64), (n, 4))
    err = action[:, :3] - s.omega
    s.i_err = s.i_err + err * c.dt
    d_err = (err - s.last_err) / c.dt
    d_err = np.where(s.first[:, None], 0.0, d_err)
    s.first = np.zeros(n, dtype=bool)
    s.last_err = err
    g = np.asarray(c.gains, dtype=np.float64)
    s.torque = g[:, 0] * err + g[:, 1] * s.i_err + g[:, 2] * d_err
    s.omega = 1 * s.omega + s.torque * c.dt / c.inertia
    E = intrinsic_xyz_matrix(s.omega[:, 0], s.omega[:, 1], s.omega[:, 2])   # angle = omega (quirk)
    s.R = s.R @ E
    if orthonormalise:
        # scipy Rotation.from_matrix (:99) projects onto SO(3); a polar step via SVD is the same
        # map to first order and both are invisible at the 1e-5 tolerance.
        u, _, vt = np.linalg.svd(s.R)
        s.R = u @ vt
    force = action[:, 3:4] * s.R[:, :, 2]
    acc = force / c.mass
    s.vel = c.vel_decay * s.vel + acc * c.dt
    s.pos = s.pos + s.vel * c.dt
    return s
