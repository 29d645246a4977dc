"""TEST INFRASTRUCTURE ONLY -- float64 NumPy model of mode C ("acro"): the inner loop BASELINE.json's north_star names
(stick-to-rate map, acro rate PID, motor mixer, per-motor thrust / torque from the T-Motor F80 bench curve) feeding the
reference's translational model.

PARITY UNPINNED: the reference has no such model (SURVEY.md section 0 -- `Drone` applies the commanded rates
kinematically and has one scalar thrust; the only rate PID is `tests/racer_drone_test.py`).  This file is OUR definition
and the only oracle for `fpv_acro_step`; everything that has a reference counterpart is reused from it:

  stick -> rate set-point, throttle low-pass   Drone.action2force            src/utils/components.py:185-194
  rate PID (P, I, D on the rate error)         racer_drone_test.PID.step     tests/racer_drone_test.py:22-32
  motor positions (X layout, 45 deg + k 90)    Drone.__init__                src/utils/components.py:120-125
  per-motor thrust = bench curve / 4           throttle2thrust               src/utils/components.py:133-136
  drag, gravity, ground spring / crash,        Drone.step                    src/utils/components.py:233-243
  semi-explicit Euler translation              update_kinematic_step         src/utils/kinematics.py:21-22

New (no reference counterpart): the mixer in throttle units with per-motor saturation, body torque from the per-motor
thrusts (arm x thrust for roll / pitch, reaction torque kappa * spin * thrust for yaw), Euler's rigid-body equation with
a diagonal inertia, and body-rate quaternion kinematics q <- q (x) exp(omega dt / 2).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import fpv_oracle as fo


@dataclass
class AcroConsts:
    base: fo.DroneConsts
    gains: np.ndarray                      # [3,3] rows roll, pitch, yaw; columns P, I, D (output in throttle units)
    inertia: np.ndarray                    # [3] diagonal body inertia, kg m^2
    kappa: float = 0.016                   # rotor reaction torque per thrust [m]
    spin: np.ndarray = field(default_factory=lambda: np.array([1.0, -1.0, 1.0, -1.0]))   # motor k spin direction
    u_min: float = -0.9                    # 5 % throttle, the reference's motor idle (components.py:138-139)
    u_max: float = 1.0
    integral_limit: float = 0.5            # anti-windup clamp on the PID's I term contribution (throttle units)
    # stick -> rate curve.  None = the reference's linear map (components.py:185).  Otherwise per axis
    # (centre sensitivity [deg/s], maximum rate [deg/s], expo in [0,1]) of the flight-controller "actual rates" curve:
    #   rate(s) = s c + max(0, m - c) |s| (s^5 e + s (1 - e)),  s = -stick (the reference's sign, components.py:185)
    rate_curve: np.ndarray | None = None

    @property
    def mix(self):
        """[4,3] mixer: motor throttle += mix @ [roll, pitch, yaw] PID sums.  roll torque = sum(y_m f_m), pitch torque
        = -sum(x_m f_m), yaw torque = kappa sum(spin_m f_m): each column is the sign pattern that produces a positive
        torque about its axis."""
        m = self.base.motor_rel
        return np.stack([np.sign(m[:, 1]), -np.sign(m[:, 0]), self.spin], axis=1)


def default_consts(base: fo.DroneConsts) -> AcroConsts:
    arm = fo.ARM_RADIUS
    ixx = 0.3 * base.mass * arm ** 2       # a 750 g 5-inch frame: ~3.6e-3 kg m^2 about roll / pitch
    return AcroConsts(base=base, gains=np.array([[0.06, 0.25, 0.0006], [0.06, 0.25, 0.0006], [0.08, 0.3, 0.0]]),
                      inertia=np.array([ixx, ixx, 1.8 * ixx]))


class AcroState:
    def __init__(self, n):
        self.pos = np.zeros((n, 3))
        self.vel = np.zeros((n, 3))
        self.q = np.tile([1.0, 0, 0, 0], (n, 1))          # w, x, y, z
        self.rate_sp = np.zeros((n, 3))                   # filtered rate set-point, deg/s (Drone.prev_rates)
        self.throttle = np.full(n, -1.0)                  # filtered collective throttle in [-1, 1]
        self.omega = np.zeros((n, 3))                     # body rates, rad/s
        self.integral = np.zeros((n, 3))
        self.e_prev = np.zeros((n, 3))
        self.first = np.ones(n, dtype=bool)
        self.done = np.zeros(n, dtype=bool)
        self.motor_thrust = np.zeros((n, 4))


def acro_reset(c: AcroConsts, position, velocity, rpy_deg) -> AcroState:
    position = np.atleast_2d(np.asarray(position, dtype=np.float64))
    n = len(position)
    s = AcroState(n)
    s.pos[:] = position
    s.vel[:] = np.asarray(velocity, dtype=np.float64)
    a = np.deg2rad(np.broadcast_to(np.asarray(rpy_deg, dtype=np.float64), (n, 3)))
    R = fo.euler_matrix(a[:, 0], a[:, 1], a[:, 2])
    s.q = np.stack([fo.matrix_to_quaternion(R[i]) for i in range(n)])
    s.q *= np.where(s.q[:, :1] < 0, -1.0, 1.0)
    return s


def _quat_mul(a, b):
    w1, x1, y1, z1 = a.T
    w2, x2, y2, z2 = b.T
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=1)


def stick_to_rate(c: AcroConsts, sticks):
    """Rate set-point [deg/s] from the three rate sticks [n,3]."""
    b = c.base
    if c.rate_curve is None:
        return np.clip(-sticks * b.max_rates, -b.max_rates, b.max_rates)          # components.py:185
    rc = np.asarray(c.rate_curve, dtype=np.float64).reshape(3, 3)
    sx = np.clip(-sticks, -1.0, 1.0)
    cen, mx, ex = rc[:, 0], rc[:, 1], rc[:, 2]
    expo = np.abs(sx) * (sx ** 5 * ex + sx * (1 - ex))
    return sx * cen + np.maximum(0.0, mx - cen) * expo


def motor_thrust_curve(c: AcroConsts, u, lut=None):
    """Single-motor thrust [N] at throttle u in [-1,1]: the 4-motor bench cubic / 4, or linear interpolation in a table
    sampled uniformly on [-1,1] (the device's shared-memory LUT)."""
    if lut is None:
        return fo.throttle2thrust(c.base, u) / 4.0
    n = len(lut)
    x = (u + 1.0) * (n - 1) * 0.5
    i = np.clip(np.floor(x).astype(int), 0, n - 2)
    f = x - i
    return (lut[i] + f * (lut[i + 1] - lut[i])) / 4.0


def rate_pid(c: AcroConsts, s: AcroState, err, dt):
    """PID.step (racer_drone_test.py:22-32: no derivative term on the first call) on the rate error, plus a clamp on the
    integrator; updates the PID state in `s`.  tests/test_acro_oracle_pieces.py pins it to the Racer restatement."""
    lim = c.integral_limit / np.maximum(c.gains[:, 1], 1e-12)
    s.integral = np.clip(s.integral + err * dt, -lim, lim)
    der = np.where(s.first[:, None], 0.0, (err - s.e_prev) / dt)
    s.e_prev = err
    s.first = np.zeros(len(err), dtype=bool)
    return c.gains[:, 0] * err + c.gains[:, 1] * s.integral + c.gains[:, 2] * der


def acro_substep(c: AcroConsts, s: AcroState, action, wind=None, dt=None, lut=None):
    b = c.base
    dt = b.dt if dt is None else dt
    n = len(s.pos)
    action = np.broadcast_to(np.asarray(action, dtype=np.float64), (n, 4))
    wind = np.zeros(3) if wind is None else np.asarray(wind, dtype=np.float64)
    # --- stick -> rate set-point and collective throttle, low-passed like action2force (components.py:185-194)
    cmd = stick_to_rate(c, action[:, :3])
    s.rate_sp = cmd * b.rtr + s.rate_sp * (1 - b.rtr)
    s.throttle = action[:, 3] * b.ttr + s.throttle * (1 - b.ttr)
    sp = np.deg2rad(s.rate_sp)
    # --- rate PID (racer_drone_test.py:22-32) with an integrator clamp
    pid = rate_pid(c, s, sp - s.omega, dt)                                                  # [n,3], throttle units
    # --- mixer in throttle units, per-motor saturation, bench curve -> per-motor thrust
    u = np.clip(s.throttle[:, None] + pid @ c.mix.T, c.u_min, c.u_max)                      # [n,4]
    f = motor_thrust_curve(c, u, lut)
    s.motor_thrust = f
    # --- body torque and Euler's equation (diagonal inertia)
    m = b.motor_rel
    tau = np.stack([f @ m[:, 1], -(f @ m[:, 0]), c.kappa * (f @ c.spin)], axis=1)
    Iw = c.inertia * s.omega
    wdot = (tau - np.cross(s.omega, Iw)) / c.inertia
    # --- translation: the reference's force model on the CURRENT attitude (components.py:233-243)
    R = fo.quaternion_to_matrix(s.q)
    thrust_vec = R[:, :, 2] * f.sum(axis=1)[:, None]
    vs = s.vel + wind
    v_body = np.einsum("nji,nj->ni", R, vs)
    drag = np.einsum("nij,nj->ni", R, b.k_drag * v_body * np.linalg.norm(vs, axis=1)[:, None])
    grav = np.array([0.0, 0.0, -b.gravity * b.mass])
    motors = s.pos[:, None, :] + np.einsum("mj,nij->nmi", m, R)
    mz = motors[:, :, 2]
    crashed = (mz < 0).any(axis=1)
    pen = mz - b.motor_radius
    spring = np.where(pen < 0, -b.spring_k * pen, 0.0).sum(axis=1)
    coll = np.zeros((n, 3))
    coll[:, 2] = np.where(crashed, 0.0, spring) if b.ground else 0.0
    s.done = crashed if b.ground else np.zeros(n, dtype=bool)
    acc = (thrust_vec + grav + drag + coll) / b.mass
    s.pos = s.pos + s.vel * dt                                                               # kinematics.py:21-22
    s.vel = s.vel + acc * dt
    # --- rotation: omega first (semi-implicit), then q <- q (x) exp(omega dt / 2), renormalised
    s.omega = s.omega + wdot * dt
    half = 0.5 * dt * s.omega
    ang = np.linalg.norm(half, axis=1, keepdims=True)
    sinc = np.where(ang > 1e-12, np.sin(ang) / np.maximum(ang, 1e-300), 1.0)
    dq = np.concatenate([np.cos(ang), half * sinc], axis=1)
    s.q = _quat_mul(s.q, dq)
    s.q /= np.linalg.norm(s.q, axis=1, keepdims=True)
    return acc


def acro_step(c: AcroConsts, s: AcroState, action, wind=None, substeps=1, dt=None, lut=None):
    done = np.zeros(len(s.pos), dtype=bool)
    acc = None
    for _ in range(substeps):
        acc = acro_substep(c, s, action, wind, dt, lut)
        done |= s.done
    s.done = done
    return acc
