"""Mode C has no reference implementation, so its oracle (oracle/acro_oracle.py) is unpinned AS A WHOLE -- but every piece of
it that has a reference counterpart can be pinned to the restatements that ARE pinned to the reference's golden vectors
(oracle/fpv_oracle.py: `Drone.step`, `Racer` + `PID`).  CPU only."""
import os

import numpy as np
import yaml

from conftest import CONFIG
from oracle import acro_oracle as ao
from oracle import fpv_oracle as fo


def base_consts(dt=1e-3):
    with open(os.path.join(CONFIG, "params.yaml")) as f:
        params = yaml.safe_load(f)
    return fo.derive_consts(params, os.path.join(CONFIG, "t_motos_f80_motor_test.csv"), dt=dt)


def test_translational_model_equals_the_pinned_drone_step():
    """With the rate loop silenced (zero gains, zero rate sticks, zero body rates) and both thrust filters at their steady
    state, mode C's translation -- per-motor bench-curve thrust summed on body z, body-frame drag with (v + wind), gravity, the
    ground plane's per-motor spring and crash rule, semi-explicit Euler -- must be `Drone.step`'s, for any fixed attitude:
    400 substeps from 0.05 .. 3 m (bounces, crashes) agree to 1e-11 with the restatement pinned to the reference."""
    b = base_consts()
    c = ao.default_consts(b)
    c.gains = np.zeros((3, 3))
    n = 256
    rng = np.random.default_rng(2)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0.05, 3.0, n)], 1)
    vel, rpy = rng.normal(0, 2, (n, 3)), rng.uniform(-50, 50, (n, 3))
    u0 = rng.uniform(c.u_min, c.u_max, n)
    wind = np.array([1.5, -0.7, 0.2])
    act = np.zeros((n, 4))
    act[:, 3] = u0
    sa = ao.acro_reset(c, pos, vel, rpy)
    sa.throttle = u0.copy()
    sd = fo.drone_reset(b, pos, vel, rpy)
    sd.prev_thrust = fo.throttle2thrust(b, u0)
    worst, crashes = 0.0, 0
    for _ in range(400):
        ao.acro_substep(c, sa, act, wind=wind)
        fo.drone_substep(b, sd, act, wind=wind)
        assert np.array_equal(sa.done, sd.done)
        crashes += int(sd.done.sum())
        worst = max(worst, np.abs(sa.pos - sd.pos).max(), np.abs(sa.vel - sd.vel).max())
        np.testing.assert_allclose(sa.motor_thrust.sum(1), sd.prev_thrust, rtol=1e-12, atol=1e-12)
    assert crashes > 0 and worst < 1e-11, (crashes, worst)
    assert np.abs(fo.quaternion_to_matrix(sa.q) - sd.R).max() < 1e-12       # nothing turned, on either side


def test_rate_pid_equals_the_pinned_racer_pid():
    """`rate_pid` against `Racer`'s PID (restatement pinned to racer_demo / racer_random goldens): the same error sequence
    gives the same output to the last bit of float64 while the integrator clamp is idle, first-call rule included."""
    b = base_consts()
    c = ao.default_consts(b)
    c.integral_limit = 1e9
    n, T = 64, 50
    rng = np.random.default_rng(4)
    rc = fo.RacerConsts(gains=c.gains.copy(), dt=b.dt)
    sr = fo.RacerState(n)
    sa = ao.AcroState(n)
    for t in range(T):
        err = rng.normal(0, 3, (n, 3))
        act = np.zeros((n, 4))
        omega0 = sr.omega.copy()
        act[:, :3] = err + omega0              # Racer forms err = set-point - omega itself
        fo.racer_step(rc, sr, act)
        pid = ao.rate_pid(c, sa, act[:, :3] - omega0, b.dt)
        np.testing.assert_allclose(pid, sr.torque, rtol=1e-13, atol=1e-13)
    # the clamp: the integrator's contribution never exceeds integral_limit
    c.integral_limit = 0.5
    sa = ao.AcroState(n)
    for t in range(2000):
        ao.rate_pid(c, sa, np.full((n, 3), 5.0), b.dt)
    assert np.allclose(c.gains[:, 1] * sa.integral, np.minimum(0.5, c.gains[:, 1] * 5.0 * 2000 * b.dt))


def test_stick_map_and_filters_are_action2force():
    """Rate set-point and its low-pass (components.py:185-192) are the pinned drone restatement's, step for step."""
    b = base_consts()
    c = ao.default_consts(b)
    c.gains = np.zeros((3, 3))
    n = 128
    rng = np.random.default_rng(6)
    sa = ao.acro_reset(c, np.tile([0, 0, 50.0], (n, 1)), np.zeros((n, 3)), np.zeros((n, 3)))
    sd = fo.drone_reset(b, np.tile([0, 0, 50.0], (n, 1)), np.zeros((n, 3)), np.zeros((n, 3)))
    for t in range(30):
        act = rng.uniform(-1.3, 1.3, (n, 4))          # beyond +-1: the clip at max_rates is exercised
        ao.acro_substep(c, sa, act)
        fo.drone_substep(b, sd, act)
        np.testing.assert_allclose(sa.rate_sp, sd.prev_rates, rtol=1e-14, atol=1e-14)
