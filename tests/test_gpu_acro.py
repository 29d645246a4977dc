"""Mode C ("acro": rate PID -> mixer -> per-motor bench-curve thrust -> rigid body, on the reference's translational
model) on the GPU against its float64 model oracle/acro_oracle.py.  PARITY UNPINNED: there is no reference
implementation of this model; the tolerance is the path's own: <= 1e-5 relative per step."""
import os

import numpy as np
import pytest
import torch

from conftest import CONFIG
from oracle import acro_oracle as ao
from oracle import fpv_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def consts(dt):
    import yaml
    with open(os.path.join(CONFIG, "params.yaml")) as f:
        params = yaml.safe_load(f)
    return ao.default_consts(fo.derive_consts(params, os.path.join(CONFIG, "t_motos_f80_motor_test.csv"), dt=dt))


def state_err(d, s):
    g = lambda a, b: np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))
    f = lambda t: t.double().cpu().numpy()
    q = f(d.quaternion)
    q = q * np.where(np.sum(q * s.q, axis=1, keepdims=True) < 0, -1.0, 1.0)
    return np.max(np.stack([g(f(d.position), s.pos), g(f(d.velocity), s.vel), g(q, s.q), g(f(d.angular_velocity), s.omega),
                            g(f(d.rate_setpoint), s.rate_sp), g(f(d.throttle)[:, None], s.throttle[:, None]),
                            g(f(d.pid_integral), s.integral)]), axis=0)


def seeded(n, seed, z_lo=2.0, z_hi=12.0):
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(z_lo, z_hi, n)], axis=1)
    return rng, pos, rng.normal(0, 1, (n, 3)), rng.uniform(-30, 30, (n, 3))


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("lut", [2049, 0])
@pytest.mark.parametrize("dt,K", [(1e-3, 1), (1e-3, 8), (1 / 60, 1)])
def test_single_control_steps_vs_model(dt, K, lut, packed):
    """Every control step restarts from the model's own float64 state (rounded to float32), so the error is one
    step's worth: <= 1e-5 relative."""
    from fpyv_b200 import BatchedAcroDrone
    n, T = 256, 25
    rng, pos, vel, rpy = seeded(n, 11)
    c = consts(dt)
    table = None
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=dt, thrust_lut=lut, packed=packed)
    if lut:
        table = d._lut.double().cpu().numpy()
    s = ao.acro_reset(c, pos, vel, rpy)
    d.reset(pos, vel, rpy)
    worst = 0.0
    f32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=DEV)
    for t in range(T):
        act = rng.uniform(-1, 1, (n, 4))
        # restart the device from the model's state
        d.position.copy_(f32(s.pos)); d.velocity.copy_(f32(s.vel)); d.quaternion.copy_(f32(s.q))
        d.throttle.copy_(f32(s.throttle)); d.rate_setpoint.copy_(f32(s.rate_sp)); d.angular_velocity.copy_(f32(s.omega))
        d.pid_integral.copy_(f32(s.integral)); d._state[6, :n, :3].copy_(f32(s.e_prev))
        d._state[3, :n, 3].copy_(f32(s.first.astype(np.float64)))
        # ... and the model from the device's float32 rounding of it
        f = lambda x: x.double().cpu().numpy()
        s.pos, s.vel, s.q, s.throttle = f(d.position), f(d.velocity), f(d.quaternion), f(d.throttle)
        s.rate_sp, s.omega, s.integral, s.e_prev = f(d.rate_setpoint), f(d.angular_velocity), f(d.pid_integral), f(d._state[6, :n, :3])
        ao.acro_step(c, s, act, substeps=K, lut=table)
        done = d.step(act).cpu().numpy().astype(bool)
        assert np.array_equal(done, s.done)
        worst = max(worst, state_err(d, s).max())
        np.testing.assert_allclose(d.motor_thrust.double().cpu().numpy(), s.motor_thrust, rtol=2e-5, atol=2e-5)
    print(f"acro dt={dt:.4g} K={K} lut={lut} packed={packed}: max single-step rel err {worst:.2e}")
    assert worst < 1e-5


def test_free_running_1s_divergence_and_physics():
    """1 s horizon (125 control steps x 8 substeps of 1 ms) free-running vs the model, then physical sanity: zero
    sticks at hover throttle hold attitude; a held roll stick converges to the commanded rate (-stick * max_rates, the
    reference's sign, components.py:185); saturated motors stay inside the bench curve's range."""
    from fpyv_b200 import BatchedAcroDrone
    n, dt, K = 128, 1e-3, 8
    rng, pos, vel, rpy = seeded(n, 12, z_lo=20, z_hi=40)
    c = consts(dt)
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=dt, thrust_lut=0)
    d.reset(pos, vel, rpy)
    s = ao.acro_reset(c, pos, vel, rpy)
    curve = []
    for t in range(125):
        act = rng.uniform(-1, 1, (n, 4)) if t % 10 == 0 else act
        ao.acro_step(c, s, act, substeps=K)
        d.step(act)
        if t + 1 in (1, 2, 5, 10, 30, 60, 125):
            curve.append((t + 1, float(np.median(state_err(d, s))), float(state_err(d, s).max())))
    print("acro divergence vs float64 model (control step: median, max):", " ".join(f"@{t}:{m:.1e},{x:.1e}" for t, m, x in curve))
    assert curve[0][2] < 1e-5 and curve[-1][1] < 1e-3
    # hover: level, zero sticks
    d = BatchedAcroDrone(None, num_envs=4, device=DEV, substeps=K, dt=dt)
    uh = d.hover_throttle()
    d.reset(np.tile([0, 0, 10.0], (4, 1)), np.zeros((4, 3)), np.zeros((4, 3)))
    d.throttle.fill_(uh)
    for _ in range(125):
        d.step(np.tile([0, 0, 0, uh], (4, 1)))
    assert torch.allclose(d.quaternion[:, 0], torch.ones(4, device=DEV), atol=1e-6)
    assert float(d.position[:, 2].sub(10).abs().max()) < 0.05 and float(d.angular_velocity.abs().max()) < 1e-4
    assert abs(float(d.motor_thrust.sum(1)[0]) - d.mass * d.gravity) < 0.02
    # roll stick 0.5 -> -100 deg/s
    for _ in range(60):
        d.step(np.tile([0.5, 0, 0, uh], (4, 1)))
    rate = np.rad2deg(d.angular_velocity.cpu().numpy()[0])
    assert abs(rate[0] + 100.0) < 5.0 and abs(rate[1]) < 1.0 and abs(rate[2]) < 1.0
    # full deflection on every axis at full throttle: motors clipped to the curve's range
    for _ in range(20):
        d.step(np.tile([1.0, -1.0, 1.0, 1.0], (4, 1)))
    mt = d.motor_thrust.cpu().numpy()
    top = float(d.constants.throttle2thrust(1.0)) / 4
    idle = float(d.constants.throttle2thrust(-0.9)) / 4
    assert (mt <= top * (1 + 1e-5)).all() and (mt >= idle * (1 - 1e-4)).all()


def test_ground_crash_auto_reset_and_errors():
    from fpyv_b200 import BatchedAcroDrone, FpvError
    n = 512
    rng, pos, vel, rpy = seeded(n, 13, z_lo=0.15, z_hi=1.0)
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=8, dt=1e-3, auto_reset=True)
    d.reset(pos, vel, rpy)
    snap = d._state.clone()
    crashed = torch.zeros(n, dtype=torch.bool, device=DEV)
    for _ in range(80):                      # motors at idle: everything falls
        done = d.step(np.tile([0, 0, 0, -1.0], (n, 1))).bool()
        just = done & ~crashed
        if just.any():                       # a crashed env restarts from the reset snapshot
            assert torch.equal(d._state[:, :n][:, just][..., :3], snap[:, :n][:, just][..., :3])
        crashed |= done
    assert crashed.float().mean() > 0.5
    assert float(d._stats[1]) == float(d._stats[2]) > 0          # crashes == episodes under auto-reset
    with pytest.raises(FpvError):
        bad = BatchedAcroDrone(None, num_envs=4, device=DEV, inertia=[0.0, 1e-3, 1e-3])
        bad.reset(np.zeros((4, 3)), np.zeros((4, 3)), np.zeros((4, 3)))
        bad.step(np.zeros((4, 4)))
    with pytest.raises(FpvError):            # the table covers throttle [-1, 1]: limits outside it are refused, not extrapolated
        bad = BatchedAcroDrone(None, num_envs=4, device=DEV, u_max=1.2, thrust_lut=2049)
        bad.reset(np.zeros((4, 3)), np.zeros((4, 3)), np.zeros((4, 3)))
        bad.step(np.zeros((4, 4)))


@pytest.mark.parametrize("lut_n", [2, 3, 2048, 2049])
def test_motor_table_ends_and_knots(lut_n):
    """The table index comes out of the float mantissa (acro_thrust_lut): full throttle lands on the LAST entry (index lut_n - 1,
    interpolation weight 0 -- the staged table carries a padding entry there), idle on the first, and a throttle exactly on a
    knot returns that entry.  Gains are zero, so every motor sees the filtered collective throttle only."""
    from fpyv_b200 import BatchedAcroDrone
    n = 64
    zero_gains = [[0.0, 0.0, 0.0]] * 3
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=1, dt=1e-3, thrust_lut=lut_n, u_min=-1.0, u_max=1.0, gains=zero_gains)
    table = d._lut.double().cpu().numpy()
    pos = np.tile([0.0, 0.0, 50.0], (n, 1))
    d.reset(pos, np.zeros((n, 3)), np.zeros((n, 3)))
    knots = np.linspace(-1.0, 1.0, lut_n)
    thr = np.resize(np.concatenate([[1.0, -1.0], knots[:: max(1, lut_n // 16)], np.random.default_rng(3).uniform(-1, 1, 32)]), n)
    d.throttle.copy_(torch.as_tensor(thr, dtype=torch.float32, device=DEV))
    act = np.zeros((n, 4))
    act[:, 3] = thr                          # the low-pass filter of a throttle equal to its input is the identity
    d.step(act)
    x = (d.throttle.double().cpu().numpy() + 1.0) * (lut_n - 1) * 0.5
    i = np.clip(np.floor(x).astype(int), 0, lut_n - 2)
    want = (table[i] + (x - i) * (table[i + 1] - table[i])) / 4.0
    got = d.motor_thrust.double().cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, np.repeat(want[:, None], 4, 1), rtol=3e-6, atol=3e-6)
    assert abs(got[0, 0] - table[-1] / 4.0) <= 1e-6 * abs(table[-1]) and abs(got[1, 0] - table[0] / 4.0) <= 1e-6


@pytest.mark.parametrize("n", [1, 127, 129, 1000])
def test_ragged_sizes_leave_padding_and_guards_untouched(n):
    """Envs beyond n (the padding up to plane_stride) and sentinel rows around the done / motor outputs stay as they
    were; the first n envs match a larger batch stepped with the same inputs."""
    import ctypes as C
    from fpyv_b200 import BatchedAcroDrone, _lib
    rng, pos, vel, rpy = seeded(1024, 21)
    big = BatchedAcroDrone(None, num_envs=1024, device=DEV, substeps=4, dt=1e-3)
    big.reset(pos, vel, rpy)
    act = rng.uniform(-1, 1, (1024, 4))
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=4, dt=1e-3)
    d.reset(pos[:n], vel[:n], rpy[:n])
    d._state[:, n:] = 123.0                                      # padding envs
    lib = _lib.load()
    G = 256
    done = torch.full((G + n + G,), 0xAB, dtype=torch.uint8, device=DEV)
    motor = torch.full((G + n + G, 4), -7.0, dtype=torch.float32, device=DEV)
    a = torch.as_tensor(act[:n], dtype=torch.float32, device=DEV).contiguous()
    for _ in range(3):
        big.step(act)
        _lib.check(lib.fpv_acro_step(C.byref(d._p), _lib.ptr(d._state), n, d._stride, _lib.ptr(a), _lib.ptr(d._lut),
                                     d._lut.numel(), C.c_void_p(done.data_ptr() + G), C.c_void_p(motor.data_ptr() + 16 * G),
                                     None, None, None, _lib.current_stream(torch.device(DEV))))
    torch.cuda.synchronize()
    assert bool((d._state[:, n:] == 123.0).all())
    assert bool((done[:G] == 0xAB).all()) and bool((done[G + n:] == 0xAB).all())
    assert bool((motor[:G] == -7.0).all()) and bool((motor[G + n:] == -7.0).all())
    assert torch.equal(d._state[:, :n], big._state[:, :n])
    assert torch.equal(motor[G:G + n], big.motor_thrust[:n])


def test_tumbling_envs_take_the_accurate_rotation_path():
    """|omega| dt / 2 >= 0.1 rad switches a lane from the series to sincosf; mixed pairs (one tumbling, one calm env in
    the same thread) must both stay within tolerance of the model."""
    from fpyv_b200 import BatchedAcroDrone
    n, dt = 512, 1e-3
    rng, pos, vel, rpy = seeded(n, 14, z_lo=30, z_hi=60)
    c = consts(dt)
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=4, dt=dt, thrust_lut=0)
    d.reset(pos, vel, rpy)
    s = ao.acro_reset(c, pos, vel, rpy)
    omega = rng.normal(0, 1, (n, 3))
    omega[::3] *= 400.0                                   # every third env tumbles at hundreds of rad/s
    d.angular_velocity.copy_(torch.as_tensor(omega, dtype=torch.float32, device=DEV))
    s.omega = d.angular_velocity.double().cpu().numpy()
    act = rng.uniform(-1, 1, (n, 4))
    ao.acro_step(c, s, act, substeps=4)
    d.step(act)
    err = state_err(d, s)
    print(f"acro tumbling: max rel err {err.max():.2e} (tumbling envs {err[::3].max():.2e}, calm {np.delete(err, np.s_[::3]).max():.2e})")
    assert err.max() < 1e-5


def test_stick_rate_curve_vs_model_and_shape():
    """FPV_F_RATE_CURVE: the flight-controller 'actual rates' stick curve (centre sensitivity, max rate, expo) in place
    of the reference's linear stick map.  Checked against the float64 model, and for its defining properties: full stick
    gives the maximum rate, small sticks the centre sensitivity, expo 0 and centre = max is the reference's linear map."""
    from fpyv_b200 import BatchedAcroDrone
    n, dt, K = 512, 1e-3, 4
    rng, pos, vel, rpy = seeded(n, 15, z_lo=20, z_hi=40)
    curve = np.array([[70.0, 670.0, 0.4], [70.0, 670.0, 0.4], [90.0, 400.0, 0.0]])
    c = consts(dt)
    c.rate_curve = curve
    d = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=dt, thrust_lut=0, rate_curve=curve)
    d.reset(pos, vel, rpy)
    s = ao.acro_reset(c, pos, vel, rpy)
    worst = 0.0
    for t in range(12):
        act = rng.uniform(-1.2, 1.2, (n, 4))             # beyond +-1: the curve clips the stick
        ao.acro_step(c, s, act, substeps=K)
        d.step(act)
        worst = max(worst, float(state_err(d, s).max()))
    print(f"acro with stick rate curve: max rel err over 12 free-running control steps {worst:.2e}")
    assert worst < 1e-5
    # shape of the curve through the filtered set-point (hold the stick long enough for the low-pass to settle)
    e = BatchedAcroDrone(None, num_envs=4, device=DEV, substeps=50, dt=dt, thrust_lut=0, rate_curve=curve)
    e.reset(np.tile([0, 0, 50.0], (4, 1)), np.zeros((4, 3)), np.zeros((4, 3)))
    sticks = np.array([[1.0, -1.0, 1.0, 0.0], [0.01, 0.01, 0.01, 0.0], [-0.5, 0.5, 0.0, 0.0], [3.0, 0.0, -2.0, 0.0]])
    e.step(sticks)
    sp = e.rate_setpoint.cpu().numpy()
    np.testing.assert_allclose(sp[0], [-670.0, 670.0, -400.0], rtol=1e-5)
    rate = lambda s_, cen, mx, ex: s_ * cen + (mx - cen) * abs(s_) * (s_ ** 5 * ex + s_ * (1 - ex))
    np.testing.assert_allclose(sp[1], [rate(-0.01, 70, 670, 0.4), rate(-0.01, 70, 670, 0.4), rate(-0.01, 90, 400, 0.0)], rtol=1e-5)
    assert abs(sp[1][0] / -0.01 - 70.0) < 4.0                     # slope at the centre ~ the centre sensitivity
    np.testing.assert_allclose(sp[2], [rate(0.5, 70, 670, 0.4), rate(-0.5, 70, 670, 0.4), 0.0], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(sp[3], [-670.0, 0.0, 400.0], rtol=1e-5, atol=1e-4)
    lin = BatchedAcroDrone(None, num_envs=4, device=DEV, substeps=50, dt=dt, thrust_lut=0, rate_curve=[200.0, 200.0, 0.0])
    ref = BatchedAcroDrone(None, num_envs=4, device=DEV, substeps=50, dt=dt, thrust_lut=0)
    for x in (lin, ref):
        x.reset(np.tile([0, 0, 50.0], (4, 1)), np.zeros((4, 3)), np.zeros((4, 3)))
        x.step(sticks)
    np.testing.assert_allclose(lin.rate_setpoint.cpu().numpy(), ref.rate_setpoint.cpu().numpy(), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("packed", [True, False])
def test_substeps_compose_bit_exactly(packed):
    """Mode B and mode C carry no per-control-step work besides the substeps themselves, so one call with K substeps is
    bit-identical to K calls with one substep and the same sticks."""
    from fpyv_b200 import BatchedAcroDrone, BatchedRacer
    n, K = 3000, 6
    rng, pos, vel, rpy = seeded(n, 16, z_lo=0.3, z_hi=6)
    act = rng.uniform(-1, 1, (n, 4))
    a = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=1e-3, packed=packed)
    b = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=1, dt=1e-3, packed=packed)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    any_done = torch.zeros(n, dtype=torch.bool, device=DEV)
    for _ in range(3):
        da = a.step(act).bool().clone()
        db = torch.zeros(n, dtype=torch.bool, device=DEV)
        for _ in range(K):
            db |= b.step(act).bool()
        assert torch.equal(da, db)
        any_done |= da
    keep = ~any_done                     # (the episode counter counts control steps, so compare the physical planes)
    for p in (0, 2, 3, 4, 5, 6):
        assert torch.equal(a._state[p, :n][keep], b._state[p, :n][keep]), p
    assert torch.equal(a._state[1, :n, :3][keep], b._state[1, :n, :3][keep])
    if packed:
        gains = {"roll": [2, 0.1, 1e-4], "pitch": [1.5, 0.2, 0], "yaw": [0.1, 0, 0]}
        ra = BatchedRacer(5, gains, num_envs=n, device=DEV, dt=1e-3, substeps=K)
        rb = BatchedRacer(5, gains, num_envs=n, device=DEV, dt=1e-3, substeps=1)
        ra.reset()
        rb.reset()
        sp = torch.as_tensor(np.concatenate([rng.uniform(-4, 4, (n, 3)), rng.uniform(0, 10, (n, 1))], 1), dtype=torch.float32, device=DEV)
        for _ in range(3):
            ra.step(sp)
            for _ in range(K):
                rb.step(sp)
        assert torch.equal(ra._state, rb._state)


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("n,K,T", [(100_003, 4, 7), (300, 1, 12)])
def test_acro_rollout_is_bit_identical_to_steps(n, K, T, packed):
    """fpv_acro_rollout (T control steps per launch, state in registers, restarts from the snapshot in registers) against
    T calls of step(): state, every step's flags, motor thrusts and statistics, bit for bit, with crashes inside."""
    from fpyv_b200 import BatchedAcroDrone
    rng, pos, vel, rpy = seeded(n, 18, z_lo=0.12, z_hi=1.5)
    a = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=1e-3, auto_reset=True, packed=packed)
    b = BatchedAcroDrone(None, num_envs=n, device=DEV, substeps=K, dt=1e-3, auto_reset=True, packed=packed)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    a._state[:, n:] = 55.0
    acts = torch.as_tensor(rng.uniform(-1, 1, (T, n, 4)), dtype=torch.float32, device=DEV).contiguous()
    acts[..., 3] = acts[..., 3] * 0.3 - 0.8                     # low throttle: many envs reach the ground within the rollout
    flags = torch.full((T + 2, n), 0xAB, dtype=torch.uint8, device=DEV)
    a.rollout(acts, done_out=flags[1:T + 1])
    ref = torch.empty((T, n), dtype=torch.uint8, device=DEV)
    for t in range(T):
        ref[t] = b.step(acts[t])
    torch.cuda.synchronize()
    assert torch.equal(a._state[:, :n].view(torch.int32), b._state[:, :n].view(torch.int32))
    assert bool((a._state[:, n:] == 55.0).all())
    assert torch.equal(flags[1:T + 1], ref) and bool((flags[0] == 0xAB).all()) and bool((flags[T + 1] == 0xAB).all())
    assert torch.equal(a.done, b.done) and torch.equal(a.motor_thrust, b.motor_thrust)
    assert torch.equal(a._stats, b._stats)
    assert n < 1000 or int(ref.sum()) > 0
