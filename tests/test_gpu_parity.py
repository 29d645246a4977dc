"""GPU parity: the CUDA path (through the C ABI, via fpyv_b200's Python mirror of `Drone`) against
(1) the golden vectors produced by the unmodified reference and (2) the float64 oracle on seeded inputs.

Tolerance (BASELINE.json north_star): <= 1e-5 relative per single step.  "relative" here is, per env and
per quantity group (position, velocity, R, rates, thrust): max|gpu - ref| / max(floor, max|ref|) -- the max-norm
relative error of SURVEY.md 8(d) -- with a scale floor TIED TO THE QUANTITY (1 cm, 1 cm/s, 1e-3 for rotation entries,
0.1 deg/s, 0.01 N: below those magnitudes a relative error stops meaning anything physical), not the blanket floor
of 1.0 the first round used (which turned "relative" into "absolute" for every quantity below 1).
The free-running divergence over the 1 s horizon is printed and bounded separately.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fpv_oracle as fo

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-5
DEV = "cuda:0"


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


FLOOR = {"position": 1e-2, "velocity": 1e-2, "R": 1e-3, "rates": 1e-1, "thrust": 1e-2}   # m, m/s, -, deg/s, N


def group_err(a, b, floor=1.0):
    """a, b: [n, ...] -> per-env max-norm relative error max|a - b| / max(floor, max|b|)."""
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    return np.max(np.abs(a - b), axis=1) / np.maximum(floor, np.max(np.abs(b), axis=1))


def drone_err(d, state, R, prev_rates, prev_thrust, vel_floor=None):
    """Per-env worst group error, every group relative to its own magnitude (floors in FLOOR).  vel_floor: contact tests
    pass the velocity scale of the contact itself (a stiff spring changes v by up to ~1 m/s per step, and the motor-surface
    distance it acts on is a difference of metre-sized float32 numbers)."""
    errs = [group_err(d.position.cpu().numpy(), state[:, :3], FLOOR["position"]),
            group_err(d.velocity.cpu().numpy(), state[:, 3:], FLOOR["velocity"] if vel_floor is None else vel_floor),
            group_err(d.rotation_matrix.cpu().numpy(), R, FLOOR["R"]),
            group_err(d.prev_rates.cpu().numpy(), prev_rates, FLOOR["rates"]),
            group_err(d.prev_thrust.cpu().numpy()[:, None], np.asarray(prev_thrust)[:, None], FLOOR["thrust"])]
    drone_err.last = {k: float(e.max()) for k, e in zip(("position", "velocity", "R", "rates", "thrust"), errs)}
    return np.max(np.stack(errs), axis=0)


def set_state(d, state, R, prev_rates, prev_thrust):
    f = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=DEV)
    d.position.copy_(f(state[:, :3]))
    d.velocity.copy_(f(state[:, 3:]))
    d.set_rotation_matrix(f(R))
    d.prev_rates.copy_(f(prev_rates))
    d.prev_thrust.copy_(f(prev_thrust))


def make(n, **kw):
    from fpyv_b200 import BatchedDrone
    d = BatchedDrone(None, num_envs=n, device=DEV, **kw)
    return d


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("name", ["drone_kat", "drone_random", "drone_ground", "drone_wind", "drone_overdrive",
                                  "drone_1ms_k8"])
def test_single_step_vs_reference(name, packed):
    """Every (t, env) pair of the golden rollout is one env: start from the reference's state at t-1, apply
    action t, compare with the reference's state at t."""
    g = load(name)
    T, n = g["actions"].shape[:2]
    d = make((T - 1) * n, dt=float(g["dt"]), packed=packed)
    d.reset()
    cat = lambda k, sl: g[k][sl].reshape((-1,) + g[k].shape[2:])
    prev, nxt = slice(0, T - 1), slice(1, T)
    set_state(d, cat("state", prev), cat("R", prev), cat("prev_rates", prev), cat("prev_thrust", prev))
    wind = g["wind"] if "wind" in g else None
    ret = d.step(cat("actions", nxt), wind_velocity_vector=wind)
    err = drone_err(d, cat("state", nxt), cat("R", nxt), cat("prev_rates", nxt), cat("prev_thrust", nxt))
    print(f"\n{name} packed={packed}: single-step rel err max {err.max():.3e} median {np.median(err):.3e}")
    assert err.max() <= TOL_STEP
    # done flag: identical, except where a motor sits within 1e-5 m of the plane in the reference itself
    done_ref = cat("done", nxt).astype(bool)
    done_gpu = d.done.cpu().numpy()
    mism = done_ref != done_gpu
    assert mism.sum() == 0, f"{mism.sum()} done flags differ"
    if "ret_Rt" in g:
        assert group_err(ret[0].cpu().numpy(), cat("ret_Rt", nxt)).max() <= TOL_STEP
        assert group_err(ret[2].cpu().numpy(), cat("ret_acc", nxt)).max() <= TOL_STEP
        # gyro matrix = euler(rates in deg/s taken as radians): |angle| up to 200 rad, so the fp32 rounding of the
        # rates themselves (rel 6e-8 * 200 rad) bounds the achievable accuracy at ~2e-5 absolute
        assert group_err(ret[1].cpu().numpy(), cat("ret_gyro", nxt)).max() <= 5e-5


@pytest.mark.parametrize("name,horizon_tol", [("drone_kat", 1e-5), ("drone_random", 1e-5), ("drone_wind", 1e-5),
                                              ("drone_overdrive", 1e-5)])
def test_free_running_divergence(name, horizon_tol):
    g = load(name)
    T, n = g["actions"].shape[:2]
    d = make(n, dt=float(g["dt"]))
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    wind = g["wind"] if "wind" in g else None
    curve = []
    for t in range(T):
        d.step(g["actions"][t], wind_velocity_vector=wind, return_obs=False)
        curve.append(drone_err(d, g["state"][t], g["R"][t], g["prev_rates"][t], g["prev_thrust"][t]).max())
    marks = [t for t in (1, 2, 5, 10, 30, 60, 120) if t <= T]
    print(f"\n{name}: divergence " + " ".join(f"@{t}:{curve[t - 1]:.2e}" for t in marks))
    assert curve[0] <= TOL_STEP
    assert max(curve) <= horizon_tol


@pytest.mark.parametrize("packed", [True, False])
def test_k8_substeps_1s_horizon(packed):
    """config 2/3 shape: 8 substeps of 1 ms per control step, 1 s horizon (1000 reference steps)."""
    g = load("drone_1ms_k8")
    T, n = g["actions"].shape[:2]
    K = int(g["hold"])
    d = make(n, dt=float(g["dt"]), substeps=K, packed=packed)
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    curve = []
    for j in range(T // K):
        d.step(g["actions"][j * K], return_obs=False)
        t = j * K + K - 1
        curve.append(drone_err(d, g["state"][t], g["R"][t], g["prev_rates"][t], g["prev_thrust"][t]).max())
        assert np.array_equal(d.done.cpu().numpy(), g["done"][j * K:t + 1].any(axis=0))
    print(f"\nK=8 packed={packed}: divergence over 1 s: first {curve[0]:.2e} @0.5s {curve[len(curve) // 2]:.2e} "
          f"@1s {curve[-1]:.2e} max {max(curve):.2e}")
    assert curve[0] <= TOL_STEP
    assert max(curve) <= 1e-4


def test_ground_contact_free_running():
    """Spring + crash branch (components.py:198-214, :239): trajectories match until the reference crashes."""
    g = load("drone_ground")
    T, n = g["actions"].shape[:2]
    d = make(n, dt=float(g["dt"]))
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    crashed = np.zeros(n, dtype=bool)
    worst = 0.0
    for t in range(T):
        d.step(g["actions"][t], return_obs=False)
        dg = d.done.cpu().numpy()
        ok = ~crashed
        assert np.array_equal(dg[ok], g["done"][t].astype(bool)[ok]), t
        e = drone_err(d, g["state"][t], g["R"][t], g["prev_rates"][t], g["prev_thrust"][t])
        worst = max(worst, e[ok].max() if ok.any() else 0.0)
        crashed |= g["done"][t].astype(bool)
    assert crashed.any() and not crashed.all()
    assert worst <= 1e-5, worst


def test_override_inputs():
    g = load("drone_override")
    T, n = g["actions"].shape[:2]
    d = make(n, dt=float(g["dt"]))
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    for t in range(T):
        d.step(g["actions"][t], rotation_matrix=g["R_override"][t], thrust_force=g["thrust_override"][t], return_obs=False)
        e = drone_err(d, g["state"][t], g["R"][t], g["prev_rates"][t], g["prev_thrust"][t])
        assert e.max() <= TOL_STEP, (t, e.max())
        assert np.array_equal(d.done.cpu().numpy(), g["done"][t].astype(bool))
    from fpyv_b200 import FpvError
    d2 = make(n, dt=float(g["dt"]), substeps=2)
    d2.reset()
    with pytest.raises(FpvError):
        d2.step(g["actions"][0], rotation_matrix=g["R_override"][0], thrust_force=g["thrust_override"][0])


@pytest.mark.parametrize("packed", [True, False])
def test_obstacles(packed):
    from fpyv_b200 import Cylinder, Ground, Target
    g = load("drone_objects")
    T, n = g["actions"].shape[:2]
    objs = [Target(g["sph"][:3], g["sph"][3]), Cylinder(g["cyl"][:3], g["cyl"][3], g["cyl"][4]), Ground()]
    # single-step form (chaotic after contact): all (t, env) pairs at once
    d = make((T - 1) * n, dt=float(g["dt"]), packed=packed)
    d.reset()
    cat = lambda k, sl: g[k][sl].reshape((-1,) + g[k].shape[2:])
    prev, nxt = slice(0, T - 1), slice(1, T)
    set_state(d, cat("state", prev), cat("R", prev), cat("prev_rates", prev), cat("prev_thrust", prev))
    d.step(cat("actions", nxt), object_list=objs, return_obs=False)
    err = drone_err(d, cat("state", nxt), cat("R", nxt), cat("prev_rates", nxt), cat("prev_thrust", nxt))
    assert err.max() <= TOL_STEP, err.max()
    assert np.array_equal(d.done.cpu().numpy(), cat("done", nxt).astype(bool))
    assert cat("done", nxt).any()


@pytest.mark.parametrize("calib", ["frsky", "calibration"])
def test_sticks(calib):
    from fpyv_b200 import Joystick
    from conftest import CONFIG
    g = load("sticks_" + calib)
    rc = Joystick(device=DEV)
    rc.calibrate(os.path.join(CONFIG, calib + ".json"))
    rc.feed(g["raw"].astype(np.int64))
    cal = rc.calib_read().cpu().numpy()
    act = rc.read_actions().cpu().numpy()
    assert np.max(np.abs(cal - g["calibrated"])) <= 2e-6
    assert np.max(np.abs(act - g["action"])) <= 2e-6


def test_step_from_sticks():
    """Drone.step(action=None): joystick -> action -> dynamics (components.py:227-228)."""
    g = load("drone_sticks")
    T = g["raw"].shape[0]
    d = make(1, dt=float(g["dt"]))
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    for t in range(T):
        d.rc.feed(g["raw"][t].astype(np.int64))
        d.step(None, return_obs=False)
        e = drone_err(d, g["state"][t], g["R"][t], g["prev_rates"][t], g["prev_thrust"][t])
        assert e.max() <= TOL_STEP, (t, e.max())
    d.rc._raw = None
    with pytest.raises(ModuleNotFoundError):
        d.step(None)


@pytest.mark.parametrize("name", ["racer_demo", "racer_random"])
def test_racer(name):
    from fpyv_b200 import BatchedRacer
    g = load(name)
    T, n = g["actions"].shape[:2]
    worst_step, curve = 0.0, []
    # free-running per env (gains differ per env -> one BatchedRacer per env, batched over nothing)
    for e in range(n):
        gains = dict(zip(("roll", "pitch", "yaw"), g["gains"][e]))
        r = BatchedRacer(5, gains, num_envs=1, device=DEV)
        r.reset()
        for t in range(T):
            r.step(g["actions"][t, e][None])
            if t in (0, 1, 9, 99, T - 1):
                err = max(group_err(r.position.cpu().numpy(), g["position"][t, e][None]).max(),
                          group_err(r.linear_velocity.cpu().numpy(), g["velocity"][t, e][None]).max(),
                          group_err(r.orientation.cpu().numpy(), g["R"][t, e][None]).max(),
                          group_err(r.angular_velocity.cpu().numpy(), g["omega"][t, e][None]).max())
                if t == 0:
                    worst_step = max(worst_step, err)
                curve.append((t + 1, err))
    by_t = {}
    for t, e in curve:
        by_t[t] = max(by_t.get(t, 0), e)
    print(f"\n{name}: divergence " + " ".join(f"@{t}:{e:.2e}" for t, e in sorted(by_t.items())))
    assert worst_step <= TOL_STEP
    assert max(by_t.values()) <= 2e-4    # |omega| ~ 80 rad feeds sin/cos: fp32 argument rounding dominates


def test_reset_and_views():
    d = make(5)
    rng = np.random.default_rng(0)
    pos, vel, rpy = rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), rng.uniform(-180, 180, (5, 3))
    d.reset(pos, vel, rpy)
    a = np.deg2rad(rpy)
    R = fo.euler_matrix(a[:, 0], a[:, 1], a[:, 2])
    assert np.max(np.abs(d.rotation_matrix.cpu().numpy() - R)) <= 5e-7   # float32 quaternion -> matrix rounding
    assert np.allclose(d.position.cpu().numpy(), pos.astype(np.float32))
    assert np.allclose(d.state.cpu().numpy()[:, 3:], vel.astype(np.float32))
    assert d.prev_rates.abs().max().item() == 0 and d.prev_thrust.abs().max().item() == 0
    assert not d.done.any()
    # masked reset leaves the other envs alone
    d.step(rng.uniform(-1, 1, (5, 4)), return_obs=False)
    before = d.position.clone()
    d.reset(pos, vel, rpy, mask=np.array([1, 0, 0, 1, 0], dtype=bool))
    after = d.position
    assert torch.equal(after[[1, 2, 4]], before[[1, 2, 4]])
    assert np.allclose(after[[0, 3]].cpu().numpy(), pos[[0, 3]].astype(np.float32))
    # stock defaults of params.yaml
    d.reset()
    assert np.allclose(d.position.cpu().numpy(), [[0, 0, 10]] * 5) and np.allclose(d.velocity.cpu().numpy(), [[1, 0, 0]] * 5)


def test_rotation_matrix_roundtrip_all_branches():
    """set_rotation_matrix / rotation_matrix through the quaternion plane, including 180-degree turns where the
    reference's own rotation_matrix_to_quaternion (helper_functions.py:65-80, trace branch only) breaks down."""
    rng = np.random.default_rng(3)
    ang = rng.uniform(-np.pi, np.pi, (200, 3))
    R = fo.euler_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    special = [np.diag([1.0, -1, -1]), np.diag([-1.0, 1, -1]), np.diag([-1.0, -1, 1]), np.eye(3),
               fo.euler_matrix(np.array([np.pi]), np.array([0.3]), np.array([-2.0]))[0]]
    R = np.concatenate([R, np.stack(special)])
    d = make(len(R))
    d.reset()
    d.set_rotation_matrix(R)
    back = d.rotation_matrix.cpu().numpy()
    assert np.max(np.abs(back - R)) <= 5e-7
    q = d.quaternion.cpu().numpy().astype(np.float64)
    assert np.max(np.abs(np.linalg.norm(q, axis=1) - 1)) <= 2e-7 and (q[:, 0] >= 0).all()
    assert np.max(np.abs(fo.quaternion_to_matrix(q) - R)) <= 5e-7
    # masked write
    d.set_rotation_matrix(np.tile(np.eye(3), (len(R), 1, 1)), mask=np.arange(len(R)) % 2 == 0)
    after = d.rotation_matrix.cpu().numpy()
    assert np.allclose(after[::2], np.eye(3), atol=1e-7) and np.max(np.abs(after[1::2] - R[1::2])) <= 5e-7


@pytest.mark.parametrize("n", [1, 2, 3, 127, 128, 129, 257, 1000])
def test_ragged_sizes_match_between_variants(n):
    """Packed (2 envs/thread) and scalar kernels agree to fp32 rounding for every tail shape, and nothing outside
    [0, n) is written."""
    rng = np.random.default_rng(n)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.05, 3, n)], 1)
    vel, rpy = rng.normal(size=(n, 3)), rng.uniform(-30, 30, (n, 3))
    acts = rng.uniform(-1, 1, (6, n, 4))
    out = []
    for packed in (True, False):
        d = make(n, packed=packed, substeps=4, dt=1e-3)
        d.reset(pos, vel, rpy)
        d._state[:, n:] = 123.0
        for a in acts:
            d.step(a, return_obs=False)
        assert (d._state[:, n:] == 123.0).all()
        out.append((d._state[:, :n].clone(), d.done.clone()))
    assert torch.allclose(out[0][0][[0, 2, 3]], out[1][0][[0, 2, 3]], rtol=1e-5, atol=1e-5)
    assert torch.allclose(out[0][0][1, :, :3], out[1][0][1, :, :3], rtol=1e-5, atol=1e-5)
    assert torch.equal(out[0][0][1, :, 3].view(torch.int32), out[1][0][1, :, 3].view(torch.int32))
    assert torch.equal(out[0][1], out[1][1])
    c = fo_consts(1e-3)
    s = fo.drone_reset(c, pos, vel, rpy)
    for a in acts:
        fo.drone_step(c, s, a, substeps=4)
    assert group_err(out[0][0][0, :, :3].cpu().numpy(), s.pos).max() <= 1e-5


def fo_consts(dt=None):
    import yaml
    from conftest import CONFIG
    with open(os.path.join(CONFIG, "params.yaml")) as f:
        params = yaml.safe_load(f)
    return fo.derive_consts(params, os.path.join(CONFIG, "t_motos_f80_motor_test.csv"), dt=dt)


def test_thrust_lut_matches_cubic():
    """north_star's shared-memory LUT with interpolation vs the reference cubic (n=2049: <= 6e-6 N)."""
    n = 4096
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (n, 4))
    res = []
    for lut in (0, 2049):
        d = make(n, thrust_lut=lut)
        d.reset()
        d.step(acts, return_obs=False)
        res.append(d.prev_thrust.cpu().numpy().astype(np.float64))
    c = fo_consts()
    ref = fo.throttle2thrust(c, acts[:, 3]) * 0.5
    assert np.max(np.abs(res[0] - ref)) <= 1e-5
    assert np.max(np.abs(res[1] - ref)) <= 2e-5
    print(f"\nLUT(2049) vs cubic: max abs {np.max(np.abs(res[1] - res[0])):.2e} N")


def test_per_env_wind():
    n = 512
    rng = np.random.default_rng(6)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(2, 12, n)], 1)
    vel, rpy = rng.normal(0, 3, (n, 3)), rng.uniform(-30, 30, (n, 3))
    wind = rng.normal(0, 2, (n, 3))
    acts = rng.uniform(-1, 1, (10, n, 4))
    c = fo_consts()
    s = fo.drone_reset(c, pos, vel, rpy)
    # oracle: per-env wind = loop-free because drone_substep broadcasts [n,3]
    for a in acts:
        fo.drone_substep(c, s, a, wind)
    for packed in (False, True):
        d = make(n, packed=packed)
        d.reset(pos, vel, rpy)
        w = torch.as_tensor(wind, dtype=torch.float32, device=DEV)
        for a in acts:
            d.step(a, wind_velocity_vector=w, return_obs=False)
        e = drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust)
        assert e.max() <= 1e-5, (packed, e.max())


def test_auto_reset_freeze_and_stats():
    n = 64
    pos = np.tile([0.0, 0.0, 0.3], (n, 1))
    pos[::2, 2] = 50.0                       # even envs fly high, odd envs start just above the ground
    vel = np.tile([0.0, 0.0, -3.0], (n, 1))
    acts = np.tile([0.0, 0.0, 0.0, -1.0], (n, 1))
    # reference behaviour: done reported per step, integration continues
    d = make(n)
    d.reset(pos, vel, 0.0)
    seen = torch.zeros(n, dtype=torch.bool, device=DEV)
    for _ in range(30):
        d.step(acts, return_obs=False)
        seen |= d.done
    assert seen[1::2].all() and not seen[::2].any()
    # freeze: crashed envs stop, done is sticky
    d = make(n, freeze_done=True)
    d.reset(pos, vel, 0.0)
    for _ in range(30):
        d.step(acts, return_obs=False)
    assert d.done[1::2].all() and not d.done[::2].any()
    z = d.position[1::2, 2].clone()
    d.step(acts, return_obs=False)
    assert torch.equal(d.position[1::2, 2], z)
    assert (d.episode_steps[1::2] < 0).all()
    st = d.episode_stats()
    assert st["episodes"] == n // 2 and st["crashes"] == n // 2 and st["nonfinite"] == 0
    assert st["env_steps"] == pytest.approx(31 * (n // 2) + st["episode_len_sum"])
    # auto-reset: crashed envs restart from the reset snapshot and keep counting episodes
    d = make(n, auto_reset=True)
    d.reset(pos, vel, 0.0)
    for _ in range(40):
        d.step(acts, return_obs=False)
    st = d.episode_stats()
    assert st["episodes"] >= n // 2 and st["env_steps"] == 40 * n
    assert (d.position[1::2, 2] > 0).all()
    assert st["mean_episode_len"] > 1


def test_fast_path_and_cuda_graph_capture():
    """The allocation-free fast path of step() (device float32 actions) is what an RL loop calls; it must be
    capturable in a CUDA graph and produce exactly the same states as the general path."""
    n, K = 4096, 8
    rng = np.random.default_rng(11)
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.5, 3, n)], 1)
    acts = torch.as_tensor(rng.uniform(-1, 1, (6, n, 4)), dtype=torch.float32, device=DEV)
    ref = make(n, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    ref.reset(pos, 0.0, 0.0)
    for a in acts:
        ref.step(a.cpu().numpy(), return_obs=False)          # general path (host array -> conversion)
    d = make(n, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    d.reset(pos, 0.0, 0.0)
    d.step(acts[0], return_obs=False)
    assert d._fast_ok
    static = torch.zeros(n, 4, device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        static.copy_(acts[1])
        d.step(static, return_obs=False)                      # warm-up on the side stream (fast path)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):                             # capture records the launch, it does not run it
        d.step(static, return_obs=False)
    for a in acts[2:]:
        static.copy_(a)
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(d._state, ref._state)
    assert torch.equal(d.done, ref.done)
    assert d.throttle is static[:, 3] or torch.equal(d.throttle, static[:, 3])


def test_drone_compat_object_matches_reference_kat():
    """num_envs=1 NumPy stand-in driven exactly like simulator.py drives the reference."""
    from fpyv_b200 import Drone, Ground
    from fpyv_b200.config import load_params
    params = load_params()
    drone = Drone(params)
    assert params["drone"]["force_multiplier_pid"]["max_output"] == pytest.approx(81.30229036293663)
    drone.reset(position=np.array(params["drone"]["initial_position"]), velocity=np.array(params["drone"]["initial_velocity"]),
                ypr=np.array(params["drone"]["initial_orientation"]))
    g = load("drone_kat")
    for t in range(60):
        Rt, gyro, acc = drone.step(action=np.array([0.3, -0.2, 0.1, 0.25]), wind_velocity_vector=np.zeros(3), object_list=[Ground()])
        assert not drone.done
    assert np.max(np.abs(drone.state - g["state"][59, 0])) / np.max(np.abs(g["state"][59, 0])) <= 1e-5
    assert np.max(np.abs(Rt - g["ret_Rt"][59, 0])) <= 1e-5
    assert drone.prev_thrust == pytest.approx(g["prev_thrust"][59, 0], rel=1e-5)


def test_error_paths():
    from fpyv_b200 import FpvError, _lib
    import ctypes as C
    lib = _lib.load()
    d = make(8)
    with pytest.raises(RuntimeError):
        d.step(np.zeros((8, 4)))            # step before reset
    d.reset()
    p, io = d._p, d._io
    d.step(np.zeros((8, 4)), return_obs=False)
    p.substeps = 0
    assert lib.fpv_drone_step(C.byref(p), C.byref(io), None) == -22
    assert b"substeps" in lib.fpv_last_error()
    p.substeps = 1
    io.actions = io.actions + 4             # misaligned
    assert lib.fpv_drone_step(C.byref(p), C.byref(io), None) == -22
    io.actions = None
    assert lib.fpv_drone_step(C.byref(p), C.byref(io), None) == -22
    with pytest.raises(FpvError):
        _lib.check(-22)
    with pytest.raises(FileNotFoundError):
        from fpyv_b200 import Joystick
        Joystick(device=DEV).calibrate("/nonexistent/calib.json")
    with pytest.raises(ValueError):
        from fpyv_b200 import Ground, Target
        d.step(np.zeros((8, 4)), object_list=[Ground(), Target([0, 0, 5], 1.0)])


def test_full_size_properties():
    """BASELINE config 3 size (1,048,576 envs, K=8, 1 ms): size-independent properties + a 4,096-env subsample
    against the oracle over the 1 s horizon (125 control steps)."""
    n, K, dt, steps = 1 << 20, 8, 1e-3, 125
    gen = torch.Generator(device=DEV).manual_seed(1234)
    pos = torch.randn(n, 3, device=DEV, generator=gen) * torch.tensor([5.0, 5.0, 2.0], device=DEV) + torch.tensor([0, 0, 10.0], device=DEV)
    pos[:, 2].clamp_(min=1.0)
    vel = torch.randn(n, 3, device=DEV, generator=gen)
    rpy = (torch.rand(n, 3, device=DEV, generator=gen) * 2 - 1) * 30
    # plant duplicates: env i+half == env i for the first 1000 envs
    half = n // 2
    for t in (pos, vel, rpy):
        t[half:half + 1000] = t[:1000]
    d = make(n, substeps=K, dt=dt)
    d.reset(pos, vel, rpy)
    sub = torch.arange(0, n, n // 4096, device=DEV)[:4096]
    c = fo_consts(dt)
    s = fo.drone_reset(c, pos[sub].cpu().numpy().astype(np.float64), vel[sub].cpu().numpy().astype(np.float64),
                       rpy[sub].cpu().numpy().astype(np.float64))
    curve = {}
    for j in range(steps):
        a = torch.rand(n, 4, device=DEV, generator=gen) * 2 - 1
        a[half:half + 1000] = a[:1000]
        d.step(a, return_obs=False)
        fo.drone_step(c, s, a[sub].cpu().numpy().astype(np.float64), substeps=K)
        if j + 1 in (1, 2, 5, 10, 30, 60, 125):
            e = np.maximum.reduce([group_err(d.position[sub].cpu().numpy(), s.pos), group_err(d.velocity[sub].cpu().numpy(), s.vel),
                                   group_err(d.rotation_matrix[sub].cpu().numpy(), s.R)])
            curve[j + 1] = e.max()
    print("\n1M envs, K=8: divergence vs float64 oracle (4096-env subsample) " + " ".join(f"@{k}:{v:.2e}" for k, v in curve.items()))
    assert curve[1] <= TOL_STEP and max(curve.values()) <= 1e-4
    # determinism across thread slots / positions in the batch
    assert torch.equal(d._state[:, half:half + 1000], d._state[:, :1000])
    # R stays a rotation: R R^T = I to fp32 drift
    R = d.rotation_matrix
    eye = torch.eye(3, device=DEV)
    assert (R @ R.transpose(1, 2) - eye).abs().max().item() <= 1e-4
    assert torch.isfinite(d._state[:, :n, :3]).all()
    st = d.episode_stats()
    assert st["env_steps"] == steps * n and st["nonfinite"] == 0


@pytest.mark.parametrize("n,K", [(1 << 20, 8), (1 << 20, 1), (200_003, 4), (4096, 8)])
def test_chained_launches_are_bit_identical(n, K):
    """FPV_F_CHAINED (launch i+1 overlaps the tail of launch i, state ordered chunk by chunk through chunk_epoch):
    an open-loop rollout with chained launches must reproduce the plain stream-ordered rollout BIT FOR BIT, on the
    same drone stepped repeatedly (true chunk dependencies) with auto-reset and crash traffic in play.  The small
    case falls back to plain stream order inside the library (grid below one full wave) and must agree as well."""
    g = torch.Generator(device=DEV).manual_seed(5)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=DEV, generator=g) * 2.95
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    steps = 40
    acts = [torch.rand(n, 4, device=DEV, generator=g) * 2 - 1 for _ in range(steps)]
    out = []
    for chained in (False, True):
        d = make(n, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
        d.reset(pos, vel, rpy)
        dones = torch.zeros(n, dtype=torch.int32, device=DEV)
        for i in range(steps):
            d.step(acts[i], return_obs=False, chained=chained)
            if i % 7 == 3:          # a reader in plain stream order between chained launches
                dones += d.done.to(torch.int32)
        torch.cuda.synchronize()
        out.append((d._state.clone(), d.done.clone(), dones, d.episode_stats(), d._chunk_epoch.clone()))
    (s0, d0, c0, st0, e0), (s1, d1, c1, st1, e1) = out
    assert torch.equal(s0.view(torch.int32), s1.view(torch.int32))
    assert torch.equal(d0, d1) and torch.equal(c0, c1)
    assert st0 == st1
    assert int(e0.max()) == 0                                     # plain launches publish nothing
    assert int(e1.min()) == steps and int(e1.max()) == steps       # every chunk saw every chained launch
    assert st0["crashes"] > 0


def test_rollout_and_stick_replay_match_stepwise():
    """BatchedDrone.rollout (chained launches) and Joystick.replay (whole log in one launch) reproduce the step-by-step
    joystick path -- checked against the golden joystick-driven reference run (drone_sticks.npz) and bit-for-bit
    against step(None) on the same raw log."""
    from fpyv_b200 import Joystick
    g = load("drone_sticks")
    raw = g["raw"][:, 0]                                    # [T,6]
    T = len(raw)
    calib = os.path.join(os.path.dirname(GOLDEN), "..", "fpyv_b200", "config", "frsky.json")
    rc = Joystick(device=DEV)
    rc.calibrate(calib, load_calibration_file=True)
    acts = rc.replay(raw)                                   # [T,1,4]
    np.testing.assert_allclose(acts[:, 0].double().cpu().numpy(), g["actions"][:, 0], rtol=0, atol=2e-6)
    d = make(1)
    d.reset(g["pos0"], g["vel0"], g["rpy0"])
    dones = torch.zeros((T, 1), dtype=torch.uint8, device=DEV)
    d.rollout(acts.contiguous(), done_out=dones)
    err = drone_err(d, g["state"][-1], g["R"][-1], g["prev_rates"][-1], g["prev_thrust"][-1])
    assert err.max() < 2e-5, err.max()                      # 30 free-running steps
    assert np.array_equal(dones.cpu().numpy().astype(bool), g["done"].astype(bool))
    # bit-for-bit against the stepwise joystick path
    d2 = make(1)
    d2.reset(g["pos0"], g["vel0"], g["rpy0"])
    for t in range(T):
        d2.rc.feed(raw[t][None])
        d2.step(None, return_obs=False)
    assert torch.equal(d._state, d2._state)
    with pytest.raises(ValueError):
        d.rollout(acts[:, :, :3])


def test_step_host_pipelined_slices_match_plain_step():
    """step_host cuts the batch into env ranges pipelined over copy / compute / copy-back streams; the result must be
    the plain step's, bit for bit, for several steps in a row (buffer reuse across steps is ordered by events)."""
    n = 300_000
    g = torch.Generator(device=DEV).manual_seed(8)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=DEV, generator=g) * 2.95
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    a = make(n, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    b = make(n, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    assert len(a._slice_bounds(4)) == 4 and a._slice_bounds(4)[-1][1] == n
    host = [torch.empty(n, 4, dtype=torch.float32, pin_memory=True).uniform_(-1, 1) for _ in range(3)]
    done_h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for i in range(7):
        a.step_host(host[i % 3], done_h)
        torch.cuda.current_stream().synchronize()
        b.step(host[i % 3].to(DEV), return_obs=False)
        assert torch.equal(done_h, b.done.cpu()), i
    assert torch.equal(a._state, b._state)
    assert a.episode_stats() == b.episode_stats()


def test_baseline_config0_single_drone_10k_steps():
    """BASELINE.json configs[0]: ONE drone, dt = 1 ms, 10,000 steps (the CPU-runnable case of the reference), both
    dynamics modes, free-running against the float64 oracle.  The single-step bound is the north_star's 1e-5; the
    divergence curve over the 10 s horizon is reported (acro dynamics are chaotic) and bounded loosely."""
    from fpyv_b200 import BatchedDrone, BatchedRacer
    rng = np.random.default_rng(31)
    # --- mode A: Drone.step with sticks changing every 50 ms
    c = fo_consts(1e-3)
    d = BatchedDrone(None, num_envs=1, device=DEV, dt=1e-3)
    pos, vel, rpy = np.array([[0.0, 0, 10]]), np.array([[1.0, 0, 0]]), np.zeros((1, 3))       # params.yaml initial state
    d.reset(pos, vel, rpy)
    s = fo.drone_reset(c, pos, vel, rpy)
    curve = {}
    act = None
    for t in range(10_000):
        if t % 50 == 0:
            act = rng.uniform(-1, 1, (1, 4)) * np.array([0.3, 0.3, 0.3, 1.0]) + np.array([0, 0, 0, -0.3])
        fo.drone_step(c, s, act)
        d.step(act, return_obs=False)
        if t + 1 in (1, 10, 100, 1000, 10_000):
            curve[t + 1] = float(drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust).max())
    print("\nconfigs[0] Drone, 1 env, dt 1 ms: divergence " + " ".join(f"@{k}:{v:.1e}" for k, v in curve.items()))
    assert curve[1] <= TOL_STEP and curve[1000] <= 1e-4 and curve[10_000] <= 5e-2
    # --- mode B: the Racer of tests/racer_drone_test.py, its own demo gains and set-point switch, 10,000 steps
    gains = {"roll": [2, 0, 0], "pitch": [2, 0, 0], "yaw": [0.1, 0, 0]}
    rc = fo.RacerConsts(gains=np.array(list(gains.values()), dtype=np.float64))
    rs = fo.RacerState(1)
    r = BatchedRacer(5, gains, num_envs=1, device=DEV)
    r.reset()
    curve = {}
    for t in range(10_000):
        a = np.array([[80.0, 10, 0, 0]]) if t <= 20 else np.array([[-30.0, -50, 0, 0]])
        fo.racer_step(rc, rs, a)
        r.step(a)
        if t + 1 in (1, 10, 100, 1000, 10_000):
            curve[t + 1] = float(max(group_err(r.angular_velocity.cpu().numpy(), rs.omega).max(),
                                     group_err(r.orientation.cpu().numpy(), rs.R).max(),
                                     group_err(r.linear_velocity.cpu().numpy(), rs.vel).max()))
    print("configs[0] Racer, 1 env, dt 1 ms: divergence " + " ".join(f"@{k}:{v:.1e}" for k, v in curve.items()))
    assert curve[1] <= TOL_STEP and curve[10_000] <= 5e-2


@pytest.mark.parametrize("n,K,T,auto_reset", [(1 << 18, 8, 6, True), (100_003, 1, 17, True), (5000, 4, 9, False), (63, 2, 5, True)])
def test_fused_rollout_is_bit_identical_to_steps(n, K, T, auto_reset):
    """fpv_drone_rollout (T control steps per launch, state in registers across the steps) against T plain steps:
    state, every step's done flags, the last acceleration and the episode statistics, bit for bit -- with crashes and
    auto-reset restarts happening inside the rollout."""
    g = torch.Generator(device=DEV).manual_seed(17)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=DEV, generator=g) * 1.5
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    acts = (torch.rand(T, n, 4, device=DEV, generator=g) * 2 - 1).contiguous()
    a = make(n, substeps=K, dt=1e-3, auto_reset=auto_reset, thrust_lut=2049)
    b = make(n, substeps=K, dt=1e-3, auto_reset=auto_reset, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    a._state[:, n:] = 77.0                       # padding up to the plane stride must stay untouched
    done_a = torch.full((T + 2, n), 0xAB, dtype=torch.uint8, device=DEV)      # guard rows around the flags
    a.rollout(acts, done_out=done_a[1:T + 1], fused=True)
    done_b = torch.empty((T, n), dtype=torch.uint8, device=DEV)
    for t in range(T):
        b.step(acts[t], return_obs=False)
        done_b[t] = b.done
    torch.cuda.synchronize()
    assert torch.equal(a._state[:, :n].view(torch.int32), b._state[:, :n].view(torch.int32))
    assert bool((a._state[:, n:] == 77.0).all())
    assert torch.equal(done_a[1:T + 1], done_b) and bool((done_a[0] == 0xAB).all()) and bool((done_a[T + 1] == 0xAB).all())
    assert torch.equal(a.done, b.done) and torch.equal(a._acc, b._acc)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert all(sa[k] == sb[k] or (sa[k] != sa[k] and sb[k] != sb[k]) for k in sa), (sa, sb)
    assert n < 5000 or int(done_b.sum()) > 0
    # a second rollout continues from the first one's state; chained per-step rollout agrees too
    a.rollout(acts, fused=True)
    b.rollout(acts, fused=False)
    assert torch.equal(a._state[:, :n].view(torch.int32), b._state[:, :n].view(torch.int32))


def test_max_size_16m_envs_single_gpu():
    """BASELINE.json configs[3]'s total batch (16,777,216 drones) on ONE GPU: 1 GiB of state, plane offsets beyond
    2^30 bytes.  Subsamples at both ends and in the middle against the float64 oracle, guard envs beyond n untouched,
    every env stepped exactly once (episode counters), K = 8 and the fused rollout agree bit for bit."""
    n, K = 1 << 24, 8
    g = torch.Generator(device=DEV).manual_seed(41)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.3 + torch.rand(n, device=DEV, generator=g) * 5
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    act = (torch.rand(2, n, 4, device=DEV, generator=g) * 2 - 1).contiguous()
    d = make(n, substeps=K, dt=1e-3, thrust_lut=0)
    d.reset(pos, vel, rpy)
    idx = torch.cat([torch.arange(0, 1024), torch.arange(n // 2 - 512, n // 2 + 512), torch.arange(n - 1024, n)]).to(DEV)
    c = fo_consts(1e-3)
    s = fo.drone_reset(c, pos[idx].double().cpu().numpy(), vel[idx].double().cpu().numpy(), rpy[idx].double().cpu().numpy())
    d.step(act[0], return_obs=False)
    d.step(act[1], return_obs=False, chained=True)
    for t in range(2):
        fo.drone_step(c, s, act[t][idx].double().cpu().numpy(), substeps=K)
    torch.cuda.synchronize()
    err = max(group_err(d.position[idx].cpu().numpy(), s.pos).max(), group_err(d.velocity[idx].cpu().numpy(), s.vel).max(),
              group_err(d.rotation_matrix[idx].cpu().numpy(), s.R).max())
    print(f"\n16,777,216 envs, 2 control steps x {K} substeps: max rel err vs oracle on 3,072 sampled envs {err:.2e}")
    assert err < 1e-5
    assert bool((d.episode_steps == 2).all())               # every env advanced exactly twice
    assert bool(torch.isfinite(d._state[:3, :n]).all())
    d2 = make(n, substeps=K, dt=1e-3, thrust_lut=0)
    d2.reset(pos, vel, rpy)
    d2.rollout(act, fused=True)
    assert torch.equal(d2._state, d._state)


@pytest.mark.parametrize("packed", [True, False])
def test_obstacle_reach_test_is_exact_against_oracle(packed):
    """The general kernel skips an obstacle whose surface is out of every motor's reach.  20,000 drones placed on shells
    around a sphere and a cylinder at distances straddling that reach (inside, touching, just in reach, just out of it,
    far) must match the float64 oracle -- forces and crash flags -- to the single-step tolerance."""
    from fpyv_b200 import Cylinder, Ground, Target
    n = 20_000
    rng = np.random.default_rng(77)
    sph_c, sph_r = np.array([2.0, -1.0, 6.0]), 1.3
    cyl_p, cyl_r, cyl_h = np.array([-4.0, 3.0, 0.0]), 0.8, 7.0
    gap = rng.choice([-0.3, -0.05, 0.02, 0.08, 0.15, 0.2, 0.224, 0.23, 0.24, 0.3, 1.0, 10.0], n) + rng.normal(0, 0.01, n)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    pos = sph_c + u * (sph_r + gap)[:, None]
    half = n // 2                                         # second half: around the cylinder (side, top cap, below the top)
    ang = rng.uniform(0, 2 * np.pi, half)
    side = rng.random(half) < 0.6
    rad = np.where(side, cyl_r + gap[half:], rng.uniform(0, cyl_r + 0.3, half))
    z = np.where(side, rng.uniform(-0.2, cyl_h + 0.4, half), cyl_h + gap[half:])
    pos[half:] = np.stack([cyl_p[0] + rad * np.cos(ang), cyl_p[1] + rad * np.sin(ang), np.maximum(z, 0.35)], axis=1)
    vel = rng.normal(0, 2, (n, 3))
    rpy = rng.uniform(-40, 40, (n, 3))
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32).astype(np.float64)   # identical inputs on both sides
    c = fo_consts()
    s = fo.drone_reset(c, pos, vel, rpy)
    d = make(n, packed=packed)
    d.reset(pos, vel, rpy)
    s.pos, s.vel = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
    s.R = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
    objs = [Target(sph_c, sph_r), Cylinder(cyl_p, cyl_r, cyl_h), Ground()]
    fo.drone_step(c, s, act, extra_objects=[fo.SphereObj(sph_c, sph_r), fo.CylinderObj(cyl_p, cyl_r, cyl_h)])
    d.step(act, np.zeros(3), objs, return_obs=False)
    err = drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust, vel_floor=1.0)
    done = d.done.cpu().numpy()
    # a motor within float32 rounding of a surface may flip the crash flag: excuse envs whose motors sit at |d| < 2e-6
    mism = done != s.done
    print(f"\nobstacle shells: max rel err {err[~mism].max():.2e}, crashed {int(s.done.sum())}, in contact without crash "
          f"{int((np.abs(s.acc).max(1) > 30).sum())}, flag mismatches {int(mism.sum())}")
    assert err[~mism].max() <= TOL_STEP and mism.sum() <= 3, drone_err.last
    assert s.done.sum() > 500 and (~s.done).sum() > 5000


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("max_rates,dt,K,ang", [(200.0, 1 / 40, 1, 2), (600.0, 1 / 60, 2, 1), (900.0, 1e-3, 8, 4),
                                                (2000.0, 1 / 30, 1, 0), (1200.0, 1 / 60, 3, 1)])
def test_every_angle_kernel_variant_vs_oracle(max_rates, dt, K, ang, packed):
    """The host picks the sin/cos evaluation of the per-substep half Euler angles from the bound max_rates*dt/2:
    series (<= 0.008 rad), degree-3/2 (<= 0.03), degree-5/4 (<= 0.05), degree-7/8 minimax (<= 0.25), sincosf beyond.
    Stock params.yaml only reaches two of them; racing rates (600-1200 deg/s) and coarse steps reach the others.
    Each variant: hot kernel, obstacle kernel and fused rollout against the float64 oracle, single steps and 20 free-running."""
    from fpyv_b200 import BatchedDrone, Cylinder, Ground, config
    half = 0.5 * np.deg2rad(max_rates) * dt
    want = 4 if half <= 0.008 else 3 if half <= 0.03 else 2 if half <= 0.05 else 1 if half <= 0.25 else 0
    assert want == ang, (half, want)
    params = config.load_params(None)
    params["drone"]["max_rates"] = max_rates
    c = fo.derive_consts(params, os.path.join(os.path.dirname(GOLDEN), "..", "fpyv_b200", "config", "t_motos_f80_motor_test.csv"), dt=dt)
    n = 2048
    rng = np.random.default_rng(int(max_rates))
    pos = np.stack([rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.05, 6, n)], 1)
    vel, rpy = rng.normal(0, 1, (n, 3)), rng.uniform(-40, 40, (n, 3))
    acts = rng.uniform(-1, 1, (20, n, 4)).astype(np.float32).astype(np.float64)   # identical inputs: float32-representable sticks
    far = Cylinder(np.array([300.0, 0.0, 0.0]), 1.0, 5.0)
    for objs in (None, [far, Ground()]):                       # hot kernel / general kernel
        d = BatchedDrone(params, num_envs=n, device=DEV, substeps=K, dt=dt, packed=packed)
        d.reset(pos, vel, rpy)
        s = fo.drone_reset(c, pos, vel, rpy)
        worst1, worst_groups = 0.0, {}
        for t in range(20):
            if t < 6:                                          # teacher-forced single steps first
                set_state(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust)
                s.pos, s.vel = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
                s.R = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
                s.prev_rates, s.prev_thrust = d.prev_rates.double().cpu().numpy(), d.prev_thrust.double().cpu().numpy()
            fo.drone_step(c, s, acts[t], substeps=K)
            d.step(acts[t], None, objs, return_obs=False)
            e = drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust)
            if t < 6:
                assert np.array_equal(d.done.cpu().numpy(), s.done)
                if float(e.max()) > worst1:
                    worst1, worst_groups = float(e.max()), dict(drone_err.last)
        flying = ~s.done & (s.pos[:, 2] > 0.5)
        assert worst1 <= TOL_STEP, (worst1, objs is not None, worst_groups)
        assert float(e[flying].max()) <= 2e-4, float(e[flying].max())      # 14 free-running steps
    if packed:                                                 # the fused rollout uses the same angle kernels
        a = BatchedDrone(params, num_envs=n, device=DEV, substeps=K, dt=dt)
        b = BatchedDrone(params, num_envs=n, device=DEV, substeps=K, dt=dt)
        a.reset(pos, vel, rpy)
        b.reset(pos, vel, rpy)
        A = torch.as_tensor(acts[:8], dtype=torch.float32, device=DEV).contiguous()
        a.rollout(A, fused=True)
        for t in range(8):
            b.step(A[t], return_obs=False)
        assert torch.equal(a._state, b._state)


@pytest.mark.parametrize("packed", [True, False])
def test_general_path_without_ground_and_with_damped_spring(packed):
    """Two configurations only the general kernel serves: no ground plane in the object list (the crash test on the
    motor heights, components.py:239, still applies) and a damped contact spring (handle_collisions' damping_constant,
    components.py:198, is 0 in the reference's call but a parameter of its spring_force, kinematics.py:56-59)."""
    n = 4096
    rng = np.random.default_rng(91)
    pos = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(-0.05, 0.6, n)], 1)
    vel, rpy = rng.normal(0, 1.5, (n, 3)), rng.uniform(-30, 30, (n, 3))
    act = rng.uniform(-1, 1, (n, 4)).astype(np.float32).astype(np.float64)       # identical inputs on both sides
    for ground, damping in ((False, 0.0), (True, 7.5)):
        c = fo_consts()
        c.ground, c.spring_c = ground, damping
        d = make(n, packed=packed, ground=ground)
        d._p.spring_c = damping
        d.reset(pos, vel, rpy)
        s = fo.drone_reset(c, pos, vel, rpy)
        s.pos, s.vel = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
        s.R = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
        fo.drone_step(c, s, act)
        d.step(act, return_obs=False)
        err = drone_err(d, np.concatenate([s.pos, s.vel], 1), s.R, s.prev_rates, s.prev_thrust)
        done = d.done.cpu().numpy()
        mism = done != s.done
        assert err[~mism].max() <= TOL_STEP, (ground, damping, err.max(), drone_err.last)
        assert mism.sum() <= 2 and s.done.sum() > 100 and (~s.done).sum() > 100
        if damping:
            assert np.abs(s.acc[~s.done & (s.pos[:, 2] < 0.15)]).max() > 15       # springs engaged


def test_two_batches_on_two_streams_concurrently():
    """Independent batches stepped from two CUDA streams at once (their persistent grids compete for the SMs): every
    launch is ordered on the caller's current stream only, results are those of the serial run, bit for bit."""
    n, K = 400_000, 8
    g = torch.Generator(device=DEV).manual_seed(23)
    mk = lambda: (torch.randn(n, 3, device=DEV, generator=g) * 5 + torch.tensor([0.0, 0.0, 12.0], device=DEV),
                  torch.randn(n, 3, device=DEV, generator=g), (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30)
    inits = [mk(), mk()]
    acts = [torch.rand(n, 4, device=DEV, generator=g) * 2 - 1 for _ in range(10)]
    ref = []
    for init in inits:
        d = make(n, substeps=K, dt=1e-3, thrust_lut=2049)
        d.reset(*init)
        for a in acts:
            d.step(a, return_obs=False)
        ref.append(d._state.clone())
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ds = []
    for init, st in zip(inits, streams):
        with torch.cuda.stream(st):
            d = make(n, substeps=K, dt=1e-3, thrust_lut=2049)
            d.reset(*init)
        ds.append(d)
    torch.cuda.synchronize()
    for a in acts:
        for d, st in zip(ds, streams):
            with torch.cuda.stream(st):
                d.step(a, return_obs=False, chained=True)
    torch.cuda.synchronize()
    for d, r in zip(ds, ref):
        assert torch.equal(d._state, r)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_api_fuzz_all_step_forms_agree(seed):
    """A random interleaving of every way to advance a batch -- plain steps, chained steps, fused rollouts, per-step
    chained rollouts, host-buffer steps, masked resets, a step with obstacles, a change of CTA slots -- against a twin
    advanced with plain steps only: bit-identical state, flags and statistics at the end."""
    from fpyv_b200 import Cylinder, Ground
    n, K = 300_001, 4
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=DEV).manual_seed(100 + seed)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=DEV, generator=g) * 2.95
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    a = make(n, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    b = make(n, substeps=K, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    host = torch.empty(n, 4, dtype=torch.float32, pin_memory=True)
    host_done = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    objs = [Cylinder(np.array([2.0, 1.0, 0.0]), 1.0, 4.0), Ground()]
    new_act = lambda: (torch.rand(n, 4, device=DEV, generator=g) * 2 - 1).contiguous()
    log = []
    for it in range(40):
        op = rng.choice(["plain", "chained", "chained", "fused", "unfused", "host", "reset", "objects", "slots"])
        log.append(op)
        if op in ("plain", "chained"):
            x = new_act()
            a.step(x, return_obs=False, chained=(op == "chained"))
            b.step(x, return_obs=False)
        elif op in ("fused", "unfused"):
            T = int(rng.integers(1, 5))
            xs = torch.stack([new_act() for _ in range(T)]).contiguous()
            a.rollout(xs, fused=(op == "fused"))
            for t in range(T):
                b.step(xs[t], return_obs=False)
        elif op == "host":
            host.copy_(new_act().cpu())
            a.step_host(host, host_done)
            torch.cuda.current_stream().synchronize()
            b.step(host.to(DEV), return_obs=False)
            assert torch.equal(host_done, b.done.to(torch.uint8).cpu())
        elif op == "reset":
            mask = torch.rand(n, device=DEV, generator=g) < 0.1
            for d in (a, b):
                d.reset(pos, vel, rpy, mask=mask)
        elif op == "objects":
            x = new_act()
            for d in (a, b):
                d.step(x, None, objs, return_obs=False)
        elif op == "slots":
            a.cta_slots = int(rng.choice([0, 1, 2, 3]))
    torch.cuda.synchronize()
    assert torch.equal(a._state.view(torch.int32), b._state.view(torch.int32)), log
    assert torch.equal(a.done, b.done), log
    sa, sb = a.episode_stats(), b.episode_stats()
    assert all(sa[k] == sb[k] or (sa[k] != sa[k] and sb[k] != sb[k]) for k in sa), (sa, sb, log)


def test_non_finite_inputs_stay_contained():
    """NaN / Inf stick commands poison only their own env: the others match a clean run bit for bit, the table lookup and
    the stores stay in bounds, and the statistics count the non-finite envs."""
    n, K = 100_000, 8
    g = torch.Generator(device=DEV).manual_seed(31)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 5 + torch.rand(n, device=DEV, generator=g) * 5
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    act = (torch.rand(n, 4, device=DEV, generator=g) * 2 - 1).contiguous()
    bad = act.clone()
    idx = torch.arange(0, n, 97, device=DEV)
    bad[idx[0::3], 3] = float("nan")
    bad[idx[1::3], 0] = float("inf")
    bad[idx[2::3], 3] = -float("inf")
    clean = make(n, substeps=K, dt=1e-3, thrust_lut=2049)
    dirty = make(n, substeps=K, dt=1e-3, thrust_lut=2049)
    for d in (clean, dirty):
        d.reset(pos, vel, rpy)
    for _ in range(3):
        clean.step(act, return_obs=False)
        dirty.step(bad, return_obs=False)
    torch.cuda.synchronize()
    ok = torch.ones(n, dtype=torch.bool, device=DEV)
    ok[idx] = False
    assert torch.equal(clean._state[:, :n][:, ok], dirty._state[:, :n][:, ok])
    assert not bool(torch.isfinite(dirty.position[idx[0::3]]).all(dim=1).any())       # NaN throttle -> NaN state
    st = dirty.episode_stats()
    assert st["nonfinite"] > 0 and clean.episode_stats()["nonfinite"] == 0


def test_step_host_sticks_matches_the_joystick_path():
    """fpv_drone_step_host_sticks (uint16 raw sticks from pinned host memory, calibration on the device, sliced pipeline)
    against `rc.feed(raw); step(None)` -- the reference's joystick path -- bit for bit, several steps in a row."""
    n = 300_000
    g = torch.Generator(device=DEV).manual_seed(12)
    pos = torch.randn(n, 3, device=DEV, generator=g) * 5
    pos[:, 2] = 0.05 + torch.rand(n, device=DEV, generator=g) * 2.95
    vel = torch.randn(n, 3, device=DEV, generator=g)
    rpy = (torch.rand(n, 3, device=DEV, generator=g) * 2 - 1) * 30
    a = make(n, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    b = make(n, substeps=8, dt=1e-3, auto_reset=True, thrust_lut=2049)
    a.reset(pos, vel, rpy)
    b.reset(pos, vel, rpy)
    rng = np.random.default_rng(3)
    done_h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for i in range(5):
        raw4 = rng.integers(0, 65536, (n, 4), dtype=np.uint16)
        host = torch.from_numpy(raw4).pin_memory()
        a.step_host_sticks(host, done_h)
        torch.cuda.current_stream().synchronize()
        raw6 = np.zeros((n, 6), dtype=np.int32)
        raw6[:, [0, 1, 2, 5]] = raw4
        b.rc.feed(raw6)
        b.step(None, return_obs=False)
        assert torch.equal(done_h, b.done.to(torch.uint8).cpu()), i
        assert torch.equal(a._last_action, b._last_action), i
    assert torch.equal(a._state, b._state)
    with pytest.raises(ValueError):
        a.step_host_sticks(torch.zeros((n, 4), dtype=torch.int32), done_h)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_driving_two_devices():
    """The library's launch facts (dynamic shared memory opt-in, occupancy, SM count, host-step copy streams) are cached
    per DEVICE: one process stepping batches on cuda:0 and cuda:1 alternately gets the single-device result on both,
    and a call made while another device is current is refused with a message."""
    from fpyv_b200 import BatchedDrone
    n, K = 100_000, 8
    g = torch.Generator().manual_seed(31)
    pos = torch.randn(n, 3, generator=g) * 5 + torch.tensor([0.0, 0.0, 3.0])
    vel, rpy = torch.randn(n, 3, generator=g), (torch.rand(n, 3, generator=g) * 2 - 1) * 30
    acts = [torch.rand(n, 4, generator=g) * 2 - 1 for _ in range(6)]
    host_done = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    pinned = [a.pin_memory() for a in acts]
    ds = []
    for i in range(2):
        with torch.cuda.device(i):
            d = BatchedDrone(None, num_envs=n, device=f"cuda:{i}", substeps=K, dt=1e-3, thrust_lut=2049, auto_reset=True)
            d.reset(pos.to(d.device), vel.to(d.device), rpy.to(d.device))
            ds.append(d)
    for t, a in enumerate(acts):
        for i, d in enumerate(ds):
            with torch.cuda.device(i):
                if t < 3:
                    d.step(a.to(d.device), return_obs=False)
                elif t < 5:
                    d.step_host(pinned[t], host_done[i])
                else:
                    d.rollout(a.to(d.device)[None].contiguous())
    for i in range(2):
        torch.cuda.synchronize(i)
    assert torch.equal(ds[0]._state.cpu(), ds[1]._state.cpu())
    assert torch.equal(host_done[0], host_done[1])
    with torch.cuda.device(0):
        with pytest.raises(RuntimeError, match="current CUDA device"):
            ds[1].reset(pos.to("cuda:1"), vel.to("cuda:1"), rpy.to("cuda:1"))
