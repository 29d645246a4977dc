"""Pins the CPU restatement (oracle/fpv_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  float64 vs float64: tolerance 1e-11 relative."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, CONFIG
from oracle import fpv_oracle as fo

TOL = 1e-11


def consts(dt=None):
    import yaml
    with open(os.path.join(CONFIG, "params.yaml")) as f:
        params = yaml.safe_load(f)
    return fo.derive_consts(params, os.path.join(CONFIG, "t_motos_f80_motor_test.csv"), dt=dt)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_consts():
    g = load("consts")
    c = consts()
    assert c.dt == float(g["dt"]) and c.mass == float(g["mass"]) and c.gravity == float(g["gravity"])
    np.testing.assert_allclose(c.k_drag, -0.5 * g["drag_coef"] * 1.2225 * g["cross_section_areas"], rtol=1e-15)
    np.testing.assert_allclose(c.motor_rel, g["motors_relative_position"], rtol=0, atol=1e-16)
    np.testing.assert_allclose(c.poly, g["poly"], rtol=1e-12)
    np.testing.assert_allclose(c.inv_poly, g["inv_poly"], rtol=1e-12)
    np.testing.assert_allclose(c.min_force, g["min_force"], rtol=1e-13)
    np.testing.assert_allclose(c.max_force, g["max_force"], rtol=1e-13)
    np.testing.assert_allclose(fo.throttle2thrust(c, g["t2t_x"]), g["t2t_y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(fo.thrust2throttle(c, g["inv_x"]), g["inv_y"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["drone_kat", "drone_random", "drone_ground", "drone_wind",
                                  "drone_1ms_k8", "drone_overdrive", "drone_sticks"])
def test_drone_trajectories(name):
    g = load(name)
    c = consts(dt=float(g["dt"]))
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    wind = g["wind"] if "wind" in g else None
    T = g["actions"].shape[0]
    worst = 0.0
    for t in range(T):
        ret = fo.drone_substep(c, s, g["actions"][t], wind)
        worst = max(worst, rel(np.concatenate([s.pos, s.vel], 1), g["state"][t]), rel(s.R, g["R"][t]),
                    rel(s.prev_rates, g["prev_rates"][t]), rel(s.prev_thrust, g["prev_thrust"][t]))
        assert np.array_equal(s.done, g["done"][t].astype(bool)), f"done differs at step {t}"
        if "ret_Rt" in g:
            worst = max(worst, rel(ret[0], g["ret_Rt"][t]), rel(ret[1], g["ret_gyro"][t]), rel(ret[2], g["ret_acc"][t]))
    assert worst < TOL, worst


def test_drone_kat_numbers():
    """The literal numbers quoted in SURVEY.md section 8(a) 'verified end-to-end KATs'."""
    g = load("drone_kat")
    np.testing.assert_allclose(g["state"][0, 0], [0.01666666666666667, 0, 10, 0.99963325, 0, 0.3175312570533787], rtol=1e-12)
    np.testing.assert_allclose(g["prev_rates"][0, 0], [-42, 28, -14], rtol=1e-13)
    np.testing.assert_allclose(g["prev_thrust"][0, 0], 21.646406567402042, rtol=1e-13)
    np.testing.assert_allclose(g["state"][59, 0], [-3.92930023046, -10.738525209773, 19.101538015046,
                                                  -7.167691714416, -23.745589333267, 2.388572941485], rtol=1e-10)


def test_drone_k_substeps_equals_k_steps():
    g = load("drone_1ms_k8")
    c = consts(dt=float(g["dt"]))
    K = int(g["hold"])
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    for j in range(g["actions"].shape[0] // K):
        fo.drone_step(c, s, g["actions"][j * K], substeps=K)
        t = j * K + K - 1
        assert rel(np.concatenate([s.pos, s.vel], 1), g["state"][t]) < TOL
        assert np.array_equal(s.done, g["done"][j * K:t + 1].any(axis=0))


def test_drone_override():
    g = load("drone_override")
    c = consts(dt=float(g["dt"]))
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    for t in range(g["actions"].shape[0]):
        fo.drone_substep(c, s, g["actions"][t], R_override=g["R_override"][t], thrust_override=g["thrust_override"][t])
        assert rel(np.concatenate([s.pos, s.vel], 1), g["state"][t]) < TOL
        assert rel(s.R, g["R"][t]) < TOL and rel(s.prev_thrust, g["prev_thrust"][t]) < TOL
        assert np.array_equal(s.done, g["done"][t].astype(bool))


def test_drone_objects():
    g = load("drone_objects")
    c = consts(dt=float(g["dt"]))
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    c.ground = False   # object order in the reference run was [sphere, cylinder, ground]
    objs = [fo.SphereObj(g["sph"][:3], g["sph"][3]), fo.CylinderObj(g["cyl"][:3], g["cyl"][3], g["cyl"][4]),
            fo._GroundPlane()]
    assert g["done"].any()
    for t in range(g["actions"].shape[0]):
        fo.drone_substep(c, s, g["actions"][t], extra_objects=objs)
        assert rel(np.concatenate([s.pos, s.vel], 1), g["state"][t]) < TOL, t
        assert np.array_equal(s.done, g["done"][t].astype(bool)), t


@pytest.mark.parametrize("calib", ["frsky", "calibration"])
def test_sticks(calib):
    g = load("sticks_" + calib)
    cal = fo.StickCalib.from_json(os.path.join(CONFIG, calib + ".json"))
    np.testing.assert_allclose(fo.calib_read(cal, g["raw"]), g["calibrated"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(fo.sticks_to_action(cal, g["raw"]), g["action"], rtol=1e-13, atol=1e-15)
    if calib == "frsky":
        np.testing.assert_allclose(g["action"][0], [0.458173335726, 0.525909423828, 0.220733642578, 0.166999183822], rtol=1e-9)


def test_drone_sticks_actions():
    g = load("drone_sticks")
    cal = fo.StickCalib.from_json(os.path.join(CONFIG, "frsky.json"))
    np.testing.assert_allclose(fo.sticks_to_action(cal, g["raw"][:, 0]), g["actions"][:, 0], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["racer_demo", "racer_random"])
def test_racer(name):
    g = load(name)
    n = g["actions"].shape[1]
    worst = 0.0
    for e in range(n):
        c = fo.RacerConsts(gains=g["gains"][e])
        s = fo.RacerState(1)
        for t in range(g["actions"].shape[0]):
            fo.racer_step(c, s, g["actions"][t, e])
            worst = max(worst, rel(s.pos[0], g["position"][t, e]), rel(s.vel[0], g["velocity"][t, e]),
                        rel(s.R[0], g["R"][t, e]), rel(s.omega[0], g["omega"][t, e]), rel(s.torque[0], g["torque"][t, e]))
    assert worst < 1e-10, worst


def test_racer_kat_numbers():
    """SURVEY.md section 8(a) row 14 KAT."""
    g = load("racer_demo")
    np.testing.assert_allclose(g["torque"][0, 0], [160, 20, 0], rtol=1e-13)
    np.testing.assert_allclose(g["omega"][0, 0], [79.36015872031744, 9.92001984003968, 0], rtol=1e-13)
    np.testing.assert_allclose(g["omega"][1, 0], [79.99488253921018, 9.999360317401273, 0], rtol=1e-13)


def test_l1_kats():
    g = load("l1_kat")
    c = consts()
    R = fo.euler_matrix(np.array([0.1]), np.array([0.2]), np.array([0.3]))
    np.testing.assert_allclose(R[0], g["R_euler_0p1_0p2_0p3"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(g["drag"], [-0.299868680248, 0.124139708619, -0.370008173779], rtol=1e-9)
    vs = np.array([[3.5, -2, 1]])
    drag = (R @ ((c.k_drag * (np.swapaxes(R, 1, 2) @ vs[..., None])[..., 0]) * np.linalg.norm(vs))[..., None])[..., 0]
    np.testing.assert_allclose(drag[0], g["drag"], rtol=1e-13)
    np.testing.assert_allclose(g["spring"], [0, 0, 5.0], atol=1e-14)
    q = fo.matrix_to_quaternion(R)
    np.testing.assert_allclose(q[0], g["quat"], rtol=1e-14)
    np.testing.assert_allclose(fo.quaternion_to_matrix(q)[0], g["quat_R"], rtol=1e-14, atol=1e-16)
    ang = np.deg2rad(np.array([[-42., 28, -14]])) / 60
    E = fo.euler_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    np.testing.assert_allclose((R @ np.swapaxes(E, 1, 2))[0], g["rot_by_rates"], rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("name", ["drone_kat", "drone_random", "drone_ground", "drone_wind", "drone_1ms_k8"])
def test_c_restatement_matches_reference(name):
    """oracle/fpv_oracle.c (the CPU-baseline code) against the same golden vectors."""
    from oracle import c_oracle
    g = load(name)
    c = consts(dt=float(g["dt"]))
    k = c_oracle.make_consts(c)
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    pos, vel, R = s.pos.copy(), s.vel.copy(), np.ascontiguousarray(s.R)
    pr, pt = np.zeros_like(pos), np.zeros(len(pos))
    wind = g["wind"] if "wind" in g else None
    worst = 0.0
    for t in range(g["actions"].shape[0]):
        done = c_oracle.drone_step(k, pos, vel, R, pr, pt, np.ascontiguousarray(g["actions"][t]), wind,
                                   threads=2 if t % 2 else 1)
        worst = max(worst, rel(np.concatenate([pos, vel], 1), g["state"][t]), rel(R, g["R"][t]),
                    rel(pr, g["prev_rates"][t]), rel(pt, g["prev_thrust"][t]))
        assert np.array_equal(done.astype(bool), g["done"][t].astype(bool))
    assert worst < 1e-11, worst


@pytest.mark.parametrize("name", ["drone_kat", "drone_random", "drone_ground", "drone_wind", "drone_1ms_k8"])
def test_numba_restatement_matches_reference(name):
    """oracle/numba_oracle.py (bench.py's `cpu_baseline_numba` leg; the reference itself ships no Numba path,
    kinematics.py:6,14 are commented out) against the same golden vectors."""
    from oracle import numba_oracle as no
    g = load(name)
    c = consts(dt=float(g["dt"]))
    k, mrel = no.make_consts(c)
    s = fo.drone_reset(c, g["pos0"], g["vel0"], g["rpy0"])
    pos, vel, R = s.pos.copy(), s.vel.copy(), np.ascontiguousarray(s.R)
    pr, pt = np.zeros_like(pos), np.zeros(len(pos))
    wind = np.ascontiguousarray(g["wind"], dtype=np.float64) if "wind" in g else np.zeros(3)
    done = np.zeros(len(pos), dtype=np.bool_)
    worst = 0.0
    for t in range(g["actions"].shape[0]):
        no.drone_step(k, mrel, pos, vel, R, pr, pt, np.ascontiguousarray(g["actions"][t], dtype=np.float64), wind, 1, done)
        worst = max(worst, rel(np.concatenate([pos, vel], 1), g["state"][t]), rel(R, g["R"][t]),
                    rel(pr, g["prev_rates"][t]), rel(pt, g["prev_thrust"][t]))
        assert np.array_equal(done, g["done"][t].astype(bool))
    assert worst < 1e-11, worst
