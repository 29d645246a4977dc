"""Pins oracle/chase_oracle.py (camera depth splat, target pixel, components.PID, point-and-shoot autopilot, the
closed loop of simulator.py:98-110) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden_chase.py).  float64 vs float64: 1e-10 relative; images bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, CONFIG
from oracle import chase_oracle as co
from oracle import fpv_oracle as fo


def params():
    import yaml
    with open(os.path.join(CONFIG, "params.yaml")) as f:
        return yaml.safe_load(f)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def world_objects(g):
    return [g[f"obj{i}"] for i in range(int(g["n_objects"]))]


def test_camera_pose_projection_and_rays():
    g = load("chase_camera")
    c = co.camera_consts(params())
    cam_pos, cam_R = co.camera_update(c, g["pos"], g["R"])
    np.testing.assert_allclose(cam_pos, g["cam_pos"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cam_R, g["cam_R"], rtol=0, atol=1e-12)
    for e in range(len(cam_pos)):
        np.testing.assert_allclose(co.projection_matrix(c, cam_pos[e], cam_R[e]), g["P"][e], rtol=1e-10, atol=1e-9)
    for k in ("world", "drone", "camera", "drone_rotation_matrix"):
        for j in range(g["pixels"].shape[1]):
            r = co.pixel2direction(c, g["pixels"][:, j], cam_R, ref_frame=k, drone_R=g["R"])
            np.testing.assert_allclose(r, g["ray_" + k][:, j], rtol=0, atol=1e-12)


@pytest.mark.parametrize("key,max_depth,objs", [("depth15", 15, "world"), ("depth25", 25, "world"), ("target15", 15, "target")])
def test_depth_images_bit_exact(key, max_depth, objs):
    g = load("chase_camera")
    c = co.camera_consts(params())
    objects = world_objects(g) if objs == "world" else [g["obj0"]]
    for e in range(len(g["pos"])):
        img = co.render_depth_image(c, g["cam_pos"][e], g["cam_R"][e], objects, max_depth)
        assert np.array_equal(img.astype(np.uint8), g[key][e]), (key, e)


def test_binary_image_and_target_pixel():
    g = load("chase_camera")
    c = co.camera_consts(params())
    for e in range(len(g["pos"])):
        img = co.render_image(c, g["cam_pos"][e], g["cam_R"][e], world_objects(g))
        assert np.array_equal(img.astype(np.uint8), g["binary"][e])
    assert co.target_pixel(g["target15"][1]) is None
    np.testing.assert_allclose(co.target_pixel(g["target15"][0]), np.array(np.where(g["target15"][0] > 0)).mean(1)[::-1])


@pytest.mark.parametrize("frame", ["world", "drone"])
@pytest.mark.parametrize("mode", ["level", "frontarget"])
def test_autopilot_single_calls(frame, mode):
    g = load("chase_autopilot")
    p = params()
    c = co.camera_consts(p)
    a = co.autopilot_consts(p, float(g["min_force"]), float(g["max_force"]))
    n = len(g["pos"])
    ang = np.deg2rad(g["rpy"])
    R = fo.euler_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    pid = co.pid_reset(n)
    for call in range(3):
        rot, f = co.needed_force_orientation(a, c, pid, g["pixel"], g["target_pos"], g["target_radius"], g["pos"], g["vel"],
                                             R, ref_frame=frame, mode=mode)
        np.testing.assert_allclose(rot, g[f"rot_{frame}_{mode}"][:, call], rtol=0, atol=1e-10)
        np.testing.assert_allclose(f, g[f"force_{frame}_{mode}"][:, call], rtol=1e-11)
    st = np.stack([pid.integral, pid.prev_derivative, pid.previous_error, pid.is_first.astype(float)], axis=1)
    np.testing.assert_allclose(st, g[f"pid_{frame}_{mode}"], rtol=1e-11, atol=1e-13)


def test_closed_loop_matches_reference():
    """simulator.py:98-110: render the target, average its pixels, autopilot, step with the override -- 40 steps."""
    g = load("chase_loop")
    p = params()
    cam = co.camera_consts(p)
    dc = fo.derive_consts(p, os.path.join(CONFIG, "t_motos_f80_motor_test.csv"))
    a = co.autopilot_consts(p, dc.min_force, dc.max_force)
    n, T = len(g["pos0"]), len(g["state"])
    pts = g["target_points"] * float(g["target_radius"])
    for e in range(n):
        s = fo.drone_reset(dc, g["pos0"][e], g["vel0"][e], g["rpy0"][e])
        pid = co.pid_reset(1)
        sphere = fo.SphereObj(g["target_pos"][e], float(g["target_radius"]))
        for t in range(T):
            cp, cR = co.camera_update(cam, s.pos, s.R)
            img = co.render_depth_image(cam, cp[0], cR[0], [pts + g["target_pos"][e]], 15)
            px = co.target_pixel(img)
            assert (px is not None) == bool(g["seen"][t, e]), (e, t)
            # the reference passes [target, ground]: extra objects come before the ground in the collision loop
            kw = dict(extra_objects=[sphere])
            if px is None:
                fo.drone_step(dc, s, g["action"], **kw)
            else:
                np.testing.assert_allclose(px, g["pixel"][t, e], rtol=0, atol=1e-9)
                rot, f = co.needed_force_orientation(a, cam, pid, px[None], g["target_pos"][e][None], float(g["target_radius"]),
                                                     s.pos, s.vel, s.R)
                np.testing.assert_allclose(rot[0], g["rot"][t, e], rtol=0, atol=1e-9)
                np.testing.assert_allclose(f[0], g["force"][t, e], rtol=1e-10)
                fo.drone_step(dc, s, g["action"], R_override=rot, thrust_override=f, **kw)
            np.testing.assert_allclose(np.concatenate([s.pos[0], s.vel[0]]), g["state"][t, e], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(s.R[0], g["R"][t, e], rtol=0, atol=1e-9)
            assert bool(s.done[0]) == bool(g["done"][t, e])


@pytest.mark.parametrize("frame", ["world", "drone"])
@pytest.mark.parametrize("mode", ["level", "frontarget"])
def test_point_and_shoot(frame, mode):
    """Drone.point_and_shoot (components.py:312-381) incl. its force-limit loop, 10 consecutive calls per env."""
    g = load("chase_point_and_shoot")
    p = params()
    c = co.camera_consts(p)
    a = co.autopilot_consts(p, float(g["min_force"]), float(g["max_force"]))
    n = len(g["pos"])
    ang = np.deg2rad(g["rpy"])
    R = fo.euler_matrix(ang[:, 0], ang[:, 1], ang[:, 2])
    assert np.array_equal(co.convert_action2position(c, g["action"]), g["position"])
    pid = co.pid_reset(n)
    prev = None
    for call in range(g[f"force_{frame}_{mode}"].shape[1]):
        rot, f, shifted = co.point_and_shoot(a, c, pid, g["pixel"] + 3.0 * call, g["action"], g["pos"], g["vel"], R,
                                             float(g["max_force"]), ref_frame=frame, mode=mode)
        np.testing.assert_allclose(rot, g[f"rot_{frame}_{mode}"][:, call], rtol=0, atol=1e-9)
        np.testing.assert_allclose(f, g[f"force_{frame}_{mode}"][:, call], rtol=1e-10)
        pv = np.zeros_like(shifted) if prev is None else (shifted - prev) / a.dt          # components.py:325-330
        prev = shifted
    np.testing.assert_allclose(np.concatenate([pv, prev], axis=1), g[f"pixvel_{frame}_{mode}"], rtol=1e-10, atol=1e-9)
    st = np.stack([pid.integral, pid.prev_derivative, pid.previous_error, pid.is_first.astype(float)], axis=1)
    np.testing.assert_allclose(st, g[f"pid_{frame}_{mode}"], rtol=1e-11, atol=1e-13)
    assert (g[f"force_{frame}_{mode}"] >= float(g["max_force"]) * (1 - 1e-6)).any()       # the limit loop ran
