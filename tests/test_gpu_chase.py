"""GPU parity of the chase pipeline (camera depth splat, target pixel, point-and-shoot autopilot + PID, step with the
per-env rotation/thrust override) -- all through the C ABI -- against (1) golden vectors produced by the unmodified
reference (oracle/make_golden_chase.py) and (2) the float64 oracle on the same inputs.
Images and pixel indices are integer work: bit-exact.  Floating-point outputs: <= 1e-5 relative per step."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import chase_oracle as co
from oracle import fpv_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def params():
    from fpyv_b200 import config
    return config.load_params(None)


def world_objects(g):
    return [g[f"obj{i}"] for i in range(int(g["n_objects"]))]


def make_cam(n):
    from fpyv_b200 import BatchedCamera
    return BatchedCamera.from_params(params(), n, DEV)


def test_camera_pose_projection_rays_vs_reference():
    g = load("chase_camera")
    n = len(g["pos"])
    cam = make_cam(n)
    cam.update(g["pos"], g["R"])
    np.testing.assert_allclose(cam.position.cpu().numpy(), g["cam_pos"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cam.rotation_matrix.cpu().numpy(), g["cam_R"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cam.projection_matrix.cpu().numpy(), g["P"], rtol=1e-10, atol=1e-9)
    assert cam.focal_length == pytest.approx(co.camera_consts(params()).focal_length, rel=1e-15)
    for k in ("world", "drone", "camera"):
        for j in range(g["pixels"].shape[1]):
            r = cam.pixel2direction(g["pixels"][:, j], ref_frame=k).cpu().numpy()
            np.testing.assert_allclose(r, g["ray_" + k][:, j], rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        cam.pixel2direction(g["pixels"][:, 0], ref_frame="moon")


@pytest.mark.parametrize("key,max_depth,objs", [("depth15", 15, "world"), ("depth25", 25, "world"), ("target15", 15, "target"),
                                                ("binary", 0, "world")])
def test_images_bit_exact_vs_reference(key, max_depth, objs):
    g = load("chase_camera")
    n = len(g["pos"])
    cam = make_cam(n)
    cam.update(g["pos"], g["R"])
    objects = world_objects(g) if objs == "world" else [g["obj0"]]
    img = cam.render_depth_image(objects, max_depth) if max_depth else cam.render_image(objects)
    assert img.dtype == torch.uint8 and tuple(img.shape) == g[key].shape
    assert np.array_equal(img.cpu().numpy(), g[key])
    assert int((g[key] > 0).sum()) > 0


def test_target_pixel_matches_reference_extraction():
    g = load("chase_camera")
    n = len(g["pos"])
    cam = make_cam(n)
    cam.update(g["pos"], g["R"])
    px, seen = cam.target_pixel([g["obj0"]], 15)
    px, seen = px.cpu().numpy(), seen.cpu().numpy()
    for e in range(n):
        ref = co.target_pixel(g["target15"][e])
        assert bool(seen[e]) == (ref is not None)
        if ref is not None:
            assert np.array_equal(px[e], ref)          # integer sums / count: exact
    assert (seen == 0).any() and (seen == 1).any()
    # per-env object offsets: the same target expressed as unit-centred vertices + a per-env translation
    verts = g["obj0"] - g["target_pos"]
    off = np.tile(g["target_pos"], (n, 1, 1))
    px2, seen2 = cam.target_pixel([verts], 15, offsets=off)
    assert np.array_equal(seen2.cpu().numpy(), seen)
    np.testing.assert_allclose(px2.cpu().numpy(), px, rtol=0, atol=0.51)   # (v + c) vs precomputed v + c: 1-ulp pixel flips only
    img = cam.render_depth_image([verts], 15, offsets=off).cpu().numpy()
    assert np.mean(img != g["target15"]) < 1e-5


def _drone_from(pos, vel, rpy, **kw):
    from fpyv_b200 import BatchedDrone
    d = BatchedDrone(None, num_envs=len(pos), device=DEV, **kw)
    d.reset(pos, vel, rpy)
    return d


@pytest.mark.parametrize("frame", ["world", "drone"])
@pytest.mark.parametrize("mode", ["level", "frontarget"])
def test_autopilot_vs_reference_and_oracle(frame, mode):
    from fpyv_b200 import Autopilot
    g = load("chase_autopilot")
    n = len(g["pos"])
    d = _drone_from(g["pos"], g["vel"], g["rpy"])
    ap = Autopilot(d)
    assert ap.force_multiplier_pid.min_output == pytest.approx(float(g["min_force"]), rel=1e-12)
    assert ap.force_multiplier_pid.max_output == pytest.approx(float(g["max_force"]), rel=1e-12)
    # the oracle on exactly the device's float32 state
    p = params()
    cam_c = co.camera_consts(p)
    a = co.autopilot_consts(p, d.min_throttle_in_force, d.max_throttle_in_force)
    pid = co.pid_reset(n)
    pos64, vel64 = d.position.double().cpu().numpy(), d.velocity.double().cpu().numpy()
    R64 = fo.quaternion_to_matrix(d.quaternion.double().cpu().numpy())
    for call in range(3):
        rot, f = ap.calculate_needed_force_orientation(g["pixel"], g["target_pos"], g["target_radius"], ref_frame=frame, mode=mode)
        rot, f = rot.double().cpu().numpy(), f.double().cpu().numpy()
        np.testing.assert_allclose(rot, g[f"rot_{frame}_{mode}"][:, call], rtol=0, atol=1e-5)
        np.testing.assert_allclose(f, g[f"force_{frame}_{mode}"][:, call], rtol=1e-5)
        ro, fo_ = co.needed_force_orientation(a, cam_c, pid, g["pixel"], g["target_pos"], g["target_radius"], pos64, vel64, R64,
                                              ref_frame=frame, mode=mode)
        np.testing.assert_allclose(rot, ro, rtol=0, atol=2e-7)
        np.testing.assert_allclose(f, fo_, rtol=2e-7)
    st = ap.force_multiplier_pid.state.cpu().numpy()
    np.testing.assert_allclose(st, g[f"pid_{frame}_{mode}"], rtol=1e-5, atol=1e-6)
    # quaternion output == the rotation output
    ap.reset()
    q, _ = ap.calculate_needed_force_orientation(g["pixel"], g["target_pos"], g["target_radius"], ref_frame=frame, mode=mode,
                                                 as_quaternion=True)
    np.testing.assert_allclose(fo.quaternion_to_matrix(q.double().cpu().numpy()), g[f"rot_{frame}_{mode}"][:, 0], rtol=0, atol=1e-5)
    with pytest.raises(ValueError):
        ap.calculate_needed_force_orientation(g["pixel"], g["target_pos"], 1.0, ref_frame="moon")
    with pytest.raises(ValueError):
        ap.calculate_needed_force_orientation(g["pixel"], g["target_pos"], 1.0, mode="sideways")


def _loop_iteration(d, cam, ap, world, tpos, tgt, action, ground):
    """One pass of simulator.py:98-110 for every env."""
    cam.update_from(d)
    px, seen = cam.target_pixel(world, 15, offsets=tpos[:, None, :])
    q, f = ap.calculate_needed_force_orientation(px, tpos, 1.0, seen=seen, as_quaternion=True)
    d.step(action, np.zeros(3), [tgt, ground], rotation_matrix=q, thrust_force=f, return_obs=False)
    return px, seen, q, f


def test_closed_loop_single_steps_from_reference_states():
    """Teacher-forced: every (t, env) of the reference's closed loop is restarted from the reference's own state at
    t-1 (drone state AND PID state), one loop iteration on the GPU, compared with the reference at t."""
    from fpyv_b200 import Autopilot, BatchedDrone, Ground, Target, World
    g = load("chase_loop")
    T, n = g["state"].shape[:2]
    N = (T - 1) * n
    # the collision loop takes ONE target per launch: the test's envs all use their own target -> run env by env
    verts = g["target_points"] * float(g["target_radius"])
    world = World([verts], DEV)
    worst = 0.0
    for e in range(n):
        m = T - 1
        d = BatchedDrone(None, num_envs=m, device=DEV)
        d.reset(g["state"][:-1, e, :3], g["state"][:-1, e, 3:], np.zeros((m, 3)))
        f32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=DEV)
        d.set_rotation_matrix(f32(g["R"][:-1, e]))
        d.prev_rates.copy_(f32(g["prev_rates"][:-1, e]))
        d.prev_thrust.copy_(f32(g["prev_thrust"][:-1, e]))
        cam = make_cam(m)
        ap = Autopilot(d, cam)
        ap.force_multiplier_pid.state.copy_(torch.as_tensor(g["pid"][:-1, e], device=DEV))
        tpos = torch.as_tensor(np.tile(g["target_pos"][e], (m, 1)), dtype=torch.float64, device=DEV)
        tgt = Target(g["target_pos"][e], float(g["target_radius"]))
        px, seen, q, f = _loop_iteration(d, cam, ap, world, tpos, tgt, np.tile(g["action"], (m, 1)), Ground())
        seen = seen.cpu().numpy().astype(bool)
        assert np.array_equal(seen, g["seen"][1:, e])
        np.testing.assert_allclose(px.cpu().numpy()[seen], g["pixel"][1:, e][seen], rtol=0, atol=0.02)
        np.testing.assert_allclose(f.double().cpu().numpy()[seen], g["force"][1:, e][seen], rtol=2e-4)
        err = np.maximum(np.abs(d.position.double().cpu().numpy() - g["state"][1:, e, :3]).max(1) /
                         np.maximum(1, np.abs(g["state"][1:, e, :3]).max(1)),
                         np.abs(d.velocity.double().cpu().numpy() - g["state"][1:, e, 3:]).max(1) /
                         np.maximum(1, np.abs(g["state"][1:, e, 3:]).max(1)))
        errR = np.abs(d.rotation_matrix.double().cpu().numpy() - g["R"][1:, e]).reshape(m, -1).max(1)
        worst = max(worst, err.max(), errR.max())
        assert not d.done.any()
    print(f"chase loop, single steps from reference states: max rel err {worst:.2e}")
    assert worst < 1e-5      # north_star tolerance: <= 1e-5 relative per single step (measured 5e-7)


def test_closed_loop_free_running_vs_reference():
    from fpyv_b200 import Autopilot, BatchedDrone, Ground, Target, World
    g = load("chase_loop")
    T, n = g["state"].shape[:2]
    verts = g["target_points"] * float(g["target_radius"])
    world = World([verts], DEV)
    worst = 0.0
    for e in range(n):
        d = BatchedDrone(None, num_envs=1, device=DEV)
        d.reset(g["pos0"][e:e + 1], g["vel0"][e:e + 1], g["rpy0"][e:e + 1])
        cam = make_cam(1)
        ap = Autopilot(d, cam)
        tpos = torch.as_tensor(g["target_pos"][e:e + 1], dtype=torch.float64, device=DEV)
        tgt = Target(g["target_pos"][e], float(g["target_radius"]))
        for t in range(T):
            px, seen, q, f = _loop_iteration(d, cam, ap, world, tpos, tgt, g["action"][None], Ground())
            assert bool(seen[0]) == bool(g["seen"][t, e]), (e, t)
            s = np.concatenate([d.position.double().cpu().numpy()[0], d.velocity.double().cpu().numpy()[0]])
            worst = max(worst, np.abs(s - g["state"][t, e]).max() / max(1.0, np.abs(g["state"][t, e]).max()))
    print(f"chase loop, {T} free-running closed-loop steps: max rel divergence {worst:.2e}")
    assert worst < 1e-3      # 40 closed-loop steps through a pixel-quantised feedback path (measured 1.2e-4)


def test_full_frame_batch_against_oracle():
    """4,096 cameras at 640x480 over the stock-sized world: a subsample of frames bit-exact against the oracle, and the
    per-env override mask (NaN thrust) leaves unseen envs on their stick command."""
    from fpyv_b200 import Autopilot, BatchedDrone, Cylinder, Ground, Target, World
    n = 4096
    rng = np.random.default_rng(3)
    ground = Ground(60, 50, random=True, rng=rng)
    cyls = [Cylinder(np.array([rng.normal(0, 10), rng.normal(0, 10), 0.0]), 2.0, 10.0, 10, 25, random=True, rng=rng) for _ in range(5)]
    tgt = Target(np.array([0.0, 0.0, 3.0]), 1.0, nu=5)
    objs = [tgt, *cyls, ground]
    gpos = torch.Generator(device=DEV).manual_seed(9)
    pos = torch.randn(n, 3, device=DEV, generator=gpos) * torch.tensor([8.0, 8.0, 0.0], device=DEV)
    pos[:, 2] = 1.0 + torch.rand(n, device=DEV, generator=gpos) * 9
    d = BatchedDrone(None, num_envs=n, device=DEV)
    d.reset(pos, torch.randn(n, 3, device=DEV, generator=gpos), (torch.rand(n, 3, device=DEV, generator=gpos) * 2 - 1) * 40)
    from fpyv_b200 import BatchedCamera
    cam = BatchedCamera.from_params(params(), n, DEV)
    cam.update_from(d)
    img = cam.render_depth_image(objs, 25)
    assert tuple(img.shape) == (n, 480, 640)
    c = co.camera_consts(params())
    cp, cR = cam.position.cpu().numpy(), cam.rotation_matrix.cpu().numpy()
    for e in (0, 1, 777, 4095):
        ref = co.render_depth_image(c, cp[e], cR[e], [np.asarray(o.points) for o in objs], 25)
        assert np.array_equal(img[e].cpu().numpy(), ref.astype(np.uint8)), e
    px, seen = cam.target_pixel([tgt], 25)
    seen_np = seen.cpu().numpy().astype(bool)
    assert seen_np.any() and (~seen_np).any()
    # unseen envs must step exactly like a plain stick step
    ap = Autopilot(d, cam)
    q, f = ap.calculate_needed_force_orientation(px, tgt.position, tgt.radius, seen=seen, as_quaternion=True)
    assert torch.isnan(f[~seen.bool()]).all() and torch.isfinite(f[seen.bool()]).all()
    act = torch.rand(n, 4, device=DEV, generator=gpos) * 2 - 1
    d2 = BatchedDrone(None, num_envs=n, device=DEV)
    d2._state.copy_(d._state)
    d2._is_reset = True
    d.step(act, np.zeros(3), [Ground()], rotation_matrix=q, thrust_force=f, return_obs=False)
    # same (general) kernel, override masked out everywhere
    d2.step(act, np.zeros(3), [Ground()], rotation_matrix=q, thrust_force=torch.full_like(f, float("nan")), return_obs=False)
    un = ~seen.bool()
    assert torch.equal(d.position[un], d2.position[un]) and torch.equal(d.quaternion[un], d2.quaternion[un])
    assert not torch.equal(d.quaternion[seen.bool()], d2.quaternion[seen.bool()])


def test_guard_cells_around_every_output_of_the_chase_kernels():
    """Raw C-ABI calls on ragged sizes with sentinel bytes on both sides of every output buffer (compute-sanitizer is
    not available on the GPU pool: out-of-bounds writes are caught by guard cells instead)."""
    from fpyv_b200 import BatchedDrone, _lib
    lib = _lib.load()
    g = load("chase_camera")
    n = 7
    cam = make_cam(n)
    p = cam._params()
    W, H = int(cam.resolution[0]), int(cam.resolution[1])
    d = BatchedDrone(None, num_envs=n, device=DEV)
    d.reset(g["pos"][:n], np.ones((n, 3)), g["rpy"][:n])
    st = _lib.current_stream(torch.device(DEV))
    G = 4096

    def guarded(nbytes, dtype):
        buf = torch.full((G + nbytes + G,), 0xAB, dtype=torch.uint8, device=DEV)
        return buf, buf[G:G + nbytes].view(dtype)

    pose_b, pose = guarded(n * 12 * 8, torch.float64)
    _lib.check(lib.fpv_camera_update(p, _lib.ptr(d._state), n, d._stride, _lib.ptr(pose), st))
    from fpyv_b200 import World
    w = World(world_objects(g), DEV)
    img_b, img = guarded(n * W * H, torch.uint8)
    keep_b, keep = guarded(n * w.n_objects, torch.uint8)
    _lib.check(lib.fpv_camera_render(p, _lib.ptr(pose), n, _lib.ptr(w.points), w.n_points, _lib.ptr(w.boxes), w.n_objects, None,
                                     15.0, _lib.ptr(keep), _lib.ptr(img), st))
    px_b, px = guarded(n * 2 * 8, torch.float64)
    seen_b, seen = guarded(n, torch.uint8)
    _lib.check(lib.fpv_camera_target_pixel(p, _lib.ptr(pose), n, _lib.ptr(w.points), w.n_points, _lib.ptr(w.boxes), w.n_objects,
                                           None, 15.0, _lib.ptr(px), _lib.ptr(seen), st))
    ray_b, ray = guarded(n * 3 * 8, torch.float64)
    _lib.check(lib.fpv_camera_rays(p, _lib.ptr(pose), n, _lib.ptr(px), 0, _lib.ptr(ray), st))
    from fpyv_b200 import Autopilot
    ap = Autopilot(d, cam)
    pid_b, pid = guarded(n * 4 * 8, torch.float64)
    pid.copy_(torch.tensor([0.0, 0, 0, 1], dtype=torch.float64, device=DEV).repeat(n))
    rot_b, rot = guarded(n * 9 * 4, torch.float32)
    q_b, q = guarded(n * 16, torch.float32)
    f_b, f = guarded(n * 4, torch.float32)
    tp = torch.as_tensor(np.tile(g["target_pos"], (n, 1)), dtype=torch.float64, device=DEV).contiguous()
    tr = torch.ones(n, dtype=torch.float64, device=DEV)
    _lib.check(lib.fpv_autopilot(ap._params("world", "level"), p, _lib.ptr(d._state), n, d._stride, _lib.ptr(px), _lib.ptr(seen),
                                 _lib.ptr(tp), _lib.ptr(tr), _lib.ptr(pid), _lib.ptr(rot), _lib.ptr(q), _lib.ptr(f), st))
    torch.cuda.synchronize()
    for name, b in (("pose", pose_b), ("image", img_b), ("keep", keep_b), ("pixel", px_b), ("seen", seen_b), ("rays", ray_b),
                    ("pid", pid_b), ("rot", rot_b), ("quat", q_b), ("force", f_b)):
        assert bool((b[:G] == 0xAB).all()) and bool((b[-G:] == 0xAB).all()), name
    # and the guarded outputs are the API's outputs
    cam.update_from(d)
    assert torch.equal(pose.view(n, 12), cam._pose)
    assert torch.equal(img.view(n, H, W), cam.render_depth_image(w, 15))


@pytest.mark.parametrize("frame", ["world", "drone"])
@pytest.mark.parametrize("mode", ["level", "frontarget"])
def test_point_and_shoot_vs_reference(frame, mode):
    """Drone.point_and_shoot (components.py:312-381): 10 consecutive calls per env against the reference's outputs,
    including the calls where the force-limit loop (:355-363) holds the force at max_throttle_in_force."""
    from fpyv_b200 import Autopilot
    g = load("chase_point_and_shoot")
    d = _drone_from(g["pos"], g["vel"], g["rpy"])
    ap = Autopilot(d)
    assert np.array_equal(ap.convert_action2position(g["action"]).cpu().numpy(), g["position"])
    calls = g[f"force_{frame}_{mode}"].shape[1]
    limited = 0
    for call in range(calls):
        rot, f = ap.point_and_shoot(g["pixel"] + 3.0 * call, g["action"], ref_frame=frame, mode=mode)
        np.testing.assert_allclose(rot.double().cpu().numpy(), g[f"rot_{frame}_{mode}"][:, call], rtol=0, atol=1e-5)
        np.testing.assert_allclose(f.double().cpu().numpy(), g[f"force_{frame}_{mode}"][:, call], rtol=1e-5)
        limited += int((g[f"force_{frame}_{mode}"][:, call] >= float(g["max_force"]) * (1 - 1e-6)).sum())
    assert limited > 0
    np.testing.assert_allclose(ap.force_multiplier_pid.state.cpu().numpy(), g[f"pid_{frame}_{mode}"], rtol=1e-5, atol=1e-6)
    pv = torch.cat([ap.pixel_velocity, ap.prev_pixel], dim=1).cpu().numpy()
    np.testing.assert_allclose(pv, g[f"pixvel_{frame}_{mode}"], rtol=1e-9, atol=1e-9)
    q, f2 = ap.point_and_shoot(g["pixel"], g["action"], ref_frame=frame, mode=mode, as_quaternion=True,
                               seen=np.arange(len(g["pos"])) % 2)
    assert torch.isnan(f2[::2]).all() and torch.isfinite(f2[1::2]).all()


@pytest.mark.parametrize("res,fov,pitch", [([160, 120], 90.0, 10.0), ([320, 200], 150.0, 0.0), ([64, 48], 60.0, 55.0)])
def test_other_cameras_bit_exact_against_oracle(res, fov, pitch):
    """Resolutions, fields of view and camera pitches other than params.yaml's: frames, binary image and target pixel
    bit-exact against the float64 oracle (which is pinned to the reference on the stock camera)."""
    import copy
    from fpyv_b200 import BatchedCamera
    g = load("chase_camera")
    p = copy.deepcopy(params())
    p["camera"].update(resolution=res, fov=fov, camera_angle=pitch)
    n = len(g["pos"])
    cam = BatchedCamera.from_params(p, n, DEV)
    cam.update(g["pos"], g["R"])
    c = co.camera_consts(p)
    cp, cR = co.camera_update(c, g["pos"], g["R"])
    np.testing.assert_allclose(cam.position.cpu().numpy(), cp, rtol=0, atol=1e-12)
    objs = world_objects(g)
    img = cam.render_depth_image(objs, 20).cpu().numpy()
    binimg = cam.render_image(objs).cpu().numpy()
    px, seen = cam.target_pixel([g["obj0"]], 20)
    lit = 0
    for e in range(n):
        ref = co.render_depth_image(c, cp[e], cR[e], objs, 20).astype(np.uint8)
        assert np.array_equal(img[e], ref), e
        assert np.array_equal(binimg[e], co.render_image(c, cp[e], cR[e], objs).astype(np.uint8)), e
        tp = co.target_pixel(co.render_depth_image(c, cp[e], cR[e], [g["obj0"]], 20))
        assert bool(seen[e]) == (tp is not None)
        if tp is not None:
            assert np.array_equal(px[e].cpu().numpy(), tp)
        lit += int((ref > 0).sum())
    assert lit > 50
