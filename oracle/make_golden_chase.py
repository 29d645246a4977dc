"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the chase pipeline (camera depth splat, target pixel, point-and-shoot
autopilot + components.PID, step with the rotation/thrust override), produced by executing the UNMODIFIED reference
through `oracle/ref_shim.py`.  Run where /root/reference is mounted:

    python oracle/make_golden_chase.py

Writes tests/golden/chase_camera.npz, chase_autopilot.npz, chase_loop.npz.
"""
from __future__ import annotations

import os
import signal
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sphere_points(n):
    """Deterministic unit-sphere point set standing in for the (absent) `icosphere` package: shape data, i.e. an
    INPUT of the pipeline -- stored in the golden files next to the outputs."""
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def main():
    ns = rs.load()
    comp, hf = ns["components"], ns["helper_functions"]
    comp.icosphere = lambda nu=1: (sphere_points(10 * nu * nu + 2), np.zeros((1, 3), dtype=int))
    params = rs.load_params()

    # ------------------------------------------------------------------ world (shared by all cases)
    rng = np.random.default_rng(21)
    with rs.quiet():
        ground = comp.Ground(size=40, resolution=24, random=False)
        cyl_a = comp.Cylinder(np.array([6.0, 2.0, 0.0]), 1.5, 8.0, 10, 12)
        cyl_b = comp.Cylinder(np.array([-4.0, -7.0, 0.0]), 2.0, 5.0, 8, 9)
        gate = comp.Gate(np.array([3.0, -5.0, 2.5]), hf.euler_angles_to_rotation_matrix(0, 0, 0.7), 5.0, shape="circle", resolution=17)
        target = comp.Target(np.array([8.0, 0.5, 4.0]), 1.0, nu=4)
    world = [target, cyl_a, cyl_b, gate, ground]
    world_pts = {f"obj{i}": np.array(o.points, dtype=np.float64) for i, o in enumerate(world)}

    # ------------------------------------------------------------------ camera: depth / binary images, rays
    n = 10
    pos = np.stack([rng.uniform(-6, 2, n), rng.uniform(-4, 4, n), rng.uniform(1.0, 7, n)], axis=1)
    rpy = np.stack([rng.uniform(-25, 25, n), rng.uniform(-25, 25, n), rng.uniform(-60, 60, n)], axis=1)
    rpy[0] = 0
    rpy[1] = [0, 0, 180]            # looks away from everything but the ground
    pos[2] = [7.6, 0.5, 4.0]        # inside the target's bounding box
    depth15, depth25, binary, tgt15, cam_pos, cam_R, Pm, Rn = [], [], [], [], [], [], [], []
    pix = rng.uniform(0, 1, (n, 5, 2)) * np.array(params["camera"]["resolution"])
    rays = {k: [] for k in ("world", "drone", "camera", "drone_rotation_matrix")}
    d = rs.make_drone(params)
    for e in range(n):
        R = hf.euler_angles_to_rotation_matrix(*np.deg2rad(rpy[e]))
        Rn.append(R)
        cam = d.camera
        cam.update(pos[e], R)
        cam_pos.append(cam.position.copy())
        cam_R.append(cam.rotation_matrix.copy())
        Pm.append(cam.projection_matrix.copy())
        with rs.quiet():
            depth15.append(np.array(cam.render_depth_image(list(world), max_depth=15)).astype(np.uint8))
            depth25.append(np.array(cam.render_depth_image(list(world), max_depth=25)).astype(np.uint8))
            tgt15.append(np.array(cam.render_depth_image([target], max_depth=15)).astype(np.uint8))
            binary.append(np.array(cam.render_image(list(world))).astype(np.uint8))
        for k in rays:
            rays[k].append(np.array([cam.pixel2direction(p, ref_frame=k, drone_rotation_matrix=R) for p in pix[e]]))
    np.savez_compressed(os.path.join(OUT, "chase_camera.npz"), pos=pos, rpy=rpy, R=np.array(Rn), cam_pos=np.array(cam_pos),
                        cam_R=np.array(cam_R), P=np.array(Pm), depth15=np.array(depth15), depth25=np.array(depth25),
                        target15=np.array(tgt15), binary=np.array(binary), pixels=pix,
                        **{"ray_" + k: np.array(v) for k, v in rays.items()}, target_pos=np.array(target.position),
                        target_radius=np.float64(target.radius), n_objects=np.int64(len(world)), **world_pts)
    print("chase_camera: non-zero depth pixels per env", [int((x > 0).sum()) for x in depth15],
          "target pixels", [int((x > 0).sum()) for x in tgt15])

    # ------------------------------------------------------------------ autopilot, single calls (all frames / modes)
    n = 24
    rng = np.random.default_rng(22)
    pos = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(0.5, 8, n)], axis=1)
    vel = rng.normal(0, 3, (n, 3))
    rpy = rng.uniform(-30, 30, (n, 3))
    tpos = pos + rng.normal(0, 6, (n, 3))
    trad = rng.uniform(0.3, 1.5, n)
    pixel = rng.uniform(0, 1, (n, 2)) * np.array(params["camera"]["resolution"])
    out = {}
    for frame in ("world", "drone"):
        for mode in ("level", "frontarget"):
            rots, forces, pid_state = [], [], []
            for e in range(n):
                dr = rs.make_drone(params)
                with rs.quiet():
                    dr.reset(pos[e], vel[e], rpy[e])
                    tg = comp.Target(tpos[e].copy(), trad[e], nu=1)
                    calls = []
                    for _ in range(3):       # three consecutive calls: the PID's integral / derivative filter evolve
                        rot, f = dr.calculate_needed_force_orientation(pixel[e], tg, ref_frame=frame, mode=mode)
                        calls.append((np.array(rot), float(f)))
                rots.append([c[0] for c in calls])
                forces.append([c[1] for c in calls])
                p = dr.force_multiplier_pid
                pid_state.append([p.integral, p.prev_derivative, p.previous_error, float(p.is_first)])
            out[f"rot_{frame}_{mode}"] = np.array(rots)
            out[f"force_{frame}_{mode}"] = np.array(forces)
            out[f"pid_{frame}_{mode}"] = np.array(pid_state)
    np.savez_compressed(os.path.join(OUT, "chase_autopilot.npz"), pos=pos, vel=vel, rpy=rpy, target_pos=tpos,
                        target_radius=trad, pixel=pixel, min_force=dr.min_throttle_in_force,
                        max_force=dr.max_throttle_in_force, **out)

    # ------------------------------------------------------------------ point_and_shoot (components.py:312-381)
    n = 24
    rng = np.random.default_rng(24)
    pos = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(0.5, 8, n)], axis=1)
    vel = rng.normal(0, 3, (n, 3))
    # (|v| is kept moderate: the reference's force-limit loop :355-363 never terminates once drag + lift + gravity
    #  alone exceed max_throttle_in_force; every call below runs under a watchdog for that reason)
    rpy = rng.uniform(-30, 30, (n, 3))
    pixel = rng.uniform(0.2, 0.8, (n, 2)) * np.array(params["camera"]["resolution"])
    action = rng.uniform(-0.6, 0.6, (n, 4))
    out = {}
    for frame in ("world", "drone"):
        for mode in ("level", "frontarget"):
            rots, forces, pid_state, pv = [], [], [], []
            for e in range(n):
                dr = rs.make_drone(params)
                with rs.quiet():
                    dr.reset(pos[e], vel[e], rpy[e])
                    calls = []
                    for k in range(10):     # the PID integrates the pixel error: later calls saturate the multiplier
                        signal.alarm(10)
                        rot, f = dr.point_and_shoot(pixel[e] + 3.0 * k, action[e], ref_frame=frame, mode=mode)
                        signal.alarm(0)
                        calls.append((np.array(rot), float(f)))
                rots.append([c[0] for c in calls])
                forces.append([c[1] for c in calls])
                p = dr.force_multiplier_pid
                pid_state.append([p.integral, p.prev_derivative, p.previous_error, float(p.is_first)])
                pv.append(np.concatenate([dr.pixel_velocity, dr.prev_pixel]))
            out[f"rot_{frame}_{mode}"] = np.array(rots)
            out[f"force_{frame}_{mode}"] = np.array(forces)
            out[f"pid_{frame}_{mode}"] = np.array(pid_state)
            out[f"pixvel_{frame}_{mode}"] = np.array(pv)
    np.savez_compressed(os.path.join(OUT, "chase_point_and_shoot.npz"), pos=pos, vel=vel, rpy=rpy, pixel=pixel, action=action,
                        max_force=dr.max_throttle_in_force, min_force=dr.min_throttle_in_force,
                        position=np.array([dr.convert_action2position(a_) for a_ in action]), **out)
    print("chase_point_and_shoot: (env, call) pairs held at the force limit:",
          int((out["force_world_level"] >= dr.max_throttle_in_force * (1 - 1e-6)).sum()), "of", out["force_world_level"].size)

    # ------------------------------------------------------------------ the closed loop of simulator.py:98-110 (dim == 3)
    n, T = 8, 40
    rng = np.random.default_rng(23)
    pos = np.stack([rng.uniform(-8, -2, n), rng.uniform(-4, 4, n), rng.uniform(2.0, 8, n)], axis=1)
    vel = np.stack([rng.uniform(0.5, 3, n), rng.normal(0, 0.5, n), rng.normal(0, 0.5, n)], axis=1)
    rpy = np.stack([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(-25, 25, n)], axis=1)
    rpy[0] = [0, 0, 170]                                # never sees the target: plain stick steps throughout
    tpos = np.stack([rng.uniform(5, 10, n), rng.uniform(-3, 3, n), rng.uniform(2, 6, n)], axis=1)
    action = np.array([-0.1, 0.0, 0.0, 0.0])            # simulator.py:88
    keys = ("state", "R", "prev_rates", "prev_thrust", "done", "pixel", "seen", "rot", "force", "pid")
    res = {k: [] for k in keys}
    gnd = rs.make_ground()
    for e in range(n):
        dr = rs.make_drone(params)
        with rs.quiet():
            dr.reset(pos[e], vel[e], rpy[e])
            tg = comp.Target(tpos[e].copy(), 1.0, nu=4)
        row = {k: [] for k in keys}
        for t in range(T):
            with rs.quiet():
                img = dr.camera.render_depth_image([tg], max_depth=15)
                tp = np.array(np.where(img > 0))
                if tp.shape[1] == 0:
                    dr.step(action=action, wind_velocity_vector=np.zeros(3), object_list=[tg, gnd])
                    px, rot, f, seen = np.full(2, np.nan), np.full((3, 3), np.nan), np.nan, False
                else:
                    px = tp.mean(1)[::-1]
                    rot, f = dr.calculate_needed_force_orientation(px, tg, ref_frame="world", mode="level")
                    dr.step(action=action, wind_velocity_vector=np.zeros(3), object_list=[tg, gnd],
                            rotation_matrix=np.array(rot).copy(), thrust_force=f)
                    seen = True
            p = dr.force_multiplier_pid
            for k, v in (("state", dr.state.copy()), ("R", np.array(dr.rotation_matrix).copy()),
                         ("prev_rates", np.array(dr.prev_rates, dtype=np.float64).copy()), ("prev_thrust", float(dr.prev_thrust)),
                         ("done", bool(dr.done)), ("pixel", np.array(px, dtype=np.float64)), ("seen", seen),
                         ("rot", np.array(rot, dtype=np.float64)), ("force", float(f)),
                         ("pid", np.array([p.integral, p.prev_derivative, p.previous_error, float(p.is_first)]))):
                row[k].append(v)
        for k in keys:
            res[k].append(np.array(row[k]))
    res = {k: np.stack(v, axis=1) for k, v in res.items()}     # [T, n, ...]
    np.savez_compressed(os.path.join(OUT, "chase_loop.npz"), pos0=pos, vel0=vel, rpy0=rpy, target_pos=tpos,
                        target_radius=np.float64(1.0), target_points=sphere_points(10 * 16 + 2), action=action,
                        dt=np.float64(1 / 60), **res)
    print("chase_loop: steps with the target in view", int(res["seen"].sum()), "of", res["seen"].size,
          "; done", int(res["done"].sum()))


if __name__ == "__main__":
    main()
