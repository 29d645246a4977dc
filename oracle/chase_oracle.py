"""TEST INFRASTRUCTURE ONLY -- CPU restatement (float64 NumPy) of the reference's chase pipeline, the caller that
sits on either side of `Drone.step` in `src/core/simulator.py:98-110`:

    depth image of the target (Camera.render_depth_image)  ->  mean target pixel (simulator.py:104-108)
    ->  Drone.calculate_needed_force_orientation (point-and-shoot autopilot + components.PID)
    ->  Drone.step(..., rotation_matrix=rot_mat, thrust_force=force_size)

Only `tests/` and `__graft_entry__.smoke()` import this module; the product never does.

Parity status: PINNED by executing the unmodified reference through `oracle/ref_shim.py`
(`oracle/make_golden_chase.py` -> `tests/golden/chase_*.npz`, re-checked by `tests/test_oracle_golden.py`).
Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

WORLD2CAM = np.array([[0.0, 1, 0], [0, 0, -1], [1, 0, 0]])   # src/utils/helper_functions.py:11-13


def _rot(angle, axis):
    """src/utils/helper_functions.py:19-36."""
    c, s = np.cos(angle), np.sin(angle)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def euler_matrix(roll, pitch, yaw):
    """src/utils/helper_functions.py:39-44: Rz(yaw) Ry(pitch) Rx(roll)."""
    return _rot(yaw, "z") @ _rot(pitch, "y") @ _rot(roll, "x")


# --------------------------------------------------------------------------------------
# Camera (src/utils/components.py:449-629)
# --------------------------------------------------------------------------------------
@dataclass
class CameraConsts:
    resolution: np.ndarray        # [W, H]
    focal_length: float
    rel_pos: np.ndarray           # position_relative_to_frame
    rel_rot: np.ndarray           # WORLD2CAM.T @ euler(deg2rad(pitch), 0, 0)      components.py:455
    K: np.ndarray                 # intrinsic matrix                                helper_functions.py:15-16
    K_inv: np.ndarray


def camera_consts(params: dict) -> CameraConsts:
    """Camera.__init__, components.py:450-470 (fov given, focal_length None)."""
    cam = params["camera"]
    res = np.array(cam["resolution"], dtype=np.float64)
    f = res[0] / (2 * np.tan(np.deg2rad(cam["fov"]) / 2))                                       # :472-474
    rel_rot = WORLD2CAM.T @ euler_matrix(np.deg2rad(cam["camera_angle"]), 0, 0)                  # :455
    K = np.array([[f, 0, res[0] / 2], [0, f, res[1] / 2], [0, 0, 1]])                            # :469-470
    return CameraConsts(resolution=res, focal_length=float(f),
                        rel_pos=np.array(cam["position_relative_to_frame"], dtype=np.float64), rel_rot=rel_rot, K=K,
                        K_inv=np.linalg.inv(K))


def camera_update(c: CameraConsts, pos, R):
    """Camera.update, components.py:501-503.  pos [n,3], R [n,3,3] -> cam_pos [n,3], cam_R [n,3,3]."""
    return pos + np.einsum("nij,j->ni", R, c.rel_pos), R @ c.rel_rot


def pixel2direction(c: CameraConsts, pixel, cam_R=None, ref_frame="world", drone_R=None):
    """Camera.pixel2direction, components.py:505-526.  pixel [n,2] -> unit vectors [n,3]."""
    p = np.concatenate([np.asarray(pixel, dtype=np.float64), np.ones((len(pixel), 1))], axis=1)
    ray = p @ c.K_inv.T
    if ref_frame == "world":
        d = np.einsum("nij,nj->ni", cam_R, ray)
    elif ref_frame == "drone":
        d = ray @ c.rel_rot.T
    elif ref_frame == "camera":
        d = ray
    elif ref_frame == "drone_rotation_matrix" and drone_R is not None:
        d = np.einsum("nij,nj->ni", drone_R @ c.rel_rot, ray)
    else:
        raise ValueError("ref_frame must be world, drone or camera")
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def projection_matrix(c: CameraConsts, cam_pos, cam_R):
    """Camera.projection_matrix, components.py:532-536: K @ inv([R t; 0 1])[:3] for ONE env."""
    E = np.eye(4)
    E[:3, :3] = cam_R
    E[:3, 3] = cam_pos
    return c.K @ np.linalg.inv(E)[:3, :]


def project(P, points):
    """Camera.project, components.py:558-568: keep depth > 0, divide, truncate toward zero (astype(int))."""
    h = P @ np.vstack([points.T, np.ones(len(points))])
    h = h.T
    depth = h[:, 2]
    keep = depth > 0
    px = (h[keep, :2] / depth[keep].reshape(-1, 1)).astype(int)
    return px, depth[keep]


def bbox3d(points):
    """helper_functions.py:120-136."""
    lo, hi = points.min(axis=0), points.max(axis=0)
    box = np.zeros((8, 3))
    box[:4, 0], box[4:, 0] = lo[0], hi[0]
    box[::2, 1], box[1::2, 1] = lo[1], hi[1]
    box[[0, 1, 4, 5], 2], box[[2, 3, 6, 7], 2] = lo[2], hi[2]
    return box


def pruned(P, objects, resolution):
    """Camera.pruned_objects_list, components.py:584-599: an object stays if at least one bbox corner is in front
    of the camera and the 2-D box of those corners overlaps the frame (max > 0 and min < resolution, per axis)."""
    keep = []
    for pts in objects:
        px, _ = project(P, bbox3d(pts))
        if len(px) == 0:
            continue
        lo, hi = px.min(axis=0), px.max(axis=0)
        if np.all(hi > 0) and np.all(lo < resolution):
            keep.append(pts)
    return keep


def render_depth_image(c: CameraConsts, cam_pos, cam_R, objects, max_depth=10.0):
    """Camera.render_depth_image, components.py:614-629, for ONE env.  objects: list of [P_o,3] world points.
    Returns uint8 [H, W]."""
    W, H = int(c.resolution[0]), int(c.resolution[1])
    img = np.zeros((H, W))
    P = projection_matrix(c, cam_pos, cam_R)
    objs = pruned(P, objects, c.resolution)
    if len(objs) == 0:
        return img            # :617-618 returns the float zero image; as uint8 it is all zeros too
    px, depth = project(P, np.vstack(objs))
    for z, p in zip(depth, px):
        if 0 <= p[0] < W and 0 <= p[1] < H and (img[p[1], p[0]] == 0 or img[p[1], p[0]] > z):
            img[p[1], p[0]] = z
    np.clip(img, 0, max_depth, out=img)
    img[img == 0] = max_depth
    return (255 * (1 - img / max_depth)).astype(np.uint8)


def render_image(c: CameraConsts, cam_pos, cam_R, objects):
    """Camera.render_image, components.py:601-612 (binary splat)."""
    W, H = int(c.resolution[0]), int(c.resolution[1])
    img = np.zeros((H, W))
    P = projection_matrix(c, cam_pos, cam_R)
    objs = pruned(P, objects, c.resolution)
    if len(objs) == 0:
        return img
    px, _ = project(P, np.vstack(objs))
    for p in px:
        if 0 <= p[0] < W and 0 <= p[1] < H:
            img[p[1], p[0]] = 1
    return img


def target_pixel(img):
    """src/core/simulator.py:104-108: mean (x, y) of the non-zero pixels, or None when the target is not seen."""
    idx = np.array(np.where(img > 0))
    if idx.shape[1] == 0:
        return None
    return idx.mean(1)[::-1]


# --------------------------------------------------------------------------------------
# components.PID (src/utils/components.py:15-54), batched: every field is an [n] array
# --------------------------------------------------------------------------------------
@dataclass
class PIDState:
    integral: np.ndarray
    prev_derivative: np.ndarray
    previous_error: np.ndarray
    is_first: np.ndarray          # bool


def pid_reset(n):
    """components.py:35-41."""
    return PIDState(np.zeros(n), np.zeros(n), np.zeros(n), np.ones(n, dtype=bool))


def pid_call(s: PIDState, current, target, kP, kI, kD, dt, integral_clip, min_output, max_output, dtr):
    """components.py:43-54.  Mutates s; returns the clipped output [n]."""
    error = current - target
    s.integral = np.clip(0.99 * s.integral + error * dt, -integral_clip, integral_clip)
    deriv = np.clip(np.where(s.is_first, 0.0, 1.0) * (error - s.previous_error) / dt, -1, 1)
    deriv = (1 - dtr) * s.prev_derivative + dtr * deriv
    s.prev_derivative = deriv
    s.is_first = np.zeros_like(s.is_first)
    s.previous_error = error
    return np.clip(kP * error + kI * s.integral + kD * deriv, min_output, max_output)


# --------------------------------------------------------------------------------------
# Drone.calculate_needed_force_orientation (src/utils/components.py:258-304)
# --------------------------------------------------------------------------------------
@dataclass
class AutopilotConsts:
    mass: float
    dt: float
    virtual_drag_coef: float
    virtual_lift_coef: float
    tof_effective_dist: float
    keep_distance: float
    uwb_max_range: float
    kP: float
    kI: float
    kD: float
    integral_clip: float
    min_output: float            # = min_throttle_in_force (components.py:143)
    max_output: float            # = max_throttle_in_force (components.py:144)
    dtr: float                   # derivative_transition_rate


def autopilot_consts(params: dict, min_force: float, max_force: float, dt: float | None = None) -> AutopilotConsts:
    """components.py:96-97, :113-118, :143-145."""
    dr, pns, pid = params["drone"], params["point_and_shoot"], params["drone"]["force_multiplier_pid"]
    return AutopilotConsts(
        mass=dr["mass"] / 1000, dt=float(1 / params["simulator"]["fps"]) if dt is None else float(dt),
        virtual_drag_coef=float(pns["virtual_drag_coefficient"]), virtual_lift_coef=float(pns["virtual_lift_coefficient"]),
        tof_effective_dist=float(pns["tof_effective_distance"]), keep_distance=float(dr["keep_distance"]),
        uwb_max_range=float(dr["UWB_sensor_max_range"]), kP=float(pid["kP"]), kI=float(pid["kI"]), kD=float(pid["kD"]),
        integral_clip=float(pid["integral_clip"]), min_output=float(min_force), max_output=float(max_force),
        dtr=float(pid["derivative_transition_rate"]))


def needed_force_orientation(a: AutopilotConsts, cam: CameraConsts, pid: PIDState, pixel, target_pos, target_radius,
                             pos, vel, R, ref_frame="world", mode="level"):
    """components.py:258-304 for n envs.  pixel [n,2], target_pos [n,3], target_radius [n] or scalar, pos/vel [n,3],
    R [n,3,3] (drone attitude at call time; the camera pose is the one Camera.update stored from it).
    Returns (rotation_to_apply_force [n,3,3], force_vector_norm [n]).  Mutates pid."""
    _, cam_R = camera_update(cam, pos, R)
    dir2t = pixel2direction(cam, pixel, cam_R)                                   # :268 -- ALWAYS the world frame
    speed = np.linalg.norm(vel, axis=1, keepdims=True)
    if ref_frame == "world":                                                     # :270-273
        gravity = np.tile([0.0, 0.0, -9.81 * a.mass], (len(pos), 1))             # kinematics.py:41-45 with g=9.81
        cosang = np.einsum("ni,ni->n", vel / speed, dir2t)[:, None]
        vdrag = -(cosang - 1) / 2 * -vel * speed
    elif ref_frame == "drone":                                                   # :274-277
        gravity = np.einsum("nij,j->ni", R, np.array([0.0, 0.0, -9.81 * a.mass]))   # :256
        rv = np.einsum("nij,nj->ni", R, vel)
        cosang = np.einsum("ni,ni->n", rv / speed, dir2t)[:, None]
        vdrag = -(cosang - 1) / 2 * -rv * speed
    else:
        raise ValueError("Unknown reference frame")
    vdrag_force = a.virtual_drag_coef * vdrag                                    # :286
    z = pos[:, 2:3]
    vlift = (z < a.tof_effective_dist) * -(a.tof_effective_dist - z) * a.virtual_lift_coef * gravity * (1 + np.abs(vel[:, 2:3]))  # :287
    dist = np.minimum(np.linalg.norm(pos - target_pos, axis=1) - target_radius, a.uwb_max_range)   # :288, :770-771
    mult = pid_call(pid, dist, a.keep_distance, a.kP, a.kI, a.kD, a.dt, a.integral_clip, a.min_output, a.max_output, a.dtr)
    mult = np.clip(mult, a.min_output, a.max_output)                             # :291
    force = mult[:, None] * dir2t + vdrag_force + vlift - gravity                # :293
    fnorm = np.linalg.norm(force, axis=1)
    if mode == "level":                                                          # :295-297
        yv = np.cross(force, gravity)
        xv = np.cross(yv, force)
    elif mode == "frontarget":                                                   # :298-300
        yv = np.cross(force, dir2t)
        xv = np.cross(yv, force)
    else:
        raise ValueError("Unknown mode")
    rot = np.stack([xv, yv, force], axis=2)                                      # columns, :303
    rot = rot / np.linalg.norm(rot, axis=1, keepdims=True)                       # :304 (column norms)
    return rot, fnorm


# --------------------------------------------------------------------------------------
# Drone.point_and_shoot (src/utils/components.py:312-381) and convert_action2position (:383-387)
# --------------------------------------------------------------------------------------
def convert_action2position(cam: CameraConsts, action):
    """components.py:383-387: (resolution[i] / 2 * (1 + action[i])).astype(int), i = 0, 1."""
    return (cam.resolution[None, :] / 2 * (1 + np.asarray(action, dtype=np.float64)[:, :2])).astype(int)


def point_and_shoot(a: AutopilotConsts, cam: CameraConsts, pid: PIDState, pixel, action, pos, vel, R, max_force,
                    ref_frame="world", mode="level", max_iter=64):
    """components.py:312-381 for n envs.  pixel [n,2], action [n,4] = (target row/column on screen in [-1,1]^2,
    virtual-target offset in [-1,1]^2).  Returns (rotation_to_apply_force [n,3,3], force_vector_norm [n], shifted
    pixel [n,2]).  The reference's `while force_vector_norm > max_throttle_in_force` loop (:350-358) does not terminate
    when the drag / lift / gravity terms alone exceed the limit; it is capped at `max_iter` passes here."""
    action = np.asarray(action, dtype=np.float64)
    pixel = np.asarray(pixel, dtype=np.float64) + action[:, 2:] * cam.resolution / 2          # :322-323
    _, cam_R = camera_update(cam, pos, R)
    dir2t = pixel2direction(cam, pixel, cam_R)                                                 # :332
    speed = np.linalg.norm(vel, axis=1, keepdims=True)
    if ref_frame == "world":
        gravity = np.tile([0.0, 0.0, -9.81 * a.mass], (len(pos), 1))
        cosang = np.einsum("ni,ni->n", vel / speed, dir2t)[:, None]
        vdrag = -(cosang - 1) / 2 * -vel * speed
    elif ref_frame == "drone":
        gravity = np.einsum("nij,j->ni", R, np.array([0.0, 0.0, -9.81 * a.mass]))
        rv = np.einsum("nij,nj->ni", R, vel)
        cosang = np.einsum("ni,ni->n", rv / speed, dir2t)[:, None]
        vdrag = -(cosang - 1) / 2 * -rv * speed
    else:
        raise ValueError("Unknown reference frame")
    vdrag_force = a.virtual_drag_coef * vdrag
    z = pos[:, 2:3]
    vlift = (z < a.tof_effective_dist) * -(a.tof_effective_dist - z) * a.virtual_lift_coef * gravity * -np.minimum(vel[:, 2:3], 0.0)  # :345
    position = convert_action2position(cam, action)                                            # :348
    mult = pid_call(pid, pixel[:, 1], position[:, 1].astype(np.float64), a.kP, a.kI, a.kD, a.dt, a.integral_clip,
                    a.min_output, a.max_output, a.dtr)                                         # :350
    rest = vdrag_force + vlift - gravity
    force = mult[:, None] * dir2t + rest
    fnorm = np.linalg.norm(force, axis=1)
    crit = np.full(len(pos), 0.9999)
    for _ in range(max_iter):                                                                  # :355-363
        over = fnorm > max_force
        if not over.any():
            break
        mult = np.where(over, np.clip(mult * crit, a.min_output, a.max_output), mult)
        force = np.where(over[:, None], mult[:, None] * dir2t + rest, force)
        fnorm = np.where(over, np.linalg.norm(force, axis=1), fnorm)
        crit = np.where(over, max_force / fnorm, crit)
    if mode == "level":
        yv = np.cross(force, gravity)
    elif mode == "frontarget":
        yv = np.cross(force, dir2t)
    else:
        raise ValueError("Unknown mode")
    xv = np.cross(yv, force)
    rot = np.stack([xv, yv, force], axis=2)
    rot = rot / np.linalg.norm(rot, axis=1, keepdims=True)
    return rot, fnorm, pixel
