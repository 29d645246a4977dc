"""`BatchedCamera`: the reference's `utils.components.Camera` (src/utils/components.py:449-629) for N drones looking at
one shared world, and `World`, the point clouds of an `object_list` packed for the device.

Name-for-name mirror: `.resolution .focal_length .fov .intrinsic_matrix .relative_rotation_matrix .relative_position`,
`update(drone_position, drone_rotation_matrix)` / `update_from(drone)`, `.position .rotation_matrix .projection_matrix`,
`pixel2direction(pixel, ref_frame)`, `render_depth_image(objects_list, max_depth)`, `render_image(objects_list)`,
plus `target_pixel(objects_list, max_depth)` = the pixel extraction of simulator.py:104-108 fused with the splat.
Geometry is float64 on the device (see fpv_api.h, "Chase pipeline"); images are uint8 [n, H, W] like the reference's
[H, W] frames.  No CPU path: every method is a launch through the C ABI."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

WORLD2CAM = np.array([[0.0, 1, 0], [0, 0, -1], [1, 0, 0]])   # helper_functions.py:11-13

_FRAMES = {"world": 0, "drone": 1, "camera": 2}


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])


class World:
    """The `.points` of every object of an object_list (Ground / Cylinder / Target / Gate, components.py:655-667,
    :697-708, :766-768, :803-805) packed as double[P][4] = x, y, z, object index, with each object's bbox3d
    (helper_functions.py:120-136).  `offsets` [n, n_objects, 3] optionally translates every object per env
    (per-env or moving targets: Target.update, components.py:770-772, shifts all vertices by the new position)."""

    def __init__(self, objects_list, device):
        if len(objects_list) < 1 or len(objects_list) > _lib.CAM_MAX_OBJECTS:
            raise ValueError(f"a world holds 1..{_lib.CAM_MAX_OBJECTS} objects (got {len(objects_list)})")
        pts, boxes = [], []
        for i, o in enumerate(objects_list):
            p = np.asarray(o if isinstance(o, np.ndarray) else o.points, dtype=np.float64).reshape(-1, 3)
            if len(p) == 0:
                raise ValueError(f"object {i} has no points")
            pts.append(np.concatenate([p, np.full((len(p), 1), float(i))], axis=1))
            boxes.append(np.concatenate([p.min(axis=0), p.max(axis=0)]))
        self.n_objects = len(objects_list)
        self.points = torch.from_numpy(np.ascontiguousarray(np.concatenate(pts))).to(device)
        self.boxes = torch.from_numpy(np.ascontiguousarray(np.stack(boxes))).to(device)
        self.n_points = int(self.points.shape[0])


class BatchedCamera:
    def __init__(self, camera_pitch_angle, position_relative_to_frame, resolution, fov=None, focal_length=None,
                 num_envs: int = 1, device="cuda:0"):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("fpyv_b200 runs on CUDA devices only (no CPU fallback)")
        self.num_envs = int(num_envs)
        self.resolution = np.array(resolution)
        self.relative_position = np.asarray(position_relative_to_frame, dtype=np.float64)
        self.relative_rotation_matrix = WORLD2CAM.T @ _rx(np.deg2rad(camera_pitch_angle))          # components.py:455
        self.focal_length, self.fov = focal_length, fov
        if focal_length is None and fov is not None:                                                # :464-465
            self.focal_length = self.convert_fov_to_focal_length(fov, self.resolution)
        elif focal_length is not None and fov is None:                                              # :466-467
            self.fov = self.convert_focal_length_to_fov(self.focal_length, self.resolution)
        self.intrinsic_matrix = None
        if self.focal_length is not None:                                                           # :468-470
            f = float(self.focal_length)
            self.intrinsic_matrix = np.array([[f, 0, self.resolution[0] / 2], [0, f, self.resolution[1] / 2], [0, 0, 1]])
        self._p = None
        self._pose = torch.zeros((self.num_envs, 12), dtype=torch.float64, device=self.device)
        self._has_pose = False

    @classmethod
    def from_params(cls, params, num_envs=1, device="cuda:0"):
        """The construction inside Drone.__init__, components.py:107-112."""
        c = params["camera"]
        return cls(camera_pitch_angle=c["camera_angle"], position_relative_to_frame=np.array(c["position_relative_to_frame"]),
                   fov=c["fov"], resolution=c["resolution"], focal_length=None, num_envs=num_envs, device=device)

    # components.py:472-477
    def convert_fov_to_focal_length(self, fov, resolution):
        return resolution[0] / (2 * np.tan(np.deg2rad(fov) / 2))

    def convert_focal_length_to_fov(self, focal_length, resolution):
        return np.rad2deg(2 * np.arctan(resolution[0] / (2 * focal_length)))

    def set_intrinsic_matrix(self, intrinsic_matrix):
        self.intrinsic_matrix = np.asarray(intrinsic_matrix, dtype=np.float64)
        self._p = None

    def _params(self) -> _lib.CameraParams:
        if self._p is None:
            if self.intrinsic_matrix is None:
                raise ValueError("the camera needs a focal length or a field of view")
            p = _lib.CameraParams()
            for i, v in enumerate(self.relative_rotation_matrix.reshape(-1)):
                p.rel_rot[i] = float(v)
            for i in range(3):
                p.rel_pos[i] = float(self.relative_position[i])
            K = self.intrinsic_matrix
            p.fx, p.fy, p.cx, p.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
            p.width, p.height = int(self.resolution[0]), int(self.resolution[1])
            self._p = p
        return self._p

    # ------------------------------------------------------------------ pose
    def update_from(self, drone):
        """Camera.update(drone.position, drone.rotation_matrix) (components.py:501-503) straight from the drone's
        device state (position + attitude quaternion planes); this is what Drone.step does at :245."""
        if drone.num_envs != self.num_envs:
            raise ValueError("camera and drone batch sizes differ")
        _lib.check(self._lib.fpv_camera_update(self._params(), _lib.ptr(drone._state), self.num_envs, drone._stride,
                                               _lib.ptr(self._pose), _lib.current_stream(self.device)))
        self._has_pose = True

    def update(self, drone_position, drone_rotation_matrix):
        """components.py:501-503 from explicit poses ([n,3], [n,3,3]; NumPy arrays or torch tensors on any device, e.g.
        `camera.update(drone.position, drone.rotation_matrix)` with BatchedDrone's CUDA tensors): one launch of
        fpv_camera_update_pose, float64 arithmetic like the reference."""
        def dev64(x, shape):
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
            return t.to(self.device, torch.float64).reshape(shape).contiguous()
        pos, R = dev64(drone_position, (self.num_envs, 3)), dev64(drone_rotation_matrix, (self.num_envs, 9))
        _lib.check(self._lib.fpv_camera_update_pose(self._params(), _lib.ptr(pos), _lib.ptr(R), self.num_envs, _lib.ptr(self._pose),
                                                    _lib.current_stream(self.device)))
        self._has_pose = True

    reset = update                                                                                   # :497-499

    @property
    def position(self):
        return self._pose[:, 9:]

    @property
    def rotation_matrix(self):
        return self._pose[:, :9].reshape(self.num_envs, 3, 3)

    @property
    def projection_matrix(self):
        """components.py:532-536: K [R t]^-1 (rows 0..2), [n,3,4] float64."""
        K = torch.as_tensor(self.intrinsic_matrix, dtype=torch.float64, device=self.device)
        Rt = self.rotation_matrix.transpose(1, 2)
        return K @ torch.cat([Rt, -(Rt @ self.position[:, :, None])], dim=2)

    def _need_pose(self):
        if not self._has_pose:
            raise RuntimeError("call update()/update_from() first (the reference's camera pose is None until reset)")

    def pixel2direction(self, pixel, ref_frame="world"):
        """components.py:505-526.  pixel [n,2] -> unit vectors [n,3] (float64)."""
        if ref_frame not in _FRAMES:
            raise ValueError("ref_frame must be world, drone or camera")
        self._need_pose()
        px = torch.as_tensor(np.asarray(pixel.cpu()) if isinstance(pixel, torch.Tensor) else np.asarray(pixel),
                             dtype=torch.float64).to(self.device).reshape(self.num_envs, 2).contiguous()
        out = torch.empty((self.num_envs, 3), dtype=torch.float64, device=self.device)
        _lib.check(self._lib.fpv_camera_rays(self._params(), _lib.ptr(self._pose), self.num_envs, _lib.ptr(px),
                                             _FRAMES[ref_frame], _lib.ptr(out), _lib.current_stream(self.device)))
        return out

    # ------------------------------------------------------------------ images
    def _world(self, objects_list):
        return objects_list if isinstance(objects_list, World) else World(objects_list, self.device)

    def _offsets(self, w, offsets):
        if offsets is None:
            return None
        o = torch.as_tensor(offsets, dtype=torch.float64, device=self.device)
        return o.reshape(self.num_envs, w.n_objects, 3).contiguous()

    def _render(self, objects_list, max_depth, offsets):
        self._need_pose()
        w = self._world(objects_list)
        n, W, H = self.num_envs, int(self.resolution[0]), int(self.resolution[1])
        img = torch.empty((n, H, W), dtype=torch.uint8, device=self.device)
        keep = torch.empty((n, w.n_objects), dtype=torch.uint8, device=self.device)
        off = self._offsets(w, offsets)
        _lib.check(self._lib.fpv_camera_render(self._params(), _lib.ptr(self._pose), n, _lib.ptr(w.points), w.n_points,
                                               _lib.ptr(w.boxes), w.n_objects, _lib.ptr(off), float(max_depth),
                                               _lib.ptr(keep), _lib.ptr(img), _lib.current_stream(self.device)))
        self.image = img
        return img

    def render_depth_image(self, objects_list, max_depth=10, offsets=None):
        """components.py:614-629 for every env: uint8 [n, H, W], nearest surface brightest, background 0."""
        if not max_depth > 0:
            raise ValueError("max_depth must be positive")
        return self._render(objects_list, max_depth, offsets)

    def render_image(self, objects_list, offsets=None):
        """components.py:601-612: binary splat, uint8 [n, H, W]."""
        return self._render(objects_list, 0.0, offsets)

    def target_pixel(self, objects_list, max_depth=10, offsets=None):
        """simulator.py:102-108: `np.where(render_depth_image(objects) > 0)` averaged and reversed to (x, y), without
        writing the image.  Returns (pixel float64 [n,2], seen uint8 [n])."""
        self._need_pose()
        w = self._world(objects_list)
        n = self.num_envs
        pixel = torch.empty((n, 2), dtype=torch.float64, device=self.device)
        seen = torch.empty(n, dtype=torch.uint8, device=self.device)
        off = self._offsets(w, offsets)
        _lib.check(self._lib.fpv_camera_target_pixel(self._params(), _lib.ptr(self._pose), n, _lib.ptr(w.points), w.n_points,
                                                     _lib.ptr(w.boxes), w.n_objects, _lib.ptr(off), float(max_depth),
                                                     _lib.ptr(pixel), _lib.ptr(seen), _lib.current_stream(self.device)))
        return pixel, seen
