"""Obstacle descriptions for `step(..., object_list=...)`.  Light-weight stand-ins for the geometry
classes of src/utils/components.py (`Ground` :649-683, `Cylinder` :685-744, `Target` :757-782, `Gate`
:784-830) carrying only what the dynamics path reads: the signed distance and the contact normal.
Objects coming from the reference itself are accepted too (duck-typed by class name / attributes)."""
from __future__ import annotations

import numpy as np

from . import _lib


def _bbox3d(points):
    """helper_functions.py:120-136: the 8 corners of the tight axis-aligned box around `points`."""
    lo, hi = points.min(axis=0), points.max(axis=0)
    box = np.zeros((8, 3))
    box[:4, 0], box[4:, 0] = lo[0], hi[0]
    box[::2, 1], box[1::2, 1] = lo[1], hi[1]
    box[[0, 1, 4, 5], 2], box[[2, 3, 6, 7], 2] = lo[2], hi[2]
    return box


def icosphere_vertices(nu=1):
    """Vertices of a geodesic icosahedron of subdivision frequency nu (10 nu^2 + 2 points on the unit sphere) -- the
    shape `icosphere.icosphere(nu)` gives the reference's Target (components.py:761); vertex order is ours."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    pts = []
    for a, b, c in f:
        for i in range(nu + 1):
            for j in range(nu + 1 - i):
                pts.append((v[a] * (nu - i - j) + v[b] * i + v[c] * j) / nu)
    pts = np.array(pts)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    _, idx = np.unique(np.round(pts, 9), axis=0, return_index=True)
    return pts[np.sort(idx)]


class Ground:
    """Plane z = 0, normal +z (components.py:674-680), with the point cloud the camera sees (:655-667)."""

    def __init__(self, size=None, resolution=None, random=False, rng=None):
        self.size, self.resolution = size, resolution
        self.points = None
        if size is not None and resolution is not None:
            self.points = self.generate_points(random, rng)

    def generate_points(self, random, rng=None):
        if random:                                                          # components.py:656-660
            rng = np.random.default_rng() if rng is None else rng
            pts = self.size * (2 * rng.random((self.resolution ** 2, 3)) - 1)
            pts[:, 2] /= self.size
            pts[:, 2] *= 0.2
            return pts
        axis = np.linspace(-self.size / 2, self.size / 2, self.resolution)   # :662-664
        x, y = np.meshgrid(axis, axis)
        return np.vstack([x.reshape(-1), y.reshape(-1), np.zeros(x.shape).reshape(-1)]).T

    position = property(lambda self: np.zeros(3))
    bbox3d = property(lambda self: _bbox3d(self.points))


class Cylinder:
    """Upright cylinder: base centre `position`, `radius`, `height` (components.py:685-729), with its surface
    point cloud (:697-708) when the two resolutions are given."""

    def __init__(self, position, radius, height, angle_resolution=None, height_resolution=None, random=False, rng=None):
        # same error class as the reference (AssertionError, components.py:688-689), our own wording
        assert radius > 0, f"Cylinder: radius {radius} is not > 0"
        assert height > 0, f"Cylinder: height {height} is not > 0"
        self.position = np.asarray(position, dtype=np.float64)
        self.radius, self.height = float(radius), float(height)
        self.angle_resolution, self.height_resolution = angle_resolution, height_resolution
        self.points = None
        if angle_resolution is not None and height_resolution is not None:
            self.points = self.position + self.generate_points(random, rng)

    def generate_points(self, random, rng=None):
        if random:                                                          # components.py:698-700
            rng = np.random.default_rng() if rng is None else rng
            angles = rng.random((self.height_resolution, self.angle_resolution)) * 2 * np.pi
            heights = rng.random((self.height_resolution, self.angle_resolution)) * self.height
        else:                                                               # :702-704
            angles = np.linspace(0, 2 * np.pi, self.angle_resolution)
            heights = np.linspace(0, self.height, self.height_resolution)
            angles, heights = np.meshgrid(angles, heights)
        return np.vstack([self.radius * np.cos(angles).reshape(-1), self.radius * np.sin(angles).reshape(-1),
                          heights.reshape(-1)]).T

    bbox3d = property(lambda self: _bbox3d(self.points))


def generate_circular_path(center, radius, resolution):
    """helper_functions.py:151-153: `resolution` way-points of a horizontal circle around `center`."""
    theta = np.linspace(0, 2 * np.pi, resolution + 1)[:-1]
    return np.vstack((np.cos(theta) * radius, np.sin(theta) * radius, np.zeros_like(theta))).T + np.array(center)


class CircularPath:
    """components.py:743-752: an endless iterator over the way-points of a circle."""

    def __init__(self, center, radius, resolution):
        self.path = generate_circular_path(center, radius, resolution)
        self.count = 0

    def __iter__(self):
        while True:
            yield self.path[self.count % len(self.path)]
            self.count += 1


class Target:
    """Sphere: centre `position`, `radius` (components.py:757-777); `vertices` (unit-sphere points scaled by the
    radius, :761-763) default to the geodesic icosahedron of frequency `nu`.  `path` = dict(radius=, resolution=) makes it
    a moving target exactly like the reference (:764-765): every `update()` (:769-771) steps to the next way-point of a
    `CircularPath` around the initial position.  `offset` is the displacement from the initial position -- the value a
    `World` built at construction time needs per env (`BatchedCamera.render_*(..., offsets=)`, fpv_camera_render's
    obj_offset); `target_offsets()` below packs it for a batch."""

    def __init__(self, position, radius, nu=None, path=None, vertices=None):
        self.position = np.asarray(position, dtype=np.float64)
        self.initial_position = self.position.copy()
        self.radius = float(radius)
        self.vertices = None
        if vertices is not None:
            self.vertices = np.asarray(vertices, dtype=np.float64) * self.radius
        elif nu is not None:
            self.vertices = icosphere_vertices(int(nu)) * self.radius
        self.path = None
        self._way_points, self._n_updates = None, 0
        if path is not None:
            cp = CircularPath(self.position, **path)
            self._way_points = cp.path
            self.path = iter(cp)

    @property
    def points(self):                                                       # components.py:766-768
        return None if self.vertices is None else self.vertices + self.position

    def update(self):                                                       # components.py:769-771
        if self.path is None:
            raise TypeError("'NoneType' object is not an iterator")         # what next(None) raises in the reference
        self.position = np.asarray(next(self.path), dtype=np.float64)
        self._n_updates += 1

    @property
    def offset(self):
        return self.position - self.initial_position

    bbox3d = property(lambda self: _bbox3d(self.points))

    def calculate_distance(self, point):                                    # :773-774
        return float(np.linalg.norm(np.asarray(point, dtype=np.float64) - self.position) - self.radius)


def target_offsets(objects_list, num_envs, per_env_phase=None):
    """obj_offset block [num_envs, n_objects, 3] (float64) for a world whose `Target`s have moved since the `World` was
    built: object j's current displacement, the same for every env -- or, with per_env_phase (int [num_envs] >= 0), env e
    sees every moving target as it will be after `per_env_phase[e]` MORE update() calls (a batch of chase envs started at
    different times).  Static objects get zeros."""
    out = np.zeros((num_envs, len(objects_list), 3))
    for j, o in enumerate(objects_list):
        if not isinstance(o, Target) or o.path is None:
            continue
        if per_env_phase is None:
            out[:, j] = o.offset
        else:
            total = o._n_updates + np.asarray(per_env_phase, dtype=np.int64)      # update() calls seen by env e
            way = o._way_points[(total - 1) % len(o._way_points)] - o.initial_position
            out[:, j] = np.where((total > 0)[:, None], way, 0.0)
    return out


class Gate:
    """Gate plane (components.py:784-822) and its outline polygon (:787-805).  Gates never collide
    (handle_collisions skips them, components.py:203); the plane is used by the gate-race reward of fpyv_b200.env."""

    def __init__(self, position, rotation_matrix, size, shape="rectangle", resolution=17):
        self.position = np.asarray(position, dtype=np.float64)
        self.rotation_matrix = np.asarray(rotation_matrix, dtype=np.float64)
        self.size = float(size)
        self.shape = shape
        size = self.size
        if shape == "rectangle":
            corners = np.array([[0, -1, -1], [0, 1, -1], [0, 1, 1], [0, -1, 1]]) * size / 2
        elif "circle" in shape:
            coef = 1 if "half" in shape else 2
            theta = np.linspace(0, coef * np.pi, resolution)
            corners = np.vstack((np.zeros_like(theta), np.cos(theta) * size / coef, np.sin(theta) * size / coef)).T
            if "half" in shape:
                corners = corners - np.array([0, 0, size / 2])
        else:
            raise NotImplementedError
        corners = (self.rotation_matrix @ corners.T).T + self.position
        self.corners = np.vstack((corners, corners[0]))

    points = property(lambda self: self.corners)
    bbox3d = property(lambda self: _bbox3d(self.points))

    @property
    def normal(self):
        return self.rotation_matrix[:, 0]

    def calculate_distance(self, point):
        return float(np.dot(self.normal, point) - np.dot(self.normal, self.position))


class Trail:
    """Placeholder so object lists written for the reference parse; ignored like components.py:203."""


def lower_object_list(object_list):
    """object_list -> (has_ground, [fpv_object_t...]) in list order (the order matters for the reference's
    early return on a crash, components.py:205-210).  The ground plane is evaluated last by the kernel, which
    is where simulator.py:84-85 puts it."""
    has_ground, out = False, []
    for i, o in enumerate(object_list):
        name = type(o).__name__
        if name in ("Gate", "Trail"):
            continue
        if name == "Ground":
            has_ground = True
            if i != len(object_list) - 1 and any(type(x).__name__ not in ("Gate", "Trail") for x in object_list[i + 1:]):
                raise ValueError("the ground plane must be the last colliding object of object_list "
                                 "(as in simulator.py:84-85)")
            continue
        pos = np.asarray(o.position, dtype=np.float64).reshape(3)
        if hasattr(o, "height"):
            out.append(_lib.Object(_lib.OBJ_CYLINDER, pos[0], pos[1], pos[2], float(o.radius), float(o.height)))
        elif hasattr(o, "radius"):
            out.append(_lib.Object(_lib.OBJ_SPHERE, pos[0], pos[1], pos[2], float(o.radius), 0.0))
        else:
            raise TypeError(f"object_list[{i}]: unsupported obstacle type {name}")
    if len(out) > _lib.MAX_OBJECTS:
        raise ValueError(f"at most {_lib.MAX_OBJECTS} obstacles per launch (got {len(out)})")
    return has_ground, out


def generate_track(count, radius, gate_size, gate_resolution=17):
    """Gate layout of the reference's generators.generate_track (src/utils/generators.py:7-18): gates on the
    ellipse [cos(t)*gate_size, sin(t)*radius, 0], yaw t + pi/2, shapes cycling rectangle / circle / half_circle.
    Reproduced as written, including its argument quirk: circle gates are raised by gate_size/2 and get size
    gate_size/2, while rectangle and half-circle gates receive `gate_resolution` as their size (:15-16)."""
    theta = np.linspace(0, 2 * np.pi, count + 1)[:-1]
    shapes = ["rectangle", "circle", "half_circle"]
    gates = []
    for i, t in enumerate(theta):
        p = np.array([np.cos(t) * gate_size, np.sin(t) * radius, 0.0])
        yaw = t + np.pi / 2
        rot = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
        if shapes[i % 3] == "circle":
            gates.append(Gate(p + np.array([0.0, 0.0, gate_size / 2]), rot, gate_size / 2, shape="circle", resolution=gate_resolution))
        else:
            gates.append(Gate(p, rot, gate_resolution, shape=shapes[i % 3], resolution=gate_resolution))
    return gates
