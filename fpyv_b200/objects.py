"""Obstacle descriptions for `step(..., object_list=...)`.  Light-weight stand-ins for the geometry
classes of src/utils/components.py (`Ground` :649-683, `Cylinder` :685-744, `Target` :757-782, `Gate`
:784-830) carrying only what the dynamics path reads: the signed distance and the contact normal.
Objects coming from the reference itself are accepted too (duck-typed by class name / attributes)."""
from __future__ import annotations

import numpy as np

from . import _lib


class Ground:
    """Plane z = 0, normal +z (components.py:674-680)."""

    def __init__(self, size=None, resolution=None, random=False):
        self.size, self.resolution = size, resolution

    position = property(lambda self: np.zeros(3))


class Cylinder:
    """Upright cylinder: base centre `position`, `radius`, `height` (components.py:685-729)."""

    def __init__(self, position, radius, height, *_, **__):
        assert radius > 0, "radius must be positive"
        assert height > 0, "height must be positive"
        self.position = np.asarray(position, dtype=np.float64)
        self.radius, self.height = float(radius), float(height)


class Target:
    """Sphere: centre `position`, `radius` (components.py:757-777)."""

    def __init__(self, position, radius, *_, **__):
        self.position = np.asarray(position, dtype=np.float64)
        self.radius = float(radius)


class Gate:
    """Gate plane (components.py:784-822).  Gates never collide (handle_collisions skips them,
    components.py:203); the plane is used by the gate-race reward of fpyv_b200.env."""

    def __init__(self, position, rotation_matrix, size, shape="rectangle", resolution=17):
        self.position = np.asarray(position, dtype=np.float64)
        self.rotation_matrix = np.asarray(rotation_matrix, dtype=np.float64)
        self.size = float(size)
        self.shape = shape

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
