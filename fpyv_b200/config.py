"""Config front-end: reads the reference's three file formats unchanged (params.yaml, the joystick
calibration JSON, the T-Motor bench CSV) and derives the constants `Drone.__init__` derives
(reference: src/utils/components.py:84-142, src/utils/flight_time_calculator.py:16-52,
src/utils/yaml_helper.py:9-12, src/utils/get_sticks.py:126-133).  Host-side, init-time only."""
from __future__ import annotations

import csv
import json
import ntpath
import os
from dataclasses import dataclass

import numpy as np

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
DEFAULT_PARAMS = os.path.join(CONFIG_DIR, "params.yaml")

AIR_DENSITY = 1.2225            # kinematics.py:33 (default argument of calculate_drag)
N_MOTORS = 4                    # components.py:120
MOTOR_RADIUS = 0.1              # components.py:121
ARM_RADIUS = 5 * 2.54 / 100     # components.py:122
COLLISION_SPRING = 100.0        # components.py:198
COLLISION_DAMPING = 0.0         # components.py:198
MIN_THROTTLE_PERCENT = 5        # components.py:139


def yaml_reader(path):
    """Same contract as the reference's yaml_helper.yaml_reader: path -> nested dict."""
    import yaml
    with open(path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def resolve_path(p: str, search_dirs=()):
    """params.yaml carries absolute Windows paths from the author's machine (params.yaml:39-40).
    Use the path if it exists, else look the basename up in `search_dirs` and the packaged config dir."""
    if os.path.isfile(p):
        return p
    base = ntpath.basename(p.replace("/", "\\"))
    for d in (*search_dirs, CONFIG_DIR):
        cand = os.path.join(d, base)
        if os.path.isfile(cand):
            return cand
        cand = os.path.join(d, p)
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError(f"cannot resolve config path {p!r} (searched {[*search_dirs, CONFIG_DIR]})")


def load_params(path: str | None = None) -> dict:
    path = path or DEFAULT_PARAMS
    params = yaml_reader(path)
    params["__config_dir__"] = os.path.dirname(os.path.abspath(path))
    return params


@dataclass
class MotorBlock:
    """One throttle sweep of the bench report (11 rows, 50..100 %)."""
    throttle_percent: np.ndarray
    thrust_grams: np.ndarray
    motor: str
    propeller: str


def read_motor_test_report(path: str) -> list[MotorBlock]:
    """T-Motor bench CSV -> sweeps.  Format facts (flight_time_calculator.py:16-40): optional header
    row starting with 'Type'; throttle like '55%'; some numeric cells use a decimal comma and are
    therefore quoted ("13,67"); a sweep ends with its 100 % row."""
    sweeps: list[MotorBlock] = []
    thr: list[float] = []
    grams: list[float] = []
    names: list[str] = []
    props: list[str] = []
    with open(path, newline="", encoding="utf-8") as f:
        for cells in csv.reader(f):
            if len(cells) < 4 or cells[0].strip() == "Type":
                continue
            thr.append(float(cells[2].strip().rstrip("%")))
            grams.append(float(cells[3].strip().replace(",", ".")))
            if cells[0].strip():
                names.append(cells[0].strip())
            if cells[1].strip():
                props.append(cells[1].strip())
            if thr[-1] == 100.0:
                sweeps.append(MotorBlock(np.array(thr), np.array(grams), names[0] if names else "",
                                         props[0] if props else ""))
                thr, grams, names, props = [], [], [], []
    if thr:
        sweeps.append(MotorBlock(np.array(thr), np.array(grams), names[0] if names else "", props[0] if props else ""))
    return sweeps


def model_xy(x, y, degree=3, origin=True) -> np.poly1d:
    """Least-squares polynomial through the points (plus the origin), flight_time_calculator.py:43-52."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if origin:
        x = np.concatenate([[0.0], x])
        y = np.concatenate([[0.0], y])
    return np.poly1d(np.polyfit(x, y, degree))


@dataclass
class StickCalibration:
    """Joystick calibration JSON (get_sticks.py:126-133): per-axis raw min/max and sign, and for the
    four sticks their axis index and resting centre (in normalised [-1,1] units)."""
    min_vals: np.ndarray
    max_vals: np.ndarray
    sign_reverse: np.ndarray
    sticks: dict
    switches: dict

    @classmethod
    def load(cls, path: str) -> "StickCalibration":
        if not os.path.isfile(path):
            raise FileNotFoundError(f"Calibration file does not exist. Calibration path given: {path}")
        with open(path) as f:
            d = json.load(f)
        return cls(np.array(d["min_vals"], dtype=np.float64), np.array(d["max_vals"], dtype=np.float64),
                   np.array(d["sign_reverse"], dtype=np.float64), d["sticks"], d.get("switches", {}))

    @property
    def stick_idx(self):
        return [int(v["idx"]) for v in self.sticks.values()]

    @property
    def stick_center(self):
        return [float(v["center"]) for v in self.sticks.values()]


@dataclass
class DroneConstants:
    """What Drone.__init__ computes from params (components.py:84-142), in float64."""
    dt: float
    gravity: float
    mass: float
    max_rates: float
    rates_transition_rate: float
    thrust_transition_rate: float
    drag_coef: np.ndarray
    dimensions: np.ndarray
    cross_section_areas: np.ndarray
    motors_relative_position: np.ndarray
    throttle_percent: np.ndarray
    thrust_newton: np.ndarray
    thrust_poly: np.poly1d          # throttle % -> N
    throttle_poly: np.poly1d        # N -> throttle %
    min_throttle_in_force: float
    max_throttle_in_force: float
    motor_name: str
    propeller: str

    @property
    def k_drag(self):
        return -0.5 * self.drag_coef * AIR_DENSITY * self.cross_section_areas     # kinematics.py:36

    def throttle2thrust(self, x):
        return self.thrust_poly(100 * (np.asarray(x, dtype=np.float64) + 1) / 2)   # components.py:136

    def thrust2throttle(self, x):
        return np.clip(self.throttle_poly(np.asarray(x, dtype=np.float64)) / 100 * 2 - 1, -1, 1)  # :137


def derive_constants(params: dict, dt: float | None = None) -> DroneConstants:
    dr, sim = params["drone"], params["simulator"]
    search = (params.get("__config_dir__"),) if params.get("__config_dir__") else ()
    gravity = float(sim["gravity"])
    dims = np.array(dr["dimensions"], dtype=np.float64) / 100
    areas = np.array([dims[1] * dims[2], dims[0] * dims[2], dims[0] * dims[1]])
    ang = np.linspace(0, 2 * np.pi, N_MOTORS + 1)[:-1]
    ang = ang + (ang[1] - ang[0]) / 2
    motors = ARM_RADIUS * np.array([np.cos(ang), np.sin(ang), np.zeros(N_MOTORS)]).T
    sweep = read_motor_test_report(resolve_path(dr["motor_test_report_path"], search))[int(dr["motor_test_report_idx"])]
    thrust_n = N_MOTORS * sweep.thrust_grams / 1000 * gravity                      # components.py:133
    c = DroneConstants(
        dt=float(1 / sim["fps"]) if dt is None else float(dt), gravity=gravity, mass=dr["mass"] / 1000,
        max_rates=float(dr["max_rates"]), rates_transition_rate=float(dr["rates_transition_rate"]),
        thrust_transition_rate=float(dr["thrust_transition_rate"]),
        drag_coef=np.array(dr["drag_coefficients"], dtype=np.float64), dimensions=dims, cross_section_areas=areas,
        motors_relative_position=motors, throttle_percent=sweep.throttle_percent, thrust_newton=thrust_n,
        thrust_poly=model_xy(sweep.throttle_percent, thrust_n), throttle_poly=model_xy(thrust_n, sweep.throttle_percent),
        min_throttle_in_force=0.0, max_throttle_in_force=0.0, motor_name=sweep.motor, propeller=sweep.propeller)
    c.min_throttle_in_force = float(c.throttle2thrust(-1 + MIN_THROTTLE_PERCENT / 100 * 2))
    assert c.min_throttle_in_force > 0, (        # same error class as components.py:141, our own wording
        f"motor curve gives {c.min_throttle_in_force:.4f} N at the 5 % idle throttle: the fitted thrust must be positive there")
    c.max_throttle_in_force = float(c.throttle2thrust(1))
    return c


def thrust_table(c: DroneConstants, n: int = 2049, source: str = "poly") -> np.ndarray:
    """Shared-memory LUT for throttle -> thrust, sampled uniformly on throttle in [-1, 1].
    source='poly': samples the reference's cubic (interpolation error ~ (2/(n-1))^2/8 * |f''|);
    source='bench': piecewise-linear through the bench points and the origin (an EXTENSION: the
    reference never evaluates the raw table; parity unpinned)."""
    x = np.linspace(-1.0, 1.0, n)
    if source == "poly":
        return c.throttle2thrust(x).astype(np.float32)
    if source == "bench":
        pct = np.concatenate([[0.0], c.throttle_percent])
        th = np.concatenate([[0.0], c.thrust_newton])
        return np.interp(100 * (x + 1) / 2, pct, th).astype(np.float32)
    raise ValueError("source must be 'poly' or 'bench'")
