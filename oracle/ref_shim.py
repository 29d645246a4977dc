"""TEST INFRASTRUCTURE ONLY -- never imported by the product (`fpyv_b200/`).

Live oracle: imports the UNMODIFIED reference (omrijsharon/FpyV) from
/root/reference through a handful of stub modules, so that its own
`utils.components.Drone.step` (src/utils/components.py:220-248) and
`tests/racer_drone_test.Racer.step` (tests/racer_drone_test.py:95-103) can be
executed in this container.  It exists to (1) pin the CPU restatement in
`oracle/fpv_oracle.py` / `oracle/fpv_oracle.c` against the reference itself and
(2) generate the golden vectors under `tests/golden/` (see
`oracle/make_golden.py`).  It only works where `/root/reference` is mounted
(this container; NOT the GPU box), which is why the vectors are committed.

What has to be shimmed, and why:
  * matplotlib / mpl_toolkits / icosphere / drawnow: plotting deps that are
    absent here and never touched by the dynamics path.
  * utils.joystickapi: ctypes.WinDLL('winmm.dll') (src/utils/joystickapi.py:5)
    -- Windows-only; replaced by a fake device whose raw axes we can inject.
  * flight_time_calculator.read_motor_test_report: raises KeyError under
    pandas 3 (`.iloc[0][0]`, src/utils/flight_time_calculator.py:26).  The
    reader below follows :16-40 step by step; `model_xy` (:43-52) is used
    unmodified.
  * the two absolute Windows paths in config/params.yaml:39-40.
"""
from __future__ import annotations

import contextlib
import copy
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("FPYV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "utils", "components.py"))


class _FakeJoystickApi(types.ModuleType):
    """Stands in for src/utils/joystickapi.py (winmm.dll bindings)."""

    class _Caps:
        szPname = "fake-radio"
        wNumButtons = 0

    class _Info:
        def __init__(self, axes):
            (self.dwXpos, self.dwYpos, self.dwZpos,
             self.dwRpos, self.dwUpos, self.dwVpos) = axes
            self.dwButtons = 0

    def __init__(self):
        super().__init__("utils.joystickapi")
        self.raw_axes = [0, 0, 0, 0, 0, 0]

    def joyGetNumDevs(self):
        return 1

    def joyGetDevCaps(self, _id):
        return True, self._Caps()

    def joyGetPosEx(self, _id):
        return True, self._Info(self.raw_axes)


_STATE: dict = {}


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    class _Any:
        def __getattr__(self, _):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = mod("matplotlib")
            mpl.pyplot = mod("matplotlib.pyplot", __getattr__=lambda n: _Any())
            mpl.tri = mod("matplotlib.tri", __getattr__=lambda n: _Any())
            mpl.ticker = mod("matplotlib.ticker", MaxNLocator=_Any())
            mt = mod("mpl_toolkits")
            mt.mplot3d = mod("mpl_toolkits.mplot3d", __getattr__=lambda n: _Any())
    mod("icosphere", icosphere=lambda nu=1: (np.zeros((1, 3)), np.zeros((1, 3), dtype=int)))
    mod("drawnow", drawnow=lambda *a, **k: None)
    if "gym" not in sys.modules:
        try:
            import gym  # noqa: F401
        except Exception:
            g = mod("gym", Env=object)
            g.spaces = mod("gym.spaces", Dict=_Any(), Box=_Any(), Discrete=_Any())


def _read_motor_test_report(path):
    """Follows src/utils/flight_time_calculator.py:16-40 (pandas-3 safe)."""
    import pandas as pd
    rep = pd.read_csv(path, header=None, index_col=False, dtype=str)
    rep.columns = ['Type', 'Propeller', 'Throttle', 'Thrust', 'Voltage', 'Current', 'RPM', 'Power',
                   'Efficiency', 'Temperture']
    if rep.iloc[0, 0] == "Type":
        rep = rep.iloc[1:]
    rep = rep.copy()
    rep['Throttle'] = rep['Throttle'].str.replace('%', '').astype(float)
    rep['Thrust'] = rep['Thrust'].str.replace(',', '.').astype(float)
    rep['Power'] = rep['Power'].str.replace(',', '.').astype(float)
    out = []
    idx = np.append(0, np.append(rep[rep['Throttle'] == 100].index.values, len(rep)))
    for b, n in zip(idx[:-1], idx[1:]):
        out.append(rep[b:n])
    if len(out[-1]) == 0:
        del out[-1]
    return out


def load():
    """Import the reference modules once; returns a namespace dict."""
    if _STATE:
        return _STATE
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    _install_stubs()
    fake = _FakeJoystickApi()
    for p in (os.path.join(REF_ROOT, "tests"), REF_ROOT, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import utils  # noqa: F401  (the reference's src/utils package)
    sys.modules["utils.joystickapi"] = fake
    utils.joystickapi = fake
    with contextlib.redirect_stdout(io.StringIO()):
        from utils import flight_time_calculator as ftc
        ftc.read_motor_test_report = _read_motor_test_report
        from utils import components, kinematics, helper_functions, get_sticks, yaml_helper
        components.read_motor_test_report = _read_motor_test_report
        import racer_drone_test
    _STATE.update(components=components, kinematics=kinematics, helper_functions=helper_functions,
                  get_sticks=get_sticks, yaml_helper=yaml_helper, ftc=ftc,
                  racer=racer_drone_test, joystickapi=fake)
    return _STATE


def load_params(calib="frsky.json"):
    """config/params.yaml through the reference's own yaml_reader, paths rewritten."""
    ns = load()
    params = ns["yaml_helper"].yaml_reader(os.path.join(REF_ROOT, "config", "params.yaml"))
    params["drone"]["joystick_calib_path"] = os.path.join(REF_ROOT, "config", calib)
    params["drone"]["motor_test_report_path"] = os.path.join(REF_ROOT, "config", "t_motos_f80_motor_test.csv")
    return params


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_drone(params=None):
    ns = load()
    params = copy.deepcopy(params) if params is not None else load_params()
    with quiet():
        d = ns["components"].Drone(params)
    return d


def make_ground():
    ns = load()
    return ns["components"].Ground(size=60, resolution=2, random=False)


def ref_drone_rollout(actions, position, velocity, rpy_deg, wind=None, params=None,
                      dt=None, objects="ground", raw_axes=None):
    """Run the reference Drone for len(actions) steps; returns float64 trajectories.

    actions: [T,4] (ignored where raw_axes is given: then action=None -> joystick path).
    Returns dict of arrays with leading dim T (values AFTER each step).
    """
    ns = load()
    d = make_drone(params)
    if dt is not None:
        d.dt = float(dt)
    object_list = [make_ground()] if objects == "ground" else list(objects)
    wind = np.zeros(3) if wind is None else np.asarray(wind, dtype=np.float64)
    with quiet():
        d.reset(np.asarray(position, dtype=np.float64), np.asarray(velocity, dtype=np.float64),
                np.asarray(rpy_deg, dtype=np.float64))
    T = len(actions) if raw_axes is None else len(raw_axes)
    out = {k: [] for k in ("state", "R", "prev_rates", "prev_thrust", "done", "ret_Rt", "ret_gyro",
                           "ret_acc", "acc", "drag", "total_forces", "action")}
    for t in range(T):
        w = wind if wind.ndim == 1 else wind[t]
        with quiet():
            if raw_axes is not None:
                ns["joystickapi"].raw_axes = list(raw_axes[t])
                ret = d.step(None, w, object_list)
                out["action"].append(np.array([-d.rc.calib_reading[1], d.rc.calib_reading[2],
                                               d.rc.calib_reading[5], d.rc.calib_reading[0]]))
            else:
                a = np.asarray(actions[t], dtype=np.float64)
                ret = d.step(a, w, object_list)
                out["action"].append(a)
        out["state"].append(d.state.copy())
        out["R"].append(np.array(d.rotation_matrix, dtype=np.float64).copy())
        out["prev_rates"].append(np.array(d.prev_rates, dtype=np.float64).copy())
        out["prev_thrust"].append(float(d.prev_thrust))
        out["done"].append(bool(d.done))
        out["ret_Rt"].append(np.array(ret[0]).copy())
        out["ret_gyro"].append(np.array(ret[1]).copy())
        out["ret_acc"].append(np.array(ret[2]).copy())
        out["acc"].append(np.array(d.acceleration).copy())
        out["drag"].append(np.array(d.drag_force).copy())
        out["total_forces"].append(np.array(d.total_forces).copy())
    return {k: np.array(v) for k, v in out.items()}


def ref_racer_rollout(actions, pid_values, prop_size_inch=5):
    """Run tests/racer_drone_test.Racer for len(actions) steps (dt = 1e-3 module constant)."""
    ns = load()
    env = ns["racer"].Racer(prop_size_inch=prop_size_inch,
                            pid_values={k: np.asarray(v, dtype=np.float64) for k, v in pid_values.items()})
    env.reset()
    out = {k: [] for k in ("position", "velocity", "R", "omega", "torque")}
    for a in actions:
        with quiet():
            env.step(np.asarray(a, dtype=np.float64))
        out["position"].append(env.position.copy())
        out["velocity"].append(env.linear_velocity.copy())
        out["R"].append(env.orientation.as_matrix().copy())
        out["omega"].append(env.angular_velocity.copy())
        out["torque"].append(env.torque.copy())
    return {k: np.array(v) for k, v in out.items()}


def ref_calib_read(raw_axes, calib="frsky.json"):
    """Joystick.calib_read (src/utils/get_sticks.py:254-265) on injected raw axes -> [6]."""
    ns = load()
    with quiet():
        rc = ns["get_sticks"].Joystick()
        rc.calibrate(os.path.join(REF_ROOT, "config", calib), load_calibration_file=True)
        ns["joystickapi"].raw_axes = list(raw_axes)
        return np.array(rc.calib_read(), dtype=np.float64)
