"""The ctypes stub INTEGRATION.md section B shows a maintainer of the reference is executed as written: its struct layouts
against fpv_sizeof() on the CPU, and a reset + steps on the GPU against BatchedDrone (same library, same parameters)."""
import os
import re
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_namespace():
    from fpyv_b200 import _lib
    _lib.load()
    text = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()
    m = re.search(r"## B\. Raw ctypes stub.*?```python\n(.*?)```", text, re.S)
    assert m, "INTEGRATION.md lost its ctypes stub"
    src = m.group(1)
    assert '"libfpyv_b200.so"' in src
    src = src.replace('"libfpyv_b200.so"', repr(os.path.join(ROOT, "fpyv_b200", "libfpyv_b200.so")))
    ns = {}
    exec(compile(src, "INTEGRATION.md:B", "exec"), ns)
    return ns


def test_stub_struct_layouts_match_the_library():
    ns = stub_namespace()      # the stub asserts fpv_abi_version() and both sizeof()s while it is imported
    from fpyv_b200 import _lib
    import ctypes as C
    assert C.sizeof(ns["DroneParams"]) == C.sizeof(_lib.DroneParams)
    assert C.sizeof(ns["DroneIO"]) == C.sizeof(_lib.DroneIO)
    for mine, theirs in ((_lib.DroneParams, ns["DroneParams"]), (_lib.DroneIO, ns["DroneIO"])):
        assert [(n, getattr(mine, n).offset) for n, _ in mine._fields_] == [(n, getattr(theirs, n).offset) for n, _ in theirs._fields_]


@pytest.mark.gpu
def test_stub_steps_like_the_package(monkeypatch):
    import pandas as pd
    import torch
    from fpyv_b200 import BatchedDrone, config
    ns = stub_namespace()
    d = BatchedDrone(None, num_envs=1, device="cuda:0", substeps=1)
    c = d.constants
    # what the stub reads from a constructed reference Drone (components.py:84-142), rebuilt from our constants
    fake = types.SimpleNamespace(
        dt=c.dt, gravity=c.gravity, mass=c.mass, max_rates=c.max_rates, rates_transition_rate=c.rates_transition_rate,
        thrust_transition_rate=c.thrust_transition_rate, drag_coef=c.drag_coef, cross_section_areas=c.cross_section_areas,
        motors_relative_position=c.motors_relative_position, motor_radius=config.MOTOR_RADIUS, n_motors=config.N_MOTORS,
        motor_test_report=pd.DataFrame({"Throttle": c.throttle_percent,
                                        "Thrust": c.thrust_newton / config.N_MOTORS * 1000 / c.gravity}))
    utils = types.ModuleType("utils")
    ftc = types.ModuleType("utils.flight_time_calculator")
    ftc.model_xy = config.model_xy
    utils.flight_time_calculator = ftc
    monkeypatch.setitem(sys.modules, "utils", utils)
    monkeypatch.setitem(sys.modules, "utils.flight_time_calculator", ftc)
    s = ns["Drone"](fake)
    pos, vel, ypr = [0.3, -0.2, 1.5], [1.0, 0.5, -0.25], [10.0, -20.0, 35.0]
    s.reset(pos, vel, ypr)
    d.reset(pos, vel, ypr)
    rng = np.random.default_rng(5)
    for _ in range(20):
        a = rng.uniform(-1, 1, 4)
        Rt = s.step(a, [0.5, -0.25, 0.0], None)
        out = d.step(torch.tensor(a[None], dtype=torch.float32, device="cuda:0"), wind_velocity_vector=[0.5, -0.25, 0.0])
        np.testing.assert_array_equal(Rt, out[0][0].cpu().numpy())
    assert torch.equal(s.state, d._state)
