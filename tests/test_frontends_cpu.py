"""CPU-side front-ends added in round 2: the blackbox-log reader (reference utils/log_reader.py:6-20), the moving-target
mirrors (`CircularPath`, `Target.update`, components.py:743-771, helper_functions.py:151-153) and the CRSF stick packing.
The stick values recovered from a log are checked against the ORACLE's restatement of `Joystick.calib_read`
(oracle/fpv_oracle.py, pinned to the reference's own get_sticks.py by tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest

from conftest import CONFIG
from oracle import fpv_oracle as fo


def write_blackbox_csv(path, t_us, rc, header_pairs=True):
    """A decoded blackbox log as blackbox_decode writes it: optional "key","value" header pairs, then the field row."""
    with open(path, "w") as f:
        if header_pairs:
            f.write('"Product","Blackbox flight data recorder by Nicholas Sherlock"\n"Firmware revision","Betaflight 4.3.0"\n')
        f.write("loopIteration, time, axisP[0], rcCommand[0], rcCommand[1], rcCommand[2], rcCommand[3], gyroADC[0]\n")
        for i, (t, r) in enumerate(zip(t_us, rc)):
            f.write(f"{i}, {int(t)}, 3, {int(r[0])}, {int(r[1])}, {int(r[2])}, {int(r[3])}, -7\n")


@pytest.mark.parametrize("calib", ["frsky.json", "calibration.json"])
def test_blackbox_log_round_trip_through_the_calibration(tmp_path, calib):
    from fpyv_b200 import config, log_reader
    rng = np.random.default_rng(4)
    T = 400
    t_us = 1_000_000 + np.arange(T) * 2000 + rng.integers(0, 50, T)          # ~500 Hz frames with jitter
    rc = np.stack([rng.integers(-500, 501, T), rng.integers(-500, 501, T), rng.integers(-500, 501, T),
                   rng.integers(1000, 2001, T)], 1)
    rc[0] = [0, 0, 0, 1500]
    rc[1] = [500, -500, 500, 2000]
    rc[2] = [-500, 500, -500, 1000]
    p = tmp_path / "log.csv"
    write_blackbox_csv(p, t_us, rc)
    data = log_reader.blackbox_parser(str(p))
    assert list(data.columns)[:4] == ["loopIteration", "time", "axisP[0]", "rcCommand[0]"] and len(data) == T
    cal = config.StickCalibration.load(os.path.join(CONFIG, calib))
    raw, t = log_reader.blackbox_sticks(data, cal)
    assert raw.shape == (T, 6) and raw.dtype == np.int32 and t[0] == 0.0 and abs(t[-1] - (t_us[-1] - t_us[0]) * 1e-6) < 1e-12
    # the oracle's calib_read + read_sticks (get_sticks.py:245-265, components.py:250-253) must give the logged sticks back
    ocal = fo.StickCalib.from_json(os.path.join(CONFIG, calib))
    act = fo.sticks_to_action(ocal, raw)
    want = np.stack([-rc[:, 0] / 500.0, rc[:, 1] / 500.0, rc[:, 2] / 500.0, (rc[:, 3] - 1500.0) / 500.0], 1)
    span = float(np.min(np.abs(cal.max_vals - cal.min_vals)[[0, 1, 2, 5]]))
    assert np.max(np.abs(act - want)) <= 2.5 / span + 1e-12          # the 16-bit axis quantisation (re-centring stretches it)
    # resampling by zero-order hold to a fixed control period
    raw2, t2 = log_reader.blackbox_sticks(data, cal, dt=0.01)
    assert np.allclose(np.diff(t2), 0.01) and len(raw2) == len(t2)
    k = np.searchsorted((t_us - t_us[0]) * 1e-6, t2, side="right") - 1
    assert np.array_equal(raw2, raw[k])
    # load_blackbox_csv = parser + sticks; a table without the rc columns is refused
    raw3, _ = log_reader.load_blackbox_csv(str(p), cal)
    assert np.array_equal(raw3, raw)
    with pytest.raises(ValueError):
        log_reader.blackbox_sticks(data.drop(columns=["rcCommand[3]"]), cal)
    bad = tmp_path / "bad.csv"
    bad.write_text("a,b\n1,2\n")
    with pytest.raises(ValueError):
        log_reader.blackbox_parser(str(bad))
    with pytest.raises(ImportError):            # raw .BBL needs orangebox exactly like the reference
        log_reader.blackbox_parser(str(tmp_path / "x.BBL"))


def test_circular_path_and_moving_target_match_the_reference_formulas():
    from fpyv_b200.objects import CircularPath, Target, generate_circular_path, target_offsets
    c, r, res = np.array([1.0, -2.0, 3.0]), 4.0, 24
    path = generate_circular_path(c, r, res)
    theta = np.linspace(0, 2 * np.pi, res + 1)[:-1]                         # helper_functions.py:151-153
    assert np.allclose(path, np.stack([np.cos(theta) * r, np.sin(theta) * r, np.zeros(res)], 1) + c, atol=0, rtol=0)
    it = iter(CircularPath(c, r, res))
    seq = [next(it) for _ in range(res + 3)]
    assert all(np.array_equal(seq[i], path[i % res]) for i in range(res + 3))  # endless, wraps (components.py:748-752)
    t = Target(c, 0.5, nu=2, path=dict(radius=r, resolution=res))
    v0 = t.points.copy()
    for i in range(5):
        t.update()                                                          # components.py:769-771
        assert np.array_equal(t.position, path[i]) and np.allclose(t.points, v0 - c + path[i])
    assert np.array_equal(t.offset, path[4] - c)
    off = target_offsets([t, Target([0, 0, 0], 1.0, nu=1)], 3, per_env_phase=[0, 1, res])
    assert np.array_equal(off[0, 0], path[4] - c) and np.array_equal(off[1, 0], path[5] - c) and np.array_equal(off[2, 0], path[4] - c)
    assert not off[:, 1].any()
    with pytest.raises(TypeError):
        Target(c, 0.5, nu=1).update()


def test_crsf_packing_round_trip():
    from fpyv_b200.sticks import crsf_to_raw16, pack_crsf
    rng = np.random.default_rng(0)
    v = rng.integers(0, 2048, (1000, 4))
    v[0], v[1] = [0, 0, 0, 0], [2047, 2047, 2047, 2047]
    b = pack_crsf(v).numpy().astype(np.uint64)
    bits = sum(b[:, i] << np.uint64(8 * i) for i in range(6))
    back = np.stack([(bits >> np.uint64(11 * c)) & np.uint64(0x7FF) for c in range(4)], 1)
    assert np.array_equal(back.astype(np.int64), v) and (bits >> np.uint64(44) == 0).all()
    r = crsf_to_raw16(v)
    assert r.min() == 0 and r.max() == 65535 and np.all(np.diff(crsf_to_raw16(np.arange(2048))) > 0)
