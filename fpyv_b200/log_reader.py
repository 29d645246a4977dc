"""Blackbox-log front-end: recorded RC sticks of a real flight -> the raw-axis block `Joystick.replay` / `BatchedDrone.rollout`
take (sim-to-real checks through the `calib_read` path).

Mirror of the reference's `utils/log_reader.py:6-20`.  There, `blackbox_parser(path)` decodes a Betaflight / iNav `.BBL`
file with the third-party `orangebox` parser into a pandas DataFrame with one column per logged field and one row per
frame.  `orangebox` is not part of this image, so `blackbox_parser` here

  * reads the DECODED form of the same log -- the CSV that Betaflight's `blackbox_decode` (or orangebox's own export)
    writes: a header row of field names (`loopIteration, time, ..., rcCommand[0], rcCommand[1], rcCommand[2],
    rcCommand[3], ...`) followed by one row per frame -- and returns the same DataFrame shape, and
  * falls back to `orangebox` exactly like the reference when it IS importable and the file is a raw `.BBL` / `.BFL`.

`blackbox_sticks` then turns the four `rcCommand` columns into raw joystick axes [T, 6] in the range of a stick
calibration (the inverse of `Joystick.calib_read`, get_sticks.py:245-265), so that the replay goes through the very
calibration arithmetic a live radio goes through.  Betaflight conventions: rcCommand[0..2] = roll, pitch, yaw in
[-500, 500]; rcCommand[3] = throttle in [1000, 2000]; `time` in microseconds."""
from __future__ import annotations

import io
import os

import numpy as np

RC_FIELDS = ("rcCommand[0]", "rcCommand[1]", "rcCommand[2]", "rcCommand[3]")


def blackbox_parser(path):
    """log_reader.py:6-20 -> pandas DataFrame (columns = logged field names, one row per frame)."""
    import pandas as pd
    ext = os.path.splitext(path)[1].lower()
    if ext in (".bbl", ".bfl"):
        try:
            from orangebox import Parser     # what the reference uses (log_reader.py:2, :12)
        except ImportError as e:
            raise ImportError("raw .BBL logs need the third-party 'orangebox' package, as in the reference "
                              "(utils/log_reader.py:2); decode the log to CSV (blackbox_decode) and pass that") from e
        parser = Parser.load(path)
        rows = []
        for frame in parser.frames():
            row = np.full(len(parser.field_names), np.nan)
            row[:len(frame.data)] = frame.data
            rows.append(row)
        return pd.DataFrame(np.asarray(rows).reshape(-1, len(parser.field_names)), columns=parser.field_names)
    with open(path, newline="", encoding="utf-8", errors="replace") as f:
        lines = f.read().splitlines()
    # blackbox_decode writes the field-name row first; some exporters put `"key","value"` header pairs before it:
    # the field row is the first one that names the rcCommand columns
    start = next((i for i, ln in enumerate(lines) if "rcCommand[0]" in ln), None)
    if start is None:
        raise ValueError(f"{path}: no 'rcCommand[0]' column -- not a decoded blackbox log")
    df = pd.read_csv(io.StringIO("\n".join(lines[start:])), skipinitialspace=True)
    df.columns = [c.strip().strip('"') for c in df.columns]
    return df


def _time_column(data):
    for name in ("time", "time (us)", "time(us)"):
        if name in data.columns:
            return np.asarray(data[name], dtype=np.float64) * 1e-6
    return None


def blackbox_sticks(data, calibration, dt=None):
    """DataFrame of `blackbox_parser` -> (raw [T, 6] int32, t [T] seconds).

    raw is what the joystick driver would have reported for these stick positions under `calibration`
    (`config.StickCalibration`, i.e. the calibration JSON of the reference): axis order dwXpos..dwVpos, throttle / roll /
    pitch / yaw on the calibration's own axis indices, switches at their minimum; `Joystick.calib_read` maps it back to the
    logged stick values (to the 16-bit quantisation of the axis).  dt: resample to a fixed control period by zero-order
    hold (a receiver holds the last frame); None keeps the log's own frames."""
    missing = [f for f in RC_FIELDS if f not in data.columns]
    if missing:
        raise ValueError(f"blackbox table lacks {missing}")
    rc = np.stack([np.asarray(data[f], dtype=np.float64) for f in RC_FIELDS], axis=1)
    keep = np.isfinite(rc).all(axis=1)
    rc = rc[keep]
    t = _time_column(data)
    t = (t[keep] - t[keep][0]) if t is not None else None
    if dt is not None:
        if t is None:
            raise ValueError("resampling needs the log's 'time' column")
        grid = np.arange(0.0, t[-1] + 1e-12, float(dt))
        rc = rc[np.clip(np.searchsorted(t, grid, side="right") - 1, 0, len(t) - 1)]
        t = grid
    roll, pitch, yaw = (np.clip(rc[:, i] / 500.0, -1.0, 1.0) for i in range(3))
    throttle = np.clip((rc[:, 3] - 1500.0) / 500.0, -1.0, 1.0)
    sticks = {"Throttle": throttle, "Roll": roll, "Pitch": pitch, "Yaw": yaw}
    cal = calibration
    raw = np.tile(np.asarray(cal.min_vals, dtype=np.float64), (len(rc), 1))       # switches / unused axes at their minimum
    names = list(cal.sticks.keys())
    for s, name in enumerate(names):
        i, c = int(cal.stick_idx[s]), float(cal.stick_center[s])
        v = sticks[name]
        n = np.where(v <= 0.0, (v + 1.0) * (c + 1.0) - 1.0, v * (1.0 - c) + c)        # undo the re-centring (:260-263)
        n = n * float(cal.sign_reverse[i])                                            # sign_reverse is +-1: its own inverse
        raw[:, i] = cal.min_vals[i] + (n + 1.0) * 0.5 * (cal.max_vals[i] - cal.min_vals[i])   # undo mapFromTo (:245-252)
    return np.rint(raw).astype(np.int32), t


def load_blackbox_csv(path, calibration, dt=None):
    """Decoded blackbox CSV -> (raw [T, 6] int32, t [T]); `Joystick.replay(raw)` gives the [T, 1, 4] action block."""
    return blackbox_sticks(blackbox_parser(path), calibration, dt)
