"""Batched stick front-end: `Joystick.calib_read` (src/utils/get_sticks.py:245-265) and
`Drone.read_sticks` (src/utils/components.py:250-253) on the GPU.  The Windows `winmm.dll` device
polling of the reference (src/utils/joystickapi.py) is out of scope; raw axis readings come from a
`source` (recorded logs, a radio bridge, a test) as int32 [n,6] in dwXpos..dwVpos order."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib, config


def pack_crsf(raw11):
    """[n,4] integer channel values 0..2047 (throttle, roll, pitch, yaw -- the resolution of a CRSF / SBUS link) -> uint8
    [n,6]: the four 11-bit channels packed little-endian, channel c in bits [11c, 11c+11) (fpv_api.h FPV_STICKS_CRSF)."""
    v = np.asarray(raw11.cpu() if isinstance(raw11, torch.Tensor) else raw11).astype(np.uint64) & 0x7FF
    bits = v[:, 0] | (v[:, 1] << np.uint64(11)) | (v[:, 2] << np.uint64(22)) | (v[:, 3] << np.uint64(33))
    out = np.empty((len(v), 6), dtype=np.uint8)
    for b in range(6):
        out[:, b] = (bits >> np.uint64(8 * b)) & np.uint64(0xFF)
    return torch.from_numpy(out)


def crsf_to_raw16(raw11):
    """The 16-bit driver reading an 11-bit channel value stands for: (v << 5) | (v >> 6) (0 -> 0, 2047 -> 65535)."""
    v = np.asarray(raw11.cpu() if isinstance(raw11, torch.Tensor) else raw11).astype(np.int64) & 0x7FF
    return (v << 5) | (v >> 6)


class Joystick:
    def __init__(self, source=None, device="cuda:0"):
        """source: callable returning raw axes [n,6] (tensor/array of ints) or None (set later with feed())."""
        self.device = torch.device(device)
        self.source = source
        self._raw = None
        self.calib = False
        self.calibration: config.StickCalibration | None = None
        self._c = None
        self.calib_reading = None

    # -- reference-compatible surface
    @property
    def status(self):
        return self.source is not None or self._raw is not None

    def calibrate(self, calibration_file_path, load_calibration_file=True):
        """get_sticks.py:101-137 with load_calibration_file=True; the interactive procedure needs the radio."""
        if not load_calibration_file:
            raise NotImplementedError("interactive calibration needs the physical radio (reference get_sticks.py:139-223)")
        if not os.path.isfile(calibration_file_path):
            raise FileNotFoundError(
                "Calibration file does not exist. Calibration path given: {}".format(calibration_file_path))
        self.load_calibration(calibration_file_path)

    def load_calibration(self, calibration_file_path):
        cal = config.StickCalibration.load(calibration_file_path)
        self.update(cal.sticks, cal.switches, cal.min_vals, cal.max_vals, cal.sign_reverse)

    def update(self, sticks, switches, min_vals, max_vals, sign_reverse):
        self.sticks, self.switches = sticks, switches
        self.min_vals, self.max_vals = np.asarray(min_vals, dtype=np.float64), np.asarray(max_vals, dtype=np.float64)
        self.sign_reverse = np.asarray(sign_reverse, dtype=np.float64)
        self.calibration = config.StickCalibration(self.min_vals, self.max_vals, self.sign_reverse, sticks, switches)
        c = _lib.StickCalib()
        for i in range(6):
            c.min_vals[i], c.max_vals[i], c.sign_reverse[i] = self.min_vals[i], self.max_vals[i], self.sign_reverse[i]
        for s, (idx, ctr) in enumerate(zip(self.calibration.stick_idx, self.calibration.stick_center)):
            c.stick_idx[s], c.stick_center[s] = idx, ctr
        self._c = c
        self.calib = True

    # -- batched data path
    def feed(self, raw):
        """Provide the next raw reading [n,6] (ints)."""
        self._raw = raw

    def read(self):
        if self.source is not None:
            self._raw = self.source()
        if self._raw is None:
            raise ModuleNotFoundError("no gamepad detected")   # same error class as get_sticks.py:34
        t = self._raw if isinstance(self._raw, torch.Tensor) else torch.as_tensor(np.asarray(self._raw))
        return t.to(self.device, torch.int32).reshape(-1, 6).contiguous()

    def _run(self, want_calibrated):
        if self._c is None:
            raise RuntimeError("Joystick is not calibrated: call calibrate(path) first")
        raw = self.read()
        n = raw.shape[0]
        actions = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        cal = torch.empty((n, 6), dtype=torch.float32, device=self.device) if want_calibrated else None
        lib = _lib.load()
        _lib.check(lib.fpv_sticks_to_actions(C.byref(self._c), _lib.ptr(raw), n, _lib.ptr(actions), _lib.ptr(cal),
                                             _lib.current_stream(self.device)))
        return actions, cal

    def replay(self, raw_log):
        """Recorded raw stick readings [T,n,6] (or [T,6] for one radio) -> actions [T,n,4] in ONE launch: the
        `calib_read` + `read_sticks` maps (get_sticks.py:254-265, components.py:250-253) applied to every logged frame,
        ready for `BatchedDrone.rollout`."""
        if self._c is None:
            raise RuntimeError("Joystick is not calibrated: call calibrate(path) first")
        t = raw_log if isinstance(raw_log, torch.Tensor) else torch.as_tensor(np.asarray(raw_log))
        if t.dim() == 2:
            t = t[:, None, :]
        if t.dim() != 3 or t.shape[2] != 6:
            raise ValueError("raw_log must be [T, n, 6] or [T, 6]")
        T, n = int(t.shape[0]), int(t.shape[1])
        raw = t.to(self.device, torch.int32).reshape(T * n, 6).contiguous()
        actions = torch.empty((T * n, 4), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().fpv_sticks_to_actions(C.byref(self._c), _lib.ptr(raw), T * n, _lib.ptr(actions), None,
                                                    _lib.current_stream(self.device)))
        return actions.reshape(T, n, 4)

    def calib_read(self):
        """[n,6] calibrated axes (get_sticks.py:254-265)."""
        _, cal = self._run(True)
        self.calib_reading = cal
        return cal

    def read_actions(self):
        """[n,4] = [-roll, pitch, yaw, throttle] (components.py:250-253)."""
        actions, _ = self._run(False)
        return actions
