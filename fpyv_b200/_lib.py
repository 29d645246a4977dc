"""ctypes binding of libfpyv_b200.so (include/fpv_api.h).  There is NO fallback: if the CUDA library is
missing, stale against the header, or cannot launch on the device, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfpyv_b200.so")
ABI_VERSION = 13

# flags (fpv_api.h)
F_GROUND, F_AUTO_RESET, F_FREEZE_DONE, F_THRUST_LUT, F_SCALAR, F_CHAINED, F_RATE_CURVE = 1, 2, 4, 8, 32, 64, 128
OBJ_SPHERE, OBJ_CYLINDER = 1, 2
MAX_OBJECTS = 16
DRONE_PLANES, RACER_PLANES = 4, 5
EXPORTS = ("fpv_abi_version", "fpv_last_error", "fpv_sizeof", "fpv_device_info", "fpv_drone_reset",
           "fpv_drone_step", "fpv_drone_step_host", "fpv_drone_step_host_sticks", "fpv_drone_rollout", "fpv_drone_observe", "fpv_drone_get_rotation", "fpv_drone_set_rotation", "fpv_matrix_to_quat", "fpv_sticks_to_actions", "fpv_racer_reset", "fpv_racer_step", "fpv_racer_observe", "fpv_gate_env_reset", "fpv_gate_env_step", "fpv_gate_race_step",
           "fpv_camera_update", "fpv_camera_update_pose", "fpv_camera_render", "fpv_camera_target_pixel", "fpv_camera_rays", "fpv_autopilot", "fpv_point_and_shoot",
           "fpv_acro_reset", "fpv_acro_step", "fpv_acro_rollout", "fpv_probe_fp32", "fpv_host_alloc", "fpv_host_free")


class FpvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfpyv_b200 error {code}: {msg}")
        self.code = code


class DroneParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("substeps", C.c_int32), ("gravity", C.c_float), ("mass", C.c_float),
                ("max_rates", C.c_float), ("rates_transition_rate", C.c_float),
                ("thrust_transition_rate", C.c_float), ("k_drag", C.c_float * 3),
                ("motor_xy", (C.c_float * 2) * 4), ("motor_radius", C.c_float), ("spring_k", C.c_float),
                ("spring_c", C.c_float), ("thrust_poly", C.c_float * 4), ("wind", C.c_float * 3),
                ("flags", C.c_uint32), ("n_objects", C.c_int32)]


class Object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("a", C.c_float),
                ("b", C.c_float)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("env_steps", "crashes", "episodes", "episode_len_sum", "reward_sum",
                                          "reward_sq_sum", "nonfinite", "reserved")]


STICKS_U16, STICKS_CRSF = 1, 2


class StickCalib(C.Structure):
    _fields_ = [("min_vals", C.c_float * 6), ("max_vals", C.c_float * 6), ("sign_reverse", C.c_float * 6),
                ("stick_idx", C.c_int32 * 4), ("stick_center", C.c_float * 4)]


class DroneIO(C.Structure):
    _fields_ = [("state", C.c_void_p), ("n", C.c_int64), ("plane_stride", C.c_int64), ("actions", C.c_void_p),
                ("sticks", C.c_void_p), ("stick_calib", C.POINTER(StickCalib)), ("stick_format", C.c_int32),
                ("wind_env", C.c_void_p), ("lut", C.c_void_p), ("lut_n", C.c_int32), ("done", C.c_void_p),
                ("done_bits", C.c_void_p), ("acc_out", C.c_void_p), ("reset_state", C.c_void_p), ("override_q", C.c_void_p),
                ("override_thrust", C.c_void_p),
                ("objects", C.POINTER(Object)), ("stats", C.c_void_p), ("work", C.c_void_p), ("chunk_epoch", C.c_void_p),
                ("epoch", C.c_uint32), ("max_ctas_per_sm", C.c_uint32), ("trace", C.c_void_p)]


class RacerParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("substeps", C.c_int32), ("mass", C.c_float), ("inertia", C.c_float * 3),
                ("gains", (C.c_float * 3) * 3), ("vel_decay", C.c_float), ("flags", C.c_uint32)]


MAX_GATES, ENV_OBS_FLOATS = 32, 16


class Gate(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("cx", "cy", "cz", "nx", "ny", "nz", "half_size", "pad")]


class GateEnvParams(C.Structure):
    _fields_ = [("n_gates", C.c_int32), ("agents_per_env", C.c_int32), ("laps_to_finish", C.c_int32),
                ("w_gate", C.c_float), ("w_progress", C.c_float), ("w_crash", C.c_float), ("gates", Gate * MAX_GATES)]


CAM_MAX_OBJECTS = 64


class CameraParams(C.Structure):
    _fields_ = [("rel_rot", C.c_double * 9), ("rel_pos", C.c_double * 3), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]


class AutopilotParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("mass", "dt", "virtual_drag_coef", "virtual_lift_coef", "tof_effective_dist",
                                          "keep_distance", "uwb_max_range", "kP", "kI", "kD", "integral_clip",
                                          "min_output", "max_output", "derivative_transition_rate")] + \
               [("ref_frame", C.c_int32), ("mode", C.c_int32), ("max_throttle_force", C.c_double),
                ("max_limit_iterations", C.c_int32), ("reserved", C.c_int32)]


ACRO_PLANES = 7


class AcroParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("substeps", C.c_int32), ("gravity", C.c_float), ("mass", C.c_float),
                ("max_rates", C.c_float), ("rates_transition_rate", C.c_float), ("thrust_transition_rate", C.c_float),
                ("k_drag", C.c_float * 3), ("motor_xy", (C.c_float * 2) * 4), ("motor_radius", C.c_float),
                ("spring_k", C.c_float), ("gains", (C.c_float * 3) * 3), ("integral_limit", C.c_float),
                ("inertia", C.c_float * 3), ("kappa", C.c_float), ("spin", C.c_float * 4), ("u_min", C.c_float),
                ("u_max", C.c_float), ("thrust_poly", C.c_float * 4), ("wind", C.c_float * 3), ("flags", C.c_uint32),
                ("rate_curve", (C.c_float * 3) * 3)]


_STRUCTS = (DroneParams, DroneIO, Object, Stats, StickCalib, RacerParams, GateEnvParams, CameraParams, AutopilotParams,
            AcroParams)
_lib = None


def load():
    """dlopen the library once, declare prototypes, verify ABI version and struct sizes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m fpyv_b200.build` "
                          "(fpyv_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it")
    lib.fpv_abi_version.restype = C.c_int
    lib.fpv_last_error.restype = C.c_char_p
    lib.fpv_sizeof.argtypes = [C.c_int]
    lib.fpv_device_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.fpv_drone_reset.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
    lib.fpv_drone_step.argtypes = [C.POINTER(DroneParams), C.POINTER(DroneIO), C.c_void_p]
    lib.fpv_drone_step_host.argtypes = [C.POINTER(DroneParams), C.POINTER(DroneIO), C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_void_p]
    lib.fpv_drone_step_host_sticks.argtypes = [C.POINTER(DroneParams), C.POINTER(DroneIO), C.POINTER(StickCalib), C.c_void_p,
                                               C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.fpv_host_alloc.argtypes = [C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.fpv_host_free.argtypes = [C.c_void_p]
    lib.fpv_drone_rollout.argtypes = [C.POINTER(DroneParams), C.POINTER(DroneIO), C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_int64, C.c_void_p]
    lib.fpv_drone_observe.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    lib.fpv_drone_get_rotation.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fpv_drone_set_rotation.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fpv_matrix_to_quat.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fpv_sticks_to_actions.argtypes = [C.POINTER(StickCalib), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    lib.fpv_racer_reset.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fpv_racer_step.argtypes = [C.POINTER(RacerParams), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fpv_racer_observe.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fpv_gate_env_reset.argtypes = [C.POINTER(GateEnvParams), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    lib.fpv_gate_env_step.argtypes = [C.POINTER(GateEnvParams), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    P, V, I64, I32, D = C.POINTER, C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.fpv_gate_race_step.argtypes = [P(DroneParams), P(DroneIO), P(GateEnvParams), V, V, V, V, V, V, V]
    lib.fpv_camera_update.argtypes = [P(CameraParams), V, I64, I64, V, V]
    lib.fpv_camera_update_pose.argtypes = [P(CameraParams), V, V, I64, V, V]
    lib.fpv_camera_render.argtypes = [P(CameraParams), V, I64, V, I32, V, I32, V, D, V, V, V]
    lib.fpv_camera_target_pixel.argtypes = [P(CameraParams), V, I64, V, I32, V, I32, V, D, V, V, V]
    lib.fpv_camera_rays.argtypes = [P(CameraParams), V, I64, V, I32, V, V]
    lib.fpv_autopilot.argtypes = [P(AutopilotParams), P(CameraParams), V, I64, I64, V, V, V, V, V, V, V, V, V]
    lib.fpv_point_and_shoot.argtypes = [P(AutopilotParams), P(CameraParams), V, I64, I64, V, V, V, V, V, V, V, V, V]
    lib.fpv_acro_reset.argtypes = [V, I64, I64, V, V, V, V, V]
    lib.fpv_acro_step.argtypes = [P(AcroParams), V, I64, I64, V, V, I32, V, V, V, V, V, V]
    lib.fpv_acro_rollout.argtypes = [P(AcroParams), V, I64, I64, V, I64, I32, V, I32, V, I64, V, V, V, V, V]
    lib.fpv_probe_fp32.argtypes = [I32, I32, V, I64, P(D), V]
    v = lib.fpv_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {v}, this package needs {ABI_VERSION}; rebuild it")
    for i, st in enumerate(_STRUCTS):
        if lib.fpv_sizeof(i) != C.sizeof(st):
            raise ImportError(f"struct {st.__name__}: library sizeof {lib.fpv_sizeof(i)} != binding {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise FpvError(rc, load().fpv_last_error().decode("utf-8", "replace"))


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def _raw_stream_getter():
    """torch's raw current-stream getter (an int handle, ~4x cheaper than building a torch.cuda.Stream object); the
    public API is the fallback."""
    import torch
    fn = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if fn is not None:
        return fn
    return lambda index: torch.cuda.current_stream(index).cuda_stream


_raw_stream = None


def raw_stream(device_index: int) -> int:
    """cudaStream_t of torch's current stream on `device_index`, as an int."""
    global _raw_stream
    if _raw_stream is None:
        _raw_stream = _raw_stream_getter()
    return _raw_stream(device_index)


def current_stream(device):
    """torch's current stream on `device` for a library call.  The library launches on the CURRENT CUDA device (the CUDA
    runtime's convention; fpv_api.h "Conventions"), so a call for buffers on another device is refused here with a clear
    message instead of failing inside the launch.  (The allocation-free fast path of BatchedDrone.step skips this check;
    there the launch itself fails, just as loudly.)"""
    import torch
    index = device.index if isinstance(device, torch.device) else torch.device(device).index
    cur = torch.cuda.current_device()
    if index is None:
        index = cur
    elif index != cur:
        raise RuntimeError(f"fpyv_b200: the buffers live on cuda:{index} but the current CUDA device is cuda:{cur}; "
                           f"wrap the call in `with torch.cuda.device({index}):` or call torch.cuda.set_device({index})")
    return C.c_void_p(raw_stream(index))
