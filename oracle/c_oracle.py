"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/fpv_oracle.c (float64; env ranges sharded over a
thread pool, ctypes releases the GIL during the C call).
Used by tests (second checker) and by bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfpv_oracle.so")


class Consts(C.Structure):
    _fields_ = [("dt", C.c_double), ("gravity", C.c_double), ("mass", C.c_double), ("max_rates", C.c_double),
                ("rtr", C.c_double), ("ttr", C.c_double), ("k_drag", C.c_double * 3),
                ("motor_rel", (C.c_double * 3) * 4), ("poly", C.c_double * 4), ("motor_radius", C.c_double),
                ("spring_k", C.c_double), ("spring_c", C.c_double), ("ground", C.c_int32)]


def build(force=False):
    src = os.path.join(HERE, "fpv_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/libfpv_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        _lib = C.CDLL(LIB)
        assert _lib.fpv_oracle_sizeof_consts() == C.sizeof(Consts)
    return _lib


def make_consts(c) -> Consts:
    """c: oracle.fpv_oracle.DroneConsts"""
    k = Consts()
    k.dt, k.gravity, k.mass, k.max_rates, k.rtr, k.ttr = c.dt, c.gravity, c.mass, c.max_rates, c.rtr, c.ttr
    for i in range(3):
        k.k_drag[i] = c.k_drag[i]
    for m in range(4):
        for j in range(3):
            k.motor_rel[m][j] = c.motor_rel[m, j]
    for i in range(4):
        k.poly[i] = c.poly[i]
    k.motor_radius, k.spring_k, k.spring_c, k.ground = c.motor_radius, c.spring_k, c.spring_c, int(c.ground)
    return k


def max_threads():
    return len(os.sched_getaffinity(0))


_pool = None


def drone_step(k: Consts, pos, vel, R, prev_rates, prev_thrust, actions, wind=None, substeps=1, acc=None, threads=1):
    """In place on float64 C-contiguous arrays pos[n,3] vel[n,3] R[n,3,3] prev_rates[n,3] prev_thrust[n];
    returns done[n] (uint8).  threads > 1 shards the env range over a thread pool."""
    global _pool
    lib = load()
    n = len(pos)
    for a in (pos, vel, R, prev_rates, prev_thrust, actions):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    wind = np.zeros(3) if wind is None else np.ascontiguousarray(wind, dtype=np.float64)
    done = np.zeros(n, dtype=np.uint8)

    def run(lo, hi):
        p = lambda a, w: C.c_void_p(a.ctypes.data + lo * w * a.itemsize)
        lib.fpv_oracle_drone_step(C.byref(k), C.c_int64(hi - lo), p(pos, 3), p(vel, 3), p(R, 9), p(prev_rates, 3),
                                  p(prev_thrust, 1), p(done, 1), p(actions, 4), C.c_void_p(wind.ctypes.data),
                                  C.c_int(substeps), p(acc, 3) if acc is not None else None)

    if threads <= 1 or n < 2 * threads:
        run(0, n)
    else:
        from concurrent.futures import ThreadPoolExecutor
        if _pool is None or _pool._max_workers != threads:
            _pool = ThreadPoolExecutor(threads)
        cuts = np.linspace(0, n, threads + 1).astype(np.int64)
        list(_pool.map(lambda i: run(int(cuts[i]), int(cuts[i + 1])), range(threads)))
    return done
