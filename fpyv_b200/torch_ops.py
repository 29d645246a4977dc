"""`torch.library` custom op over the C ABI -- the second binding form SURVEY.md section 8(b) names next to ctypes.

`torch.ops.fpyv_b200.drone_step(state, actions, done, acc, handle)` is `BatchedDrone.step(actions, return_obs=False)`
(Drone.step, components.py:220-248, for every env) declared to PyTorch as an op that MUTATES the drone's state planes, done
flags and acceleration buffer.  What that buys over the plain method: a policy + env step written as one function can go
through `torch.compile(fullgraph=True)` / `torch.export` without a graph break, and functionalisation knows which buffers
the step writes.  The op is a thin shell: the launch itself is the same `fpv_drone_step` call on the caller's current
stream, with the same fast path; there is no CPU implementation (calling it with CPU tensors raises).

    from fpyv_b200 import BatchedDrone, torch_ops
    d = BatchedDrone(None, num_envs=n, device="cuda:0", substeps=8, dt=1e-3)
    d.reset(pos, vel, ypr)
    step = torch_ops.bind(d)                    # -> callable(actions) -> None, usable inside torch.compile
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib

_DRONES: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next_handle = 1


def register(drone) -> int:
    """Handle of `drone` for the `handle` argument of the op (custom ops take tensors and scalars, not objects)."""
    global _next_handle
    h = getattr(drone, "_op_handle", None)
    if h is None:
        h = drone._op_handle = _next_handle
        _next_handle += 1
    _DRONES[h] = drone
    return h


@torch.library.custom_op("fpyv_b200::drone_step", mutates_args=("state", "done", "acc"), device_types="cuda")
def drone_step(state: torch.Tensor, actions: torch.Tensor, done: torch.Tensor, acc: torch.Tensor, handle: int) -> None:
    """One control step ON THE TENSORS PASSED IN (a functionalising backend may hand the op copies of the drone's buffers
    and copy them back afterwards, so the op never reaches for the drone's own state behind its arguments); the drone
    behind `handle` supplies the parameters, the motor-curve table, the reset snapshot and the scratch counters."""
    d = _DRONES.get(handle)
    if d is None:
        raise RuntimeError(f"fpyv_b200::drone_step: unknown drone handle {handle} (torch_ops.register(drone) first)")
    if not d._is_reset:
        raise RuntimeError("call reset() before step() (the reference's state is None until reset)")
    for name, t, like in (("state", state, d._state), ("done", done, d._done), ("acc", acc, d._acc), ("actions", actions, d._actions)):
        if t.shape != like.shape or t.dtype != like.dtype or t.device != like.device or not t.is_contiguous():
            raise RuntimeError(f"fpyv_b200::drone_step: `{name}` must be a contiguous {like.dtype} tensor of shape "
                               f"{tuple(like.shape)} on {like.device}")
    if d._static is not None:
        raise RuntimeError("fpyv_b200::drone_step serves the hot-path configuration; a drone with set_static_objects() steps with step()")
    if not d._fast_ok:
        d._configure_plain_io()
    io = _lib.DroneIO.from_buffer_copy(d._io)
    io.state, io.actions, io.done, io.acc_out = state.data_ptr(), actions.data_ptr(), done.data_ptr(), acc.data_ptr()
    io.chunk_epoch = None
    d._p.flags = d._flags
    d._last_action, d._chain_ready = actions, False
    _lib.check(d._lib.fpv_drone_step(C.byref(d._p), C.byref(io), _lib.current_stream(d.device)))


@drone_step.register_fake
def _(state, actions, done, acc, handle):
    return None


def bind(drone):
    """callable(actions) -> None stepping `drone` through the custom op."""
    h = register(drone)
    state, done, acc = drone._state, drone._done, drone._acc

    def step(actions: torch.Tensor) -> None:
        torch.ops.fpyv_b200.drone_step(state, actions, done, acc, h)

    step.drone = drone      # the registry holds weak references: the callable keeps its drone alive
    return step
