"""Page-locked host buffers as torch tensors (fpv_host_alloc / fpv_host_free): what the host-buffer entries
(`BatchedDrone.step_host`, `step_host_sticks`) and their zero-copy forms read from.  `write_combined=True` allocates
write-combining memory for buffers a CPU producer REWRITES every step: its lines never sit dirty in the CPU caches, so the
device's reads (DMA or zero-copy) do not have to snoop them; the CPU must only write such a buffer, never read it back."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class _Block:
    def __init__(self, nbytes, write_combined):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        _lib.check(self.lib.fpv_host_alloc(int(nbytes), int(bool(write_combined)), C.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.fpv_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:       # noqa: BLE001  (interpreter shutdown)
            pass


def pinned(shape, dtype=torch.float32, write_combined=False, pad_to=16):
    """A page-locked CPU tensor of `shape` / `dtype`, its storage padded to a multiple of `pad_to` bytes (the step's bulk
    copies move multiples of 16 bytes).  The tensor keeps its allocation alive; like any buffer handed to an asynchronous
    launch it must outlive the work that reads or writes it (synchronise the stream before dropping the last reference --
    a zero-copy step reads it from the kernel itself)."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
    n = int(np.prod(shape)) if shape else 1
    item = torch.empty((), dtype=dtype).element_size()
    nbytes = max(pad_to, (n * item + pad_to - 1) // pad_to * pad_to)
    block = _Block(nbytes, write_combined)
    buf = (C.c_char * nbytes).from_address(block.ptr.value)
    t = torch.frombuffer(buf, dtype=dtype, count=n).view(shape)
    t._fpv_block = block        # (torch.frombuffer holds `buf`, which does not own the memory: the block does)
    return t
