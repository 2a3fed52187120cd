"""Page-locked host staging buffers (kws_host_alloc) exposed as torch CPU tensors.

`upload_buffer` returns write-combined pinned memory: the host writes PCM into it and the copy engine reads it at the
PCIe line rate (55 GB/s measured on the B200 box, against 17-35 GB/s from cacheable pinned pages, which is what
`torch.Tensor.pin_memory()` gives).  Do not read such a tensor on the CPU in a loop: write-combined reads are
uncached.  `download_buffer` is ordinary pinned memory for results the host reads."""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from . import _lib


class _HostBlock:
    """Owns one kws_host_alloc allocation; freed when the last tensor view that references it is gone."""

    def __init__(self, nbytes: int, write_combined: bool):
        self.ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().kws_host_alloc(ctypes.byref(self.ptr), int(nbytes), 1 if write_combined else 0),
                   "kws_host_alloc")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if getattr(self, "ptr", None) is not None and self.ptr.value:
                _lib.lib().kws_host_free(self.ptr)
                self.ptr = ctypes.c_void_p()
        except Exception:
            pass




def _tensor(shape: Sequence[int], dtype: torch.dtype, write_combined: bool) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= int(d)
    itemsize = torch.empty((), dtype=dtype).element_size()
    if n == 0:
        return torch.empty(tuple(shape), dtype=dtype)
    block = _HostBlock(n * itemsize, write_combined)
    raw = (ctypes.c_uint8 * block.nbytes).from_address(block.ptr.value)
    raw._kws_block = block                      # the buffer object torch holds on to keeps the allocation alive
    t = torch.frombuffer(raw, dtype=dtype, count=n).view(tuple(shape))
    return t


def upload_buffer(shape: Sequence[int], dtype: torch.dtype = torch.int16) -> torch.Tensor:
    """Write-combined pinned host tensor for host -> device staging (write-only from the CPU's point of view)."""
    return _tensor(shape, dtype, True)


def download_buffer(shape: Sequence[int], dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Pinned host tensor for device -> host results."""
    return _tensor(shape, dtype, False)
