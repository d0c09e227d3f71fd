"""Neighbour list for CFConv (mirror of the reference's NNPOps.CFConvNeighbors, src/pytorch/CFConvNeighbors.py:27-45 over the
Holder of src/pytorch/CFConvNeighbors.cpp:37-85)."""
import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from ._lib import lib, check, ptr, current_stream, register

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
register({
    "nnpops_cfconv_neighbors_create": [C.POINTER(_vp), _i, _f],
    "nnpops_cfconv_neighbors_build": [_vp, _vp, _vp, _vp],
    "nnpops_cfconv_neighbors_num_pairs": [_vp, C.POINTER(C.c_longlong)],
    "nnpops_cfconv_create": [C.POINTER(_vp), _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _i],
    "nnpops_cfconv_compute": [_vp, _vp, _vp, _vp, _vp],
    "nnpops_cfconv_backprop": [_vp, _vp, _vp, _vp, _vp, _vp, _vp],
}, destroyers=("nnpops_cfconv_neighbors_destroy", "nnpops_cfconv_destroy"))


class CFConvNeighbors(torch.nn.Module):
    """``CFConvNeighbors(cutoff)``; ``build(positions)`` like the reference.  The reference's torch surface is hard-wired
    non-periodic (CFConvNeighbors.cpp:52,57,74); ``build`` additionally accepts the periodic box that the C++ interface
    (CFConv.h:57) supports."""

    def __init__(self, cutoff: float) -> None:
        super().__init__()
        self.cutoff = float(cutoff)
        self._h = None
        self._n = -1
        self._device = None

    def __getstate__(self):   # pickle = the cutoff, like the reference Holder (CFConvNeighbors.cpp:81-84)
        return {"cutoff": self.cutoff}

    def __setstate__(self, st):
        self.__init__(st["cutoff"])

    def __del__(self):
        if getattr(self, "_h", None):
            lib.nnpops_cfconv_neighbors_destroy(self._h)
            self._h = None

    def build(self, positions: Tensor, box_vectors: Optional[Tensor] = None) -> None:
        if positions.dtype != torch.float32:
            raise RuntimeError('The type of "positions" has to be float32')
        if positions.dim() != 2 or positions.shape[1] != 3:
            raise RuntimeError('The shape of "positions" has to be (numAtoms, 3)')
        if not positions.is_cuda:
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
        n = positions.shape[0]
        if self._h is None:
            h = C.c_void_p()
            with torch.cuda.device(positions.device):
                check(lib.nnpops_cfconv_neighbors_create(C.byref(h), n, self.cutoff))
            self._h, self._n, self._device = h, n, positions.device
        if n != self._n:
            raise RuntimeError('The size of the 1nd dimension of "positions" has to be %d' % self._n)
        if positions.device != self._device:
            raise RuntimeError('The device of "positions" has changed')
        pos = positions.detach().contiguous()
        box = None
        if box_vectors is not None:
            box = box_vectors.detach().to(device=pos.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(pos.device):
            check(lib.nnpops_cfconv_neighbors_build(self._h, ptr(pos), ptr(box), current_stream(pos.device)))

    def num_pairs(self) -> int:
        v = C.c_longlong(0)
        check(lib.nnpops_cfconv_neighbors_num_pairs(self._h, C.byref(v)))
        return v.value
