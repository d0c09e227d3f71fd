"""getNeighborPairs on the B200 cell-list kernels.

Same signature, output conventions and error behaviour as the reference's ``NNPOps.neighbors.getNeighborPairs``
(src/pytorch/neighbors/getNeighborPairs.py:5-147 over the op ``neighbors::getNeighborPairs``, neighbors.cpp:4):

* ``neighbors`` int32 ``(2, num_pairs)`` with ``neighbors[0] > neighbors[1]``; ``deltas`` ``(num_pairs, 3)`` pointing from
  ``neighbors[1]`` to ``neighbors[0]``; ``distances`` ``(num_pairs,)``; pairs beyond the cutoff are ``-1`` / ``NaN``;
* ``max_num_pairs=-1`` -> ``num_pairs = N(N-1)/2`` with the triangular slot order, ``max_num_pairs>0`` -> compacted and padded;
* ``check_errors=True`` raises ``RuntimeError`` when more pairs are found than fit (this synchronises); with
  ``check_errors=False`` nothing synchronises and the call can be captured in a CUDA graph;
* float32 and float64 positions; gradients flow to ``positions`` through ``deltas`` and ``distances``.

Differences, all within what the reference leaves undefined: the compacted order is deterministic (the reference CUDA path uses
an atomic slot counter), and ``number_found_pairs`` is the number of pairs inside the cutoff in both modes.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from .._lib import lib, check, ptr, current_stream, register
import ctypes as C

_vp, _i, _ll = C.c_void_p, C.c_int, C.c_longlong
register({
    "nnpops_neighbor_pairs_f32": [_vp, _vp, _i, C.c_float, _ll, _vp, _vp, _vp, _vp, _vp],
    "nnpops_neighbor_pairs_f64": [_vp, _vp, _i, C.c_double, _ll, _vp, _vp, _vp, _vp, _vp],
    "nnpops_neighbor_pairs_backward_f32": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp],
    "nnpops_neighbor_pairs_backward_f64": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp, _vp],
})


class _NeighborPairs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, cutoff, max_num_pairs, box_vectors, check_errors):
        if positions.dim() != 2:
            raise RuntimeError('Expected "positions" to have two dimensions')
        if positions.shape[0] <= 0:
            raise RuntimeError('Expected the 1nd dimension size of "positions" to be more than 0')
        if positions.shape[1] != 3:
            raise RuntimeError('Expected the 2nd dimension size of "positions" to be 3')
        if not positions.is_contiguous():
            raise RuntimeError('Expected "positions" to be contiguous')
        if not positions.is_cuda:
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
        if positions.dtype not in (torch.float32, torch.float64):
            raise RuntimeError('Expected "positions" to be float32 or float64')
        if not float(cutoff) > 0:
            raise RuntimeError('Expected "cutoff" to be positive')
        max_num_pairs = int(max_num_pairs)
        if not (max_num_pairs > 0 or max_num_pairs == -1):
            raise RuntimeError('Expected "max_num_pairs" to be positive or equal to -1')
        box = None
        if box_vectors is not None and box_vectors.numel() != 0:
            if box_vectors.dim() != 2 or box_vectors.shape != (3, 3):
                raise RuntimeError('Expected "box_vectors" to have shape (3, 3)')
            box = box_vectors.detach().to(device=positions.device, dtype=positions.dtype).contiguous()
        n = positions.shape[0]
        num_pairs = n * (n - 1) // 2 if max_num_pairs == -1 else max_num_pairs
        pos = positions.detach()
        dev, dt = pos.device, pos.dtype
        neighbors = torch.empty((2, num_pairs), dtype=torch.int32, device=dev)
        deltas = torch.empty((num_pairs, 3), dtype=dt, device=dev)
        distances = torch.empty((num_pairs,), dtype=dt, device=dev)
        found = torch.empty((1,), dtype=torch.int32, device=dev)
        fn = lib.nnpops_neighbor_pairs_f32 if dt == torch.float32 else lib.nnpops_neighbor_pairs_f64
        with torch.cuda.device(dev):
            check(fn(ptr(pos), ptr(box), n, float(cutoff), max_num_pairs, ptr(neighbors), ptr(deltas), ptr(distances), ptr(found),
                     current_stream(dev)))
        if check_errors and max_num_pairs != -1:
            if int(found.item()) > max_num_pairs:   # synchronises, like the reference (getNeighborPairsCUDA.cu:157-160)
                raise RuntimeError('The maximum number of pairs has been exceed! Increase "max_num_pairs"')
        ctx.save_for_backward(neighbors, deltas, distances)
        ctx.num_atoms = n
        ctx.mark_non_differentiable(neighbors, found)
        return neighbors, deltas, distances, found

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _gn, grad_deltas, grad_distances, _gf):
        neighbors, deltas, distances = ctx.saved_tensors
        dt, dev = deltas.dtype, deltas.device
        if grad_deltas is None:
            grad_deltas = torch.zeros_like(deltas)
        if grad_distances is None:
            grad_distances = torch.zeros_like(distances)
        gd = grad_deltas.contiguous()
        gr = grad_distances.contiguous()
        grad_pos = torch.empty((ctx.num_atoms, 3), dtype=dt, device=dev)
        fn = lib.nnpops_neighbor_pairs_backward_f32 if dt == torch.float32 else lib.nnpops_neighbor_pairs_backward_f64
        with torch.cuda.device(dev):
            check(fn(ptr(neighbors), ptr(deltas), ptr(distances), ptr(gd), ptr(gr), distances.shape[0], ctx.num_atoms, ptr(grad_pos),
                     current_stream(dev)))
        return grad_pos, None, None, None, None


def getNeighborPairs(positions: Tensor, cutoff: float, max_num_pairs: int = -1, box_vectors: Optional[Tensor] = None,
                     check_errors: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns indices, displacement vectors and distances of atom pairs within ``cutoff`` (see module docstring)."""
    return _NeighborPairs.apply(positions, cutoff, max_num_pairs, box_vectors, check_errors)
