"""Continuous-filter convolution of SchNet (mirror of the reference's NNPOps.CFConv, src/pytorch/CFConv.py:29-83 over the Holder
and operation of src/pytorch/CFConv.cpp:50-290).  Same constructor
``CFConv(gaussianWidth, activation, weights1[G, W], biases1[W], weights2[W, W], biases2[W])`` and call
``conv(neighbors, positions, input) -> output``; gradients with respect to positions and input only."""
import ctypes as C

import torch
from torch import Tensor

from ._lib import lib, check, ptr, current_stream
from .CFConvNeighbors import CFConvNeighbors


class _Operation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv, neighbors, positions, input):
        if positions.dtype != torch.float32 or input.dtype != torch.float32:
            raise RuntimeError('The type of "positions" and "input" has to be float32')
        if input.dim() != 2 or input.shape[1] != conv.numFilters:
            raise RuntimeError('The shape of "input" has to be (numAtoms, %d)' % conv.numFilters)
        if neighbors._h is None:
            raise RuntimeError("CFConvNeighbors.build() has to be called before CFConv")
        if input.shape[0] != neighbors._n or positions.shape[0] != neighbors._n:
            raise RuntimeError('The size of the 1nd dimension of "positions" and "input" has to be %d' % neighbors._n)
        conv._ensure(neighbors.cutoff, input.device)
        x = input.detach().contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(lib.nnpops_cfconv_compute(conv._h, neighbors._h, ptr(x), ptr(out), current_stream(x.device)))
        ctx.conv, ctx.neighbors = conv, neighbors
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        go = grad_output.contiguous()
        input_grad = torch.empty_like(x)
        pos_grad = torch.empty((x.shape[0], 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.nnpops_cfconv_backprop(ctx.conv._h, ctx.neighbors._h, ptr(x), ptr(go), ptr(input_grad), ptr(pos_grad),
                                             current_stream(x.device)))
        return None, None, pos_grad, input_grad


class CFConv(torch.nn.Module):
    def __init__(self, gaussianWidth: float, activation: str, weights1: Tensor, biases1: Tensor, weights2: Tensor,
                 biases2: Tensor) -> None:
        super().__init__()
        if activation not in ("ssp", "tanh"):
            raise ValueError('Invalid value of "activation"')   # CFConv.cpp:98-103
        if weights1.dim() != 2 or weights2.dim() != 2 or biases1.dim() != 1 or biases2.dim() != 1:
            raise RuntimeError("weights have to be 2-D and biases 1-D")
        self.numGaussians, self.numFilters = weights1.shape          # documented [numGaussians, numFilters] (CFConv.py:48)
        if weights2.shape != (self.numFilters, self.numFilters) or biases1.shape[0] != self.numFilters or biases2.shape[0] != self.numFilters:
            raise RuntimeError("inconsistent weight shapes")          # CFConv.cpp:107-128
        self.gaussianWidth = float(gaussianWidth)
        self.activation = activation
        # the storage is handed over unchanged and indexed [numFilters][numGaussians] by the kernels, as in CFConv.cpp:131-132
        self.register_buffer("weights1", weights1.detach().to(torch.float32).contiguous().clone())
        self.register_buffer("biases1", biases1.detach().to(torch.float32).contiguous().clone())
        self.register_buffer("weights2", weights2.detach().to(torch.float32).contiguous().clone())
        self.register_buffer("biases2", biases2.detach().to(torch.float32).contiguous().clone())
        self._h = None
        self._key = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib.nnpops_cfconv_destroy(self._h)
            self._h = None

    def _ensure(self, cutoff: float, device) -> None:
        key = (float(cutoff), str(device))
        if self._h is not None and self._key == key:
            return
        if device.type != "cuda":
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
        if self._h is not None:
            lib.nnpops_cfconv_destroy(self._h)
        w1, b1, w2, b2 = (t.to(device) for t in (self.weights1, self.biases1, self.weights2, self.biases2))
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.nnpops_cfconv_create(C.byref(h), self.numFilters, self.numGaussians, float(cutoff), self.gaussianWidth,
                                           0 if self.activation == "ssp" else 1, ptr(w1), ptr(b1), ptr(w2), ptr(b2), 0))
        self._h, self._key = h, key

    def forward(self, neighbors: CFConvNeighbors, positions: Tensor, input: Tensor) -> Tensor:
        return _Operation.apply(self, neighbors, positions, input)
