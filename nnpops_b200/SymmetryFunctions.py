"""ANI symmetry functions on the B200 kernels.

Mirrors the reference's ``NNPOps.SymmetryFunctions`` (src/pytorch/SymmetryFunctions.py:32-123) and the Holder/operation pair behind
it (src/pytorch/SymmetryFunctions.cpp:52-263): same constructor arguments, same ``forward((species, positions), cell, pbc)``
contract, same error behaviour (float32 positions of shape [1, N, 3], fully periodic or not at all, no batches), gradients with
respect to positions only.  The arithmetic runs in libnnpops_b200.so through the C ABI (include/nnpops_b200.h).
"""
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import lib, check, ptr, current_stream


def function_tables(EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ):
    """Expand the TorchANI constant lists into {eta, rs} / {eta, rs, zeta, thetas} tables in the reference's order
    (src/pytorch/SymmetryFunctions.cpp:110-120); values are narrowed to float exactly as there."""
    radial = np.array([[e, s] for e in EtaR for s in ShfR], np.float32).reshape(-1, 2)
    angular = np.array([[e, s, z, t] for e in EtaA for z in Zeta for s in ShfA for t in ShfZ], np.float32).reshape(-1, 4)
    return radial, angular


class Holder:
    """Counterpart of torch.classes.NNPOpsANISymmetryFunctions.Holder (SymmetryFunctions.cpp:52-234)."""

    def __init__(self, numSpecies: int, Rcr: float, Rca: float, EtaR: Sequence[float], ShfR: Sequence[float], EtaA: Sequence[float],
                 Zeta: Sequence[float], ShfA: Sequence[float], ShfZ: Sequence[float], atomSpecies: Sequence[int], torchani: bool = True,
                 maxRadialNeighbors: int = 0, maxAngularNeighbors: int = 0):
        self.args = (int(numSpecies), float(Rcr), float(Rca), list(map(float, EtaR)), list(map(float, ShfR)), list(map(float, EtaA)),
                     list(map(float, Zeta)), list(map(float, ShfA)), list(map(float, ShfZ)), list(map(int, atomSpecies)))
        self.torchani = bool(torchani)
        self.caps = (int(maxRadialNeighbors), int(maxAngularNeighbors))
        self.numSpecies = int(numSpecies)
        self.numAtoms = len(self.args[9])
        self.radial_fn, self.angular_fn = function_tables(*self.args[3:9])
        self._h = None
        self._device = None
        self._periodic = None

    @classmethod
    def from_function_lists(cls, numSpecies, Rcr, Rca, radial_fn, angular_fn, atomSpecies, torchani=True):
        """The C++-level interface takes arbitrary function lists (ANISymmetryFunctions.h:60-64)."""
        self = cls(numSpecies, Rcr, Rca, [], [], [], [], [], [], atomSpecies, torchani)
        self.radial_fn = np.ascontiguousarray(radial_fn, np.float32).reshape(-1, 2)
        self.angular_fn = np.ascontiguousarray(angular_fn, np.float32).reshape(-1, 4)
        return self

    # pickling = constructor arguments only, like Holder::serialize (SymmetryFunctions.cpp:177-218)
    def __getstate__(self):
        return {"args": self.args, "torchani": self.torchani, "caps": self.caps, "radial_fn": self.radial_fn, "angular_fn": self.angular_fn}

    def __setstate__(self, st):
        self.__init__(*st["args"], torchani=st["torchani"], maxRadialNeighbors=st["caps"][0], maxAngularNeighbors=st["caps"][1])
        self.radial_fn, self.angular_fn = st["radial_fn"], st["angular_fn"]

    def __del__(self):
        if getattr(self, "_h", None):
            lib.nnpops_ani_destroy(self._h)
            self._h = None

    @property
    def radial_width(self):
        return self.numSpecies * len(self.radial_fn)

    @property
    def angular_width(self):
        return self.numSpecies * (self.numSpecies + 1) // 2 * len(self.angular_fn)

    def _create(self, device):
        if not device.type == "cuda":
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback); got device %s" % device)
        species = np.ascontiguousarray(self.args[9], np.int32)
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.nnpops_ani_create(C.byref(h), self.numAtoms, self.numSpecies, self.args[1], self.args[2], ptr(species),
                                        len(self.radial_fn), ptr(self.radial_fn), len(self.angular_fn), ptr(self.angular_fn),
                                        int(self.torchani), self.caps[0], self.caps[1]))
        self._h, self._device = h, device

    def forward(self, positions: Tensor, cell: Optional[Tensor]) -> List[Tensor]:
        # validation mirrors Holder::forward (SymmetryFunctions.cpp:76-102)
        if positions.dtype != torch.float32:
            raise RuntimeError('The type of "positions" has to be float32')
        if positions.dim() != 2:
            raise RuntimeError('The shape of "positions" has to have 2 dimensions')
        if positions.shape[0] != self.numAtoms:
            raise RuntimeError('The size of the 1nd dimension of "positions" has to be %d' % self.numAtoms)
        if positions.shape[1] != 3:
            raise RuntimeError('The size of the 2nd dimension of "positions" has to be 3')
        if cell is not None:
            if cell.dtype != torch.float32:
                raise RuntimeError('The type of "cell" has to be float32')
            if cell.dim() != 2 or cell.shape[0] != 3 or cell.shape[1] != 3:
                raise RuntimeError('The shape of "cell" has to be (3, 3)')
            if cell.device != positions.device:
                raise RuntimeError('"cell" has to be on the same device as "positions"')
        if self._h is None:
            self._create(positions.device)
            self._periodic = cell is not None   # periodic-ness is frozen at the first call (SymmetryFunctions.cpp:122)
        if positions.device != self._device:
            raise RuntimeError('The device of "positions" has changed')
        if (cell is not None) != self._periodic:
            raise RuntimeError("The periodicity of the system has changed")
        pos = positions.detach().contiguous()
        box = cell.detach().contiguous() if cell is not None else None
        radial = torch.empty((self.numAtoms, self.radial_width), dtype=torch.float32, device=pos.device)
        angular = torch.empty((self.numAtoms, self.angular_width), dtype=torch.float32, device=pos.device)
        # The reference has no neighbour limit (N x N table, CudaANISymmetryFunctions.cu:44); here rows have a capacity.  An overflow is
        # detected right after the call and the call repeated with rows twice as long, so a result is never silently truncated.
        # Nothing may synchronise while a CUDA graph is being captured: there the check is left to the next eager call.
        capturing = torch.cuda.is_current_stream_capturing()
        with torch.cuda.device(pos.device):
            for attempt in range(9):
                check(lib.nnpops_ani_forward(self._h, ptr(pos), ptr(box), ptr(radial), ptr(angular), current_stream(pos.device)))
                if capturing:
                    break
                flags = self.overflowed()          # synchronises, like the reference's host read of the box
                if not flags:
                    break
                if attempt == 8:
                    raise RuntimeError("nnpops_b200: neighbour rows still overflow after growing them 8 times")
                caps = list(self.caps)
                if flags & 1:
                    caps[0] = 2 * (caps[0] or 256)
                if flags & 2:
                    caps[1] = 2 * (caps[1] or 64)
                self.caps = tuple(caps)
                lib.nnpops_ani_destroy(self._h)
                self._h = None
                self._create(pos.device)
        return [radial, angular]

    def backward(self, grads: List[Tensor]) -> Tensor:
        rg = grads[0].contiguous().float()
        ag = grads[1].contiguous().float()
        out = torch.empty((self.numAtoms, 3), dtype=torch.float32, device=rg.device)
        with torch.cuda.device(rg.device):
            check(lib.nnpops_ani_backward(self._h, ptr(rg), ptr(ag), ptr(out), current_stream(rg.device)))
        return out

    def overflowed(self) -> int:
        f = C.c_int(0)
        check(lib.nnpops_ani_overflowed(self._h, C.byref(f)))
        return f.value

    def work(self):
        t, p = C.c_longlong(0), C.c_longlong(0)
        check(lib.nnpops_ani_work(self._h, C.byref(t), C.byref(p), current_stream(self._device)))
        return t.value, p.value


class _Operation(torch.autograd.Function):
    """Counterpart of AutogradFunctions (SymmetryFunctions.cpp:236-256): gradient w.r.t. positions only."""

    @staticmethod
    def forward(ctx, holder, positions, cell):
        ctx.holder = holder
        radial, angular = holder.forward(positions, cell)
        return radial, angular

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_radial, grad_angular):
        holder = ctx.holder
        if grad_radial is None:
            grad_radial = torch.zeros((holder.numAtoms, holder.radial_width), dtype=torch.float32, device=holder._device)
        if grad_angular is None:
            grad_angular = torch.zeros((holder.numAtoms, holder.angular_width), dtype=torch.float32, device=holder._device)
        return None, holder.backward([grad_radial, grad_angular]), None


def operation(holder: Holder, positions: Tensor, cell: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    return _Operation.apply(holder, positions, cell)


class TorchANISymmetryFunctions(torch.nn.Module):
    """Drop-in for torchani.AEVComputer with the reference's constructor (SymmetryFunctions.py:66-91): ``converter`` and
    ``symmFunc`` are duck-typed (num_species, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ and a callable species converter)."""

    def __init__(self, converter, symmFunc, atomicNumbers: Tensor) -> None:
        super().__init__()
        self.num_species = symmFunc.num_species
        EtaR = symmFunc.EtaR[:, 0].tolist()
        ShfR = symmFunc.ShfR[0, :].tolist()
        EtaA = symmFunc.EtaA[:, 0, 0, 0].tolist()
        Zeta = symmFunc.Zeta[0, :, 0, 0].tolist()
        ShfA = symmFunc.ShfA[0, 0, :, 0].tolist()
        ShfZ = symmFunc.ShfZ[0, 0, 0, :].tolist()
        species = converter((atomicNumbers, torch.empty(0))).species[0].tolist()
        self.holder = Holder(self.num_species, symmFunc.Rcr, symmFunc.Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, species)

    @classmethod
    def from_constants(cls, num_species, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, species):
        """torchani-free constructor: raw AEV constants and species indices (0..num_species-1)."""
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.num_species = num_species
        self.holder = Holder(num_species, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, list(species))
        return self

    def forward(self, species_positions: Tuple[Tensor, Tensor], cell: Optional[Tensor] = None,
                pbc: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        species, positions = species_positions
        if species.shape[0] != 1:
            raise ValueError('Batched computation of molecules is not supported')
        if cell is not None:
            if pbc is None:
                raise ValueError('"pbc" has to be defined')
            if pbc.tolist() != [True, True, True]:
                raise ValueError('Only fully periodic systems are supported, i.e. pbc = [True, True, True]')
        radial, angular = operation(self.holder, positions[0], cell)
        features = torch.cat((radial, angular), dim=1).unsqueeze(0)
        return species, features
