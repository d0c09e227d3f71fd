"""PME on the B200 kernels: same class, constructor arguments and methods as the reference's ``NNPOps.pme.PME``
(src/pytorch/pme/pme.py:5-196): ``PME(gridx, gridy, gridz, order, alpha, coulomb, exclusions)`` with ``compute_direct`` and
``compute_reciprocal``; gradients with respect to positions and charges only; second derivatives raise."""
import ctypes as C
import math
from typing import Optional

import torch
from torch import Tensor

from .._lib import lib, check, ptr, current_stream, register
from ..neighbors import getNeighborPairs

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
register({
    "nnpops_pme_direct": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _f, _f, _vp, _vp, _vp, _vp],
    "nnpops_pme_direct_fused": [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _i, _i, _vp, _vp, _vp, _vp],
    "nnpops_pme_reciprocal_forward": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "nnpops_pme_reciprocal_backward": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "nnpops_pme_spread": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp],
    "nnpops_pme_solve": [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
})


def _f32(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError('"%s" has to be float32' % what)
    return t.detach().contiguous()


class _Direct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, charges, neighbors, deltas, distances, exclusions, alpha, coulomb):
        pos, q = _f32(positions, "positions"), _f32(charges, "charges")
        d, r = _f32(deltas, "deltas"), _f32(distances, "distances")
        nb = neighbors.detach().to(torch.int32).contiguous()
        ex = exclusions.detach().to(device=pos.device, dtype=torch.int32).contiguous()
        n = q.shape[0]
        energy = torch.empty((), dtype=torch.float32, device=pos.device)
        pos_deriv = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
        charge_deriv = torch.empty((n,), dtype=torch.float32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.nnpops_pme_direct(ptr(pos), ptr(q), ptr(nb), ptr(d), ptr(r), ptr(ex) if ex.numel() else None, n, nb.shape[1],
                                        ex.shape[1] if ex.dim() == 2 else 0, float(alpha), float(coulomb), ptr(energy), ptr(pos_deriv),
                                        ptr(charge_deriv), current_stream(pos.device)))
        ctx.save_for_backward(pos_deriv, charge_deriv)
        return energy

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        pos_deriv, charge_deriv = ctx.saved_tensors
        return pos_deriv * grad, charge_deriv * grad, None, None, None, None, None, None


class _DirectFused(torch.autograd.Function):
    """Direct space straight from positions (nnpops_pme_direct_fused: cell list + centre-owned erfc sum, no pair list); `shard` =
    (index, count) restricts the centres to one slab of the cell-sorted atoms."""

    @staticmethod
    def forward(ctx, positions, charges, box_vectors, exclusions, cutoff, alpha, coulomb, shard):
        pos, q, box = _f32(positions, "positions"), _f32(charges, "charges"), _f32(box_vectors, "box_vectors")
        ex = exclusions.detach().to(device=pos.device, dtype=torch.int32).contiguous()
        n = q.shape[0]
        energy = torch.empty((), dtype=torch.float32, device=pos.device)
        pos_deriv = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
        charge_deriv = torch.empty((n,), dtype=torch.float32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.nnpops_pme_direct_fused(ptr(pos), ptr(q), ptr(box), ptr(ex) if ex.numel() else None, n,
                                              ex.shape[1] if ex.dim() == 2 else 0, float(cutoff), float(alpha), float(coulomb),
                                              int(shard[0]), int(shard[1]), ptr(energy), ptr(pos_deriv), ptr(charge_deriv),
                                              current_stream(pos.device)))
        ctx.save_for_backward(pos_deriv, charge_deriv)
        return energy

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        pos_deriv, charge_deriv = ctx.saved_tensors
        return pos_deriv * grad, charge_deriv * grad, None, None, None, None, None, None


class _Reciprocal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, charges, box_vectors, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli):
        pos, q, box = _f32(positions, "positions"), _f32(charges, "charges"), _f32(box_vectors, "box_vectors")
        dev = pos.device
        xm, ym, zm = (m.detach().to(device=dev, dtype=torch.float32).contiguous() for m in (xmoduli, ymoduli, zmoduli))
        energy = torch.empty((), dtype=torch.float32, device=dev)
        recip = torch.empty((gridx, gridy, gridz // 2 + 1, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.nnpops_pme_reciprocal_forward(ptr(pos), ptr(q), ptr(box), q.shape[0], gridx, gridy, gridz, order, float(alpha),
                                                    float(coulomb), ptr(xm), ptr(ym), ptr(zm), ptr(energy), ptr(recip), current_stream(dev)))
        ctx.save_for_backward(pos, q, box, recip)
        ctx.params = (gridx, gridy, gridz, order, float(coulomb))
        return energy

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        pos, q, box, recip = ctx.saved_tensors
        gridx, gridy, gridz, order, coulomb = ctx.params
        dev = pos.device
        pos_deriv = torch.empty_like(pos)
        charge_deriv = torch.empty_like(q)
        with torch.cuda.device(dev):
            check(lib.nnpops_pme_reciprocal_backward(ptr(pos), ptr(q), ptr(box), q.shape[0], gridx, gridy, gridz, order, coulomb, ptr(recip),
                                                     ptr(pos_deriv), ptr(charge_deriv), current_stream(dev)))
        return (pos_deriv * grad, charge_deriv * grad) + (None,) * 10


def pme_direct(positions, charges, neighbors, deltas, distances, exclusions, alpha, coulomb):
    """Functional form of the op pme::pme_direct (pme.cpp:3-4)."""
    return _Direct.apply(positions, charges, neighbors, deltas, distances, exclusions, alpha, coulomb)


def pme_direct_fused(positions, charges, box_vectors, exclusions, cutoff, alpha, coulomb, shard=(0, 1)):
    """getNeighborPairs + pme::pme_direct in one traversal (same accepted pairs, same terms; sums in a different order).  `exclusions`
    as PME prepares them: int32 [atoms, max_exclusions], rows sorted descending, padded with -1, symmetric."""
    return _DirectFused.apply(positions, charges, box_vectors, exclusions, cutoff, alpha, coulomb, shard)


def pme_reciprocal(positions, charges, box_vectors, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli):
    """Functional form of the op pme::pme_reciprocal (pme.cpp:5-6)."""
    return _Reciprocal.apply(positions, charges, box_vectors, int(gridx), int(gridy), int(gridz), int(order), alpha, coulomb, xmoduli,
                             ymoduli, zmoduli)


def shard_range(num_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `rank` when num_items atoms (or pairs) are dealt to `world` ranks."""
    per = (num_items + world - 1) // world
    return min(rank * per, num_items), min((rank + 1) * per, num_items)


def pme_spread(positions, charges, box_vectors, gridx, gridy, gridz, order, coulomb) -> Tensor:
    """Charge grid float [gridx, gridy, gridz] of the atoms given (first stage of pme_reciprocal, nnpops_pme_spread)."""
    pos, q, box = _f32(positions, "positions"), _f32(charges, "charges"), _f32(box_vectors, "box_vectors")
    grid = torch.empty((gridx, gridy, gridz), dtype=torch.float32, device=pos.device)
    with torch.cuda.device(pos.device):
        check(lib.nnpops_pme_spread(ptr(pos), ptr(q), ptr(box), q.shape[0], gridx, gridy, gridz, order, float(coulomb), ptr(grid),
                                    current_stream(pos.device)))
    return grid


def pme_solve(grid, box_vectors, alpha, xmoduli, ymoduli, zmoduli):
    """FFT + Ewald convolution + energy of a charge grid (second stage, nnpops_pme_solve) -> (energy [], convolved half-complex grid)."""
    g, box = _f32(grid, "grid"), _f32(box_vectors, "box_vectors")
    gridx, gridy, gridz = g.shape
    dev = g.device
    xm, ym, zm = (m.detach().to(device=dev, dtype=torch.float32).contiguous() for m in (xmoduli, ymoduli, zmoduli))
    energy = torch.empty((), dtype=torch.float32, device=dev)
    recip = torch.empty((gridx, gridy, gridz // 2 + 1, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.nnpops_pme_solve(ptr(g), ptr(box), gridx, gridy, gridz, float(alpha), ptr(xm), ptr(ym), ptr(zm), ptr(energy), ptr(recip),
                                   current_stream(dev)))
    return energy, recip


class _ShardedReciprocal(torch.autograd.Function):
    """pme_reciprocal with the atoms dealt to the ranks of a torch.distributed group (PME row): every rank
    spreads its block of atoms into a private full-size grid, ONE all-reduce sums the grids (8.4 MB at 128^3, NCCL over NVLink),
    every rank solves the same grid, and in backward interpolates the forces of its own atoms; the per-atom derivatives are
    assembled with a second all-reduce of zero-padded arrays.  Inputs are replicated, the energy and the gradients come out
    replicated.  `emulate` = (rank, world, reduce) runs one rank of the scheme without a process group (tests)."""

    @staticmethod
    def forward(ctx, positions, charges, box_vectors, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli, group, emulate):
        import torch.distributed as dist
        if emulate is not None:
            rank, world, reduce = emulate
        else:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)   # noqa: E731
        pos, q, box = _f32(positions, "positions"), _f32(charges, "charges"), _f32(box_vectors, "box_vectors")
        lo, hi = shard_range(q.shape[0], rank, world)
        grid = pme_spread(pos[lo:hi], q[lo:hi], box, gridx, gridy, gridz, order, coulomb)
        if world > 1:
            reduce(grid)
        energy, recip = pme_solve(grid, box, alpha, xmoduli, ymoduli, zmoduli)
        ctx.save_for_backward(pos, q, box, recip)
        ctx.params = (gridx, gridy, gridz, order, float(coulomb), lo, hi, world, reduce)
        return energy

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        pos, q, box, recip = ctx.saved_tensors
        gridx, gridy, gridz, order, coulomb, lo, hi, world, reduce = ctx.params
        dev = pos.device
        packed = torch.zeros((q.shape[0], 4), dtype=torch.float32, device=dev)      # dE/dx | dE/dq, zero outside the rank's block
        if hi > lo:
            pd = torch.empty((hi - lo, 3), dtype=torch.float32, device=dev)
            cd = torch.empty((hi - lo,), dtype=torch.float32, device=dev)
            pl, ql = pos[lo:hi].contiguous(), q[lo:hi].contiguous()
            with torch.cuda.device(dev):
                check(lib.nnpops_pme_reciprocal_backward(ptr(pl), ptr(ql), ptr(box), hi - lo, gridx, gridy, gridz, order, coulomb, ptr(recip),
                                                         ptr(pd), ptr(cd), current_stream(dev)))
            packed[lo:hi, :3] = pd
            packed[lo:hi, 3] = cd
        if world > 1:
            reduce(packed)
        return (packed[:, :3] * grad, packed[:, 3] * grad) + (None,) * 12


def pme_reciprocal_sharded(positions, charges, box_vectors, gridx, gridy, gridz, order, alpha, coulomb, xmoduli, ymoduli, zmoduli,
                           group=None, emulate=None):
    """pme_reciprocal for a system sharded over the ranks of `group` (default group if None); see _ShardedReciprocal."""
    return _ShardedReciprocal.apply(positions, charges, box_vectors, int(gridx), int(gridy), int(gridz), int(order), alpha, coulomb,
                                    xmoduli, ymoduli, zmoduli, group, emulate)


def pme_direct_sharded(positions, charges, neighbors, deltas, distances, exclusions, alpha, coulomb, group=None):
    """pme_direct with the PAIRS of a (replicated) neighbour list dealt to the ranks: every rank evaluates its block of pairs, the
    exclusion correction (a per-atom term) is kept on rank 0 only, and energy and derivatives are summed by all-reduce through
    autograd-transparent wrappers.  Returns the total energy on every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(neighbors.shape[1], rank, world)
    e = pme_direct(positions, charges, neighbors[:, lo:hi].contiguous(), deltas[lo:hi].contiguous(), distances[lo:hi].contiguous(),
                   exclusions, alpha, coulomb)
    if rank != 0 and exclusions.numel() > 0:
        # the op adds the exclusion correction whenever exclusions are passed: take it out again on every rank but one
        empty = neighbors[:, :0].contiguous()
        e = e - pme_direct(positions, charges, empty, deltas[:0].contiguous(), distances[:0].contiguous(), exclusions, alpha, coulomb)
    return _AllReduceSum.apply(e, group)


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x; dL/dx = dL/dy (the loss is the same replicated scalar on every rank, each rank differentiates its
    own term, and the callers sum position gradients over the ranks)."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, grad):
        return grad, None


def bspline_moduli(order: int, sizes):
    """|DFT of the order-`order` cardinal B-spline sampled at the integers|^2 for each grid size, in float32 exactly as the
    reference builds them in PME.__init__ (pme.py:94-129); moduli below 1e-7 are replaced by the mean of their neighbours."""
    max_size = max(sizes)
    data = torch.zeros(order, dtype=torch.float32)
    bsplines_data = torch.zeros(max(max_size, order + 1), dtype=torch.float32)
    data[0] = 1
    for i in range(3, order):
        data[i - 1] = 0
        for j in range(1, i - 1):
            data[i - j - 1] = (j * data[i - j - 2] + (i - j) * data[i - j - 1]) / (i - 1)
        data[0] /= i - 1
    for i in range(1, order - 1):
        data[order - i - 1] = (i * data[order - i - 2] + (order - i) * data[order - i - 1]) / (order - 1)
    data[0] /= order - 1
    bsplines_data[1:order + 1] = data
    out = []
    for n in sizes:
        scale = torch.tensor([2 * math.pi * i / n for i in range(n)], dtype=torch.float64).to(torch.float32)
        arg = scale[:, None] * torch.arange(n).to(torch.float32)[None, :]
        b = bsplines_data[:n]
        sc = torch.sum(b * torch.cos(arg), dim=1)
        ss = torch.sum(b * torch.sin(arg), dim=1)
        m = sc * sc + ss * ss
        for i in range(n):
            if m[i] < 1e-7:
                m[i] = (m[(i - 1 + n) % n] + m[(i + 1) % n]) * 0.5
        out.append(m)
    return out


class PME:
    """Particle-mesh Ewald for a periodic system of point charges; see the reference docstring (pme.py:5-51) for the physics."""

    def __init__(self, gridx: int, gridy: int, gridz: int, order: int, alpha: float, coulomb: float, exclusions: Tensor):
        if gridx < 1 or gridy < 1 or gridz < 1:
            raise ValueError('The grid dimensions must be positive')
        if order < 1:
            raise ValueError('order must be positive')
        if alpha <= 0:
            raise ValueError('alpha must be positive')
        if coulomb <= 0:
            raise ValueError('coulomb must be positive')
        if exclusions.dim() != 2:
            raise ValueError('exclusions must be 2D')
        self.gridx, self.gridy, self.gridz = gridx, gridy, gridz
        self.order, self.alpha, self.coulomb = order, alpha, coulomb
        self.exclusions, _ = torch.sort(exclusions.to(torch.int32), descending=True)
        self.moduli = bspline_moduli(order, (gridx, gridy, gridz))

    def _validate(self, positions, charges, box_vectors):
        if positions.dim() != 2 or positions.shape[1] != 3:
            raise ValueError('positions must have shape (atoms, 3)')
        if charges.dim() != 1:
            raise ValueError('charges must be 1D')
        if positions.shape[0] != self.exclusions.shape[0] or charges.shape[0] != self.exclusions.shape[0]:
            raise ValueError('positions, charges, and exclusions must all have the same length')
        if box_vectors.dim() != 2 or box_vectors.shape[0] != 3 or box_vectors.shape[1] != 3:
            raise ValueError('box_vectors must have shape (3, 3)')

    def compute_direct(self, positions: Tensor, charges: Tensor, cutoff: float, box_vectors: Tensor, max_num_pairs: int = -1):
        """Direct-space energy (pme.py:131-165): neighbour list + erfc sum + exclusion correction.  With the default
        max_num_pairs = -1 (the reference then allocates N (N - 1) / 2 pair slots) nothing about the list is observable, and the sum
        is evaluated by the fused kernel without materialising it; a positive max_num_pairs keeps the reference's two steps and
        their capacity semantics (pairs beyond it are dropped)."""
        self._validate(positions, charges, box_vectors)
        if cutoff <= 0:
            raise ValueError('cutoff must be positive')
        self.exclusions = self.exclusions.to(positions.device)
        if max_num_pairs == -1:
            return pme_direct_fused(positions, charges, box_vectors, self.exclusions, cutoff, self.alpha, self.coulomb)
        neighbors, deltas, distances, _ = getNeighborPairs(positions, cutoff, max_num_pairs, box_vectors)
        return pme_direct(positions, charges, neighbors, deltas, distances, self.exclusions, self.alpha, self.coulomb)

    def compute_direct_sharded(self, positions: Tensor, charges: Tensor, cutoff: float, box_vectors: Tensor, group=None, emulate=None):
        """compute_direct for one box over the ranks of a torch.distributed group (PME row): positions and
        charges are replicated, rank r evaluates the centres of slab r of the cell-sorted atoms with the fused kernel (each is a
        complete sum over its neighbours: no halo exchange, no atomics), and the energy is summed by one scalar all-reduce.
        Returns the total energy on every rank; a rank's autograd gradient holds the derivatives of ITS centres (zeros elsewhere),
        so the caller sums position / charge gradients over the ranks (one all-reduce of [atoms, 4]), as for
        pme_reciprocal_sharded.  `emulate` = (rank, world) evaluates one rank's term without a process group (tests)."""
        self._validate(positions, charges, box_vectors)
        if cutoff <= 0:
            raise ValueError('cutoff must be positive')
        self.exclusions = self.exclusions.to(positions.device)
        if emulate is not None:
            return pme_direct_fused(positions, charges, box_vectors, self.exclusions, cutoff, self.alpha, self.coulomb, emulate)
        import torch.distributed as dist
        shard = (dist.get_rank(group), dist.get_world_size(group))
        e = pme_direct_fused(positions, charges, box_vectors, self.exclusions, cutoff, self.alpha, self.coulomb, shard)
        return _AllReduceSum.apply(e, group)

    def energy_and_derivatives(self, positions: Tensor, charges: Tensor, cutoff: float, box_vectors: Tensor, group=None, world=None):
        """Total PME energy (direct + reciprocal + self term) with dE/dpositions and dE/dcharges in ONE call, without autograd: the
        form an MD engine needs (forces = -dE/dx), and the form that shards.  Returns (energy [] , dE/dx [atoms, 3], dE/dq [atoms]).

        When torch.distributed is initialised (ranks of `group`, default group if None; pass world=1 for a rank-local evaluation) ONE
        box is evaluated by all N ranks; positions, charges and the results are replicated:
          direct space      rank r evaluates the centres of slab r of the cell-sorted atoms (fused kernel, no halo, no atomics);
          reciprocal space  rank r spreads ITS block of atoms, the charge grids are summed by one all-reduce (8.4 MB at 128^3), every
                            rank solves the same grid (FFT, convolution, energy) and interpolates the derivatives of its own atoms;
          assembly          ONE all-reduce of the packed [atoms, 4] derivatives (direct + reciprocal + self term together) and one
                            of the two energy scalars.
        Same arithmetic as compute_direct + compute_reciprocal + backward (tests/test_neighbors_pme_gpu.py)."""
        import torch.distributed as dist
        self._validate(positions, charges, box_vectors)
        if cutoff <= 0:
            raise ValueError('cutoff must be positive')
        pos, q, box = _f32(positions, "positions"), _f32(charges, "charges"), _f32(box_vectors, "box_vectors")
        dev = pos.device
        self.exclusions = self.exclusions.to(dev)
        for i in range(3):
            self.moduli[i] = self.moduli[i].to(device=dev, dtype=torch.float32).contiguous()
        if world is None:
            world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        n = q.shape[0]
        ex = self.exclusions.contiguous()
        gx, gy, gz = self.gridx, self.gridy, self.gridz
        packed = torch.empty((n, 4), dtype=torch.float32, device=dev)        # dE/dx | dE/dq
        pos_deriv = torch.empty((n, 3), dtype=torch.float32, device=dev)
        charge_deriv = torch.empty((n,), dtype=torch.float32, device=dev)
        energies = torch.empty((2,), dtype=torch.float32, device=dev)       # direct (this rank's slab), reciprocal (replicated)
        stream = current_stream(dev)
        with torch.cuda.device(dev):
            check(lib.nnpops_pme_direct_fused(ptr(pos), ptr(q), ptr(box), ptr(ex) if ex.numel() else None, n, ex.shape[1] if ex.dim() == 2 else 0,
                                              float(cutoff), float(self.alpha), float(self.coulomb), rank, world, ptr(energies[0:1]),
                                              ptr(pos_deriv), ptr(charge_deriv), stream))
            packed[:, :3] = pos_deriv
            packed[:, 3] = charge_deriv
            lo, hi = shard_range(n, rank, world)
            grid = torch.empty((gx, gy, gz), dtype=torch.float32, device=dev)
            pl, ql = pos[lo:hi].contiguous(), q[lo:hi].contiguous()
            check(lib.nnpops_pme_spread(ptr(pl), ptr(ql), ptr(box), hi - lo, gx, gy, gz, self.order, float(self.coulomb), ptr(grid), stream))
            if world > 1:
                dist.all_reduce(grid, group=group)
            recip = torch.empty((gx, gy, gz // 2 + 1, 2), dtype=torch.float32, device=dev)
            check(lib.nnpops_pme_solve(ptr(grid), ptr(box), gx, gy, gz, float(self.alpha), ptr(self.moduli[0]), ptr(self.moduli[1]),
                                       ptr(self.moduli[2]), ptr(energies[1:2]), ptr(recip), stream))
            if hi > lo:
                pd = torch.empty((hi - lo, 3), dtype=torch.float32, device=dev)
                cd = torch.empty((hi - lo,), dtype=torch.float32, device=dev)
                check(lib.nnpops_pme_reciprocal_backward(ptr(pl), ptr(ql), ptr(box), hi - lo, gx, gy, gz, self.order, float(self.coulomb),
                                                         ptr(recip), ptr(pd), ptr(cd), stream))
                self_scale = self.coulomb * self.alpha / math.sqrt(math.pi)
                packed[lo:hi, :3] += pd
                packed[lo:hi, 3] += cd - (2.0 * self_scale) * ql           # d/dq of the self term -k alpha / sqrt(pi) sum q^2
        if world > 1:
            dist.all_reduce(packed, group=group)
            e_direct = energies[0:1].clone()
            dist.all_reduce(e_direct, group=group)
        else:
            e_direct = energies[0:1]
        self_energy = -torch.sum(q * q) * (self.coulomb * self.alpha / math.sqrt(math.pi))
        energy = (e_direct[0] + energies[1] + self_energy)
        return energy, packed[:, :3], packed[:, 3]

    def compute_reciprocal(self, positions: Tensor, charges: Tensor, box_vectors: Tensor):
        """Reciprocal-space energy including the self term (pme.py:167-196)."""
        self._validate(positions, charges, box_vectors)
        for i in range(3):
            self.moduli[i] = self.moduli[i].to(positions.device)
        self_energy = -torch.sum(charges ** 2) * self.coulomb * self.alpha / math.sqrt(math.pi)
        return self_energy + pme_reciprocal(positions, charges, box_vectors, self.gridx, self.gridy, self.gridz, self.order, self.alpha,
                                            self.coulomb, self.moduli[0], self.moduli[1], self.moduli[2])
