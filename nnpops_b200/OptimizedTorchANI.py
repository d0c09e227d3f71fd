"""ANI energy + forces on the B200 kernels.

``FusedANI`` is the scalable module behind the BASELINE.json headline metric: one C-ABI call evaluates
AEV -> species-grouped ensemble MLP -> dE/dx for a whole system (src/pytorch/OptimizedTorchANI.py:45-54 plus the autograd
backward of the reference, without per-atom weight replication and without materialising anything on the host).
``OptimizedTorchANI`` keeps the reference's constructor (a TorchANI-like model + atomic numbers) and forward signature and is
built on FusedANI; the TorchANI objects are duck-typed, so no torchani import is needed.
"""
import ctypes as C
from typing import List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import lib, check, ptr, current_stream
from .SymmetryFunctions import function_tables


class SpeciesEnergies(NamedTuple):
    species: Tensor
    energies: Tensor


def pack_network_params(networks: Sequence[Sequence[Sequence[Tuple[np.ndarray, np.ndarray]]]]):
    """networks[s][e][l] = (W [out, in], b [out]) -> (dims int32 [S, L+1], flat float32 params) in the C-ABI order."""
    S = len(networks)
    L = len(networks[0][0])
    dims = np.zeros((S, L + 1), np.int32)
    chunks = []
    for s in range(S):
        for e, member in enumerate(networks[s]):
            assert len(member) == L
            for l, (W, b) in enumerate(member):
                W = np.ascontiguousarray(W, np.float32)
                b = np.ascontiguousarray(b, np.float32).reshape(-1)
                if e == 0:
                    dims[s, l], dims[s, l + 1] = W.shape[1], W.shape[0]
                assert W.shape == (dims[s, l + 1], dims[s, l]) and b.shape[0] == W.shape[0]
                chunks += [W.reshape(-1), b]
    return dims, np.concatenate(chunks)


class _EnergyGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, positions, cell):
        energy, grad = module._evaluate(positions, cell)
        ctx.save_for_backward(grad)
        return energy

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_energy):
        (grad,) = ctx.saved_tensors
        return None, grad * grad_energy, None


class FusedANI(torch.nn.Module):
    """AEV + ensemble of per-species networks + analytic position gradient in one fused pipeline.

    species: int sequence [N] with values in [0, num_species); networks[s][e][l] = (W, b) numpy arrays.
    mlp_impl: "tcgen05" (tensor cores) or "simt" (fp32 validation path).
    shard: (rank, world) -- one box sharded over `world` GPUs: this instance evaluates only the centres i with i % world == rank
    (all atoms stay neighbour candidates); energy and gradient are then PARTIAL and must be summed over the ranks
    (ShardedFusedANI does that with one all-reduce).
    """

    def __init__(self, num_species: int, Rcr: float, Rca: float, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, species: Sequence[int],
                 networks, mlp_impl: str = "tcgen05", device: str = "cuda", max_radial_neighbors: int = 0,
                 max_angular_neighbors: int = 0, shard: Tuple[int, int] = (0, 1), owned: Optional[Sequence[int]] = None, skin: float = 0.0):
        super().__init__()
        self.num_atoms = len(species)
        self.num_species = int(num_species)
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
        radial_fn, angular_fn = function_tables(EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ)
        dims, params = pack_network_params(networks)
        sp = np.ascontiguousarray(species, np.int32)
        h = C.c_void_p()
        with torch.cuda.device(self.device_):
            if owned is not None:
                # spatial decomposition (nnpops_b200.halo.HaloBoxANI): the atoms are a brick's own atoms + its ghosts; only the atoms with
                # owned[i] != 0 are centres, energy and gradient are this rank's PARTIAL results
                mask = np.ascontiguousarray(owned, np.uint8)
                assert mask.shape == (self.num_atoms,)
                check(lib.nnpops_ani_model_create_owned(C.byref(h), self.num_atoms, self.num_species, float(Rcr), float(Rca), ptr(sp),
                                                        len(radial_fn), ptr(radial_fn), len(angular_fn), ptr(angular_fn), len(networks[0]),
                                                        dims.shape[1] - 1, ptr(dims), ptr(params), {"simt": 0, "tcgen05": 1}[mlp_impl],
                                                        max_radial_neighbors, max_angular_neighbors, ptr(mask)))
            else:
                check(lib.nnpops_ani_model_create_sharded(C.byref(h), self.num_atoms, self.num_species, float(Rcr), float(Rca), ptr(sp),
                                                          len(radial_fn), ptr(radial_fn), len(angular_fn), ptr(angular_fn), len(networks[0]),
                                                          dims.shape[1] - 1, ptr(dims), ptr(params), {"simt": 0, "tcgen05": 1}[mlp_impl],
                                                          max_radial_neighbors, max_angular_neighbors, int(shard[0]), int(shard[1])))
        self.shard = (int(shard[0]), int(shard[1]))
        self._h = h
        if skin > 0:
            self.set_skin(skin)
        self.mlp_impl = mlp_impl
        self.aev_length = int(dims[0, 0])

    def __del__(self):
        h = self.__dict__.get("_h")
        if h:
            self.__dict__["_h"] = None      # plain dict access: Module.__setattr__ is not usable during interpreter shutdown
            try:
                lib.nnpops_ani_model_destroy(h)
            except Exception:               # noqa: BLE001  (library already unloaded at exit)
                pass

    def _evaluate(self, positions: Tensor, cell: Optional[Tensor]):
        if positions.dtype != torch.float32 or positions.dim() != 2 or positions.shape != (self.num_atoms, 3):
            raise RuntimeError('"positions" has to be a float32 tensor of shape (%d, 3)' % self.num_atoms)
        if positions.device.type != "cuda":
            raise RuntimeError("nnpops_b200 runs on CUDA devices only (no CPU fallback)")
        self._raise_if_overflowed()
        pos = positions.detach().contiguous()
        box = None
        if cell is not None:
            if cell.dtype != torch.float32 or cell.shape != (3, 3) or cell.device != positions.device:
                raise RuntimeError('"cell" has to be a float32 (3, 3) tensor on the device of "positions"')
            box = cell.detach().contiguous()
        energy = torch.empty(1, dtype=torch.float32, device=pos.device)
        grad = torch.empty_like(pos)
        with torch.cuda.device(pos.device):
            check(lib.nnpops_ani_model_energy_grad(self._h, ptr(pos), ptr(box), ptr(energy), ptr(grad), current_stream(pos.device)))
        return energy, grad

    def set_skin(self, skin: float):
        """Verlet skin of the neighbour search in the length unit of the positions (0 = rebuild every call).  For time-stepping callers:
        the candidate rows are reused until an atom has moved more than skin / 2; the rows the kernels see stay exact."""
        with torch.cuda.device(self.device_):
            check(lib.nnpops_ani_model_set_skin(self._h, float(skin)))

    def skin_stats(self):
        """(calls that rebuilt the candidate rows, calls that reused them) since set_skin; synchronises."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        check(lib.nnpops_ani_model_skin_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _raise_if_overflowed(self):
        """Never blocks: a neighbour row that overflowed in an EARLIER evaluation is reported now (the reference has no neighbour
        limit; this path must stay asynchronous, so it cannot check the evaluation it is about to launch)."""
        f, r, a = C.c_int(0), C.c_int(0), C.c_int(0)
        check(lib.nnpops_ani_model_overflow_poll(self._h, C.byref(f), C.byref(r), C.byref(a)))
        if f.value:
            raise RuntimeError("nnpops_b200: a neighbour row overflowed in an earlier evaluation (capacities %d radial / %d angular): "
                               "results since then are wrong. Pass larger max_radial_neighbors / max_angular_neighbors." % (r.value, a.value))

    def energy_and_gradient(self, positions: Tensor, cell: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """(energy [1], dE/dx [N, 3]) without autograd bookkeeping."""
        return self._evaluate(positions, cell)

    def energy_and_gradient_host(self, positions: np.ndarray, cell: Optional[np.ndarray], energy_out: np.ndarray, grad_out: np.ndarray):
        """Host-buffer entry point (copies inside the call): the end-to-end path measured by bench.py."""
        check(lib.nnpops_ani_model_energy_grad_host(self._h, ptr(positions), ptr(cell), ptr(energy_out), ptr(grad_out),
                                                    current_stream(self.device_)))

    def forward(self, positions: Tensor, cell: Optional[Tensor] = None) -> Tensor:
        """positions [N, 3] -> energy [1]; a batch of conformers of the same system [B, N, 3] -> energies [B] (evaluated one after the
        other on the same stream; the reference rejects batches, SymmetryFunctions.py:110-111).  Forces by autograd."""
        if positions.dim() == 3:
            return torch.cat([_EnergyGrad.apply(self, positions[b], cell) for b in range(positions.shape[0])])
        return _EnergyGrad.apply(self, positions, cell)

    # -- introspection for tests / benchmarks
    def buffers(self):
        f, g, st = C.c_void_p(), C.c_void_p(), C.c_int()
        rows = np.zeros(self.num_atoms, np.int32)
        check(lib.nnpops_ani_model_buffers(self._h, C.byref(f), C.byref(g), C.byref(st), ptr(rows)))
        return f.value, g.value, st.value, rows

    def _read(self, which: int) -> Tensor:
        out = torch.empty((self.num_atoms, self.aev_length), dtype=torch.float32, device=self.device_)
        with torch.cuda.device(self.device_):
            check(lib.nnpops_ani_model_read_features(self._h, which, ptr(out), current_stream(self.device_)))
        return out

    def features(self) -> Tensor:
        """Copy of the AEV matrix of the last evaluation, atom order, [N, aev_length]."""
        return self._read(0)

    def feature_grad(self) -> Tensor:
        """Copy of dE/dAEV of the last evaluation, atom order, [N, aev_length]."""
        return self._read(1)

    def work(self):
        t, p, fl = C.c_longlong(0), C.c_longlong(0), C.c_double(0)
        check(lib.nnpops_ani_model_work(self._h, C.byref(t), C.byref(p), C.byref(fl), current_stream(self.device_)))
        full, active, ex = C.c_int(0), C.c_int(0), C.c_double(0)
        check(lib.nnpops_ani_model_info(self._h, C.byref(full), C.byref(active), C.byref(ex)))
        fused = C.c_int(0)
        check(lib.nnpops_ani_model_mlp_fused(self._h, C.byref(fused)))
        return {"triples": t.value, "radial_pairs": p.value, "mlp_flops_forward": fl.value, "aev_length": full.value,
                "active_features": active.value, "mlp_flops_forward_executed": ex.value, "mlp_fused": bool(fused.value)}

    STAGES = ("cells+rows", "radial_fwd", "angular_fwd", "mlp_fwd", "mlp_bwd", "radial_bwd", "angular_bwd")

    def timing_begin(self, max_steps: int):
        check(lib.nnpops_ani_model_timing_begin(self._h, int(max_steps)))

    def timing_end(self):
        """-> (dict stage -> mean ms per evaluation, evaluations recorded); synchronises."""
        ms = np.zeros(7, np.float32)
        steps = C.c_int(0)
        check(lib.nnpops_ani_model_timing_end(self._h, ptr(ms), C.byref(steps)))
        k = max(steps.value, 1)
        return {name: float(ms[i]) / k for i, name in enumerate(self.STAGES)}, steps.value

    def overflowed(self) -> int:
        f = C.c_int(0)
        check(lib.nnpops_ani_model_overflowed(self._h, C.byref(f)))
        return f.value


def shard_mask(num_atoms: int, rank: int, world: int) -> np.ndarray:
    """Ownership of the centres when one box is sharded over `world` ranks: atom i belongs to rank i mod world (the rule of
    nnpops_ani_model_create_sharded).  Interleaving balances species and density without looking at the geometry."""
    return (np.arange(num_atoms) % world) == rank


class ShardedFusedANI(torch.nn.Module):
    """One periodic box evaluated cooperatively by all ranks of a torch.distributed group (one process per GPU): every rank holds
    the full position array, evaluates the AEVs and networks of its own centres (FusedANI with shard=(rank, world)) and the
    partial energy and dE/dx are summed with ONE all-reduce of 4 + 12 N bytes (NCCL over NVLink) -- the exchange step of
    SURVEY.md section 8e, variant (ii).  ``energy_and_gradient`` returns the TOTAL energy [1] and gradient [N, 3] on every rank.
    The reference has no multi-GPU path."""

    def __init__(self, *args, group=None, local_factory=None, **kwargs):
        super().__init__()
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # local_factory(rank, world) -> object with energy_and_gradient(positions, cell): lets the host logic run without a GPU
        self.local = (local_factory(self.rank, self.world) if local_factory is not None
                      else FusedANI(*args, shard=(self.rank, self.world), **kwargs))
        self._packed = None

    def energy_and_gradient(self, positions: Tensor, cell: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        e, g = self.local.energy_and_gradient(positions, cell)
        if self.world == 1:
            return e, g
        import torch.distributed as dist
        n = g.shape[0]
        if self._packed is None or self._packed.numel() != 3 * n + 1:
            self._packed = torch.empty(3 * n + 1, dtype=torch.float32, device=g.device)
        self._packed[:3 * n] = g.reshape(-1)      # one message: gradient + energy
        self._packed[3 * n:] = e
        dist.all_reduce(self._packed, op=dist.ReduceOp.SUM, group=self.group)
        return self._packed[3 * n:].clone(), self._packed[:3 * n].reshape(n, 3).clone()

    def forward(self, positions: Tensor, cell: Optional[Tensor] = None) -> Tensor:
        return _ShardedEnergyGrad.apply(self, positions, cell)


class _ShardedEnergyGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, positions, cell):
        e, g = model.energy_and_gradient(positions, cell)
        ctx.save_for_backward(g)
        return e

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return None, g * grad_out, None


def _linear_layers(sequential) -> List[Tuple[np.ndarray, np.ndarray]]:
    """(W, b) of the nn.Linear members of a TorchANI per-species Sequential (items 0, 2, 4, 6: BatchedNN.py:55-59)."""
    out = []
    for layer in sequential:
        if hasattr(layer, "weight") and hasattr(layer, "bias") and getattr(layer, "weight").dim() == 2:
            out.append((layer.weight.detach().cpu().numpy(), layer.bias.detach().cpu().numpy()))
    return out


class OptimizedTorchANI(torch.nn.Module):
    """Reference-compatible wrapper (src/pytorch/OptimizedTorchANI.py:33-54): ``model`` is a TorchANI-like object with
    ``species_converter``, ``aev_computer``, ``neural_networks`` and ``energy_shifter``; ``atomicNumbers`` is [[Z...]].
    forward((species, coordinates[1, N, 3]), cell, pbc) -> SpeciesEnergies(species, energies[1])."""

    def __init__(self, model, atomicNumbers: Tensor, mlp_impl: str = "tcgen05") -> None:
        super().__init__()
        conv = model.species_converter
        aev = model.aev_computer
        conv_device = getattr(getattr(conv, "conv_tensor", None), "device", atomicNumbers.device)
        species = conv((atomicNumbers.to(conv_device), torch.empty(0))).species
        self.register_buffer("species", species)
        nets = model.neural_networks
        ensemble = list(nets) if isinstance(nets, torch.nn.ModuleList) else [nets]
        members = [list(m.values()) for m in ensemble]               # members[e][s] = Sequential
        num_species = aev.num_species
        networks = [[_linear_layers(members[e][s]) for e in range(len(members))] for s in range(num_species)]
        self.fused = FusedANI(num_species, aev.Rcr, aev.Rca, aev.EtaR[:, 0].tolist(), aev.ShfR[0, :].tolist(),
                              aev.EtaA[:, 0, 0, 0].tolist(), aev.Zeta[0, :, 0, 0].tolist(), aev.ShfA[0, 0, :, 0].tolist(),
                              aev.ShfZ[0, 0, 0, :].tolist(), species[0].tolist(), networks, mlp_impl=mlp_impl,
                              device=str(atomicNumbers.device) if atomicNumbers.device.type == "cuda" else "cuda")
        # self energies are a constant (EnergyShifter.py:42-52)
        self.register_buffer("self_energies", model.energy_shifter.sae(species.cpu()))

    def forward(self, species_coordinates: Tuple[Tensor, Tensor], cell: Optional[Tensor] = None,
                pbc: Optional[Tensor] = None) -> SpeciesEnergies:
        _, coordinates = species_coordinates
        if cell is not None:
            if pbc is None:
                raise ValueError('"pbc" has to be defined')
            if pbc.tolist() != [True, True, True]:
                raise ValueError('Only fully periodic systems are supported, i.e. pbc = [True, True, True]')
        # A batch is a set of conformers of the system this module was built for (one Holder = one list of species,
        # SymmetryFunctions.cpp:52-92): coordinates [B, N, 3] -> energies [B].  The reference stops at B = 1 (SymmetryFunctions.py:110-111).
        batch = coordinates.shape[0]
        energies = self.fused(coordinates if batch > 1 else coordinates[0], cell)
        return SpeciesEnergies(self.species.expand(batch, -1), energies.double() + self.self_energies.to(energies.device))


class ScriptableFusedANI(torch.nn.Module):
    """TorchScript-able form of FusedANI: the same fused pipeline behind the custom class ``NNPOpsFusedANI::Holder`` and the autograd
    op ``NNPOpsFusedANI::operation`` of libNNPOpsPyTorch.so, so a scripted / saved module (e.g. for openmm-torch's TorchForce) runs the
    species-grouped tensor-core path.  ``forward(positions[N, 3], cell[3, 3] or None) -> energy[1]``; forces by autograd."""

    def __init__(self, num_species: int, Rcr: float, Rca: float, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, species: Sequence[int], networks,
                 mlp_impl: str = "tcgen05"):
        super().__init__()
        from . import torch_ops
        torch_ops.load()
        dims, params = pack_network_params(networks)
        self.holder = torch.classes.NNPOpsFusedANI.Holder(int(num_species), float(Rcr), float(Rca), [float(x) for x in EtaR],
                                                          [float(x) for x in ShfR], [float(x) for x in EtaA], [float(x) for x in Zeta],
                                                          [float(x) for x in ShfA], [float(x) for x in ShfZ], [int(x) for x in species],
                                                          len(networks[0]), [int(x) for x in dims.reshape(-1)], torch.from_numpy(params),
                                                          {"simt": 0, "tcgen05": 1}[mlp_impl])

    def forward(self, positions: Tensor, cell: Optional[Tensor] = None) -> Tensor:
        if positions.dim() == 3:      # a batch of conformers of the same system: [B, N, 3] -> energies [B]
            out: List[Tensor] = []
            for b in range(positions.shape[0]):
                out.append(torch.ops.NNPOpsFusedANI.operation(self.holder, positions[b], cell))
            return torch.cat(out)
        return torch.ops.NNPOpsFusedANI.operation(self.holder, positions, cell)
