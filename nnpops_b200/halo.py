"""ONE periodic box over several GPUs by spatial domain decomposition with ghost-atom halos (variant i; the
reference is single-device, src/pytorch/SymmetryFunctions.cpp:129).

The box is cut into a grid of bricks, one per rank.  A rank evaluates the AEVs and networks of the atoms inside its brick (its
CENTRES) and needs, as neighbour candidates only, the atoms of other bricks within the radial cutoff (+ skin) of its brick: its
GHOSTS.  Per evaluation:

    forward  halo : every rank sends the positions of those of its atoms that are ghosts elsewhere          (grouped NCCL send/recv)
    local         : the fused model over centres + ghosts (nnpops_ani_model_create_owned) in the REAL periodic box: the ghosts keep
                    their original coordinates and the kernels form the minimum-image displacement exactly as on one GPU, so cutoff
                    decisions and AEVs are bit-identical to the unsharded evaluation; cell list, neighbour rows, AEV, network and
                    backward touch n_brick + n_ghost atoms, not N
    reverse halo  : the gradient rows a rank accumulated on its ghosts go back to the owners and are added    (grouped NCCL send/recv)
    energy        : one all-reduce of a scalar

`HaloPlan` is the host-side bookkeeping (pure numpy, identical on every rank because it is computed from the same reference
positions); it stays valid while no atom has moved further than skin / 2 from the positions it was built for, like a Verlet list.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np


def brick_grid(world: int) -> Tuple[int, int, int]:
    """Factorisation of `world` into bricks per dimension with the smallest brick surface for a cubic box
    (8 -> 2x2x2, 4 -> 2x2x1, 2 -> 2x1x1, 6 -> 3x2x1)."""
    best, best_cost = (world, 1, 1), None
    for gx in range(1, world + 1):
        if world % gx:
            continue
        for gy in range(1, world // gx + 1):
            if (world // gx) % gy:
                continue
            gz = world // gx // gy
            a, b, c = 1.0 / gx, 1.0 / gy, 1.0 / gz                    # brick edges in units of the box edge
            cost = a * b + b * c + a * c
            if best_cost is None or cost < best_cost - 1e-12 or (abs(cost - best_cost) <= 1e-12 and (gx, gy, gz) > best):
                best, best_cost = (gx, gy, gz), cost
    return best


class HaloPlan:
    """Ownership, ghost lists and exchange lists of every rank for an ORTHORHOMBIC periodic box.

    positions: [N, 3] reference positions (any periodic images); box: [3] edge lengths; halo: radial cutoff + skin; grid: bricks per
    dimension.  For rank r:
        owned[r]        atom indices inside brick r (ascending)
        ghost_atom[r]   atom index of every ghost (an atom of another brick within `halo` of brick r, measured per dimension with
                        periodic wrap-around), grouped by source rank, ascending inside a group
        ghost_range[r][p]  slice of ghost_atom[r] that comes from rank p
        send_idx[r][p]  positions in owned[r] of the atoms rank r sends to rank p, in the order rank p stores them
    """

    def __init__(self, positions: np.ndarray, box: Sequence[float], halo: float, grid: Tuple[int, int, int]):
        pos = np.asarray(positions, np.float64)
        self.box = np.asarray(box, np.float64).reshape(3)
        self.grid = tuple(int(g) for g in grid)
        self.world = self.grid[0] * self.grid[1] * self.grid[2]
        self.halo = float(halo)
        n = len(pos)
        g = np.asarray(self.grid)
        width = self.box / g
        for d in range(3):
            if self.grid[d] > 1 and width[d] < self.halo:
                raise ValueError("bricks must be at least one halo width wide in every cut dimension")
        wrapped = pos - np.floor(pos / self.box) * self.box
        cell = np.minimum((wrapped / width).astype(np.int64), g - 1)
        self.owner = (cell[:, 0] * self.grid[1] + cell[:, 1]) * self.grid[2] + cell[:, 2]
        self.owned: List[np.ndarray] = [np.nonzero(self.owner == r)[0] for r in range(self.world)]
        self.ghost_atom, self.ghost_range, self.send_idx = [], [], []
        for r in range(self.world):
            c = np.array([r // (self.grid[1] * self.grid[2]), (r // self.grid[2]) % self.grid[1], r % self.grid[2]])
            lo, hi = c * width, (c + 1) * width
            near = self.owner != r
            for d in range(3):
                if self.grid[d] == 1:
                    continue                                   # an uncut dimension: the brick spans the whole period
                x = wrapped[:, d]
                dist = np.full(n, np.inf)
                for k in (-1.0, 0.0, 1.0):                     # distance to the interval [lo, hi] over the periodic images of the atom
                    xs = x + k * self.box[d]
                    dist = np.minimum(dist, np.maximum(np.maximum(lo[d] - xs, xs - hi[d]), 0.0))
                near &= dist < self.halo
            atoms = np.nonzero(near)[0]
            src = self.owner[atoms]
            order = np.lexsort((atoms, src))
            atoms, src = atoms[order], src[order]
            self.ghost_atom.append(atoms)
            bounds = np.searchsorted(src, np.arange(self.world + 1))
            self.ghost_range.append([slice(int(bounds[p]), int(bounds[p + 1])) for p in range(self.world)])
        for r in range(self.world):
            pos_in_owned = np.full(n, -1, np.int64)
            pos_in_owned[self.owned[r]] = np.arange(len(self.owned[r]))
            self.send_idx.append([pos_in_owned[self.ghost_atom[p][self.ghost_range[p][r]]] for p in range(self.world)])

    def local_atoms(self, r: int) -> np.ndarray:
        """Atom index of every local atom of rank r: centres first, then ghosts."""
        return np.concatenate([self.owned[r], self.ghost_atom[r]])

    @staticmethod
    def still_valid(positions: np.ndarray, reference: np.ndarray, skin: float) -> bool:
        """True while no atom has moved more than skin / 2 from the positions the plan was built for."""
        d = np.asarray(positions, np.float64) - np.asarray(reference, np.float64)
        return float(np.sqrt((d * d).sum(1)).max()) <= 0.5 * skin


class HaloBoxANI:
    """FusedANI for ONE periodic orthorhombic box sharded over the ranks of a torch.distributed group by bricks with ghost halos.

    Every rank constructs it with the same arguments (species of ALL atoms and reference positions of ALL atoms: used once, on the
    host, to build the plan).  `energy_and_gradient(pos_owned, cell)` takes the CURRENT positions of this rank's own atoms (device
    [n_owned, 3], in the order of `plan.owned[rank]`) and the box, and returns the total energy [1] and dE/dx of this rank's atoms.

    local_factory(species_local, owned_mask) -> object with energy_and_gradient(positions[n_local, 3], cell): lets the host logic run
    without a GPU (tests).  rank / world / plan may be given explicitly to emulate several ranks in one process; the two exchange
    phases then take an `exchange(send, recv)` callable instead of the process group."""

    def __init__(self, num_species, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, species, networks, positions, box, skin: float = 0.0,
                 grid: Optional[Tuple[int, int, int]] = None, group=None, mlp_impl: str = "tcgen05", device: str = "cuda",
                 local_factory=None, rank: Optional[int] = None, world: Optional[int] = None, plan: Optional[HaloPlan] = None, **kwargs):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        live = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if live else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if live else 1)
        edges = np.diag(np.asarray(box, np.float64).reshape(3, 3)) if np.asarray(box).size == 9 else np.asarray(box, np.float64)
        self.plan = plan if plan is not None else HaloPlan(positions, edges, float(Rcr) + float(skin), grid or brick_grid(self.world))
        assert self.plan.world == self.world, "the brick grid must have one brick per rank"
        r = self.rank
        local = self.plan.local_atoms(r)
        self.n_owned, self.n_ghost = len(self.plan.owned[r]), len(self.plan.ghost_atom[r])
        sp = np.asarray(species, np.int32)[local]
        owned_mask = np.zeros(len(local), np.uint8)
        owned_mask[:self.n_owned] = 1
        if local_factory is not None:
            self.local = local_factory(sp, owned_mask)
            self.device = torch.device("cpu")
        else:
            from .OptimizedTorchANI import FusedANI
            self.local = FusedANI(num_species, Rcr, Rca, EtaR, ShfR, EtaA, Zeta, ShfA, ShfZ, sp, networks, mlp_impl=mlp_impl, device=device,
                                  owned=owned_mask, **kwargs)
            self.device = torch.device(device)
        self.send_idx = [torch.tensor(np.ascontiguousarray(ix), dtype=torch.int64, device=self.device) for ix in self.plan.send_idx[r]]
        self.ghost_range = self.plan.ghost_range[r]
        self.local_pos = torch.empty((self.n_owned + self.n_ghost, 3), dtype=torch.float32, device=self.device)
        self.recv_pos = [torch.empty((s.stop - s.start, 3), dtype=torch.float32, device=self.device) for s in self.ghost_range]
        self.recv_grad = [torch.empty((len(ix), 3), dtype=torch.float32, device=self.device) for ix in self.plan.send_idx[r]]
        # all-to-all form of the two halo phases (NCCL): ONE gather + ONE collective each way.  The ghosts are stored grouped by source
        # rank, so the receive buffer of the forward phase IS the ghost block of local_pos and the send buffer of the reverse phase IS
        # the ghost block of the local gradient.
        self.send_all = torch.cat(self.send_idx) if self.world > 0 else None
        self.send_splits = [len(ix) for ix in self.plan.send_idx[r]]
        self.ghost_splits = [s.stop - s.start for s in self.ghost_range]
        self.recv_grad_all = torch.empty((int(sum(self.send_splits)), 3), dtype=torch.float32, device=self.device)
        self.halo_bytes_forward = 12 * sum(len(ix) for ix in self.plan.send_idx[r])      # sent by this rank per evaluation
        self.halo_bytes_reverse = 12 * self.n_ghost
        self.graph = None

    def _exchange(self, send, recv):
        """recv[p] <- rank p's send[self.rank], for every peer, as ONE grouped NCCL launch (ncclGroupStart ... ncclSend/ncclRecv ...)."""
        dist = self.dist
        ops = []
        for p in range(self.world):
            if p == self.rank:
                continue
            if recv[p].numel():
                ops.append(dist.P2POp(dist.irecv, recv[p], p, group=self.group))
            if send[p].numel():
                ops.append(dist.P2POp(dist.isend, send[p], p, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def assemble(self, pos_owned, exchange=None):
        """Forward halo: local positions [n_local, 3] = own atoms followed by the ghosts, all in their original coordinates."""
        self.local_pos[:self.n_owned] = pos_owned
        send = [pos_owned.index_select(0, ix) for ix in self.send_idx]
        (exchange or self._exchange)(send, self.recv_pos)
        ghosts = self.local_pos[self.n_owned:]
        for p in range(self.world):
            if p != self.rank and self.recv_pos[p].numel():
                ghosts[self.ghost_range[p]] = self.recv_pos[p]
        return self.local_pos

    def scatter_back(self, g_local, exchange=None):
        """Reverse halo: dE/dx of this rank's atoms = its own rows + the rows other ranks accumulated on their ghost copies."""
        g_own = g_local[:self.n_owned].clone()
        g_ghost = g_local[self.n_owned:]
        send = [g_ghost[self.ghost_range[p]].contiguous() for p in range(self.world)]
        (exchange or self._exchange)(send, self.recv_grad)
        for p in range(self.world):
            if p != self.rank and len(self.send_idx[p]):
                g_own.index_add_(0, self.send_idx[p], self.recv_grad[p])
        return g_own

    def _use_all_to_all(self):
        return (self.world > 1 and self.dist.is_initialized() and self.dist.get_backend(self.group) == "nccl")

    def _step(self, pos_owned, cell):
        dist = self.dist
        if self._use_all_to_all():
            self.local_pos[:self.n_owned] = pos_owned
            send = pos_owned.index_select(0, self.send_all)
            dist.all_to_all_single(self.local_pos[self.n_owned:], send, self.ghost_splits, self.send_splits, group=self.group)
            e, g = self.local.energy_and_gradient(self.local_pos, cell)
            dist.all_to_all_single(self.recv_grad_all, g[self.n_owned:].contiguous(), self.send_splits, self.ghost_splits, group=self.group)
            g_own = g[:self.n_owned].clone()
            g_own.index_add_(0, self.send_all, self.recv_grad_all)
        else:
            local = self.assemble(pos_owned)
            e, g = self.local.energy_and_gradient(local, cell)
            g_own = self.scatter_back(g)
        if self.world > 1 and dist.is_initialized():
            e = e.clone()
            dist.all_reduce(e, op=dist.ReduceOp.SUM, group=self.group)
        return e, g_own

    def capture(self, cell):
        """Capture the WHOLE step -- halo exchange (NCCL), brick-local model, reverse halo, energy all-reduce -- into one CUDA graph;
        afterwards energy_and_gradient copies the positions into the graph's input buffer and replays it (every rank must call this
        at the same point: the collectives are captured collectively).  At eight bricks the step is a few dozen short kernels and
        three small collectives; launched one by one the host is the bottleneck."""
        torch = self.torch
        self.static_pos = torch.zeros((self.n_owned, 3), dtype=torch.float32, device=self.device)
        self.static_cell = cell.detach().clone()
        return self

    def _capture_now(self, pos_owned):
        torch = self.torch
        self.static_pos.copy_(pos_owned)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                                   # warm-up: allocations, communicator set-up, one-time attributes
                self._step(self.static_pos, self.static_cell)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_e, self.static_g = self._step(self.static_pos, self.static_cell)

    def energy_and_gradient(self, pos_owned, cell):
        if getattr(self, "static_pos", None) is not None:
            if self.graph is None:
                self._capture_now(pos_owned)                     # the first call after capture() records the graph with real positions
            self.static_pos.copy_(pos_owned)
            self.graph.replay()
            return self.static_e, self.static_g
        return self._step(pos_owned, cell)
