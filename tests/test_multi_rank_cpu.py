"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo group, conformer sharding and max-over-ranks timing."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.shard_conformers(64, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    slow = bench.max_over_ranks(10.0 + 5.0 * rank, dist, torch.device("cpu"))
    dist.barrier()
    if rank == 0:
        out.put((gathered, slow))
    dist.destroy_process_group()


def test_sharding_and_timing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, slow = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(gathered[0] + gathered[1]) == list(range(64))          # every conformer exactly once
    assert len(gathered[0]) == len(gathered[1]) == 32 and not set(gathered[0]) & set(gathered[1])
    assert slow == 15.0                                                   # the slowest rank defines the step time


def test_single_rank_sharding_is_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.shard_conformers(5, 0, 1) == [0, 1, 2, 3, 4]
    assert bench.max_over_ranks(3.5, None, torch.device("cpu")) == 3.5


class _FakeLocal:
    """Stand-in for FusedANI(shard=(rank, world)) on CPU: a pair potential whose centres are split with the library's own rule."""

    def __init__(self, rank, world, n):
        from nnpops_b200.OptimizedTorchANI import shard_mask
        self.mask = torch.from_numpy(shard_mask(n, rank, world))

    def energy_and_gradient(self, positions, cell=None):
        with torch.enable_grad():                                                # also called inside an autograd.Function
            pos = positions.detach().clone().requires_grad_(True)
            d = (pos[:, None, :] - pos[None, :, :]).norm(dim=-1) + torch.eye(len(pos))
            e_atom = (torch.exp(-d) * (1 - torch.eye(len(pos)))).sum(1)      # energy of centre i: depends on all atoms
            e = (e_atom * self.mask).sum().reshape(1)
            (g,) = torch.autograd.grad(e.sum(), pos)
        return e.detach().float(), g.float()


def _box_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from nnpops_b200.OptimizedTorchANI import ShardedFusedANI
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37
    pos = torch.linspace(0, 5, 3 * n).reshape(n, 3).float() ** 1.3
    m = ShardedFusedANI(local_factory=lambda r, w: _FakeLocal(r, w, n))
    e, g = m.energy_and_gradient(pos)
    p2 = pos.clone().requires_grad_(True)
    m(p2).sum().backward()                                                  # the autograd face gives the same total gradient
    ref = _FakeLocal(0, 1, n).energy_and_gradient(pos)
    out.put((rank, float((e - ref[0]).abs().max()), float((g - ref[1]).abs().max()), float((p2.grad - ref[1]).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_box_allreduce_world2():
    """One box sharded over 2 ranks (gloo): partial energies / gradients of the owned centres sum to the single-rank result on
    EVERY rank; the ownership masks partition the atoms."""
    sys.path.insert(0, ROOT)
    from nnpops_b200.OptimizedTorchANI import shard_mask
    masks = [shard_mask(37, r, 3) for r in range(3)]
    assert (sum(m.astype(int) for m in masks) == 1).all()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_box_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for _, de, dg, dga in res:
        assert de < 1e-4 and dg < 1e-5 and dga < 1e-5


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    import importlib
    import types
    # the PME module needs the CUDA library at import: test the pure helper through its source
    src = open(os.path.join(ROOT, "nnpops_b200", "pme", "pme.py")).read()
    start = src.index("def shard_range"); end = src.index("def pme_spread")
    ns = {}
    exec(src[start:end], ns)
    for n, w in ((200000, 8), (7, 3), (5, 8), (0, 2)):
        blocks = [ns["shard_range"](n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(b[1] == blocks[i + 1][0] for i, b in enumerate(blocks[:-1])) and all(lo <= hi for lo, hi in blocks)


# ------------------------------------------------------------------------------------------------ spatial decomposition with halos
class _FakePeriodicLocal:
    """Stand-in for FusedANI(owned=mask) on CPU: a short-ranged pair potential in the periodic box, centre-owned -- the energy of
    an owned centre depends on every local atom within the cutoff (minimum image), exactly the data dependence of the AEV."""

    def __init__(self, owned_mask, rc):
        self.mask = torch.tensor(owned_mask, dtype=torch.float64)
        self.rc = rc

    def energy_and_gradient(self, positions, cell):
        with torch.enable_grad():
            pos = positions.detach().double().clone().requires_grad_(True)
            L = torch.diagonal(cell.double())
            d = pos[:, None, :] - pos[None, :, :]
            d = d - torch.round(d / L) * L
            r = d.norm(dim=-1) + torch.eye(len(pos))
            phi = torch.where(r < self.rc, (1 - r / self.rc) ** 2, torch.zeros_like(r)) * (1 - torch.eye(len(pos)))
            e = (phi.sum(1) * self.mask).sum().reshape(1)
            (g,) = torch.autograd.grad(e.sum(), pos)
        return e.detach().float(), g.float()


def _halo_problem():
    import numpy as np
    rng = np.random.default_rng(12)
    n, L, rc = 400, 12.0, 2.5
    pos = rng.uniform(-3.0, L + 3.0, (n, 3)).astype(np.float32)      # some atoms outside the primary cell: the plan wraps them
    return n, L, rc, pos


def test_halo_plan_covers_every_neighbour():
    """Pure numpy: bricks partition the atoms; every atom within the halo distance of an owned atom is local to the owner's rank;
    the send lists mirror the ghost lists."""
    import numpy as np
    sys.path.insert(0, ROOT)
    src = open(os.path.join(ROOT, "nnpops_b200", "halo.py")).read()
    ns = {}
    exec(src[:src.index("class HaloBoxANI")], ns)          # the plan is pure numpy; importing the package needs the CUDA library
    n, L, rc, pos = _halo_problem()
    assert ns["brick_grid"](8) == (2, 2, 2) and sorted(ns["brick_grid"](4)) == [1, 2, 2] and sorted(ns["brick_grid"](2)) == [1, 1, 2]
    for grid in ((2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2)):
        plan = ns["HaloPlan"](pos, [L, L, L], rc, grid)
        assert sorted(np.concatenate(plan.owned).tolist()) == list(range(n))
        d = pos[:, None, :].astype(np.float64) - pos[None, :, :]
        d -= np.round(d / L) * L
        close = np.sqrt((d * d).sum(-1)) < rc
        for r in range(plan.world):
            local = set(plan.local_atoms(r).tolist())
            assert len(local) == len(plan.local_atoms(r))                      # no duplicates: ghosts are unique atoms
            for i in plan.owned[r]:
                assert set(np.nonzero(close[i])[0].tolist()) <= local
            for p in range(plan.world):
                sent = plan.owned[r][plan.send_idx[r][p]]
                assert np.array_equal(sent, plan.ghost_atom[p][plan.ghost_range[p][r]])
    assert ns["HaloPlan"].still_valid(pos + 0.1, pos, 0.5) and not ns["HaloPlan"].still_valid(pos + 0.3, pos, 0.5)


def _halo_worker(rank, world, port, out):
    import numpy as np
    sys.path.insert(0, ROOT)
    from nnpops_b200.halo import HaloBoxANI
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, L, rc, pos = _halo_problem()
    cell = torch.diag(torch.tensor([L, L, L]))
    m = HaloBoxANI(1, rc, rc, [], [], [], [], [], [], np.zeros(n, np.int32), None, pos, [L, L, L],
                   local_factory=lambda sp, mask: _FakePeriodicLocal(mask, rc))
    own = m.plan.owned[rank]
    e, g = m.energy_and_gradient(torch.tensor(pos[own]), cell)
    ref_e, ref_g = _FakePeriodicLocal(np.ones(n), rc).energy_and_gradient(torch.tensor(pos), cell)
    out.put((rank, float((e - ref_e).abs().max() / ref_e.abs().max()), float((g - ref_g[own]).abs().max() / ref_g.abs().max()), len(own),
             m.n_ghost))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_box_exchange_world2():
    """One periodic box over 2 ranks (gloo) by bricks with ghost halos: positions go out, ghost gradient rows come back, the
    energies are summed -- every rank ends with the total energy and the exact gradient of ITS atoms."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sum(r[3] for r in res) == 400 and all(0 < r[4] < 400 for r in res)
    for _, de, dg, _, _ in res:
        assert de < 1e-6 and dg < 1e-5
