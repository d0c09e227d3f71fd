"""GPU parity tests for getNeighborPairs and PME against the numpy oracles and the reference's golden values.
Modelled on the reference's own suites (src/pytorch/neighbors/TestNeighbors.py, src/pytorch/pme/TestPme.py)."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from systems import lattice, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("neighbors_pme_oracle", os.path.join(HERE, "..", "oracle", "neighbors_pme_oracle.py"))
NP = importlib.util.module_from_spec(spec)
spec.loader.exec_module(NP)
G = json.load(open(os.path.join(HERE, "golden", "pme_openmm.json")))


def sort_pairs(nb, deltas, dist):
    nb = np.asarray(nb); keep = nb[0] >= 0
    nb, deltas, dist = nb[:, keep], np.asarray(deltas)[keep], np.asarray(dist)[keep]
    order = np.lexsort((nb[1], nb[0]))
    return nb[:, order], deltas[order], dist[order]


def run(pos, cutoff, max_num_pairs=-1, box=None, check_errors=False):
    from nnpops_b200.neighbors import getNeighborPairs
    p = torch.tensor(pos, device="cuda")
    b = torch.tensor(np.asarray(box), device="cuda", dtype=p.dtype) if box is not None else None
    nb, d, r, f = getNeighborPairs(p, cutoff, max_num_pairs, b, check_errors)
    return nb.cpu().numpy(), d.cpu().numpy(), r.cpu().numpy(), int(f.cpu()[0])


def test_doctest_examples():
    pos = np.array([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]], np.float32)
    nb, d, r, f = run(pos, 3.0)
    assert nb.tolist() == [[1, 2, 2], [0, 0, 1]] and r.tolist() == [1.0, 2.0, 1.0] and d.tolist() == [[1, 0, 0], [2, 0, 0], [1, 0, 0]]
    nb, d, r, f = run(pos, 1.5)
    assert nb.tolist() == [[1, -1, 2], [0, -1, 1]] and np.isnan(r[1]) and np.isnan(d[1]).all()
    nb, d, r, f = run(pos, 3.0, 6)
    assert sort_pairs(nb, d, r)[0].tolist() == [[1, 2, 2], [0, 0, 1]] and (nb[:, 3:] == -1).all() and np.isnan(r[3:]).all()
    nb, d, r, f = run(pos, 1.5, 6)
    assert sort_pairs(nb, d, r)[0].tolist() == [[1, 2], [0, 1]] and f == 2


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 10, 100, 1000])
@pytest.mark.parametrize("cutoff", [1, 10, 100])
@pytest.mark.parametrize("all_pairs", [True, False])
def test_neighbor_values(dtype, n, cutoff, all_pairs):
    """TestNeighbors.py:32-90: exact indices, deltas and distances, both output modes; bit-exact against the oracle."""
    rng = np.random.default_rng(n + cutoff)
    pos = (10 * rng.standard_normal((n, 3))).astype(dtype)
    ref = NP.neighbor_pairs(pos, cutoff)
    found = ref[3]
    max_pairs = -1 if all_pairs else max(found, 1)
    nb, d, r, f = run(pos, cutoff, max_pairs)
    assert f == found
    if all_pairs:
        assert np.array_equal(nb, ref[0])
        assert np.array_equal(np.isnan(r), np.isnan(ref[2]))
        ok = ~np.isnan(r)
        assert np.array_equal(d[ok], ref[1][ok]) and np.array_equal(r[ok], ref[2][ok])
    else:
        a = sort_pairs(nb, d, r); b = sort_pairs(*ref[:3])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert nb.shape[1] == max_pairs and (nb[:, found:] == -1).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["cubic", "triclinic"])
def test_neighbor_periodic(dtype, kind):
    """TestNeighbors.py:209-270: periodic boxes (reduced triclinic form), positions also outside the primary cell."""
    rng = np.random.default_rng(12)
    n, cutoff = 700, 2.5
    if kind == "cubic":
        box = np.diag([11.0, 12.0, 13.0])
    else:
        box = np.array([[12.0, 0, 0], [3.0, 11.0, 0], [-2.5, 4.0, 13.0]])
    pos = ((rng.uniform(-1.5, 2.5, (n, 3))) @ box).astype(dtype)
    ref = NP.neighbor_pairs(pos, cutoff, -1, box.astype(dtype))
    nb, d, r, f = run(pos, cutoff, ref[3] + 5, box.astype(dtype))
    assert f == ref[3]
    a = sort_pairs(nb, d, r); b = sort_pairs(*ref[:3])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_neighbor_grads(dtype):
    """TestNeighbors.py:92-140: gradients through deltas and distances vs a pure-torch construction."""
    from nnpops_b200.neighbors import getNeighborPairs
    n = 60
    rng = np.random.default_rng(2)
    pos_np = 3 * rng.standard_normal((n, 3))
    pos = torch.tensor(pos_np, dtype=dtype, device="cuda", requires_grad=True)
    nb, d, r, _ = getNeighborPairs(pos, 4.0, -1)
    mask = nb[0] >= 0
    (d[mask].pow(2).sum() + r[mask].sum() * 3).backward()
    ref = torch.tensor(pos_np, dtype=torch.float64, requires_grad=True)
    rows, cols = np.tril_indices(n, -1)
    dd = ref[rows] - ref[cols]
    rr = dd.norm(dim=1)
    m = rr <= 4.0
    (dd[m].pow(2).sum() + rr[m].sum() * 3).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert rel_err(pos.grad.cpu().numpy(), ref.grad.numpy()) < tol


def test_too_many_neighbors_semantics():
    """TestNeighbors.py:143-168."""
    from nnpops_b200.neighbors import getNeighborPairs
    pos = torch.zeros((4, 3), device="cuda")
    with pytest.raises(RuntimeError):
        getNeighborPairs(pos, cutoff=1, max_num_pairs=1, check_errors=True)
    nb, d, r, f = getNeighborPairs(pos, cutoff=1, max_num_pairs=1, check_errors=False)
    assert int(f) == 6
    with pytest.raises(RuntimeError):
        getNeighborPairs(pos, cutoff=1, max_num_pairs=5, check_errors=True)
    getNeighborPairs(pos, cutoff=1, max_num_pairs=6, check_errors=True)
    with pytest.raises(RuntimeError):
        getNeighborPairs(pos, cutoff=-1.0)
    with pytest.raises(RuntimeError):
        getNeighborPairs(pos, cutoff=1.0, max_num_pairs=0)


def test_neighbors_cuda_graph():
    """TestNeighbors.py:170-206: capturable with check_errors=False (no sync, shapes independent of the data)."""
    from nnpops_b200.neighbors import getNeighborPairs
    rng = np.random.default_rng(4)
    pos = torch.tensor(rng.standard_normal((50, 3)).astype(np.float32), device="cuda", requires_grad=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            nb, d, r, _ = getNeighborPairs(pos, 2.0, 2000)
            r[nb[0] >= 0].sum().backward()
            pos.grad = None
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        nb, d, r, f = getNeighborPairs(pos, 2.0, 2000)
    pos.data.copy_(torch.tensor(rng.standard_normal((50, 3)).astype(np.float32)))
    graph.replay()
    torch.cuda.synchronize()
    ref = NP.neighbor_pairs(pos.detach().cpu().numpy(), 2.0)
    a = sort_pairs(nb.cpu().numpy(), d.detach().cpu().numpy(), r.detach().cpu().numpy()); b = sort_pairs(*ref[:3])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_neighbors_full_size_properties():
    """BASELINE config 5 size: 200 000 atoms, 0.9 nm cutoff, periodic.  Size-independent checks: every pair has row > col, is
    unique, lies inside the cutoff, the count matches an independent KD-tree count, and padding is -1 / NaN."""
    from scipy.spatial import cKDTree
    n = 200000
    pos, L = lattice(n, 0.2154, 0.3, 5005)
    box = np.diag([L, L, L]).astype(np.float32)
    cap = 33_000_000
    nb, d, r, f = run(pos, 0.9, cap, box)
    assert f < cap
    assert (nb[0, :f] > nb[1, :f]).all() and (nb[1, :f] >= 0).all()
    assert (r[:f] <= np.float32(0.9)).all() and (nb[:, f:] == -1).all() and np.isnan(r[f:]).all()
    key = nb[0, :f].astype(np.int64) * n + nb[1, :f]
    assert len(np.unique(key)) == f
    assert np.allclose(np.linalg.norm(d[:f].astype(np.float64), axis=1), r[:f], rtol=1e-6)
    tree = cKDTree(np.mod(pos.astype(np.float64), L), boxsize=L)
    count = tree.count_neighbors(tree, 0.9) - n   # ordered pairs incl. self
    assert abs(count // 2 - f) <= 50          # float32 vs float64 disagree only on pairs within an ulp of the cutoff


# ------------------------------------------------------------------------------------------------------------------- PME
def pme_case(case):
    c = G["cases"][case]
    gx, gy, gz, order, alpha, coulomb = c["pme_args"]
    pos = np.array(c["pos"], np.float32)
    q = np.array([(i - 4) * 0.1 for i in range(9)], np.float32)
    excl = np.array(c["excl"], np.int32) if "excl" in c else np.zeros((9, 0), np.int32)
    return c, (int(gx), int(gy), int(gz), int(order), alpha, coulomb), pos, q, np.array(c["box"], np.float32), excl


@pytest.mark.parametrize("case", ["rectangular", "triclinic", "exclusions"])
def test_pme_golden(case):
    """TestPme.py:17-171: energies and forces computed with OpenMM, rtol 1e-4; charge derivatives vs the fp64 oracle."""
    from nnpops_b200.pme import PME
    c, (gx, gy, gz, order, alpha, coulomb), pos, q, box, excl = pme_case(case)
    pme = PME(gx, gy, gz, order, alpha, coulomb, torch.tensor(excl))
    p = torch.tensor(pos, device="cuda", requires_grad=True)
    ch = torch.tensor(q, device="cuda", requires_grad=True)
    b = torch.tensor(box, device="cuda")
    ed = pme.compute_direct(p, ch, c["cutoff"], b)
    assert np.allclose(c["edirect"], ed.item(), rtol=1e-4)
    er = pme.compute_reciprocal(p, ch, b)
    assert np.allclose(c["erecip"], er.item(), rtol=1e-4)
    ed.backward()
    assert np.allclose(c["expected_ddirect"], p.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    dq_direct = ch.grad.cpu().numpy().copy()
    p.grad = None; ch.grad = None
    er.backward()
    assert np.allclose(c["expected_drecip"], p.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    dq_recip = ch.grad.cpu().numpy()
    _, _, dq0 = NP.pme_direct(pos, q, box, c["cutoff"], alpha, coulomb, excl if excl.size else None)
    _, _, dq1 = NP.pme_reciprocal(pos, q, box, (gx, gy, gz), order, alpha, coulomb)
    assert rel_err(dq_direct, dq0) < 1e-4 and rel_err(dq_recip, dq1) < 1e-4


def test_pme_random_system_vs_oracle():
    from nnpops_b200.pme import PME
    rng = np.random.default_rng(3)
    n = 300
    box = np.array([[3.0, 0, 0], [0.4, 3.1, 0], [-0.3, 0.5, 2.9]], np.float32)
    pos = (rng.uniform(-0.5, 1.5, (n, 3)) @ box).astype(np.float32)
    q = rng.uniform(-0.5, 0.5, n).astype(np.float32); q -= q.mean()
    pme = PME(32, 30, 36, 4, 3.2, 138.935, torch.zeros((n, 0), dtype=torch.int32))
    p = torch.tensor(pos, device="cuda", requires_grad=True); ch = torch.tensor(q, device="cuda", requires_grad=True)
    b = torch.tensor(box, device="cuda")
    e = pme.compute_direct(p, ch, 1.2, b, max_num_pairs=60000) + pme.compute_reciprocal(p, ch, b)
    e.backward()
    e0d, f0d, q0d = NP.pme_direct(pos, q, box, 1.2, 3.2, 138.935)
    e0r, f0r, q0r = NP.pme_reciprocal(pos, q, box, (32, 30, 36), 4, 3.2, 138.935)
    errs = dict(energy=abs(e.item() - (e0d + e0r)) / abs(e0d + e0r), forces=rel_err(p.grad.cpu().numpy(), f0d + f0r),
                dq=rel_err(ch.grad.cpu().numpy(), q0d + q0r))
    print("PME random 300-atom triclinic system vs fp64 oracle:", errs)
    # north-star tolerance 1e-5 (the OpenMM goldens above carry five digits and keep the reference's own rtol 1e-4, TestPme.py:38-63)
    assert errs["energy"] < 1e-5 and errs["forces"] < 1e-5 and errs["dq"] < 1e-5


def _random_pme_system(n, seed, with_exclusions):
    rng = np.random.default_rng(seed)
    box = np.array([[3.0, 0, 0], [0.4, 3.1, 0], [-0.3, 0.5, 2.9]], np.float32)
    pos = (rng.uniform(-0.5, 1.5, (n, 3)) @ box).astype(np.float32)
    q = rng.uniform(-0.5, 0.5, n).astype(np.float32); q -= q.mean()
    if with_exclusions:   # "molecules" of three consecutive atoms exclude each other (symmetric, padded with -1)
        excl = -np.ones((n, 2), np.int32)
        for i in range(n):
            mates = [j for j in range(3 * (i // 3), min(3 * (i // 3) + 3, n)) if j != i]
            excl[i, :len(mates)] = mates
    else:
        excl = np.zeros((n, 0), np.int32)
    return pos, q, box, excl


@pytest.mark.parametrize("with_exclusions", [False, True])
@pytest.mark.parametrize("n", [3, 50, 700])
def test_pme_direct_fused_matches_list_path(n, with_exclusions):
    """The fused direct-space kernel (cell list + centre-owned erfc sum, no pair list) against the reference's two steps
    (getNeighborPairs + pme_direct, the path the OpenMM goldens pin): same accepted pairs and terms, sums in a different order."""
    from nnpops_b200.pme import PME
    pos, q, box, excl = _random_pme_system(n, 100 + n, with_exclusions)
    pme = PME(32, 30, 36, 5, 3.2, 138.935, torch.tensor(excl))
    b = torch.tensor(box, device="cuda")
    out = []
    for max_pairs in (n * n, -1):     # list path, fused path
        p = torch.tensor(pos, device="cuda", requires_grad=True); ch = torch.tensor(q, device="cuda", requires_grad=True)
        e = pme.compute_direct(p, ch, 1.2, b, max_num_pairs=max_pairs)
        e.backward()
        out.append((e.item(), p.grad.cpu().numpy(), ch.grad.cpu().numpy()))
    (e0, g0, q0), (e1, g1, q1) = out
    assert abs(e1 - e0) <= 2e-6 * abs(e0) + 1e-5
    assert rel_err(g1, g0) < 5e-6 and rel_err(q1, q0) < 5e-6
    e_ref, f_ref, q_ref = NP.pme_direct(pos, q, box, 1.2, 3.2, 138.935, excl if excl.size else None)
    errs = dict(energy_abs=abs(e1 - e_ref), energy_ref=e_ref, forces=rel_err(g1, f_ref), dq=rel_err(q1, q_ref))
    print("fused direct space vs fp64 oracle, n=%d exclusions=%s:" % (n, with_exclusions), errs)
    # the energy is returned as ONE fp32 number (kJ/mol, up to 1e4 here): 2e-3 is its resolution
    assert errs["energy_abs"] <= 1e-5 * abs(e_ref) + 2e-3 and errs["forces"] < 1e-5 and errs["dq"] < 1e-5, errs


@pytest.mark.parametrize("world", [2, 3, 8])
def test_pme_direct_sharded_partials_sum_to_the_whole(world):
    """One box over `world` ranks (emulated on one GPU): rank r owns slab r of the cell-sorted atoms as centres; energies and
    derivatives of the ranks add up to the unsharded result, and every atom's derivative comes from exactly one rank."""
    from nnpops_b200.pme import PME
    n = 901
    pos, q, box, excl = _random_pme_system(n, 7, True)
    pme = PME(32, 30, 36, 5, 3.2, 138.935, torch.tensor(excl))
    b = torch.tensor(box, device="cuda")
    p0 = torch.tensor(pos, device="cuda", requires_grad=True); c0 = torch.tensor(q, device="cuda", requires_grad=True)
    e0 = pme.compute_direct(p0, c0, 1.2, b)
    e0.backward()
    esum = 0.0
    gp = torch.zeros_like(p0); gq = torch.zeros_like(c0); owners = torch.zeros(n, device="cuda")
    for r in range(world):
        p = torch.tensor(pos, device="cuda", requires_grad=True); c = torch.tensor(q, device="cuda", requires_grad=True)
        e = pme.compute_direct_sharded(p, c, 1.2, b, emulate=(r, world))
        e.backward()
        esum += e.item()
        gp += p.grad; gq += c.grad
        owners += (p.grad.abs().sum(dim=1) > 0).float()
    assert abs(esum - e0.item()) <= 2e-6 * abs(e0.item()) + 1e-5
    assert torch.equal(gp, p0.grad) and torch.equal(gq, c0.grad)      # a centre's sums do not depend on the sharding
    assert float(owners.max()) == 1.0


def test_pme_energy_and_derivatives_matches_autograd_path():
    """PME.energy_and_derivatives (one call, no autograd: fused direct space + reciprocal forward and backward + self term) against
    compute_direct + compute_reciprocal differentiated by autograd."""
    from nnpops_b200.pme import PME
    n = 901
    pos, q, box, excl = _random_pme_system(n, 21, True)
    pme = PME(32, 30, 36, 5, 3.2, 138.935, torch.tensor(excl))
    b = torch.tensor(box, device="cuda")
    p = torch.tensor(pos, device="cuda", requires_grad=True); c = torch.tensor(q, device="cuda", requires_grad=True)
    e0 = pme.compute_direct(p, c, 1.2, b) + pme.compute_reciprocal(p, c, b)
    e0.backward()
    e1, gx, gq = pme.energy_and_derivatives(p.detach(), c.detach(), 1.2, b, world=1)
    assert abs(e1.item() - e0.item()) <= 2e-6 * abs(e0.item()) + 1e-3
    assert rel_err(gx.cpu().numpy(), p.grad.cpu().numpy()) < 2e-6 and rel_err(gq.cpu().numpy(), c.grad.cpu().numpy()) < 2e-6


@pytest.mark.parametrize("world", [2, 3])
def test_pme_reciprocal_sharded_matches_single(world):
    """Reciprocal PME with the atoms dealt to `world` ranks, emulated on one GPU: every rank spreads its block, the
    grids are summed (the all-reduce), every rank solves the same grid and differentiates its own atoms; energy and the assembled
    derivatives equal the single-rank op."""
    from nnpops_b200.pme import PME
    from nnpops_b200.pme.pme import pme_reciprocal, pme_reciprocal_sharded, pme_spread, shard_range
    rng = np.random.default_rng(11)
    n = 501
    box = np.array([[3.0, 0, 0], [0.4, 3.1, 0], [-0.3, 0.5, 2.9]], np.float32)
    pos = (rng.uniform(0, 1, (n, 3)) @ box).astype(np.float32)
    q = rng.uniform(-0.5, 0.5, n).astype(np.float32); q -= q.mean()
    pme = PME(32, 30, 36, 5, 3.2, 138.935, torch.zeros((n, 0), dtype=torch.int32))
    mod = [m.cuda() for m in pme.moduli]
    b = torch.tensor(box, device="cuda")
    p0 = torch.tensor(pos, device="cuda", requires_grad=True); c0 = torch.tensor(q, device="cuda", requires_grad=True)
    e0 = pme_reciprocal(p0, c0, b, 32, 30, 36, 5, 3.2, 138.935, *mod)
    e0.backward()
    # what the all-reduce would deliver: the sum of the ranks' private grids
    total = sum(pme_spread(p0.detach()[slice(*shard_range(n, r, world))], c0.detach()[slice(*shard_range(n, r, world))], b, 32, 30, 36, 5, 138.935)
                for r in range(world))
    gp = torch.zeros_like(p0); gq = torch.zeros_like(c0)
    for r in range(world):
        def reduce(t, r=r):                       # grid: replace by the global sum; packed derivatives: keep the local block (summed below)
            if t.dim() == 3:
                t.copy_(total)
        p = torch.tensor(pos, device="cuda", requires_grad=True); c = torch.tensor(q, device="cuda", requires_grad=True)
        e = pme_reciprocal_sharded(p, c, b, 32, 30, 36, 5, 3.2, 138.935, *mod, emulate=(r, world, reduce))
        assert abs(e.item() - e0.item()) <= 2e-6 * abs(e0.item()) + 1e-5
        e.backward()
        lo, hi = shard_range(n, r, world)
        assert not p.grad[:lo].any() and not p.grad[hi:].any()          # a rank differentiates its own atoms only
        gp += p.grad; gq += c.grad
    assert rel_err(gp.cpu().numpy(), p0.grad.cpu().numpy()) < 2e-6 and rel_err(gq.cpu().numpy(), c0.grad.cpu().numpy()) < 2e-6


def test_pme_double_derivative_raises():
    """TestPme.py:296-318."""
    from nnpops_b200.pme import PME
    c, (gx, gy, gz, order, alpha, coulomb), pos, q, box, excl = pme_case("rectangular")
    pme = PME(gx, gy, gz, order, alpha, coulomb, torch.tensor(excl))
    p = torch.tensor(pos, device="cuda", requires_grad=True); ch = torch.tensor(q, device="cuda"); b = torch.tensor(box, device="cuda")
    for fn in (lambda: pme.compute_direct(p, ch, 0.5, b), lambda: pme.compute_reciprocal(p, ch, b)):
        e = fn()
        (g,) = torch.autograd.grad(e, p, create_graph=True)
        with pytest.raises(RuntimeError):
            g.sum().backward()


def test_pme_cuda_graph():
    """TestPme.py:260-293."""
    from nnpops_b200.pme import PME
    c, (gx, gy, gz, order, alpha, coulomb), pos, q, box, excl = pme_case("rectangular")
    pme = PME(gx, gy, gz, order, alpha, coulomb, torch.tensor(excl))
    p = torch.tensor(pos, device="cuda", requires_grad=True); ch = torch.tensor(q, device="cuda"); b = torch.tensor(box, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            e = pme.compute_direct(p, ch, 0.5, b) + pme.compute_reciprocal(p, ch, b)
            e.backward(); p.grad = None
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        e = pme.compute_direct(p, ch, 0.5, b) + pme.compute_reciprocal(p, ch, b)
        e.backward()
    graph.replay()
    torch.cuda.synchronize()
    assert np.allclose(c["edirect"] + c["erecip"], e.item(), rtol=1e-4)
    assert np.allclose(np.array(c["expected_ddirect"]) + np.array(c["expected_drecip"]), p.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
