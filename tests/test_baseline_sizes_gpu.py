"""GPU parity at the sizes BASELINE.json names (configs 2-5): the whole 50 000-atom box against the compiled reference, the fused
model on the 5 000-atom protein with the full ANI-2x networks, CFConv at 100 000 atoms (cutoff 5 and 10 A) and PME at 200 000
charges / 128^3 / order 5 against oracle evaluations of sampled atoms, with EXACT neighbour counts (the reference's own fp32
cutoff arithmetic restated in numpy on KD-tree candidates)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle_lib as O
from mlp_ref import mlp_energy_and_grad, random_networks
from systems import ANI2X, ANI2X_HIDDEN, cubic_box, lattice, protein_species, rel_err, water_species

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("neighbors_pme_oracle", os.path.join(HERE, "..", "oracle", "neighbors_pme_oracle.py"))
NP = importlib.util.module_from_spec(spec)
spec.loader.exec_module(NP)
TOL = 1e-5


def dev(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def ani2x_tables():
    return O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])


def exact_pair_count_fp32(pos, L, cutoff, mode):
    """Pairs i < j inside the cutoff of a cubic periodic box, decided with the REFERENCE's fp32 arithmetic (no FMA, same operation
    order) on candidates from a float64 KD-tree with a margin.
    mode "mul": delta - round(delta * inv) * L with inv = fp32(1 / L), r2 < cutoff^2     (CpuANISymmetryFunctions.cpp:355-379, CpuCFConv.cpp:30-55)
    mode "div": delta - round(delta / L) * L, r = sqrt(r2), r <= cutoff (inclusive)       (getNeighborPairsCPU.cpp:60-81)"""
    from scipy.spatial import cKDTree
    w = np.mod(pos.astype(np.float64), L)
    tree = cKDTree(w, boxsize=L)
    cand = tree.query_pairs(cutoff * (1 + 1e-4) + 1e-4, output_type="ndarray")
    p = pos.astype(np.float32)
    Lf = np.float32(L)
    count = 0
    for lo in range(0, len(cand), 4_000_000):
        c = cand[lo:lo + 4_000_000]
        d = p[c[:, 1]] - p[c[:, 0]]
        if mode == "mul":
            inv = np.float32(1.0) / Lf
            d = d - np.rint(d * inv) * Lf
            r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            count += int(np.count_nonzero(r2 < np.float32(cutoff) * np.float32(cutoff)))
        else:
            d = d - np.rint(d / Lf) * Lf
            r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            count += int(np.count_nonzero(np.sqrt(r2) <= np.float32(cutoff)))
    return count


# --------------------------------------------------------------------------------------------------------------- config 3
def test_config3_whole_box_aev_and_gradient_vs_reference():
    """BASELINE config 3: the complete AEV matrix and dE/dx (random upstream gradient) of the 50 000-atom periodic water box, Rcr 5.2,
    against the UNMODIFIED reference class CpuANISymmetryFunctions (oracle/_ref, O(N^2): about a minute on one core) -- or, where
    that library is absent, its C restatement.  The north-star bar: max|delta| / max|ref| <= 1e-5; forces max|delta| printed."""
    from nnpops_b200.SymmetryFunctions import Holder
    n = 50000
    pos, L = lattice(n, 2.154, 0.3, 3000)
    species = water_species(n)
    box = cubic_box(L)
    rfn, afn = ani2x_tables()
    rng = np.random.default_rng(50)
    rg = rng.standard_normal((n, 112)).astype(np.float32)
    ag = rng.standard_normal((n, 896)).astype(np.float32)
    h = Holder.from_function_lists(7, 5.2, 3.5, rfn, afn, list(species))
    radial, angular = h.forward(dev(pos), dev(box))
    g = h.backward([dev(rg), dev(ag)]).cpu().numpy()
    assert h.overflowed() == 0
    triples, pairs = h.work()
    assert pairs == exact_pair_count_fp32(pos, L, 5.2, "mul")          # the cell list misses nothing and adds nothing
    impl = "ref" if O.ref_lib() is not None else "oracle"
    r0, a0 = O.ani_forward(pos, species, 7, 5.2, 3.5, rfn, afn, box=box, impl=impl)
    g0 = O.ani_backward(pos, species, 7, 5.2, 3.5, rfn, afn, rg, ag, box=box, impl=impl)
    errs = dict(radial=rel_err(radial.cpu().numpy(), r0), angular=rel_err(angular.cpu().numpy(), a0), grad=rel_err(g, g0))
    print("config 3 whole box vs %s:" % impl, errs, "forces max|delta| = %.3e of max|ref| %.3e, triples %d, pairs %d" %
          (np.abs(g - g0).max(), np.abs(g0).max(), triples, pairs))
    assert errs["radial"] < TOL and errs["angular"] < TOL and errs["grad"] < TOL


# --------------------------------------------------------------------------------------------------------------- config 2
def test_config2_protein5000_fused_energy_forces_full_ani2x():
    """BASELINE config 2: energy and forces of the fused model on the 5 000-atom non-periodic protein-like system with the full
    ANI-2x network shapes (5 species present -> 560 active AEV columns, first-layer K = 560 on the tensor-core path) against the
    fp64 chain oracle AEV -> ATen MLP -> oracle backward."""
    from nnpops_b200.OptimizedTorchANI import FusedANI
    n = 5000
    pos, _ = lattice(n, 2.154, 0.3, 2002)
    species = protein_species(n)
    nets = random_networks(7, ANI2X_HIDDEN, 8, 1008, 42)
    m = FusedANI(7, 5.1, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species, nets)
    e, g = m.energy_and_gradient(dev(pos), None)
    e = float(e.cpu()[0]); g = g.cpu().numpy()
    assert m.overflowed() == 0
    assert m.work()["active_features"] == 5 * 16 + 15 * 32
    rfn, afn = ani2x_tables()
    r0, a0 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, bits=64)
    e0, dA = mlp_energy_and_grad(np.concatenate([r0, a0], axis=1), species, nets, torch.float64)
    g0 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, dA[:, :112], dA[:, 112:], bits=64)
    errs = dict(energy=abs(e - e0) / abs(e0), forces=rel_err(g, g0))
    print("config 2 protein5000:", errs, "forces max|delta| = %.3e" % np.abs(g - g0).max())
    assert errs["energy"] < TOL and errs["forces"] < TOL


# --------------------------------------------------------------------------------------------------------------- config 4
@pytest.mark.parametrize("cutoff,samples", [(5.0, 96), (10.0, 6)])
def test_config4_cfconv_100k_sampled_rows_vs_oracle(cutoff, samples):
    """BASELINE config 4 (100 000 atoms periodic, width 128, 50 Gaussians, sigma 0.2) at cutoff 5 A and at the reference benchmark's
    10 A: exact pair count, and for sampled centres the output row, the input-gradient row and the position gradient against the fp64
    oracle evaluated on the centre's neighbourhood (explicit periodic images; every pair that touches the centre is in the cluster,
    and the upstream gradient is non-zero on the centre only, so the centre's gradients are complete)."""
    from scipy.spatial import cKDTree
    from nnpops_b200.CFConv import CFConv
    from nnpops_b200.CFConvNeighbors import CFConvNeighbors
    rng = np.random.default_rng(9)
    n, W, Gn, sigma = 100000, 128, 50, 0.2
    pos, L = lattice(n, 2.154, 0.3, 4004)
    w1 = rng.normal(0, 0.1, (W, Gn)).astype(np.float32); b1 = rng.normal(0, 0.1, W).astype(np.float32)
    w2 = rng.normal(0, 0.1, (W, W)).astype(np.float32); b2 = rng.normal(0, 0.1, W).astype(np.float32)
    x = rng.standard_normal((n, W)).astype(np.float32)
    # sampled centres at least two cutoffs (+ margin) apart: then no atom is a neighbour of two of them and the upstream gradients,
    # which are non-zero on the centres only, do not mix inside an oracle cluster
    wrapped = np.mod(pos.astype(np.float64), L)
    centres = []
    for i in rng.permutation(n):
        d = wrapped[centres] - wrapped[i] if centres else np.zeros((0, 3))
        d -= np.round(d / L) * L
        if not len(d) or np.sqrt((d * d).sum(1)).min() > 2.1 * cutoff + 0.2:
            centres.append(int(i))
        if len(centres) == samples:
            break
    centres = np.array(centres)
    assert len(centres) == samples
    go = np.zeros((n, W), np.float32)
    go[centres] = rng.standard_normal((samples, W)).astype(np.float32)
    nb = CFConvNeighbors(cutoff)
    conv = CFConv(sigma, "ssp", torch.tensor(w1.reshape(Gn, W)), torch.tensor(b1), torch.tensor(w2), torch.tensor(b2))
    p = torch.tensor(pos, device="cuda", requires_grad=True)
    xx = torch.tensor(x, device="cuda", requires_grad=True)
    nb.build(p, dev(cubic_box(L)))
    y = conv(nb, p, xx)
    y.backward(dev(go))
    y = y.detach().cpu().numpy(); ig = xx.grad.cpu().numpy(); pg = p.grad.cpu().numpy()
    assert nb.num_pairs() == exact_pair_count_fp32(pos, L, cutoff, "mul")
    tree = cKDTree(wrapped, boxsize=L)
    worst = dict(out=0.0, input_grad=0.0, pos_grad=0.0)
    scale = dict(out=np.abs(y[centres]).max(), input_grad=np.abs(ig[centres]).max(), pos_grad=np.abs(pg[centres]).max())
    for i in centres:
        near = [j for j in tree.query_ball_point(wrapped[i], cutoff * 1.02 + 0.05) if j != i]
        d = wrapped[near] - wrapped[i]
        d -= np.round(d / L) * L
        cluster = np.vstack([np.zeros((1, 3)), d]).astype(np.float32)
        ci = np.concatenate([[i], near])
        gc = np.zeros((len(ci), W), np.float32); gc[0] = go[i]
        y0, ig0, pg0, _ = O.cfconv(cluster, W, Gn, cutoff, sigma, "ssp", w1, b1, w2, b2, x[ci], out_grad=gc, bits=64)
        # with an upstream gradient on the centre only: d/d(input of the centre) comes from pairs (centre, j) with gradient on j = 0, so
        # compare the centre's OUTPUT row, the neighbours' input-gradient rows and the centre's position gradient
        worst["out"] = max(worst["out"], np.abs(y[i] - y0[0]).max() / scale["out"])
        worst["input_grad"] = max(worst["input_grad"], np.abs(ig[near] - ig0[1:]).max() / max(np.abs(ig0).max(), 1e-30))
        worst["pos_grad"] = max(worst["pos_grad"], np.abs(pg[i] - pg0[0]).max() / max(np.abs(pg0).max(), 1e-30))
    print("config 4 cutoff %.0f A: pairs %d, worst errors over sampled centres %s" % (cutoff, nb.num_pairs(), worst))
    assert worst["out"] < TOL and worst["input_grad"] < TOL and worst["pos_grad"] < TOL


# --------------------------------------------------------------------------------------------------------------- config 5
def test_config5_pme_200k_vs_oracle():
    """BASELINE config 5 (200 000 charges, cubic 12.71 nm box, 128^3 grid, order 5, alpha 2.92 / nm, direct cutoff 0.9 nm):
    reciprocal energy, dE/dx and dE/dq against the vectorised fp64 numpy oracle over ALL atoms (O(125 N)); direct-space dE/dx and
    dE/dq on sampled atoms against an fp64 KD-tree evaluation; exact pair count of getNeighborPairs."""
    from nnpops_b200.neighbors import getNeighborPairs
    from nnpops_b200.pme import PME
    n = 200000
    pos, L = lattice(n, 0.2154, 0.3, 5005)
    rng = np.random.default_rng(5005)
    q = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    q -= q.mean(dtype=np.float64).astype(np.float32)
    box = cubic_box(L)
    alpha, coulomb, cutoff = 2.92, 138.935, 0.9
    pme = PME(128, 128, 128, 5, alpha, coulomb, torch.zeros((n, 0), dtype=torch.int32))
    p = torch.tensor(pos, device="cuda", requires_grad=True); ch = torch.tensor(q, device="cuda", requires_grad=True)
    b = dev(box)
    er = pme.compute_reciprocal(p, ch, b)
    er.backward()
    gr = p.grad.cpu().numpy().copy(); qr = ch.grad.cpu().numpy().copy()
    e0, f0, q0 = NP.pme_reciprocal_vec(pos, q, box, (128, 128, 128), 5, alpha, coulomb)
    # the value returned includes the self term -k alpha / sqrt(pi) sum q^2 (pme.py:194), which cancels nine tenths of the reciprocal sum
    # here; the error of the fp32 grid arithmetic is measured against the reciprocal sum itself
    e_self = -coulomb * alpha / np.sqrt(np.pi) * float(np.sum(q.astype(np.float64) ** 2))
    errs = dict(erecip=abs(er.item() - e0) / abs(e0 - e_self), frecip=rel_err(gr, f0), qrecip=rel_err(qr, q0))
    p.grad = None; ch.grad = None
    npairs_exact = exact_pair_count_fp32(pos, L, cutoff, "div")
    ed = pme.compute_direct(p, ch, cutoff, b, max_num_pairs=int(npairs_exact * 1.02) + 1024)
    ed.backward()
    gd = p.grad.cpu().numpy(); qd = ch.grad.cpu().numpy()
    _, _, _, found = getNeighborPairs(p.detach(), cutoff, int(npairs_exact * 1.02) + 1024, b)
    assert int(found.item()) == npairs_exact
    sample = rng.choice(n, 400, replace=False)
    fd0, qd0, _ = NP.pme_direct_sampled(pos, q, L, cutoff, alpha, coulomb, sample)
    errs.update(fdirect=float(np.abs(gd[sample] - fd0).max() / np.abs(fd0).max()), qdirect=float(np.abs(qd[sample] - qd0).max() / np.abs(qd0).max()))
    # the fused direct-space kernel (no pair list; PME.compute_direct with the default max_num_pairs) on the same samples, and
    # against the list path over all atoms
    p.grad = None; ch.grad = None
    ef = pme.compute_direct(p, ch, cutoff, b)
    ef.backward()
    gf = p.grad.cpu().numpy(); qf = ch.grad.cpu().numpy()
    errs.update(fdirect_fused=float(np.abs(gf[sample] - fd0).max() / np.abs(fd0).max()),
                qdirect_fused=float(np.abs(qf[sample] - qd0).max() / np.abs(qd0).max()),
                fused_vs_list_f=rel_err(gf, gd), fused_vs_list_q=rel_err(qf, qd), fused_vs_list_e=abs(ef.item() - ed.item()) / abs(ed.item()))
    assert errs["fdirect_fused"] < TOL and errs["qdirect_fused"] < TOL
    assert errs["fused_vs_list_f"] < TOL and errs["fused_vs_list_q"] < TOL and errs["fused_vs_list_e"] < TOL
    print("config 5 PME 200k / 128^3 / order 5:", errs, "pairs", npairs_exact)
    assert errs["erecip"] < TOL and errs["frecip"] < TOL and errs["qrecip"] < TOL
    assert errs["fdirect"] < TOL and errs["qdirect"] < TOL
