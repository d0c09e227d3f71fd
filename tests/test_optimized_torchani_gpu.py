"""This package's OptimizedTorchANI (same constructor as the reference's, built on the fused model) on a torchani stand-in, against
oracle AEV + ATen MLP (cf. TestOptimizedTorchANI.py:59-100 of the reference, which compares against torchani itself).  The literal
chain -- the reference's own SymmetryFunctions.py + BatchedNN.py on the drop-in ops -- is run from the reference's files in
tests/test_reference_package_gpu.py."""
import numpy as np
import pytest
import torch

import oracle_lib as O
from fake_torchani import Model
from mlp_ref import mlp_energy_and_grad
from systems import ANI2X, lattice, rel_err

pytestmark = pytest.mark.gpu
HIDDEN = [(96, 64, 48), (80, 64, 48), (64, 48, 32), (64, 48, 32), (48, 32, 32), (48, 32, 32), (48, 32, 32)]


def reference_values(model, pos, species):
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    r0, a0 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, bits=64)
    e0, dA = mlp_energy_and_grad(np.concatenate([r0, a0], 1), species, model.networks_numpy(), torch.float64)
    g0 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, dA[:, :112], dA[:, 112:], bits=64)
    sae = float(model.energy_shifter.sae(torch.tensor(species)[None])[0])
    return e0 + sae, g0


def test_module_chain():
    from nnpops_b200.OptimizedTorchANI import OptimizedTorchANI
    n = 46   # size of the reference's benchmark ligand 2iuz
    pos, _ = lattice(n, 1.9, 0.3, 46)
    species = np.random.default_rng(3).integers(0, 7, n)
    numbers = torch.tensor([[Model(HIDDEN, 1, 0).species_converter.ELEMENTS[s] for s in species]])
    model = Model(HIDDEN, 4, seed=11)
    e_ref, g_ref = reference_values(model, pos, species)
    p = torch.tensor(pos, device="cuda").unsqueeze(0).requires_grad_(True)
    nnp = OptimizedTorchANI(model, numbers.cuda())
    energy = nnp((numbers.cuda(), p)).energies
    energy.sum().backward()
    e = float(energy.detach().cpu().double()[0])
    assert abs(e - e_ref) < 5e-6 * abs(e_ref)
    assert rel_err(p.grad.cpu().numpy()[0], g_ref) < 1e-5


def test_batch_of_conformers():
    """the scalable modules take a batch of conformers of one system, coordinates [B, N, 3] -> energies [B] (the
    reference raises at B > 1, SymmetryFunctions.py:110-111); energies and autograd forces equal the one-at-a-time evaluations, also
    through the TorchScript class."""
    from nnpops_b200.OptimizedTorchANI import OptimizedTorchANI, ScriptableFusedANI
    n, B = 46, 3
    species = np.random.default_rng(3).integers(0, 7, n)
    model = Model(HIDDEN, 2, seed=5)
    numbers = torch.tensor([[model.species_converter.ELEMENTS[s] for s in species]])
    confs = np.stack([lattice(n, 1.9, 0.3, 100 + b)[0] for b in range(B)])
    nnp = OptimizedTorchANI(model, numbers.cuda())
    p = torch.tensor(confs, device="cuda", requires_grad=True)
    out = nnp((numbers.cuda().expand(B, -1), p))
    assert out.energies.shape == (B,) and out.species.shape == (B, n)
    out.energies.sum().backward()
    for b in range(B):
        q = torch.tensor(confs[b:b + 1], device="cuda", requires_grad=True)
        e = nnp((numbers.cuda(), q)).energies
        e.sum().backward()
        assert abs(float(e[0]) - float(out.energies[b])) <= 1e-9 * abs(float(e[0]))
        assert rel_err(p.grad[b].cpu().numpy(), q.grad[0].cpu().numpy()) < 1e-6
    nets = model.networks_numpy()
    sm = torch.jit.script(ScriptableFusedANI(7, 5.1, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"],
                                             [int(s) for s in species], nets))
    eb = sm(torch.tensor(confs, device="cuda"), None)
    sae = float(model.energy_shifter.sae(torch.tensor(species)[None])[0])
    assert eb.shape == (B,)
    assert np.allclose(eb.double().cpu().numpy() + sae, out.energies.detach().cpu().numpy(), rtol=1e-6)
