"""The reference's torch custom-op surface (libNNPOpsPyTorch.so): registration / TorchScript round trips on CPU, numerics on GPU.
Mirrors the structure of the reference suites TestSymmetryFunctions.py (serialisation, non-default stream), TestBatchedNN.py,
TestCFConv.py, TestNeighbors.py (jit) and TestPme.py (jit)."""
import json
import os
import tempfile

import numpy as np
import pytest
import torch

import oracle_lib as O
from systems import ANI2X, lattice, cubic_box, rel_err, water_species

from nnpops_b200 import torch_ops

torch_ops.build()
torch_ops.load()

HERE = os.path.dirname(os.path.abspath(__file__))
EXPECTED_SCHEMAS = {
    # schemas as registered by the reference build
    "NNPOpsANISymmetryFunctions::operation": "NNPOpsANISymmetryFunctions::operation(__torch__.torch.classes.NNPOpsANISymmetryFunctions.Holder? _0, Tensor _1, Tensor? _2) -> Tensor[] _0",
    "NNPOpsBatchedNN::BatchedLinear": "NNPOpsBatchedNN::BatchedLinear(Tensor _0, Tensor _1, Tensor _2) -> Tensor _0",
    "NNPOpsCFConv::operation": "NNPOpsCFConv::operation(__torch__.torch.classes.NNPOpsCFConv.Holder? _0, Any _1, Tensor _2, Tensor _3) -> Tensor _0",
    "neighbors::getNeighborPairs": "neighbors::getNeighborPairs(Tensor positions, Scalar cutoff, Scalar max_num_neighbors, Tensor box_vectors, bool checkErrors) -> (Tensor neighbors, Tensor deltas, Tensor distances, Tensor num_pairs)",
    "pme::pme_direct": "pme::pme_direct(Tensor positions, Tensor charges, Tensor neighbors, Tensor deltas, Tensor distances, Tensor exclusions, Scalar alpha, Scalar coulomb) -> Tensor",
    "pme::pme_reciprocal": "pme::pme_reciprocal(Tensor positions, Tensor charges, Tensor box_vectors, Scalar gridx, Scalar gridy, Scalar gridz, Scalar order, Scalar alpha, Scalar coulomb, Tensor xmoduli, Tensor ymoduli, Tensor zmoduli) -> Tensor",
}


def test_schemas_match_reference_registration():
    for name, expected in EXPECTED_SCHEMAS.items():
        got = [str(s) for s in torch._C._jit_get_schemas_for_operator(name)]
        assert expected in got, (name, got)


class AevModule(torch.nn.Module):
    """Shaped like the reference's TorchANISymmetryFunctions (SymmetryFunctions.py:66-123), without torchani."""

    def __init__(self, species):
        super().__init__()
        self.holder = torch.classes.NNPOpsANISymmetryFunctions.Holder(7, 5.1, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"],
                                                                     ANI2X["ShfA"], ANI2X["ShfZ"], species)

    def forward(self, positions: torch.Tensor, cell: torch.Tensor):
        radial, angular = torch.ops.NNPOpsANISymmetryFunctions.operation(self.holder, positions, cell)
        return torch.cat((radial, angular), dim=1)


class CFModule(torch.nn.Module):
    def __init__(self, w1, b1, w2, b2):
        super().__init__()
        self.neighbors = torch.classes.NNPOpsCFConvNeighbors.Holder(2.0)
        self.conv = torch.classes.NNPOpsCFConv.Holder(0.5, "ssp", w1, b1, w2, b2)

    def forward(self, positions: torch.Tensor, x: torch.Tensor):
        self.neighbors.build(positions)
        return torch.ops.NNPOpsCFConv.operation(self.conv, self.neighbors, positions, x)


class NeighborModule(torch.nn.Module):
    def forward(self, positions: torch.Tensor):
        nb, d, r, f = torch.ops.neighbors.getNeighborPairs(positions, 2.0, 200, torch.empty((0, 0), device=positions.device), False)
        return r


def test_torchscript_roundtrip_cpu():
    """script -> save -> load works without a GPU (custom classes pickle by their constructor arguments); running needs CUDA."""
    m = torch.jit.script(AevModule([0, 3, 0, 0, 3, 0]))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        m.save(f.name)
        m2 = torch.jit.load(f.name)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m2(torch.zeros(6, 3), torch.eye(3))
    G = json.load(open(os.path.join(HERE, "golden", "cfconv_water18.json")))
    cf = torch.jit.script(CFModule(torch.tensor(G["w1"]).reshape(5, 8), torch.tensor(G["b1"]), torch.tensor(G["w2"]).reshape(8, 8),
                                   torch.tensor(G["b2"])))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        cf.save(f.name)
        torch.jit.load(f.name)
    torch.jit.script(NeighborModule())
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, NotImplementedError)):
            torch.ops.neighbors.getNeighborPairs(torch.zeros(3, 3), 1.0, -1, torch.empty(0, 0), False)


# ----------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_aev_op_scripted_saved_loaded_and_on_a_side_stream():
    pos, L = lattice(300, 2.154, 0.3, 3000)
    species = water_species(300)
    box = cubic_box(L)
    m = torch.jit.script(AevModule(species.tolist()))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        m.save(f.name)
        m = torch.jit.load(f.name)
    m = m.to("cuda")
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    r0, a0 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, box=box)
    w = np.random.default_rng(1).standard_normal((300, 1008)).astype(np.float32)
    g0 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, w[:, :112], w[:, 112:], box=box)
    stream = torch.cuda.Stream()   # TestSymmetryFunctions.py:145-178: must honour the current (non-default) stream
    with torch.cuda.stream(stream):
        p = torch.tensor(pos, device="cuda", requires_grad=True)
        aev = m(p, torch.tensor(box, device="cuda"))
        (aev * torch.tensor(w, device="cuda")).sum().backward()
    stream.synchronize()
    assert rel_err(aev.detach().cpu().numpy(), np.concatenate([r0, a0], 1)) < 1e-5
    assert rel_err(p.grad.cpu().numpy(), g0) < 1e-5


@pytest.mark.gpu
def test_batched_linear_op_matches_aten():
    """BatchedLinear = matmul(weights, vectors) + biases with gradient w.r.t. the vectors only (BatchedNN.cpp:30-42), including the
    broadcast ensemble axis of layer 0."""
    rng = np.random.default_rng(0)
    N, M, nout, nin = 37, 8, 96, 1008
    W = torch.tensor(rng.standard_normal((1, N, M, nout, nin)).astype(np.float32) * 0.05, device="cuda")
    b = torch.tensor(rng.standard_normal((1, N, M, nout, 1)).astype(np.float32), device="cuda")
    for vm in (1, M):
        v = torch.tensor(rng.standard_normal((1, N, vm, nin, 1)).astype(np.float32), device="cuda", requires_grad=True)
        out = torch.ops.NNPOpsBatchedNN.BatchedLinear(v, W, b)
        go = torch.tensor(rng.standard_normal(out.shape).astype(np.float32), device="cuda")
        out.backward(go)
        v2 = v.detach().clone().double().requires_grad_(True)
        ref = torch.matmul(W.double(), v2) + b.double()
        ref.backward(go.double())
        assert out.shape == (1, N, M, nout, 1)
        assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 2e-6
        assert v.grad.shape == v.shape and rel_err(v.grad.cpu().numpy(), v2.grad.cpu().numpy()) < 2e-6


@pytest.mark.gpu
def test_cfconv_op_golden_and_script():
    G = json.load(open(os.path.join(HERE, "golden", "cfconv_water18.json")))
    pos = torch.tensor(np.array(G["positions"], np.float32).reshape(-1, 3), device="cuda", requires_grad=True)
    x = torch.tensor((0.1 * np.arange(144)).astype(np.float32).reshape(18, 8), device="cuda", requires_grad=True)
    mod = torch.jit.script(CFModule(torch.tensor(G["w1"]).reshape(5, 8), torch.tensor(G["b1"]), torch.tensor(G["w2"]).reshape(8, 8),
                                    torch.tensor(G["b2"])))
    y = mod(pos, x)
    exp = np.array(G["cases"]["nonperiodic"]["output"]).reshape(18, 8)
    diff = np.abs(exp - y.detach().cpu().numpy())
    assert not ((diff > 1e-4) & (diff / np.abs(exp) > 1e-3)).any()
    y.sum().backward()
    y0, ig0, pg0, _ = O.cfconv(pos.detach().cpu().numpy(), 8, 5, 2.0, 0.5, "ssp", G["w1"], G["b1"], G["w2"], G["b2"], x.detach().cpu().numpy(),
                               out_grad=np.ones((18, 8)), bits=64)
    assert rel_err(x.grad.cpu().numpy(), ig0) < 1e-5 and rel_err(pos.grad.cpu().numpy(), pg0) < 1e-5


@pytest.mark.gpu
def test_neighbor_and_pme_ops():
    G = json.load(open(os.path.join(HERE, "golden", "pme_openmm.json")))["cases"]["triclinic"]
    gx, gy, gz, order, alpha, coulomb = G["pme_args"]
    from nnpops_b200.pme.pme import bspline_moduli
    mod = [m.cuda() for m in bspline_moduli(int(order), (int(gx), int(gy), int(gz)))]
    pos = torch.tensor(G["pos"], dtype=torch.float32, device="cuda", requires_grad=True)
    q = torch.tensor([(i - 4) * 0.1 for i in range(9)], dtype=torch.float32, device="cuda")
    box = torch.tensor(G["box"], dtype=torch.float32, device="cuda")
    nb, d, r, f = torch.ops.neighbors.getNeighborPairs(pos, G["cutoff"], -1, box, False)
    ed = torch.ops.pme.pme_direct(pos, q, nb, d, r, torch.zeros((9, 0), dtype=torch.int32, device="cuda"), alpha, coulomb)
    er = torch.ops.pme.pme_reciprocal(pos, q, box, int(gx), int(gy), int(gz), int(order), alpha, coulomb, mod[0], mod[1], mod[2])
    self_e = -float((q ** 2).sum()) * coulomb * alpha / np.sqrt(np.pi)
    assert np.allclose(G["edirect"], ed.item(), rtol=1e-4) and np.allclose(G["erecip"], er.item() + self_e, rtol=1e-4)
    (ed + er).backward()
    assert np.allclose(np.array(G["expected_ddirect"]) + np.array(G["expected_drecip"]), pos.grad.cpu().numpy(), rtol=1e-4, atol=1e-3)
    with pytest.raises(RuntimeError):
        torch.ops.neighbors.getNeighborPairs(torch.zeros((4, 3), device="cuda"), 1.0, 1, torch.empty((0, 0), device="cuda"), True)


@pytest.mark.gpu
def test_scriptable_fused_ani_roundtrip_matches_ctypes_path():
    """The fused model as a TorchScript custom class: script -> save -> load -> energy + autograd forces equal the ctypes FusedANI."""
    from mlp_ref import random_networks
    from nnpops_b200.OptimizedTorchANI import FusedANI, ScriptableFusedANI
    n = 500
    pos, L = lattice(n, 2.154, 0.3, 3000)
    species = water_species(n)
    box = cubic_box(L)
    nets = random_networks(7, [(64, 64, 32)] * 7, 2, 1008, seed=3)
    args = (7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species, nets)
    mod = torch.jit.script(ScriptableFusedANI(*args))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        mod.save(f.name)
        mod = torch.jit.load(f.name)
    p = torch.tensor(pos, device="cuda", requires_grad=True)
    b = torch.tensor(box, device="cuda")
    e = mod(p, b)
    e.sum().backward()
    ref = FusedANI(*args)
    e0, g0 = ref.energy_and_gradient(torch.tensor(pos, device="cuda"), b)
    assert abs(float(e) - float(e0)) <= 1e-6 * abs(float(e0))
    assert rel_err(p.grad.cpu().numpy(), g0.cpu().numpy()) < 1e-6
