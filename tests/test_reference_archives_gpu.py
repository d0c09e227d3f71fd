"""Pickle byte-compatibility: TorchScript archives SAVED BY THE REFERENCE LIBRARY (tests/golden/ref_saved/*.pt,
made in the build container by scripts/make_ref_archives.py with the unmodified reference extension on the CPU) load on this library,
move to the GPU and reproduce the outputs and gradients the reference recorded next to them.  This is what an openmm-torch user has on
disk: custom-class state written by the reference's __getstate__ (SymmetryFunctions.cpp:177-218, CFConv.cpp:191-241,
CFConvNeighbors.cpp:54-75) and graphs that bind ops by their registered schemas."""
import os

import numpy as np
import pytest
import torch

from systems import rel_err

pytestmark = pytest.mark.gpu
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_saved")
E = np.load(os.path.join(D, "expected.npz"))


@pytest.fixture(scope="module", autouse=True)
def library():
    from nnpops_b200 import torch_ops
    torch_ops.load()


def t(name, grad=False, dtype=torch.float32):
    return torch.tensor(E[name], dtype=dtype, device="cuda", requires_grad=grad)


def test_symmetry_functions_archive():
    m = torch.jit.load(os.path.join(D, "symmfunc.pt")).to("cuda")
    species = t("symmfunc.species", dtype=torch.int64)
    pos = t("symmfunc.positions", grad=True)
    _, aev = m((species, pos))
    assert rel_err(aev.detach().cpu().numpy(), E["symmfunc.aev"]) < 1e-5
    (aev * t("symmfunc.weights")).sum().backward()
    assert rel_err(pos.grad.cpu().numpy(), E["symmfunc.grad"]) < 1e-5
    m2 = torch.jit.load(os.path.join(D, "symmfunc.pt")).to("cuda")      # periodicity is frozen at a holder's first call
    _, aev_p = m2((species, pos.detach()), t("symmfunc.cell"), torch.tensor([True, True, True], device="cuda"))
    assert rel_err(aev_p.cpu().numpy(), E["symmfunc.aev_periodic"]) < 1e-5


def test_cfconv_archives():
    nb = torch.jit.load(os.path.join(D, "cfconv_nb.pt")).to("cuda")
    cf = torch.jit.load(os.path.join(D, "cfconv.pt")).to("cuda")
    pos = t("cfconv.positions", grad=True); x = t("cfconv.input", grad=True)
    nb.build(pos)
    y = cf(nb, pos, x)
    y.sum().backward()
    assert rel_err(y.detach().cpu().numpy(), E["cfconv.output"]) < 1e-5
    assert rel_err(pos.grad.cpu().numpy(), E["cfconv.pos_grad"]) < 1e-5 and rel_err(x.grad.cpu().numpy(), E["cfconv.input_grad"]) < 1e-5


def test_pme_and_neighbors_archives():
    pm = torch.jit.load(os.path.join(D, "pme.pt")).to("cuda")
    pos = t("pme.positions", grad=True)
    e = pm(pos, t("pme.charges"), t("pme.box"))
    e.backward()
    assert abs(e.item() - float(E["pme.energy"])) <= 1e-5 * abs(float(E["pme.energy"]))
    assert rel_err(pos.grad.cpu().numpy(), E["pme.grad"]) < 1e-4     # the reference's CPU path differentiates in fp32 as well
    nm = torch.jit.load(os.path.join(D, "neighbors.pt")).to("cuda")
    p = t("neighbors.positions", grad=True)
    s = nm(p, t("neighbors.box"))
    s.backward()
    assert abs(s.item() - float(E["neighbors.value"])) <= 1e-6 * abs(float(E["neighbors.value"]))
    assert rel_err(p.grad.cpu().numpy(), E["neighbors.grad"]) < 1e-5
