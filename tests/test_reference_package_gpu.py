"""The reference's OWN Python package and pytest suites, unchanged, on top of libNNPOpsPyTorch.so (reference
src/pytorch/__init__.py:14 loads the library by name).  scripts/make_ref_package.py assembles baseline/_ref/ (git-ignored, travels with
gpurun) from /root/reference: NNPOps/ = the reference's *.py byte for byte with the one load_library line re-pointed, ref_tests/ = its
TestNeighbors.py, TestPme.py, TestCFConv.py, TestCFConvNeighbors.py.  Each suite runs in its own interpreter; parametrisations on
device 'cpu' are deselected -- this library has, by design, no CPU implementation."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle_lib as O
from mlp_ref import mlp_energy_and_grad
from systems import ANI2X, lattice, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFPKG = os.path.join(ROOT, "baseline", "_ref")
needs_pkg = pytest.mark.skipif(not os.path.isdir(os.path.join(REFPKG, "NNPOps")),
                               reason="baseline/_ref/NNPOps not assembled (python scripts/make_ref_package.py needs /root/reference)")


def ref_env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REFPKG, os.path.join(HERE, "torchani_stub"), HERE, env.get("PYTHONPATH", "")])
    return env


@needs_pkg
@pytest.mark.parametrize("suite", ["TestNeighbors.py", "TestPme.py", "TestCFConv.py", "TestCFConvNeighbors.py"])
def test_reference_suite_runs_unchanged(suite):
    # One deselection besides the CPU parametrisations: TestNeighbors.py::test_is_cuda_graph_compatible passes max_num_pairs as a
    # numpy.int64, which torch 2.11 refuses to bind to a `Scalar` schema argument ("Cannot cast 48 to number") before any library
    # code runs -- the reference's own library rejects the same call in this image (checked with oracle/_ref/libNNPOpsPyTorch_refcpu.so).
    # CUDA-graph capture of the op is covered with a Python int by tests/test_neighbors_pme_gpu.py::test_neighbors_cuda_graph.
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REFPKG, "ref_tests", suite), "-q", "-k", "not cpu and not test_is_cuda_graph_compatible", "-p", "no:cacheprovider"],
                       env=ref_env(), capture_output=True, text=True, cwd=REFPKG, timeout=1500)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    print("reference %s on libNNPOpsPyTorch.so: %s" % (suite, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]))
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]


SCRIPT = r'''
import json, sys
import numpy as np, torch
import NNPOps                                            # the reference package (baseline/_ref/NNPOps), library = libNNPOpsPyTorch.so
from NNPOps.OptimizedTorchANI import OptimizedTorchANI   # reference src/pytorch/OptimizedTorchANI.py:33-54
from NNPOps.SymmetryFunctions import TorchANISymmetryFunctions
from NNPOps.BatchedNN import TorchANIBatchedNN
from fake_torchani import Model
cfg = json.loads(sys.argv[1])
pos = np.array(cfg["pos"], np.float32); species = cfg["species"]
model = Model(cfg["hidden"], cfg["ensemble"], seed=cfg["seed"])
numbers = torch.tensor([[model.species_converter.ELEMENTS[s] for s in species]])
out = {}
for path in ("optimized", "scripted"):
    nnp = OptimizedTorchANI(model, numbers).to("cuda")
    if path == "scripted":
        import tempfile
        nnp = torch.jit.script(nnp)
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            nnp.save(f.name); nnp = torch.jit.load(f.name)
    p = torch.tensor(pos, device="cuda").unsqueeze(0).requires_grad_(True)
    e = nnp((numbers.cuda(), p)).energies
    e.sum().backward()
    out[path] = {"energy": float(e.detach().double().cpu()[0]), "grad": p.grad.cpu().numpy()[0].tolist()}
print("RESULT" + json.dumps(out))
'''


@needs_pkg
def test_reference_optimized_torchani_on_this_library():
    """The reference's OptimizedTorchANI module chain (its SymmetryFunctions.py, BatchedNN.py, EnergyShifter.py, SpeciesConverter.py,
    unchanged) over a torchani stand-in, eager and TorchScript save/load, against oracle AEV + fp64 ATen MLP -- the comparison
    TestOptimizedTorchANI.py:59-100 makes against torchani itself."""
    from fake_torchani import Model
    hidden = [(96, 64, 48), (80, 64, 48), (64, 48, 32), (64, 48, 32), (48, 32, 32), (48, 32, 32), (48, 32, 32)]
    n = 46
    pos, _ = lattice(n, 1.9, 0.3, 46)
    species = np.random.default_rng(3).integers(0, 7, n)
    cfg = dict(pos=pos.tolist(), species=[int(s) for s in species], hidden=hidden, ensemble=4, seed=11)
    r = subprocess.run([sys.executable, "-c", SCRIPT, json.dumps(cfg)], env=ref_env(), capture_output=True, text=True, cwd=REFPKG, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1][6:])
    model = Model(hidden, 4, seed=11)
    rfn, afn = O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
    r0, a0 = O.ani_forward(pos, species, 7, 5.1, 3.5, rfn, afn, bits=64)
    e0, dA = mlp_energy_and_grad(np.concatenate([r0, a0], 1), species, model.networks_numpy(), torch.float64)
    g0 = O.ani_backward(pos, species, 7, 5.1, 3.5, rfn, afn, dA[:, :112], dA[:, 112:], bits=64)
    e0 += float(model.energy_shifter.sae(torch.tensor(species)[None])[0])
    for path, v in res.items():
        print("reference OptimizedTorchANI (%s): energy rel %.2e, forces rel %.2e" % (path, abs(v["energy"] - e0) / abs(e0), rel_err(np.array(v["grad"]), g0)))
        assert abs(v["energy"] - e0) < 5e-6 * abs(e0)
        assert rel_err(np.array(v["grad"]), g0) < 1e-5
