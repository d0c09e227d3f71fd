"""TorchScript archives SAVED BY THE REFERENCE LIBRARY, for the pickle byte-compatibility test (a module saved
with the reference's libNNPOpsPyTorch.so must load and run on this library, because that is what openmm-torch users have on disk).

Runs in THIS container only (needs /root/reference): loads oracle/_ref/libNNPOpsPyTorch_refcpu.so (oracle/build_ref_torch.py), imports
the reference's Python wrappers from baseline/_ref/NNPOps without executing its __init__ (which would load this repository's library),
scripts and saves small modules on the CPU and records the reference's own outputs next to them:

    tests/golden/ref_saved/symmfunc.pt      TorchANISymmetryFunctions (SymmetryFunctions.py; Holder pickled by SymmetryFunctions.cpp:177-218)
    tests/golden/ref_saved/cfconv_nb.pt     CFConvNeighbors          (CFConvNeighbors.cpp:54-75)
    tests/golden/ref_saved/cfconv.pt        CFConv                   (CFConv.cpp:191-241)
    tests/golden/ref_saved/pme.pt           module over NNPOps.pme.PME (pme.py)
    tests/golden/ref_saved/neighbors.pt     module over getNeighborPairs
    tests/golden/ref_saved/expected.npz     inputs and the reference's outputs / gradients (CPU, fp32)
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "torchani_stub"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

REFLIB = os.path.join(ROOT, "oracle", "_ref", "libNNPOpsPyTorch_refcpu.so")
REFPKG = os.path.join(ROOT, "baseline", "_ref", "NNPOps")
OUT = os.path.join(ROOT, "tests", "golden", "ref_saved")


def main():
    assert os.path.exists(REFLIB), "run oracle/build_ref_torch.py first"
    assert os.path.isdir(REFPKG), "run scripts/make_ref_package.py first"
    torch.ops.load_library(REFLIB)
    pkg = types.ModuleType("NNPOps")
    pkg.__path__ = [REFPKG]
    sys.modules["NNPOps"] = pkg
    from NNPOps.SymmetryFunctions import TorchANISymmetryFunctions
    from NNPOps.CFConv import CFConv
    from NNPOps.CFConvNeighbors import CFConvNeighbors
    from NNPOps.neighbors import getNeighborPairs
    from NNPOps.pme import PME
    from fake_torchani import AEVComputer, SpeciesConverter
    os.makedirs(OUT, exist_ok=True)
    exp = {}
    rng = np.random.default_rng(2024)

    # --- ANI symmetry functions: 14 atoms, all seven species, non-periodic and periodic
    conv = SpeciesConverter()
    numbers = torch.tensor([[1, 6, 7, 8, 16, 9, 17, 1, 1, 6, 8, 7, 1, 6]])
    pos = torch.tensor(rng.uniform(0.0, 5.0, (1, 14, 3)), dtype=torch.float32, requires_grad=True)
    sf = TorchANISymmetryFunctions(conv, AEVComputer(), numbers)
    species = conv((numbers, torch.empty(0))).species
    _, aev = sf((species, pos))
    w = torch.tensor(rng.standard_normal(aev.shape), dtype=torch.float32)
    (aev * w).sum().backward()
    torch.jit.script(sf).save(os.path.join(OUT, "symmfunc.pt"))
    cell = torch.tensor([[11.0, 0, 0], [0, 12.0, 0], [0, 0, 10.5]])
    sf2 = TorchANISymmetryFunctions(conv, AEVComputer(), numbers)     # periodicity is frozen at the first call: a second holder
    _, aev_p = sf2((species, pos.detach()), cell, torch.tensor([True, True, True]))
    exp["symmfunc"] = dict(species=species.tolist(), positions=pos.detach().tolist(), weights=w.tolist(), aev=aev.detach().tolist(),
                           grad=pos.grad.tolist(), cell=cell.tolist(), aev_periodic=aev_p.detach().tolist())

    # --- CFConv + neighbours (the construction of TestCFConv.py:35-47)
    nA, nF, nG = 9, 5, 7
    p2 = torch.tensor(rng.uniform(-3.0, 3.0, (nA, 3)), dtype=torch.float32, requires_grad=True)
    x = torch.tensor(rng.uniform(0, 1, (nA, nF)), dtype=torch.float32, requires_grad=True)
    nb = CFConvNeighbors(5.0)
    cf = CFConv(0.5, "ssp",
                torch.tensor(rng.uniform(0, 1, (nG, nF)), dtype=torch.float32), torch.tensor(rng.uniform(0, 1, nF), dtype=torch.float32),
                torch.tensor(rng.uniform(0, 1, (nF, nF)), dtype=torch.float32), torch.tensor(rng.uniform(0, 1, nF), dtype=torch.float32))
    nb.build(p2)
    y = cf(nb, p2, x)
    y.sum().backward()
    torch.jit.script(nb).save(os.path.join(OUT, "cfconv_nb.pt"))
    torch.jit.script(cf).save(os.path.join(OUT, "cfconv.pt"))
    exp["cfconv"] = dict(positions=p2.detach().tolist(), input=x.detach().tolist(), output=y.detach().tolist(), pos_grad=p2.grad.tolist(),
                         input_grad=x.grad.tolist())

    # --- PME and getNeighborPairs as scripted modules (TestPme.py:196-258, TestNeighbors.py:273-289)
    class PmeModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pme = PME(14, 16, 15, 5, 4.985823141035867, 138.935, torch.tensor([[1], [0], [-1], [-1], [-1], [-1], [-1], [-1], [-1]]))

        def forward(self, positions, charges, box_vectors):
            return self.pme.compute_direct(positions, charges, 0.5, box_vectors) + self.pme.compute_reciprocal(positions, charges, box_vectors)

    p3 = torch.tensor(3 * rng.uniform(0, 1, (9, 3)) - 1, dtype=torch.float32, requires_grad=True)
    q = torch.tensor([(i - 4) * 0.1 for i in range(9)], dtype=torch.float32)
    box = torch.tensor([[1, 0, 0], [0, 1.1, 0], [0, 0, 1.2]], dtype=torch.float32)
    pm = torch.jit.script(PmeModule())
    e = pm(p3, q, box)
    e.backward()
    pm.save(os.path.join(OUT, "pme.pt"))
    exp["pme"] = dict(positions=p3.detach().tolist(), charges=q.tolist(), box=box.tolist(), energy=float(e), grad=p3.grad.tolist())

    class PairModule(torch.nn.Module):
        def forward(self, positions, box_vectors):
            neighbors, deltas, distances, _ = getNeighborPairs(positions, cutoff=1.0, max_num_pairs=64, box_vectors=box_vectors)
            mask = torch.isnan(distances)
            return torch.sum(distances[~mask] ** 2)

    p4 = torch.tensor(rng.uniform(0, 3, (20, 3)), dtype=torch.float32, requires_grad=True)
    b4 = torch.tensor([[3.0, 0, 0], [0, 3.0, 0], [0, 0, 3.0]])
    nm = torch.jit.script(PairModule())
    s = nm(p4, b4)
    s.backward()
    nm.save(os.path.join(OUT, "neighbors.pt"))
    exp["neighbors"] = dict(positions=p4.detach().tolist(), box=b4.tolist(), value=float(s), grad=p4.grad.tolist())

    flat = {"%s.%s" % (k, kk): np.asarray(vv, np.float32 if kk != "species" else np.int64) for k, v in exp.items() for kk, vv in v.items()}
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **flat)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
