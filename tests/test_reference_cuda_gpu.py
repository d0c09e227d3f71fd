"""Our AEV kernels against the reference's own CUDA kernels (CudaANISymmetryFunctions, compiled unmodified for sm_100a into
oracle/_ref/libnnpops_ref_cuda.so by oracle/Makefile) on the same GPU and the same inputs: forward AEVs and the position gradient of a
random upstream gradient, at the two BASELINE sizes the reference's N x N neighbour table allows (5 000-atom protein, 40 000-atom
periodic water box)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O
from systems import ANI2X, cubic_box, lattice, protein_species, rel_err, water_species

pytestmark = pytest.mark.gpu


def _tables():
    return O.fn_tables(ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])


@pytest.mark.parametrize("name,n,periodic", [("protein5000", 5000, False), ("water40000", 40000, True)])
def test_aev_forward_backward_match_the_reference_cuda_kernels(name, n, periodic):
    if O.ref_cuda_lib() is None:
        pytest.skip("oracle/_ref/libnnpops_ref_cuda.so was not built (needs /root/reference at build time)")
    from nnpops_b200.SymmetryFunctions import Holder
    if periodic:
        pos, L = lattice(n, 2.154, 0.3, 4000)
        species, box = water_species(n), cubic_box(L)
    else:
        pos, _ = lattice(n, 2.0, 0.3, 11)
        species, box = protein_species(n), None
    rfn, afn = _tables()
    p = torch.tensor(pos, device="cuda")
    b = torch.tensor(box, device="cuda") if box is not None else None
    ref = O.RefCudaANI(species, 7, 5.2, 3.5, rfn, afn, periodic)
    r0, a0 = ref.forward(p, b)
    ours = Holder(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species.tolist())
    r1, a1 = ours.forward(p, b)
    g = torch.Generator(device="cuda").manual_seed(3)
    rg = torch.randn(r0.shape, device="cuda", generator=g)
    ag = torch.randn(a0.shape, device="cuda", generator=g)
    d0 = ref.backward(rg, ag)
    d1 = ours.backward([rg, ag])
    torch.cuda.synchronize()
    errs = {"radial": rel_err(r1.cpu().numpy(), r0.cpu().numpy()), "angular": rel_err(a1.cpu().numpy(), a0.cpu().numpy()),
            "grad": rel_err(d1.cpu().numpy(), d0.cpu().numpy())}
    print(name, "vs reference CUDA kernels:", errs)
    ref.close()
    # both sides are fp32 with different summation orders (the reference accumulates with float atomics): 1e-5 relative, as north_star asks
    assert errs["radial"] < 1e-5, errs
    if periodic:
        assert errs["angular"] < 1e-5 and errs["grad"] < 1e-5, errs
    else:
        # Measured on B200 (sm_100a build of the unmodified source): the reference's angular CUDA kernel returns values that differ from
        # the reference's own CPU implementation by O(1) on this 5 000-atom non-periodic input (it agrees at 600 atoms and on the
        # periodic 40 000-atom box).  The arbiter here is therefore the CPU reference restatement (pinned to the compiled reference CPU
        # class and the TorchANI goldens in tests/test_oracle_ani.py); the deviation of the CUDA reference is printed, not asserted.
        r_o, a_o = O.ani_forward(pos, species, 7, 5.2, 3.5, rfn, afn)
        g_o = O.ani_backward(pos, species, 7, 5.2, 3.5, rfn, afn, rg.cpu().numpy(), ag.cpu().numpy())
        ours_vs_cpu = {"radial": rel_err(r1.cpu().numpy(), r_o), "angular": rel_err(a1.cpu().numpy(), a_o), "grad": rel_err(d1.cpu().numpy(), g_o)}
        refcuda_vs_cpu = {"angular": rel_err(a0.cpu().numpy(), a_o), "grad": rel_err(d0.cpu().numpy(), g_o)}
        print(name, "ours vs CPU reference:", ours_vs_cpu, " reference CUDA vs CPU reference:", refcuda_vs_cpu)
        assert ours_vs_cpu["radial"] < 1e-5 and ours_vs_cpu["angular"] < 1e-5 and ours_vs_cpu["grad"] < 1e-5, ours_vs_cpu
