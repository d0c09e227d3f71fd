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


@pytest.mark.parametrize("periodic", [False, True])
def test_cfconv_matches_the_reference_cuda_kernels(periodic):
    """SchNet CFConv (neighbour build, forward, backprop) against the reference's own CUDA classes CudaCFConvNeighbors / CudaCFConv
    (unmodified source, sm_100a) on 6 000 atoms, width 128, 50 Gaussians, cutoff 5 A: same pair count, outputs and gradients to 1e-5.
    (The reference's N x N pair table in managed memory limits it to about 16 000 atoms.)"""
    l = O.ref_cuda_lib()
    if l is None or not hasattr(l, "refcuda_cfconv_create"):
        pytest.skip("oracle/_ref/libnnpops_ref_cuda.so was not built with the CFConv classes (needs /root/reference at build time)")
    from nnpops_b200.CFConv import CFConv
    from nnpops_b200.CFConvNeighbors import CFConvNeighbors
    n, W, Gn, cutoff, sigma = 6000, 128, 50, 5.0, 0.2
    rng = np.random.default_rng(17)
    pos, L = lattice(n, 2.154, 0.3, 4004)
    box = cubic_box(L) if periodic else None
    w1 = rng.normal(0, 0.1, (W, Gn)).astype(np.float32); b1 = rng.normal(0, 0.1, W).astype(np.float32)
    w2 = rng.normal(0, 0.1, (W, W)).astype(np.float32); b2 = rng.normal(0, 0.1, W).astype(np.float32)
    x = torch.tensor(rng.standard_normal((n, W)).astype(np.float32), device="cuda")
    go = torch.tensor(rng.standard_normal((n, W)).astype(np.float32), device="cuda")
    p = torch.tensor(pos, device="cuda")
    b = torch.tensor(box, device="cuda") if box is not None else None
    ref = O.RefCudaCFConv(n, W, Gn, cutoff, periodic, sigma, "ssp", w1, b1, w2, b2)
    ref.build(p, b)
    y0 = ref.compute(p, b, x)
    ig0, pg0 = ref.backprop(p, b, x, go)
    torch.cuda.synchronize()
    pairs0 = ref.num_pairs()
    nb = CFConvNeighbors(cutoff)
    conv = CFConv(sigma, "ssp", torch.tensor(w1.reshape(Gn, W)), torch.tensor(b1), torch.tensor(w2), torch.tensor(b2))
    pr = p.clone().requires_grad_(True); xr = x.clone().requires_grad_(True)
    nb.build(pr, b)
    y1 = conv(nb, pr, xr)
    y1.backward(go)
    errs = {"out": rel_err(y1.detach().cpu().numpy(), y0.cpu().numpy()), "input_grad": rel_err(xr.grad.cpu().numpy(), ig0.cpu().numpy()),
            "pos_grad": rel_err(pr.grad.cpu().numpy(), pg0.cpu().numpy())}
    print("CFConv vs reference CUDA kernels (periodic=%s): pairs %d / %d" % (periodic, nb.num_pairs(), pairs0), errs)
    ref.close()
    assert nb.num_pairs() == pairs0
    assert errs["out"] < 1e-5 and errs["input_grad"] < 1e-5 and errs["pos_grad"] < 1e-5, errs
