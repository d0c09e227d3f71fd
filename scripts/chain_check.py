"""Development aid for the fused layer-chain kernel (csrc/mlp_chain.cu): on water boxes of several sizes compare energy, dE/dAEV and
forces of the fused path with the per-layer tcgen05 path (NNPOPS_NO_CHAIN=1) and, for the small box, with the fp64 ATen chain;
then time both on the BASELINE-size box.  Run under `timeout` on the GPU box: a barrier bug in a persistent kernel hangs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from systems import ANI2X, ANI2X_HIDDEN, cubic_box, lattice, water_species
from mlp_ref import random_networks
from nnpops_b200.OptimizedTorchANI import FusedANI


def build(species, nets, chain, rcr=5.2):
    if chain:
        os.environ.pop("NNPOPS_NO_CHAIN", None)
    else:
        os.environ["NNPOPS_NO_CHAIN"] = "1"
    return FusedANI(7, rcr, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], species, nets)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "300,1000,5000,50000").split(",")]
    nets = random_networks(7, ANI2X_HIDDEN, 8, 1008, 42)
    for n in sizes:
        pos, L = lattice(n, 2.154, 0.3, 3000)
        species = water_species(n)
        box = torch.tensor(cubic_box(L), device="cuda")
        p = torch.tensor(pos, device="cuda")
        out = {}
        for chain in (False, True):
            m = build(species, nets, chain)
            e, g = m.energy_and_gradient(p, box)
            torch.cuda.synchronize()
            out[chain] = (float(e.cpu()[0]), g.cpu().numpy().astype(np.float64), m.feature_grad().cpu().numpy().astype(np.float64))
            if n >= int(os.environ.get('CHAIN_TIMING_MIN', '20000')):
                for _ in range(3):
                    m.energy_and_gradient(p, box)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(20):
                    m.energy_and_gradient(p, box)
                torch.cuda.synchronize()
                print("  n=%d chain=%s: %.3f ms per evaluation" % (n, chain, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
                m.timing_begin(10)
                for _ in range(10):
                    m.energy_and_gradient(p, box)
                st, cnt = m.timing_end()
                print("  stage ms:", {k: round(v, 4) for k, v in st.items()}, flush=True)
            del m
        (e0, g0, d0), (e1, g1, d1) = out[False], out[True]
        print("n=%d  energy %.9g vs %.9g (rel %.2e)  dE/dAEV rel %.2e  forces rel %.2e" % (n, e1, e0, abs(e1 - e0) / abs(e0), rel(d1, d0), rel(g1, g0)),
              flush=True)


if __name__ == "__main__":
    main()
