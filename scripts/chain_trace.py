"""Development aid: run the CHAIN_TRACE variant of the chain kernel (scripts/build_variant.sh trace mlp_chain.cu -DCHAIN_TRACE) for two
evaluations of the 50k-atom water box; the kernel prints CTA 0's event log ("T role tag clock") on stdout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species, lattice, cubic_box
from mlp_ref import random_networks
from nnpops_b200.OptimizedTorchANI import FusedANI
n = 50000
pos, L = lattice(n, 2.154, 0.3, 3000)
nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
m = FusedANI(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], water_species(n), nets)
p = torch.tensor(pos, device="cuda"); b = torch.tensor(cubic_box(L), device="cuda")
for _ in range(2):
    e, g = m.energy_and_gradient(p, b)
    torch.cuda.synchronize()
    print("EVAL", float(e.cpu()[0]), flush=True)
