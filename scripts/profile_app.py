"""Small driver for ncu: a few energy+force evaluations of the 50k-atom benchmark system (no timing, no baseline)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species, lattice, cubic_box
from mlp_ref import random_networks
from nnpops_b200.OptimizedTorchANI import FusedANI

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pos, L = lattice(n, 2.154, 0.3, 3000)
nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
m = FusedANI(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], water_species(n), nets,
             mlp_impl=os.environ.get("NNPOPS_MLP", "tcgen05"))
p = torch.tensor(pos, device="cuda"); b = torch.tensor(cubic_box(L), device="cuda")
for _ in range(evals):
    e, g = m.energy_and_gradient(p, b)
torch.cuda.synchronize()
print("energy", float(e.cpu()[0]), "work", m.work())
