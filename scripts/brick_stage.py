"""Development aid: stage times of ONE brick's local model of the 8-GPU box decomposition, emulated on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species, lattice, cubic_box
from mlp_ref import random_networks
from nnpops_b200.halo import HaloPlan
from nnpops_b200.OptimizedTorchANI import FusedANI
n = 50000
pos, L = lattice(n, 2.154, 0.3, 3000)
species = water_species(n); box = cubic_box(L)
nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
args = (7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"])
for grid in ((2, 2, 2), (2, 2, 1)):
    plan = HaloPlan(pos, [L, L, L], 5.2, grid)
    local = plan.local_atoms(0)
    mask = np.zeros(len(local), np.uint8); mask[:len(plan.owned[0])] = 1
    m = FusedANI(*args, species[local], nets, owned=mask)
    p = torch.tensor(pos[local], device="cuda"); b = torch.tensor(box, device="cuda")
    for _ in range(5): m.energy_and_gradient(p, b)
    torch.cuda.synchronize()
    steps = 50
    m.timing_begin(steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): m.energy_and_gradient(p, b)
    e1.record(); torch.cuda.synchronize()
    st, _ = m.timing_end()
    print(grid, "owned", len(plan.owned[0]), "local", len(local), "ms/step %.4f" % (e0.elapsed_time(e1) / steps), {k: round(v, 4) for k, v in st.items()})
