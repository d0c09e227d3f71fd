"""Development aid: stage times of the fused model on the 50k-atom water box (one model, no comparison); pairs with NNPOPS_LIB_PATH."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species, lattice, cubic_box
from mlp_ref import random_networks
from nnpops_b200.OptimizedTorchANI import FusedANI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
pos, L = lattice(n, 2.154, 0.3, 3000)
nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
m = FusedANI(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], water_species(n), nets)
p = torch.tensor(pos, device="cuda"); b = torch.tensor(cubic_box(L), device="cuda")
for _ in range(5):
    e, g = m.energy_and_gradient(p, b)
m.timing_begin(20)
for _ in range(20):
    e, g = m.energy_and_gradient(p, b)
st, cnt = m.timing_end()
print(os.environ.get("NNPOPS_LIB_PATH", "default").split("/")[-1], "energy %.6f" % float(e.cpu()[0]), "mlp %.4f ms" % (st["mlp_fwd"] + st["mlp_bwd"]),
      "total %.4f" % sum(st.values()), flush=True)
