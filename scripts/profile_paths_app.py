"""ncu driver for the neighbour / PME path at config-5 size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from systems import lattice, cubic_box
from nnpops_b200.neighbors import getNeighborPairs
from nnpops_b200.pme.pme import pme_direct
n = 200000
pos_np, L = lattice(n, 0.2154, 0.3, 5005)
pos = torch.tensor(pos_np, device="cuda"); box = torch.tensor(cubic_box(L), device="cuda")
q = torch.tensor(np.random.default_rng(5).uniform(-0.5, 0.5, n).astype(np.float32), device="cuda")
for _ in range(2):
    nb, d, r, f = getNeighborPairs(pos, 0.9, 33_000_000, box)
    e = pme_direct(pos, q, nb, d, r, torch.zeros((n, 0), dtype=torch.int32, device="cuda"), 2.92, 138.935)
torch.cuda.synchronize()
print(int(f), float(e))
