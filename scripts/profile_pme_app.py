"""ncu driver for the fused PME direct-space kernel at config-5 size (200 000 charges, cutoff 0.9 nm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from systems import lattice, cubic_box
from nnpops_b200.pme.pme import pme_direct_fused
n = 200000
pos_np, L = lattice(n, 0.2154, 0.3, 5005)
pos = torch.tensor(pos_np, device="cuda"); box = torch.tensor(cubic_box(L), device="cuda")
q = torch.tensor(np.random.default_rng(5).uniform(-0.5, 0.5, n).astype(np.float32), device="cuda")
excl = torch.zeros((n, 0), dtype=torch.int32, device="cuda")
for _ in range(3):
    e = pme_direct_fused(pos, q, box, excl, 0.9, 2.92, 138.935)
torch.cuda.synchronize()
print(float(e))
