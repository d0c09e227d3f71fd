"""ncu driver for PME reciprocal space (forward + backward) at config-5 size (200 000 charges, 128^3 grid, order 5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from systems import lattice, cubic_box
from nnpops_b200.pme import PME
n = 200000
pos_np, L = lattice(n, 0.2154, 0.3, 5005)
pos = torch.tensor(pos_np, device="cuda", requires_grad=True); box = torch.tensor(cubic_box(L), device="cuda")
q = torch.tensor(np.random.default_rng(5).uniform(-0.5, 0.5, n).astype(np.float32), device="cuda", requires_grad=True)
pme = PME(128, 128, 128, 5, 2.92, 138.935, torch.zeros((n, 0), dtype=torch.int32))
for _ in range(3):
    pos.grad = None; q.grad = None
    e = pme.compute_reciprocal(pos, q, box)
    e.backward()
torch.cuda.synchronize()
print(float(e))
