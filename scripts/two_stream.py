"""Experiment: two FusedANI instances on two CUDA streams evaluating alternate conformers (fills launch gaps and kernel tails)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from systems import ANI2X, ANI2X_HIDDEN, ANI2X_ENSEMBLE, water_species, lattice, cubic_box
from mlp_ref import random_networks
from nnpops_b200.OptimizedTorchANI import FusedANI
n = 50000
nets = random_networks(7, ANI2X_HIDDEN, ANI2X_ENSEMBLE, 1008, seed=42)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
models = [FusedANI(7, 5.2, 3.5, ANI2X["EtaR"], ANI2X["ShfR"], ANI2X["EtaA"], ANI2X["Zeta"], ANI2X["ShfA"], ANI2X["ShfZ"], water_species(n), nets) for _ in range(K)]
streams = [torch.cuda.Stream() for _ in range(K)]
confs = []
for c in range(4):
    pos, L = lattice(n, 2.154, 0.3, 3000 + c)
    confs.append((torch.tensor(pos, device="cuda"), torch.tensor(cubic_box(L), device="cuda")))
def run(steps):
    for i in range(steps):
        k = i % K
        with torch.cuda.stream(streams[k]):
            models[k].energy_and_gradient(*confs[i % 4])
run(8); torch.cuda.synchronize()
steps = 40
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for s in streams: s.wait_stream(torch.cuda.current_stream())
run(steps)
for s in streams: torch.cuda.current_stream().wait_stream(s)
t1.record(); torch.cuda.synchronize()
print("streams", K, "evals/s", steps / (t0.elapsed_time(t1) / 1e3))
