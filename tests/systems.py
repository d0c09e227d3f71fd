"""Seeded synthetic systems shared by tests, bench.py and smoke()."""
import math

import numpy as np

# ANI-2x AEV constants: /root/reference/src/ani/BenchmarkCudaANISymmetryFunctions.cu:101-153 (values restated, ShfZ at full precision
# as TorchANI supplies them through SymmetryFunctions.py:78-83)
ANI2X = dict(
    num_species=7, Rcr=5.1, Rca=3.5,
    EtaR=[19.7], ShfR=[0.8 + 0.26875 * k for k in range(16)],
    EtaA=[12.5], Zeta=[14.1], ShfA=[0.8 + 0.3375 * a for a in range(8)],
    ShfZ=[(2 * z + 1) * math.pi / 8 for z in range(4)],
)
# per-species hidden widths of the ANI-2x networks, species order H C N O S F Cl
ANI2X_HIDDEN = [(256, 192, 160), (224, 192, 160), (192, 160, 128), (192, 160, 128), (160, 128, 96), (160, 128, 96), (160, 128, 96)]
ANI2X_ENSEMBLE = 8


def lattice(n, a, jitter, seed):
    """Jittered simple-cubic lattice with random vacancies; returns (positions float32 [n,3], cubic box edge)."""
    rng = np.random.default_rng(seed)
    m = math.ceil(n ** (1.0 / 3.0) - 1e-9)
    while m ** 3 < n:
        m += 1
    sites = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), -1).reshape(-1, 3)
    sites = sites[rng.permutation(m ** 3)[:n]]
    pos = ((sites + 0.5) * a + rng.uniform(-jitter, jitter, (n, 3)) * a).astype(np.float32)
    return pos, float(m * a)


def water_species(n):
    s = np.where(np.arange(n) % 3 == 2, 3, 0).astype(np.int32)
    np.random.default_rng(1).shuffle(s)
    return s


def protein_species(n, seed=2003):
    return np.random.default_rng(seed).choice(5, size=n, p=[0.5, 0.32, 0.085, 0.09, 0.005]).astype(np.int32)


def cubic_box(edge):
    return np.diag([edge, edge, edge]).astype(np.float32)


def rel_err(a, b):
    """max|a-b| / max|b| : the normalised parity metric of SURVEY.md section 8d."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
