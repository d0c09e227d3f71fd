"""Plain PyTorch restatement of the reference BatchedNN arithmetic (src/pytorch/BatchedNN.py:90-111, BatchedNN.cpp:30-42):
per atom and ensemble member a chain Linear -> CELU(0.1) x3 -> Linear, energies = sum / num_models.  The reference's arithmetic
lives in ATen, so the oracle for this op is ATen itself on the CPU (fp32, and fp64 as arbiter).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch


def random_networks(num_species, hidden, ensemble, n_in, seed, dtype=np.float32):
    """networks[s][e][l] = (W [out, in], b [out]) with nn.Linear's default init (uniform +-1/sqrt(fan_in))."""
    rng = np.random.default_rng(seed)
    nets = []
    for s in range(num_species):
        dims = [n_in] + list(hidden[s]) + [1]
        members = []
        for e in range(ensemble):
            layers = []
            for l in range(len(dims) - 1):
                bound = 1.0 / np.sqrt(dims[l])
                W = rng.uniform(-bound, bound, (dims[l + 1], dims[l])).astype(dtype)
                b = rng.uniform(-bound, bound, (dims[l + 1],)).astype(dtype)
                layers.append((W, b))
            members.append(layers)
        nets.append(members)
    return nets


def mlp_energy_and_grad(aev, species, networks, dtype=torch.float32):
    """aev [N, F] (numpy), species [N] -> (energy, dE/dAEV [N, F]) via autograd on the CPU."""
    x = torch.tensor(np.asarray(aev), dtype=dtype, requires_grad=True)
    species = np.asarray(species)
    M = len(networks[0])
    total = torch.zeros((), dtype=dtype)
    for s in range(len(networks)):
        idx = np.nonzero(species == s)[0]
        if len(idx) == 0:
            continue
        xs = x[torch.from_numpy(idx)]
        for e in range(M):
            h = xs
            for l, (W, b) in enumerate(networks[s][e]):
                h = h @ torch.tensor(W, dtype=dtype).T + torch.tensor(b, dtype=dtype)
                if l < len(networks[s][e]) - 1:
                    h = torch.nn.functional.celu(h, alpha=0.1)
            total = total + h.sum()
    energy = total / M
    (g,) = torch.autograd.grad(energy, x)
    return float(energy), g.numpy()
