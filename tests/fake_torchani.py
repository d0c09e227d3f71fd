"""TEST-ONLY stand-in for the handful of torchani attributes the reference wrappers duck-type:
AEVComputer constants, SpeciesConverter, an Ensemble of ANIModel-like per-species Sequentials, EnergyShifter.sae.
Random ANI-2x-shaped weights; never shipped as product."""
from collections import OrderedDict, namedtuple

import numpy as np
import torch

from systems import ANI2X

SpeciesCoordinates = namedtuple("SpeciesCoordinates", ["species", "coordinates"])


class AEVComputer:
    def __init__(self, Rcr=5.1):
        self.num_species = 7
        self.Rcr, self.Rca = Rcr, ANI2X["Rca"]
        self.EtaR = torch.tensor(ANI2X["EtaR"]).view(-1, 1)
        self.ShfR = torch.tensor(ANI2X["ShfR"]).view(1, -1)
        self.EtaA = torch.tensor(ANI2X["EtaA"]).view(-1, 1, 1, 1)
        self.Zeta = torch.tensor(ANI2X["Zeta"]).view(1, -1, 1, 1)
        self.ShfA = torch.tensor(ANI2X["ShfA"]).view(1, 1, -1, 1)
        self.ShfZ = torch.tensor(ANI2X["ShfZ"], dtype=torch.float64).view(1, 1, 1, -1)


class SpeciesConverter(torch.nn.Module):
    ELEMENTS = [1, 6, 7, 8, 16, 9, 17]

    def __init__(self):
        super().__init__()
        conv = torch.full((120,), -1, dtype=torch.long)
        for i, z in enumerate(self.ELEMENTS):
            conv[z] = i
        self.register_buffer("conv_tensor", conv)

    def forward(self, input_):
        numbers, coords = input_
        return SpeciesCoordinates(self.conv_tensor[numbers], coords)


class ANIModel(torch.nn.ModuleDict):
    def __init__(self, hidden, rng):
        mods = OrderedDict()
        for name, h in zip("H C N O S F Cl".split(), hidden):
            dims = [1008] + list(h) + [1]
            layers = []
            for i in range(len(dims) - 1):
                lin = torch.nn.Linear(dims[i], dims[i + 1])
                bound = 1 / np.sqrt(dims[i])
                with torch.no_grad():
                    lin.weight.copy_(torch.tensor(rng.uniform(-bound, bound, (dims[i + 1], dims[i])), dtype=torch.float32))
                    lin.bias.copy_(torch.tensor(rng.uniform(-bound, bound, dims[i + 1]), dtype=torch.float32))
                layers.append(lin)
                if i < len(dims) - 2:
                    layers.append(torch.nn.CELU(0.1))
            mods[name] = torch.nn.Sequential(*layers)
        super().__init__(mods)


class EnergyShifter:
    def __init__(self):
        self.self_energies = torch.tensor([-0.5, -38.0, -54.7, -75.2, -398.1, -99.8, -460.1], dtype=torch.float64)

    def sae(self, species):
        return self.self_energies[species].sum(dim=1)


class Model:
    def __init__(self, hidden, ensemble, seed, Rcr=5.1):
        rng = np.random.default_rng(seed)
        self.species_converter = SpeciesConverter()
        self.aev_computer = AEVComputer(Rcr)
        self.neural_networks = torch.nn.ModuleList([ANIModel(hidden, rng) for _ in range(ensemble)])
        self.energy_shifter = EnergyShifter()

    def networks_numpy(self):
        """networks[s][e][l] = (W, b) for the oracle / FusedANI."""
        out = []
        for s in range(7):
            members = []
            for model in self.neural_networks:
                seq = list(model.values())[s]
                members.append([(m.weight.detach().numpy(), m.bias.detach().numpy()) for m in seq if isinstance(m, torch.nn.Linear)])
            out.append(members)
        return out
