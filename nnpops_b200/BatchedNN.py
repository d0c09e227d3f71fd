"""Literal BatchedNN (mirror of the reference's NNPOps.BatchedNN, src/pytorch/BatchedNN.py:41-122): per-atom replicated, zero-padded
weights driven through NNPOpsBatchedNN::BatchedLinear, CELU(0.1) between layers, energies = sum / num_models.  Kept for drop-in
compatibility with small molecules; it inherits the reference's memory cost (weights[1, N, M, out, in]).  The scalable path is
nnpops_b200.OptimizedTorchANI.FusedANI (species-grouped tensor-core MLP)."""
from typing import List, NamedTuple, Tuple, Union

import torch
from torch import nn, Tensor
from torch.nn import functional as F

from . import torch_ops

torch_ops.load()
batchedLinear = torch.ops.NNPOpsBatchedNN.BatchedLinear


class SpeciesEnergies(NamedTuple):
    species: Tensor
    energies: Tensor


class _BatchedNN(torch.nn.Module):
    def __init__(self, converter, ensemble, atomicNumbers: Tensor):
        super().__init__()
        species_list = converter((atomicNumbers, torch.empty(0))).species[0].tolist()
        members = list(ensemble) if isinstance(ensemble, torch.nn.ModuleList) else [ensemble]
        models = [list(model.values()) for model in members]
        for ilayer in [0, 2, 4, 6]:   # the nn.Linear members of each per-species Sequential (BatchedNN.py:55-59)
            layers = [[model[s][ilayer] for s in species_list] for model in models]
            weights, biases = self.batchLinearLayers(layers)
            self.register_buffer(f'layer{ilayer}_weights', weights)
            self.register_buffer(f'layer{ilayer}_biases', biases)

    @staticmethod
    def batchLinearLayers(layers: List[List[nn.Linear]]) -> Tuple[Tensor, Tensor]:
        num_models, num_atoms = len(layers), len(layers[0])
        flat = [l for sub in layers for l in sub]
        max_out = max(l.out_features for l in flat)
        max_in = max(l.in_features for l in flat)
        weights = torch.zeros((1, num_atoms, num_models, max_out, max_in), dtype=torch.float32)
        biases = torch.zeros((1, num_atoms, num_models, max_out, 1), dtype=torch.float32)
        for imodel, sublayers in enumerate(layers):
            for iatom, layer in enumerate(sublayers):
                num_out, num_in = layer.weight.shape
                weights[0, iatom, imodel, :num_out, :num_in] = layer.weight.detach()
                biases[0, iatom, imodel, :num_out, 0] = layer.bias.detach()
        return weights, biases

    def forward(self, species_aev: Tuple[Tensor, Tensor]) -> SpeciesEnergies:
        species, aev = species_aev
        vectors = aev.unsqueeze(-2).unsqueeze(-1)     # [mols, atoms, features] -> [mols, atoms, 1, features, 1]
        vectors = batchedLinear(vectors, self.layer0_weights, self.layer0_biases)
        vectors = F.celu(vectors, alpha=0.1)
        vectors = batchedLinear(vectors, self.layer2_weights, self.layer2_biases)
        vectors = F.celu(vectors, alpha=0.1)
        vectors = batchedLinear(vectors, self.layer4_weights, self.layer4_biases)
        vectors = F.celu(vectors, alpha=0.1)
        vectors = batchedLinear(vectors, self.layer6_weights, self.layer6_biases)
        energies = torch.sum(vectors, (1, 2, 3, 4)) / vectors.shape[2]   # one fused sum / mean (BatchedNN.py:105-109)
        return SpeciesEnergies(species, energies)


class TorchANIBatchedNN(torch.nn.ModuleList):
    def __init__(self, converter, ensemble, atomicNumbers: Tensor):
        super().__init__([_BatchedNN(converter, ensemble, atomicNumbers)])

    def forward(self, species_aev: Tuple[Tensor, Tensor]) -> SpeciesEnergies:
        return self[0].forward(species_aev)
