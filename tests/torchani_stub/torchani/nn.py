import torch
from fake_torchani import ANIModel, SpeciesConverter  # noqa: F401

Ensemble = torch.nn.ModuleList
