"""TEST-ONLY `torchani` stand-in (never shipped): exactly the names the reference's wrappers import in their class bodies
(src/pytorch/SymmetryFunctions.py:63-64, BatchedNN.py:39,116, EnergyShifter.py:34-35, SpeciesConverter.py:26), backed by
tests/fake_torchani.py.  Put this directory and tests/ on PYTHONPATH to import the reference's NNPOps package without torchani."""
from fake_torchani import AEVComputer, SpeciesConverter  # noqa: F401
from . import nn, utils  # noqa: F401
