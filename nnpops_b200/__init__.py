"""nnpops_b200 -- B200-native (sm_100a) implementation of the NNPOps per-atom hot path.

Importing the package loads the hand-written CUDA library (libnnpops_b200.so) through its C ABI and fails loudly when it has not
been built.  Sub-modules mirror the reference's Python surface (src/pytorch/*.py):

    nnpops_b200.SymmetryFunctions   TorchANISymmetryFunctions        (reference: SymmetryFunctions.py)
    (BatchedNN.py has no mirror: the reference's own file runs unchanged on the NNPOpsBatchedNN::BatchedLinear op of libNNPOpsPyTorch.so)
    nnpops_b200.OptimizedTorchANI   OptimizedTorchANI, FusedANI      (reference: OptimizedTorchANI.py)
    nnpops_b200.neighbors           getNeighborPairs                 (reference: neighbors/getNeighborPairs.py)
    nnpops_b200.CFConv / CFConvNeighbors                             (reference: CFConv.py, CFConvNeighbors.py)
    nnpops_b200.pme                 PME                              (reference: pme/pme.py)
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)

__all__ = ["_lib"]
