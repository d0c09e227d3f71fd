"""Neighbor operations (mirror of the reference's NNPOps.neighbors package, src/pytorch/neighbors/__init__.py)."""
from .getNeighborPairs import getNeighborPairs

__all__ = ["getNeighborPairs"]
