"""Particle-mesh Ewald (mirror of the reference's NNPOps.pme package, src/pytorch/pme/__init__.py)."""
from .pme import PME

__all__ = ["PME"]
