"""pywatershed_b200: B200-native (sm_100a CUDA, fp64) implementation of
pywatershed's per-timestep PRMS/NHM process chain, behind the reference's
Process / Model API.  See DESIGN.md."""
from .engine import Engine, Registry, calendar  # noqa: F401

__all__ = ["Engine", "Registry", "calendar"]
