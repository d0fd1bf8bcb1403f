"""Multi-GPU strip decomposition (placeholder until the strip path lands)."""
from __future__ import annotations


def init(world, rank, local):
    raise NotImplementedError("multi-GPU strips not built yet")
