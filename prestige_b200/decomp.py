"""Slab decomposition along x for multi-GPU runs: host-side planning.

Every rank owns a whole number of cell layers, so neighbouring slabs share cell boundaries and a
rank's outermost owned layer is exactly the ghost layer its neighbour needs (cutoff <= cell edge).
The device side (csrc/halo.cu) ships those layers with NCCL send/recv; this module only decides
who owns what, and is what the CPU (gloo) tests exercise.
"""
from __future__ import annotations

import math

import numpy as np


def slab_bounds(lo_x: float, cell: float, layers_per_rank: int, rank: int):
    """[lo, hi) of rank's slab: layers_per_rank cell layers starting at lo_x + rank * layers_per_rank * cell."""
    return lo_x + rank * layers_per_rank * cell, lo_x + (rank + 1) * layers_per_rank * cell


def split_layers(n_layers: int, world: int):
    """Near-equal split of n_layers cell layers over world ranks -> list of (first_layer, n_layers)."""
    base, rem = divmod(n_layers, world)
    out, first = [], 0
    for r in range(world):
        k = base + (1 if r < rem else 0)
        out.append((first, k))
        first += k
    return out


def owner_mask(x: np.ndarray, lo: float, hi: float, first: bool, last: bool) -> np.ndarray:
    """Particles a slab owns: x in [lo, hi); the outer slabs also take whatever lies beyond the box."""
    m = np.ones(len(x), bool)
    if not first:
        m &= x >= lo
    if not last:
        m &= x < hi
    return m


def layer_index(x: np.ndarray, lo: float, cell: float, n_layers: int) -> np.ndarray:
    """Cell layer of each owned particle inside its slab, clamped like the device key kernel."""
    c = np.floor((x - lo) * (1.0 / cell)).astype(np.int64)
    return np.clip(c, 0, n_layers - 1)


def edge_layers(x: np.ndarray, lo: float, cell: float, n_layers: int):
    """Boolean masks (to_left, to_right): the owned particles a neighbour needs as ghosts."""
    c = layer_index(x, lo, cell, n_layers)
    return c == 0, c == n_layers - 1


def ghost_capacity(ny_particles: int, nz_particles: int, cell: float, dx: float, slack: float = 1.5) -> int:
    """Upper estimate of one ghost layer's particle count for a lattice of spacing dx."""
    return int(math.ceil(cell / dx + 1) * ny_particles * nz_particles * slack)


def uniform_across_ranks(values: np.ndarray, dist=None, device=None) -> bool:
    """True iff every element of `values` on EVERY rank is the same number (one all-reduce of (min, -max)).
    What the option "uniform_mass_global" asks the caller to vouch for: each rank's library only sees its own uploads,
    ghosts and migrants come from the others.  `dist` = torch.distributed (initialised) or None for a single rank."""
    if values.size == 0:
        lo, hi = math.inf, -math.inf
    else:
        lo, hi = float(values.min()), float(values.max())
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.tensor([lo, -hi], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        lo, hi = float(t[0]), -float(t[1])
    return lo == hi
