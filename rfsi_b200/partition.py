"""Pixel-grid x day partitioner: one process per GPU, no collective on the compute path.

Every (pixel, day) is independent in both stages, so the raster is cut
into contiguous bands of rows (spatial coherence keeps a warp's tree paths together) or,
when there are more days than useful rows, into day ranges.  Forest and station tables
are replicated.  Results stay with the rank that made them (each rank writes its own
band); `gather_to_root` assembles them on rank 0 when a caller wants one array -- that is
the only communication, and it is outside the compute path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np


@dataclass(frozen=True)
class WorkUnit:
    rank: int
    pix_begin: int      # first raster cell (row-major)
    npix: int
    day_begin: int
    ndays: int


def _split(n: int, parts: int) -> List[tuple]:
    """Contiguous, balanced: the first n % parts pieces get one extra element."""
    base, extra = divmod(n, parts)
    out, a = [], 0
    for i in range(parts):
        b = a + base + (1 if i < extra else 0)
        out.append((a, b))
        a = b
    return out


def partition(nrow: int, ncol: int, ndays: int, world: int, mode: str = "auto") -> List[WorkUnit]:
    """mode 'rows': bands of raster rows x all days; 'days': full raster x day ranges;
    'auto': rows when every rank gets at least 16 rows, else days when ndays >= world, else rows."""
    if world < 1 or nrow < 0 or ncol < 1 or ndays < 1:
        raise ValueError("partition: bad sizes")
    if mode == "auto":
        mode = "rows" if nrow >= 16 * world or ndays < world else "days"
    units = []
    if mode == "rows":
        for r, (a, b) in enumerate(_split(nrow, world)):
            units.append(WorkUnit(r, a * ncol, (b - a) * ncol, 0, ndays))
    elif mode == "days":
        for r, (a, b) in enumerate(_split(ndays, world)):
            units.append(WorkUnit(r, 0, nrow * ncol, a, b - a))
    else:
        raise ValueError(f"partition: unknown mode {mode!r}")
    return units


def run_partitioned(compute: Callable[[WorkUnit], np.ndarray], nrow: int, ncol: int, ndays: int, rank: int,
                    world: int, mode: str = "auto") -> tuple:
    """Run this rank's unit.  compute(unit) -> (unit.ndays, unit.npix) array."""
    unit = partition(nrow, ncol, ndays, world, mode)[rank]
    if unit.npix == 0 or unit.ndays == 0:
        return unit, np.empty((unit.ndays, unit.npix))
    out = np.ascontiguousarray(compute(unit))
    if out.shape != (unit.ndays, unit.npix):
        raise ValueError(f"compute returned {out.shape}, expected {(unit.ndays, unit.npix)}")
    return unit, out


def gather_to_root(unit: WorkUnit, local: np.ndarray, nrow: int, ncol: int, ndays: int, group=None,
                   mode: str = "auto") -> Optional[np.ndarray]:
    """Assemble the (ndays, nrow*ncol) result on rank 0 with torch.distributed (gloo or nccl).
    Returns the full array on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    units = partition(nrow, ncol, ndays, world, mode)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    if rank == 0:
        full = np.empty((ndays, nrow * ncol), dtype=local.dtype)
        bufs = [torch.empty((u.ndays, u.npix), dtype=mine.dtype, device=dev) for u in units]
        bufs[0] = mine
        reqs = [dist.irecv(bufs[u.rank], src=u.rank, group=group) for u in units[1:] if u.npix and u.ndays]
        for r in reqs:
            r.wait()
        for u in units:
            full[u.day_begin:u.day_begin + u.ndays, u.pix_begin:u.pix_begin + u.npix] = bufs[u.rank].cpu().numpy()
        return full
    if unit.npix and unit.ndays:
        dist.send(mine, dst=0, group=group)
    return None
