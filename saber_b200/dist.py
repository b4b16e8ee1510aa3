"""One-process-per-GPU plumbing for the slice-wise path (torch.distributed; NCCL on B200, gloo in CPU tests).

* Tomogram-level data parallelism keeps the reference's GPUPool rule: task i -> worker i % n
  (REF saber/utils/parallelization.py:139-141).
* Within a tomogram the slice-wise path shards by z-slab: rank r owns slices [r*Z/N, (r+1)*Z/N) (remainder spread
  over the first ranks). Slices are independent, so the data path has no collective; only the final label gather
  (SURVEY §8e Phase C) moves data: uint16 label slabs -> rank 0, which runs the whole-volume 3-D connected components.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def tasks_for_rank(n_tasks: int, rank: int, world: int) -> List[int]:
    return [i for i in range(n_tasks) if i % world == rank]


def zslab_range(Z: int, rank: int, world: int) -> Tuple[int, int]:
    base, rem = divmod(Z, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def gather_label_slabs(slab: torch.Tensor, Z: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """slab: this rank's (z1-z0, Y, X) int16/uint16-payload labels. Returns the full (Z, Y, X) volume on rank
    ``dst`` (None elsewhere). Slabs are padded to the largest slab so one all_gather suffices."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return slab
    zmax = max(zslab_range(Z, r, world)[1] - zslab_range(Z, r, world)[0] for r in range(world))
    pad = torch.zeros((zmax,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    pad[:slab.shape[0]] = slab
    raw = pad.view(torch.uint8)  # 16-bit integer types are not collective dtypes (gloo / NCCL): move bytes
    out = [torch.empty_like(raw) for _ in range(world)] if rank == dst else None
    dist.gather(raw, out, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        z0, z1 = zslab_range(Z, r, world)
        parts.append(out[r].view(slab.dtype)[:z1 - z0])
    return torch.cat(parts, 0)


def slice_by_slice_sharded(label_fn, separate_fn, volume_slab: torch.Tensor, Z: int, group=None):
    """z-slab sharded ``slice_by_slice``: ``label_fn(volume_slab) -> labels slab`` runs on every rank,
    the slabs are gathered on rank 0 and ``separate_fn(full labels)`` runs there. Returns the separated volume on
    rank 0, None elsewhere."""
    labels = label_fn(volume_slab)
    full = gather_label_slabs(labels, Z, 0, group)
    if full is None:
        return None
    return separate_fn(full)
