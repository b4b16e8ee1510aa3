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


# ---------------------------------------------------------------------------------------------------------------------
# z-axis propagation across GPUs: Phase A shards the frame encodes by z-slab and exchanges the
# cached features; Phase B shards the tracked OBJECTS (GPUPool rule k -> rank k % world: every chain is sequential in
# z, objects are independent); Phase C merges label volumes with an element-wise max ("higher object id wins").
# ---------------------------------------------------------------------------------------------------------------------
def world_rank(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_max_labels(labels: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise max of a uint16-payload (int16 storage) label volume over ranks, in place ("higher object id wins";
    SURVEY 8e Phase C). 16-bit integers are not collective dtypes, so the labels travel as BYTES:
    NCCL: reduce-scatter by z-chunk built from all_to_all_single (every rank receives the N versions of ITS chunk, 2 B per
          voxel), a local max, then all_gather_into_tensor — (N-1)/N x 2 B sent and received per voxel and phase instead
          of the 4-byte widened all-reduce of round 1 (1.07 GB -> 0.47 GB per rank and phase for 300 x 928 x 960 at N = 8);
    gloo (CPU tests): the widened int32 all-reduce (all_to_all is not a gloo collective)."""
    world, rank = world_rank(group)
    if world == 1:
        return labels
    flat = labels.view(-1)
    if dist.get_backend(group) != "nccl":
        chunk = 64 * 1024 * 1024
        for s in range(0, flat.numel(), chunk):
            part = flat[s:s + chunk]
            wide = part.to(torch.int32) & 0xFFFF  # uint16 payload
            dist.all_reduce(wide, op=dist.ReduceOp.MAX, group=group)
            part.copy_(wide.to(torch.int16))  # values < 2^16 wrap back to the same 16 bits
        return labels
    n = flat.numel()
    per = (n + world - 1) // world
    per += (-per) % 8
    send = torch.zeros((world, per), dtype=torch.int16, device=labels.device)
    send.view(-1)[:n] = flat
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(torch.uint8), send.view(torch.uint8), group=group)
    # object ids are < 2^15 (SABER tracks a few hundred objects per seed slice): the signed maximum is the unsigned one
    mine = recv.amax(dim=0).contiguous()
    dist.all_gather_into_tensor(send.view(torch.uint8), mine.view(torch.uint8), group=group)
    flat.copy_(send.view(-1)[:n])
    return labels


def allreduce_any(flags: torch.Tensor, group=None) -> torch.Tensor:
    """Logical OR of a small per-slice flag vector over ranks (int32 MAX): which z-slices any rank has already labelled."""
    world, _ = world_rank(group)
    if world == 1:
        return flags
    f = flags.to(torch.int32)
    dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)
    return f


def exchange_frame_features(cached: dict, Z: int, group=None) -> None:
    """Phase A exchange: every rank holds `cached[f] = {"feat","s1","s0"}` for the frames of its z-slab; after the call
    every rank holds all Z frames. ONE all_gather_into_tensor per feature level (slabs padded to the largest slab);
    fp32 on the wire so that a sharded run is bit-identical to the single-GPU run (16 MB per frame: 4.8 GB for Z = 300,
    7/8 of it received per rank = 6 ms at the measured 770 GB/s NVLink peer bandwidth; round 1 issued one broadcast per
    slab and level)."""
    world, rank = world_rank(group)
    if world == 1:
        return
    some = next(iter(cached.values()))
    dev = some["feat"].device
    ranges = [zslab_range(Z, r, world) for r in range(world)]
    zmax = max(z1 - z0 for z0, z1 in ranges)
    z0, z1 = ranges[rank]
    for key, width, rows in (("feat", 256, 4096), ("s1", 64, 16384), ("s0", 32, 65536)):
        mine = torch.zeros((zmax, rows, width), dtype=torch.float32, device=dev)
        for j, f in enumerate(range(z0, z1)):
            mine[j] = cached[f][key]
        full = torch.empty((world, zmax, rows, width), dtype=torch.float32, device=dev)
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(full, mine, group=group)
        else:
            dist.all_gather(list(full.unbind(0)), mine, group=group)
        for r, (a, b) in enumerate(ranges):
            if r == rank:
                continue
            for j, f in enumerate(range(a, b)):
                cached.setdefault(f, {})[key] = full[r, j]


def merge_captured_scores(per_rank: List[dict], n_local: List[int]) -> dict:
    """Rebuild the single-process hook log from per-rank logs. per_rank[r][fidx] is rank r's flat list of object-score
    values filed under frame key `fidx` (one value per local object per decoder call group, in call order); n_local[r]
    is rank r's number of tracked objects; global object k lives on rank k % world at local index k // world. Every
    rank saw the same sequence of call groups, so group g of the merged log lists the objects in global order."""
    world = len(per_rank)
    keys = []
    for d in per_rank:
        for k in d:
            if k not in keys:
                keys.append(k)
    total = sum(n_local)
    merged = {}
    for k in keys:
        groups = None
        for r in range(world):
            if n_local[r] == 0:
                continue
            vals = per_rank[r].get(k, [])
            assert len(vals) % n_local[r] == 0, "ranks disagree on the call-group structure"
            g = len(vals) // n_local[r]
            groups = g if groups is None else groups
            assert groups == g, "ranks disagree on the number of call groups"
        out = []
        for g in range(groups or 0):
            for obj in range(total):
                r, li = obj % world, obj // world
                out.append(per_rank[r][k][g * n_local[r] + li])
        merged[k] = out
    return merged
