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


# ---------------------------------------------------------------------------------------------------------------------
# z-slab RELAY of the propagation (BASELINE north_star: "NCCL ... only for halo slices of the memory bank and the final
# label gather"; SURVEY §8e Phase B). Every rank keeps the features of ITS z-slab only. The chain of a seed slice runs
# through the slabs in order; what crosses a slab boundary is the part of the predictor's per-object memory bank the
# next frames can still see: the conditioning frame(s) (maskmem_features [4096,64] bf16 + object pointer), the object
# pointers of the last 15 tracked frames and the memory features of the last num_maskmem-1 frames — ~0.5 MB per object
# and boundary, sent point-to-point to the neighbour. Stored tensors travel bit for bit, so a relayed run produces the
# single-GPU label volume exactly.
# ---------------------------------------------------------------------------------------------------------------------
def owner_of_frame(frame: int, Z: int, world: int) -> int:
    for r in range(world):
        z0, z1 = zslab_range(Z, r, world)
        if z0 <= frame < z1:
            return r
    raise ValueError(f"frame {frame} outside [0, {Z})")


def halo_pack(state: dict, n_obj: int, ptr_frames: List[int], mem_frames: List[int]):
    """Serialise the memory-bank halo of objects 0..n_obj-1 of a video-predictor state: all conditioning frames
    (memory + pointer), `ptr_frames` (pointer; frames missing from the bank are skipped) and `mem_frames` (memory).
    Returns (meta int64 [..], mem blob [n, 4096*64] in the bank's storage dtype or None, ptr blob fp32 [m, 256] or None)."""
    per = state["output_dict_per_obj"]
    cond_frames = sorted(per[0]["cond_frame_outputs"]) if n_obj else []
    pf = [f for f in ptr_frames if all(f in per[i]["non_cond_frame_outputs"] for i in range(n_obj))]
    mf = [f for f in mem_frames if f in pf]
    mems, ptrs = [], []
    for i in range(n_obj):
        for f in cond_frames:
            out = per[i]["cond_frame_outputs"][f]
            mems.append(out["maskmem_features"].reshape(1, -1))
            ptrs.append(out["obj_ptr"].reshape(1, -1))
        for f in pf:
            ptrs.append(per[i]["non_cond_frame_outputs"][f]["obj_ptr"].reshape(1, -1))
        for f in mf:
            mems.append(per[i]["non_cond_frame_outputs"][f]["maskmem_features"].reshape(1, -1))
    meta = torch.tensor([n_obj, len(cond_frames), len(pf), len(mf)] + cond_frames + pf + mf, dtype=torch.int64)
    return meta, (torch.cat(mems, 0).contiguous() if mems else None), (torch.cat(ptrs, 0).contiguous() if ptrs else None)


def halo_unpack(meta: torch.Tensor, mem: Optional[torch.Tensor], ptr: Optional[torch.Tensor]):
    """Inverse of halo_pack -> list over objects of {"cond": {f: out}, "non_cond": {f: out}} (out dicts in the
    predictor's layout; entries without memory carry no "maskmem_features" key)."""
    m = [int(v) for v in meta.tolist()]
    n_obj, nc, npf, nmf = m[:4]
    cond_frames, pf, mf = m[4:4 + nc], m[4 + nc:4 + nc + npf], m[4 + nc + npf:4 + nc + npf + nmf]
    objs = []
    mi = pi = 0
    for _ in range(n_obj):
        cond, non = {}, {}
        for f in cond_frames:
            cond[f] = {"maskmem_features": mem[mi].view(-1, 64), "maskmem_pos_enc": True, "pred_masks": None,
                       "obj_ptr": ptr[pi].view(1, -1), "object_score_logits": None}
            mi += 1
            pi += 1
        for f in pf:
            non[f] = {"maskmem_pos_enc": True, "pred_masks": None, "obj_ptr": ptr[pi].view(1, -1),
                      "object_score_logits": None}
            pi += 1
        for f in mf:
            non[f]["maskmem_features"] = mem[mi].view(-1, 64)
            mi += 1
        objs.append({"cond": cond, "non_cond": non})
    return objs


def halo_send(dst: int, meta, mem, ptr, device, group=None) -> int:
    """Point-to-point send of a packed halo; returns the payload bytes. 16-bit memory travels as bytes (uint8 view)."""
    comm_dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mem_b = mem.view(torch.uint8).reshape(-1) if mem is not None else torch.empty(0, dtype=torch.uint8)
    ptr_b = ptr.reshape(-1) if ptr is not None else torch.empty(0, dtype=torch.float32)
    head = torch.tensor([meta.numel(), mem_b.numel(), ptr_b.numel(), 0 if mem is None else mem.element_size()],
                        dtype=torch.int64)
    dist.send(head.to(comm_dev), dst, group=group)
    dist.send(meta.to(comm_dev), dst, group=group)
    if mem_b.numel():
        dist.send(mem_b.to(comm_dev).contiguous(), dst, group=group)
    if ptr_b.numel():
        dist.send(ptr_b.to(comm_dev).contiguous(), dst, group=group)
    return int(mem_b.numel() + 4 * ptr_b.numel() + 8 * meta.numel() + 32)


def halo_recv(src: int, device, group=None):
    comm_dev = device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    head = torch.empty(4, dtype=torch.int64, device=comm_dev)
    dist.recv(head, src, group=group)
    n_meta, n_mem, n_ptr, esz = [int(v) for v in head.tolist()]
    meta = torch.empty(n_meta, dtype=torch.int64, device=comm_dev)
    dist.recv(meta, src, group=group)
    mem = ptr = None
    if n_mem:
        raw = torch.empty(n_mem, dtype=torch.uint8, device=comm_dev)
        dist.recv(raw, src, group=group)
        mem = raw.to(device).view(torch.bfloat16 if esz == 2 else torch.float32).view(-1, 4096 * 64)
    if n_ptr:
        p = torch.empty(n_ptr, dtype=torch.float32, device=comm_dev)
        dist.recv(p, src, group=group)
        ptr = p.to(device).view(-1, 256)
    return meta.cpu(), mem, ptr


def halo_install(predictor, state: dict, obj_ids: List[int], objs: List[dict]) -> None:
    """Put an unpacked halo into a predictor state (creating the objects on first use). Existing entries are kept:
    the receiving rank's own results are authoritative."""
    for obj_id, h in zip(obj_ids, objs):
        idx = predictor._obj_id_to_idx(state, obj_id)
        od = state["output_dict_per_obj"][idx]
        for f, out in h["cond"].items():
            od["cond_frame_outputs"].setdefault(f, out)
        for f, out in h["non_cond"].items():
            if f not in od["cond_frame_outputs"]:
                od["non_cond_frame_outputs"].setdefault(f, out)


def allgather_label_slabs(labels: torch.Tensor, Z: int, group=None) -> torch.Tensor:
    """Every rank holds valid labels for its z-slab of the full-size volume `labels` [Z,Y,X] (16-bit): after the call every
    rank holds the whole volume. One all_gather of the slabs as bytes (2 B per voxel, (N-1)/N of the volume received)."""
    world, rank = world_rank(group)
    if world == 1:
        return labels
    ranges = [zslab_range(Z, r, world) for r in range(world)]
    zmax = max(b - a for a, b in ranges)
    z0, z1 = ranges[rank]
    plane = labels[0].numel()
    mine = torch.zeros((zmax * plane,), dtype=labels.dtype, device=labels.device)
    mine[:(z1 - z0) * plane] = labels[z0:z1].reshape(-1)
    full = torch.empty((world, zmax * plane), dtype=labels.dtype, device=labels.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(full.view(torch.uint8), mine.view(torch.uint8), group=group)
    else:
        dist.all_gather(list(full.view(torch.uint8).unbind(0)), mine.view(torch.uint8), group=group)
    for r, (a, b) in enumerate(ranges):
        if r != rank:
            labels[a:b] = full[r, :(b - a) * plane].view(b - a, *labels.shape[1:])
    return labels


def relay_propagate(predictor, state: dict, obj_ids: List[int], start: int, Z: int, seed_fn, on_frame, set_frame_key,
                    device, group=None, n_ptr_frames: int = 15) -> dict:
    """Bidirectional propagation of one seed slice with the frames sharded by z-slab (this rank tracks only its own
    frames). `seed_fn()` registers the seed prompts (called on the rank that owns `start`), `on_frame(pass_id,
    frame_idx, obj_ids, logits)` consumes a tracked frame, `set_frame_key(k, pass_id)` tells the caller which frame key
    the single-process loop would hold at the next decoder call and which pass (0 forward, 1 backward) it belongs to
    (SABER's hook files scores under the key). Returns {"bytes_sent": n, "frames": tracked here}.

    Dependencies (upstream SAM2 memory bank): a forward frame f reads the conditioning frame, the memory of f-1 and the
    object pointers of f-1 .. f-15; a backward frame reads f+1 and the pointers of f+1 .. f+15 — which, next to the seed
    slice, include FORWARD frames start+1 .. start+14. Nothing in the forward pass reads a backward result once it is 15
    frames past the seed. Schedule: the seed rank tracks forward 15 frames, then its part of the backward pass, hands the
    backward halo down, and only then finishes its forward frames and hands the forward halo up — so the backward chain
    (lower ranks) and the forward chain (higher ranks) run CONCURRENTLY. When the seed slab ends within 15 frames of the
    seed, the missing forward pointers come back from the next rank first (sequential fall-back)."""
    world, rank = world_rank(group)
    ranges = [zslab_range(Z, r, world) for r in range(world)]
    z0, z1 = ranges[rank]
    r_seed = owner_of_frame(start, Z, world)
    n_mem = max(int(getattr(predictor, "num_maskmem", 2)) - 1, 0)
    stats = {"bytes_sent": 0, "frames": 0}
    n_obj = len(obj_ids)
    seed_z1 = ranges[r_seed][1]
    need_hi = min(start + n_ptr_frames - 1, Z - 1)  # last forward frame the backward pass can see
    overlap = need_hi + 1 <= seed_z1 - 1             # the seed slab holds them all (+1: forward must be past them)

    def nonempty(r):
        return 0 <= r < world and ranges[r][1] > ranges[r][0]

    def run(first, count, reverse, pass_id, key_before):
        if count < 0:
            return
        set_frame_key(key_before, pass_id)
        for frame_idx, ids, logits in predictor.propagate_in_video(state, start_frame_idx=first,
                                                                   max_frame_num_to_track=count, reverse=reverse):
            set_frame_key(frame_idx, pass_id)
            on_frame(pass_id, frame_idx, ids, logits)
            stats["frames"] += 1

    def send_up():
        if rank >= r_seed and nonempty(rank + 1):
            pk = halo_pack(state, n_obj, list(range(max(z1 - n_ptr_frames, 0), z1)), list(range(z1 - n_mem, z1)))
            stats["bytes_sent"] += halo_send(rank + 1, *pk, device, group)

    def send_down():
        if rank <= r_seed and nonempty(rank - 1) and z1 > z0:
            pk = halo_pack(state, n_obj, list(range(z0, min(z0 + n_ptr_frames, Z))), list(range(z0, min(z0 + n_mem, Z))))
            stats["bytes_sent"] += halo_send(rank - 1, *pk, device, group)

    if rank == r_seed:
        seed_fn()
        if overlap:
            run(start, need_hi + 1 - start, False, 0, None)          # forward: start .. start+15
            run(start, start - z0, True, 1, need_hi + 1)             # backward part of this slab
            send_down()
            run(need_hi + 2, z1 - 1 - (need_hi + 2), False, 0, need_hi + 1)  # rest of the forward frames
            send_up()
        else:
            run(start, z1 - 1 - start, False, 0, None)
            send_up()
            if need_hi >= seed_z1 and nonempty(r_seed + 1):
                if ranges[r_seed + 1][1] <= need_hi and nonempty(r_seed + 2):
                    raise RuntimeError("z-slab relay needs slabs of at least 15 frames next to the seed slice")
                objs = halo_unpack(*halo_recv(r_seed + 1, device, group))
                for o in objs:
                    o["cond"] = {}
                halo_install(predictor, state, obj_ids, objs)
            run(start, start - z0, True, 1, z1 - 1)
            send_down()
    elif rank > r_seed and z1 > z0:
        halo_install(predictor, state, obj_ids, halo_unpack(*halo_recv(rank - 1, device, group)))
        if rank == r_seed + 1 and not overlap and need_hi >= seed_z1:
            # the seed rank waits for the forward pointers seed_z1 .. need_hi before its backward pass: track them first
            if z1 <= need_hi and nonempty(rank + 1):
                raise RuntimeError("z-slab relay needs slabs of at least 15 frames next to the seed slice")
            last = min(need_hi, z1 - 1)
            run(z0, last - z0, False, 0, z0 - 1)
            pk = halo_pack(state, n_obj, list(range(seed_z1, last + 1)), [])
            stats["bytes_sent"] += halo_send(r_seed, *pk, device, group)
            run(last + 1, z1 - 1 - (last + 1), False, 0, last)
        else:
            run(z0, z1 - 1 - z0, False, 0, z0 - 1)
        send_up()
    elif rank < r_seed and z1 > z0:
        halo_install(predictor, state, obj_ids, halo_unpack(*halo_recv(rank + 1, device, group)))
        run(z1 - 1, z1 - 1 - z0, True, 1, z1)
        send_down()
    return stats
