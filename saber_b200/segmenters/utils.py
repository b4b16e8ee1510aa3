"""B200 twin of REF saber/segmenters/utils.py: ``remove_duplicate_masks`` (:5-86) and ``separate_masks``
(:88-131). numpy in / numpy out keeps the reference call shape; the ``*_device`` variants are the resident
path. Integer results are bit-exact with the reference functions (tests/test_integer_stages.py)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops

_I32 = torch.int32


def _greedy_groups(area: np.ndarray, stab: np.ndarray, inter: np.ndarray, iou_threshold: float) -> List[int]:
    """Sequential grouping of REF :47-86 from the pairwise intersection matrix (inter[i,j] = -1 when the area
    ratio test already failed). Returns the index kept for every duplicate group, in group order."""
    m = len(area)
    processed = np.zeros(m, dtype=bool)
    keep: List[int] = []
    a64 = area.astype(np.int64)
    for i in range(m):
        if processed[i]:
            continue
        row = inter[i, i + 1:].astype(np.int64)
        union = a64[i] + a64[i + 1:] - row
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = np.where(union == 0, 0.0, row / np.where(union == 0, 1, union))
        dup = (row >= 0) & ~(iou < iou_threshold) & ~processed[i + 1:]
        js = np.flatnonzero(dup) + i + 1
        processed[js] = True
        processed[i] = True
        best = i
        if js.size:
            group = [i] + js.tolist()
            best = max(group, key=lambda g: stab[g])  # first maximum, like max() over the group list
        keep.append(int(best))
    return keep


def remove_duplicate_indices(bits: torch.Tensor, bbox: torch.Tensor, area: torch.Tensor, stability: Sequence[float],
                             W: int, iou_threshold: float = 0.9, area_threshold: float = 0.9,
                             area_union: Optional[torch.Tensor] = None) -> List[int]:
    """Device core of remove_duplicate_masks: packed masks [m,H,WW] (+ boxes, areas) -> indices to keep.
    ``area`` feeds the area-ratio test (the dict's 'area'); ``area_union`` (default: the same) is the mask
    popcount used for |A or B| = |A| + |B| - |A and B|."""
    m = bits.shape[0]
    if m == 0:
        return []
    if m == 1:
        return [0]
    inter = ops.pair_intersections(bits, bbox, area, W, area_threshold).cpu().numpy()
    au = area if area_union is None else area_union
    return _greedy_groups(au.cpu().numpy(), np.asarray(stability, dtype=np.float64), inter, iou_threshold)


def pack_masks(masks: np.ndarray, device) -> tuple:
    """bool [m,H,W] host masks -> (bits [m,H,WW] int32, bbox [m,4] int32 xyxy inclusive, area [m] int32) on device."""
    m, H, W = masks.shape
    WW = (W + 31) // 32
    padded = np.zeros((m, H, WW * 32), dtype=np.uint8)
    padded[:, :, :W] = masks
    words = np.packbits(padded.reshape(m, H, WW, 32), axis=-1, bitorder="little").view(np.uint32).reshape(m, H, WW)
    bbox = np.zeros((m, 4), dtype=np.int32)
    area = masks.reshape(m, -1).sum(axis=1).astype(np.int32)
    for i in range(m):
        ys, xs = np.nonzero(masks[i])
        if ys.size:
            bbox[i] = (xs.min(), ys.min(), xs.max(), ys.max())
    dev = torch.device(device)
    return (torch.from_numpy(words.view(np.int32)).to(dev), torch.from_numpy(bbox).to(dev),
            torch.from_numpy(area).to(dev))


def remove_duplicate_masks(masks: List[Dict[str, Any]], iou_threshold: float = 0.9, area_threshold: float = 0.9,
                           verbose: bool = False, device="cuda") -> List[Dict[str, Any]]:
    """REF saber/segmenters/utils.py:5-86 on host mask dicts (pairwise intersections computed on the GPU)."""
    if len(masks) < 2:
        return list(masks)
    seg = np.stack([np.asarray(m["segmentation"], dtype=bool) for m in masks])
    bits, bbox, area_pop = pack_masks(seg, device)
    area = torch.tensor([int(m["area"]) for m in masks], dtype=_I32, device=bits.device)
    stab = [m.get("stability_score", 0) for m in masks]
    keep = remove_duplicate_indices(bits, bbox, area, stab, seg.shape[2], iou_threshold, area_threshold, area_pop)
    return [masks[i] for i in keep]


def separate_masks_device(combined: torch.Tensor, min_mask_area: int = 100) -> torch.Tensor:
    """CUDA (Z,Y,X) integer volume -> int32 labels (non-negative; reinterpret as uint32)."""
    assert combined.is_cuda and combined.dim() == 3
    min_vol = min_mask_area * 10
    labels, _ = ops.ccl3d_26(combined.contiguous(), min_vol)
    return labels


def separate_masks(combined_mask: np.ndarray, min_mask_area: int = 100, device="cuda") -> np.ndarray:
    """REF saber/segmenters/utils.py:88-131: 26-connected 3-D components, components smaller than
    ``min_mask_area * 10`` voxels dropped, compact relabel in raster order -> uint32."""
    arr = np.ascontiguousarray(combined_mask)
    if arr.dtype == bool:
        arr = arr.view(np.uint8)
    if arr.dtype.itemsize not in (1, 2, 4):
        arr = (arr != 0).astype(np.uint8)
    if arr.dtype.itemsize == 2:
        t = torch.from_numpy(arr.view(np.int16))
    elif arr.dtype.itemsize == 4:
        t = torch.from_numpy(arr.view(np.int32))
    else:
        t = torch.from_numpy(arr.view(np.uint8))
    labels = separate_masks_device(t.to(device), min_mask_area)
    return labels.cpu().numpy().view(np.uint32)
