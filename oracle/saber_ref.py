"""ORACLE (test infrastructure only — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg; never by the product path).

CPU restatement (numpy + scipy) of the SABER-owned functions on the slice-wise hot path. Each function cites
the reference lines it follows. These restatements are pinned against outputs of the *reference's own code*
(`saber.utils.preprocessing`, `saber.segmenters.utils` import cleanly from /root/reference in the build
container): see oracle/make_golden.py and tests/golden/saber_*.npz, checked by tests/test_oracle_pins.py.
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np
from scipy import ndimage as ndi
from scipy.ndimage import uniform_filter


# ---- REF saber/utils/preprocessing.py:4-18 ----------------------------------------------------
def contrast(image: np.ndarray, std_cutoff: int = 5) -> np.ndarray:
    image_mean = uniform_filter(image, size=500)
    image_sq = uniform_filter(image ** 2, size=500)
    image_var = np.clip(image_sq - image_mean ** 2, a_min=0, a_max=None)
    image_std = np.sqrt(image_var)
    image = (image - image_mean) / (image_std + 1e-8)
    return np.clip(image, -std_cutoff, std_cutoff)


# ---- REF saber/utils/preprocessing.py:20-37 ---------------------------------------------------
def normalize(image: np.ndarray, rgb: bool = False) -> np.ndarray:
    if rgb:
        mn = image.min(axis=(0, 1), keepdims=True)
        mx = image.max(axis=(0, 1), keepdims=True)
    else:
        mn, mx = image.min(), image.max()
    return (image - mn) / (mx - mn + 1e-8)


# ---- REF saber/utils/preprocessing.py:67-81 ---------------------------------------------------
def prepare(image: np.ndarray, to_rgb: bool = False) -> np.ndarray:
    image = contrast(image, std_cutoff=3)
    image = normalize(image, rgb=False)
    if to_rgb and image.ndim == 2:
        image = np.repeat(image[..., None], 3, axis=2).astype(np.float32)
    return image


# ---- REF saber/segmenters/utils.py:5-86 -------------------------------------------------------
def remove_duplicate_masks(masks: List[Dict[str, Any]], iou_threshold: float = 0.9,
                           area_threshold: float = 0.9) -> List[Dict[str, Any]]:
    def iou_of(a, b):
        inter = np.logical_and(a, b).sum()
        union = np.logical_or(a, b).sum()
        return 0.0 if union == 0 else inter / union

    def duplicate(m1, m2):
        a1, a2 = m1["area"], m2["area"]
        ratio = min(a1, a2) / max(a1, a2) if max(a1, a2) > 0 else 0
        iou = iou_of(m1["segmentation"], m2["segmentation"])
        return not (ratio < area_threshold or iou < iou_threshold)

    unique, done = [], set()
    for i, m1 in enumerate(masks):
        if i in done:
            continue
        group = [(i, m1)]
        for j in range(i + 1, len(masks)):
            if j in done:
                continue
            if duplicate(m1, masks[j]):
                group.append((j, masks[j]))
                done.add(j)
        if len(group) > 1:
            unique.append(max(group, key=lambda x: x[1].get("stability_score", 0))[1])
        else:
            unique.append(m1)
        done.add(i)
    return unique


# ---- REF saber/segmenters/utils.py:88-131 -----------------------------------------------------
def separate_masks(combined_mask: np.ndarray, min_mask_area: int = 100) -> np.ndarray:
    m = np.ascontiguousarray(combined_mask.astype(bool))
    if not m.any():
        return np.zeros_like(m, dtype=np.uint32)
    z, y, x = np.where(m)
    z0, z1, y0, y1, x0, x1 = z.min(), z.max() + 1, y.min(), y.max() + 1, x.min(), x.max() + 1
    sub = m[z0:z1, y0:y1, x0:x1]
    labels_sub, _ = ndi.label(sub, structure=ndi.generate_binary_structure(rank=3, connectivity=3))
    min_vol = min_mask_area * 10
    if min_vol > 1:
        counts = np.bincount(labels_sub.ravel())
        small = np.flatnonzero((counts < min_vol) & (np.arange(counts.size) != 0))
        if small.size:
            labels_sub[np.isin(labels_sub, small)] = 0
    counts = np.bincount(labels_sub.ravel())
    keep = counts > 0
    keep[0] = False
    new_ids = np.cumsum(keep).astype(np.uint32)
    remap = np.zeros_like(new_ids, dtype=np.uint32)
    remap[keep] = new_ids[keep]
    labeled = np.zeros_like(m, dtype=np.uint32)
    labeled[z0:z1, y0:y1, x0:x1] = remap[labels_sub]
    return labeled


# ---- REF saber/segmenters/base.py:159-176 (classifier is None branch) -------------------------
def apply_classifier_none(masks: List[Dict[str, Any]], min_mask_area: int) -> List[Dict[str, Any]]:
    masks = [m for m in masks if m["area"] >= min_mask_area]
    masks = remove_duplicate_masks(masks)
    return sorted(masks, key=lambda m: m["area"], reverse=False)


# ---- REF saber/segmenters/propagation.py:177-187 (one iteration of slice_by_slice) ------------
def stitch_slice(mask_list: List[np.ndarray], shape) -> np.ndarray:
    masks3d = np.zeros(shape, dtype=np.uint16)
    for idx, mask in enumerate(mask_list):
        masks3d[mask] = idx + 1
    return masks3d
