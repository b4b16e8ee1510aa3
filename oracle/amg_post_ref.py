"""ORACLE (test infrastructure only). CPU restatement of the integer / indexing stages of upstream
``SAM2AutomaticMaskGenerator._process_batch`` / ``_process_crop`` / ``_generate_masks`` that follow the mask
decoder (sam2/automatic_mask_generator.py, sam2/utils/amg.py, SAM2Transforms.postprocess_masks,
torchvision.ops.nms) — the twin of saber_b200/csrc/amg_post.cu. Call sites in the reference:
REF saber/adapters/sam2/automask.py:66-78 (thresholds) and REF saber/adapters/sam2/amg.py:163.

``upsample_bilinear`` evaluates, in numpy fp32 with float64-emulated fused multiply-adds, the expression tree
of torch's CPU ``F.interpolate(mode="bilinear", align_corners=False)``; tests/test_oracle_pins.py pins it
bitwise against torch itself, and pins ``nms`` against ``torchvision.ops.nms``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

f32 = np.float32


def _fma(a, b, c):
    """fp32 fused multiply-add (single rounding): exact in float64 for fp32 operands, then rounded once."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _src_index(in_size: int, out_size: int):
    scale = f32(in_size) / f32(out_size)
    d = np.arange(out_size).astype(np.float32) + f32(0.5)
    s = _fma(np.full_like(d, scale), d, np.full_like(d, -0.5))
    s = np.maximum(s, f32(0))
    i0 = np.minimum(np.floor(s).astype(np.int64), in_size - 1)
    l1 = np.clip(s - i0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    i1 = i0 + (i0 < in_size - 1)
    l0 = (f32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def upsample_bilinear(planes: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """[..., Hi, Wi] fp32 -> [..., Ho, Wo] fp32."""
    planes = np.asarray(planes, np.float32)
    Hi, Wi = planes.shape[-2:]
    Ho, Wo = out_hw
    y0, y1, ly0, ly1 = _src_index(Hi, Ho)
    x0, x1, lx0, lx1 = _src_index(Wi, Wo)
    r0, r1 = planes[..., y0, :], planes[..., y1, :]
    v00, v01, v10, v11 = r0[..., x0], r0[..., x1], r1[..., x0], r1[..., x1]
    t0 = _fma(v00, lx0, v01 * lx1)
    t1 = _fma(v10, lx0, v11 * lx1)
    return _fma(t0, ly0[:, None], t1 * ly1[:, None])


def mask_post(planes: np.ndarray, ious: np.ndarray, crop_box, frame_hw, pred_iou_thresh: float, mask_thresh: float,
              stab_offset: float, stab_thresh: float, edge_atol: float = 20.0,
              only_iou_survivors: bool = False) -> Dict[str, np.ndarray]:
    """planes [n,S,S] fp32 low-res logits, ious [n]. Returns per-candidate keep / stability / bbox (full-frame
    xyxy, inclusive max) / area and bool masks in the full frame — the records of the kept candidates are what
    upstream's MaskData holds after _process_batch (filters: iou > thr, stability >= thr, not near crop edge).
    only_iou_survivors: evaluate only the candidates that pass the predicted-IoU filter, as upstream does (it filters
    by IoU BEFORE computing stability scores and boxes); the other rows come back with keep = False and zeros — what
    makes the checker affordable at SABER's default AMG size (9 216 candidates per image)."""
    if only_iou_survivors and pred_iou_thresh > 0.0:
        n = planes.shape[0]
        H, W = frame_hw
        sel = np.flatnonzero(np.asarray(ious, np.float32) > f32(pred_iou_thresh))
        out = {"keep": np.zeros(n, bool), "stability": np.zeros(n, np.float32), "bbox": np.zeros((n, 4), np.int32),
               "area": np.zeros(n, np.int32), "masks": None, "rows": sel}
        if sel.size:
            r = mask_post(planes[sel], np.asarray(ious)[sel], crop_box, frame_hw, pred_iou_thresh, mask_thresh, stab_offset,
                          stab_thresh, edge_atol)
            for k in ("keep", "stability", "bbox", "area"):
                out[k][sel] = r[k]
            out["masks_rows"] = r["masks"]  # row j belongs to candidate sel[j]
        return out
    x0, y0, x1, y1 = crop_box
    H, W = frame_hw
    n = planes.shape[0]
    up = upsample_bilinear(planes, (y1 - y0, x1 - x0))
    keep = np.ones(n, bool)
    if pred_iou_thresh > 0.0:
        keep &= ious > f32(pred_iou_thresh)
    hi, lo = f32(mask_thresh) + f32(stab_offset), f32(mask_thresh) - f32(stab_offset)
    inter = (up > hi).reshape(n, -1).sum(1).astype(np.int32)
    union = (up > lo).reshape(n, -1).sum(1).astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        stab = inter.astype(np.float32) / union.astype(np.float32)
    if stab_thresh > 0.0:
        keep &= stab >= f32(stab_thresh)
    binm = up > f32(mask_thresh)
    bbox = np.zeros((n, 4), np.int32)
    area = binm.reshape(n, -1).sum(1).astype(np.int32)
    for i in range(n):
        ys, xs = np.nonzero(binm[i])
        if ys.size:
            bbox[i] = (xs.min() + x0, ys.min() + y0, xs.max() + x0, ys.max() + y0)
        else:
            bbox[i] = (x0, y0, x0, y0)  # batched_mask_to_box -> [0,0,0,0] in the crop frame, then uncrop_boxes_xyxy
    fb = bbox.astype(np.float32)
    cb = np.array([x0, y0, x1, y1], np.float32)[None]
    ob = np.array([0, 0, W, H], np.float32)[None]
    near_crop = np.abs(fb - cb) <= f32(edge_atol)
    near_img = np.abs(fb - ob) <= f32(edge_atol)
    keep &= ~np.any(near_crop & ~near_img, axis=1)
    full = np.zeros((n, H, W), bool)
    full[:, y0:y1, x0:x1] = binm
    return {"keep": keep, "stability": stab, "bbox": bbox, "area": area, "masks": full}


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """torchvision.ops.nms: stable sort by score descending; box j is suppressed by a kept box i when
    IoU(i, j) > thr, IoU = inter / (area_i + area_j - inter) in fp32. Returns kept indices in score order."""
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores, kind="stable")
    areas = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    suppressed = np.zeros(n, bool)
    keep = []
    thr = f32(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(boxes[i, 0], boxes[rest, 0])
        yy1 = np.maximum(boxes[i, 1], boxes[rest, 1])
        xx2 = np.minimum(boxes[i, 2], boxes[rest, 2])
        yy2 = np.minimum(boxes[i, 3], boxes[rest, 3])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)
