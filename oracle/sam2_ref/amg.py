"""ORACLE (test infrastructure only). Restatement of upstream sam2/automatic_mask_generator.py and
sam2/utils/amg.py as configured by REF saber/adapters/sam2/automask.py:66-78 (points_per_side 32,
points_per_batch 64, crop_n_layers 2, downscale 2, use_m2m, multimask; REF saber/adapters/sam2/amg.py:7-17),
including torchvision.ops.batched_nms semantics restated in numpy. SURVEY §8a U5.
"""
from __future__ import annotations

import math
from itertools import product
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .image_predictor import SAM2ImagePredictor


# ----------------------------------------------------------------------------- helpers
def build_point_grid(n_per_side: int) -> np.ndarray:
    offset = 1 / (2 * n_per_side)
    pts = np.linspace(offset, 1 - offset, n_per_side)
    px = np.tile(pts[None, :], (n_per_side, 1))
    py = np.tile(pts[:, None], (1, n_per_side))
    return np.stack([px, py], axis=-1).reshape(-1, 2)


def build_all_layer_point_grids(n_per_side, n_layers, scale_per_layer) -> List[np.ndarray]:
    return [build_point_grid(int(n_per_side / (scale_per_layer ** i))) for i in range(n_layers + 1)]


def generate_crop_boxes(im_size, n_layers, overlap_ratio):
    crop_boxes, layer_idxs = [], []
    im_h, im_w = im_size
    short_side = min(im_h, im_w)
    crop_boxes.append([0, 0, im_w, im_h])
    layer_idxs.append(0)

    def crop_len(orig_len, n_crops, overlap):
        return int(math.ceil((overlap * (n_crops - 1) + orig_len) / n_crops))

    for i_layer in range(n_layers):
        n = 2 ** (i_layer + 1)
        overlap = int(overlap_ratio * short_side * (2 / n))
        crop_w = crop_len(im_w, n, overlap)
        crop_h = crop_len(im_h, n, overlap)
        x0s = [int((crop_w - overlap) * i) for i in range(n)]
        y0s = [int((crop_h - overlap) * i) for i in range(n)]
        for x0, y0 in product(x0s, y0s):
            crop_boxes.append([x0, y0, min(x0 + crop_w, im_w), min(y0 + crop_h, im_h)])
            layer_idxs.append(i_layer + 1)
    return crop_boxes, layer_idxs


def calculate_stability_score(masks: torch.Tensor, mask_threshold: float, threshold_offset: float) -> torch.Tensor:
    inter = (masks > (mask_threshold + threshold_offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    union = (masks > (mask_threshold - threshold_offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return inter / union


def batched_mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    if torch.numel(masks) == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    shape = masks.shape
    h, w = shape[-2:]
    masks = masks.flatten(0, -3) if len(shape) > 2 else masks.unsqueeze(0)
    in_height, _ = torch.max(masks, dim=-1)
    in_height_coords = in_height * torch.arange(h, device=in_height.device)[None, :]
    bottom_edges, _ = torch.max(in_height_coords, dim=-1)
    in_height_coords = in_height_coords + h * (~in_height)
    top_edges, _ = torch.min(in_height_coords, dim=-1)
    in_width, _ = torch.max(masks, dim=-2)
    in_width_coords = in_width * torch.arange(w, device=in_width.device)[None, :]
    right_edges, _ = torch.max(in_width_coords, dim=-1)
    in_width_coords = in_width_coords + w * (~in_width)
    left_edges, _ = torch.min(in_width_coords, dim=-1)
    empty = (right_edges < left_edges) | (bottom_edges < top_edges)
    out = torch.stack([left_edges, top_edges, right_edges, bottom_edges], dim=-1)
    out = out * (~empty).unsqueeze(-1)
    return out.reshape(*shape[:-2], 4) if len(shape) > 2 else out[0]


def uncrop_boxes_xyxy(boxes, crop_box):
    x0, y0, _, _ = crop_box
    offset = torch.tensor([[x0, y0, x0, y0]], device=boxes.device)
    if len(boxes.shape) == 3:
        offset = offset.unsqueeze(1)
    return boxes + offset


def uncrop_points(points, crop_box):
    x0, y0, _, _ = crop_box
    offset = torch.tensor([[x0, y0]], device=points.device)
    if len(points.shape) == 3:
        offset = offset.unsqueeze(1)
    return points + offset


def uncrop_masks(masks, crop_box, orig_h, orig_w):
    x0, y0, x1, y1 = crop_box
    if x0 == 0 and y0 == 0 and x1 == orig_w and y1 == orig_h:
        return masks
    pad_x, pad_y = orig_w - (x1 - x0), orig_h - (y1 - y0)
    return F.pad(masks, (x0, pad_x - x0, y0, pad_y - y0), value=0)


def is_box_near_crop_edge(boxes, crop_box, orig_box, atol=20.0):
    crop_box_t = torch.as_tensor(crop_box, dtype=torch.float, device=boxes.device)
    orig_box_t = torch.as_tensor(orig_box, dtype=torch.float, device=boxes.device)
    boxes = uncrop_boxes_xyxy(boxes, crop_box).float()
    near_crop = torch.isclose(boxes, crop_box_t[None, :], atol=atol, rtol=0)
    near_image = torch.isclose(boxes, orig_box_t[None, :], atol=atol, rtol=0)
    near_crop = torch.logical_and(near_crop, ~near_image)
    return torch.any(near_crop, dim=1)


def box_xyxy_to_xywh(box):
    b = box.clone() if isinstance(box, torch.Tensor) else np.array(box).copy()
    b[2] = b[2] - b[0]
    b[3] = b[3] - b[1]
    return b


def mask_to_rle(mask: np.ndarray) -> Dict[str, Any]:
    """Column-major uncompressed RLE of one bool mask (upstream mask_to_rle_pytorch, per mask)."""
    h, w = mask.shape
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)
    diff = flat[1:] ^ flat[:-1]
    change = np.flatnonzero(diff)
    cur = np.concatenate([[0], change + 1, [h * w]])
    btw = cur[1:] - cur[:-1]
    counts = [] if flat[0] == 0 else [0]
    counts.extend(int(c) for c in btw)
    return {"size": [h, w], "counts": counts}


def rle_to_mask(rle: Dict[str, Any]) -> np.ndarray:
    h, w = rle["size"]
    mask = np.empty(h * w, dtype=bool)
    idx, parity = 0, False
    for c in rle["counts"]:
        mask[idx: idx + c] = parity
        idx += c
        parity ^= True
    return mask.reshape(w, h).T


def area_from_rle(rle) -> int:
    return int(sum(rle["counts"][1::2]))


def nms_numpy(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """torchvision.ops.nms restated: greedy by score desc (stable), suppress iff IoU > thr, fp32 IoU
    = inter / (a + b - inter) with no +1. Returns kept indices in score-descending order."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


class MaskData:
    def __init__(self, **kwargs):
        self._stats = dict(**kwargs)

    def __setitem__(self, k, v):
        self._stats[k] = v

    def __delitem__(self, k):
        del self._stats[k]

    def __getitem__(self, k):
        return self._stats[k]

    def items(self):
        return self._stats.items()

    def filter(self, keep: torch.Tensor):
        for k, v in self._stats.items():
            if v is None:
                self._stats[k] = None
            elif isinstance(v, torch.Tensor):
                self._stats[k] = v[torch.as_tensor(keep, device=v.device)]
            elif isinstance(v, np.ndarray):
                self._stats[k] = v[keep.detach().cpu().numpy()]
            elif isinstance(v, list) and keep.dtype == torch.bool:
                self._stats[k] = [a for i, a in enumerate(v) if keep[i]]
            elif isinstance(v, list):
                self._stats[k] = [v[i] for i in keep]
            else:
                raise TypeError(k)

    def cat(self, new):
        for k, v in new.items():
            if k not in self._stats or self._stats[k] is None:
                self._stats[k] = v.clone() if isinstance(v, torch.Tensor) else list(v) if isinstance(v, list) else v.copy()
            elif isinstance(v, torch.Tensor):
                self._stats[k] = torch.cat([self._stats[k], v], dim=0)
            elif isinstance(v, np.ndarray):
                self._stats[k] = np.concatenate([self._stats[k], v], axis=0)
            elif isinstance(v, list):
                self._stats[k] = self._stats[k] + list(v)
            else:
                raise TypeError(k)

    def to_numpy(self):
        for k, v in self._stats.items():
            if isinstance(v, torch.Tensor):
                self._stats[k] = v.float().detach().cpu().numpy()


def batch_iterator(batch_size, *args):
    n = len(args[0])
    for b in range(n // batch_size + int(n % batch_size != 0)):
        yield [arg[b * batch_size: (b + 1) * batch_size] for arg in args]


class SAM2AutomaticMaskGenerator:
    def __init__(self, model, points_per_side=32, points_per_batch=64, pred_iou_thresh=0.8,
                 stability_score_thresh=0.95, stability_score_offset=1.0, mask_threshold=0.0, box_nms_thresh=0.7,
                 crop_n_layers=0, crop_nms_thresh=0.7, crop_overlap_ratio=512 / 1500,
                 crop_n_points_downscale_factor=1, point_grids=None, min_mask_region_area=0,
                 output_mode="binary_mask", use_m2m=False, multimask_output=True, **kwargs):
        assert min_mask_region_area == 0, "postprocess_small_regions is not on SABER's path"
        self.point_grids = build_all_layer_point_grids(points_per_side, crop_n_layers, crop_n_points_downscale_factor)
        self.predictor = SAM2ImagePredictor(model, max_hole_area=min_mask_region_area,
                                            max_sprinkle_area=min_mask_region_area)
        self.points_per_batch = points_per_batch
        self.pred_iou_thresh = pred_iou_thresh
        self.stability_score_thresh = stability_score_thresh
        self.stability_score_offset = stability_score_offset
        self.mask_threshold = mask_threshold
        self.box_nms_thresh = box_nms_thresh
        self.crop_n_layers = crop_n_layers
        self.crop_nms_thresh = crop_nms_thresh
        self.crop_overlap_ratio = crop_overlap_ratio
        self.crop_n_points_downscale_factor = crop_n_points_downscale_factor
        self.output_mode = output_mode
        self.use_m2m = use_m2m
        self.multimask_output = multimask_output

    @torch.no_grad()
    def generate(self, image: np.ndarray) -> List[Dict[str, Any]]:
        data = self._generate_masks(image)
        anns = []
        for idx in range(len(data["rles"])):
            rle = data["rles"][idx]
            anns.append({
                "segmentation": rle_to_mask(rle),
                "area": area_from_rle(rle),
                "bbox": box_xyxy_to_xywh(data["boxes"][idx]).tolist(),
                "predicted_iou": data["iou_preds"][idx].item(),
                "point_coords": [data["points"][idx].tolist()],
                "stability_score": data["stability_score"][idx].item(),
                "crop_box": box_xyxy_to_xywh(data["crop_boxes"][idx]).tolist(),
            })
        return anns

    def _generate_masks(self, image):
        orig_size = image.shape[:2]
        crop_boxes, layer_idxs = generate_crop_boxes(orig_size, self.crop_n_layers, self.crop_overlap_ratio)
        data = MaskData()
        for crop_box, layer_idx in zip(crop_boxes, layer_idxs):
            data.cat(self._process_crop(image, crop_box, layer_idx, orig_size))
        if len(crop_boxes) > 1 and len(data["rles"]) > 0:
            cb = data["crop_boxes"].float()
            scores = 1 / ((cb[:, 2] - cb[:, 0]) * (cb[:, 3] - cb[:, 1]))
            keep = nms_numpy(data["boxes"].float().numpy(), scores.numpy(), self.crop_nms_thresh)
            data.filter(torch.as_tensor(keep))
        data.to_numpy()
        return data

    def _process_crop(self, image, crop_box, crop_layer_idx, orig_size):
        x0, y0, x1, y1 = crop_box
        cropped_im = image[y0:y1, x0:x1, :]
        cropped_im_size = cropped_im.shape[:2]
        self.predictor.set_image(cropped_im)
        points_scale = np.array(cropped_im_size)[None, ::-1]
        points_for_image = self.point_grids[crop_layer_idx] * points_scale
        data = MaskData()
        for (points,) in batch_iterator(self.points_per_batch, points_for_image):
            data.cat(self._process_batch(points, cropped_im_size, crop_box, orig_size, normalize=True))
        self.predictor.reset_predictor()
        keep = nms_numpy(data["boxes"].float().numpy(), data["iou_preds"].numpy(), self.box_nms_thresh)
        data.filter(torch.as_tensor(keep))
        data["boxes"] = uncrop_boxes_xyxy(data["boxes"], crop_box)
        data["points"] = uncrop_points(data["points"], crop_box)
        data["crop_boxes"] = torch.tensor([crop_box for _ in range(len(data["rles"]))]).reshape(-1, 4)
        return data

    def _process_batch(self, points, im_size, crop_box, orig_size, normalize=False):
        orig_h, orig_w = orig_size
        dev = self.predictor.device
        points = torch.as_tensor(points, dtype=torch.float32, device=dev)
        in_points = self.predictor._transforms.transform_coords(points, normalize=normalize, orig_hw=im_size)
        in_labels = torch.ones(in_points.shape[0], dtype=torch.int, device=dev)
        masks, iou_preds, low_res_masks = self.predictor._predict(
            in_points[:, None, :], in_labels[:, None], multimask_output=self.multimask_output, return_logits=True)
        data = MaskData(masks=masks.flatten(0, 1), iou_preds=iou_preds.flatten(0, 1),
                        points=points.repeat_interleave(masks.shape[1], dim=0),
                        low_res_masks=low_res_masks.flatten(0, 1))
        del masks
        if self.use_m2m:
            in_points = self.predictor._transforms.transform_coords(data["points"], normalize=normalize, orig_hw=im_size)
            labels = torch.ones(in_points.shape[0], dtype=torch.int, device=dev)
            masks, ious = self.refine_with_m2m(in_points, labels, data["low_res_masks"], self.points_per_batch)
            data["masks"] = masks.squeeze(1)
            data["iou_preds"] = ious.squeeze(1)
        if self.pred_iou_thresh > 0.0:
            data.filter(data["iou_preds"] > self.pred_iou_thresh)
        data["stability_score"] = calculate_stability_score(data["masks"], self.mask_threshold, self.stability_score_offset)
        if self.stability_score_thresh > 0.0:
            data.filter(data["stability_score"] >= self.stability_score_thresh)
        data["masks"] = data["masks"] > self.mask_threshold
        data["boxes"] = batched_mask_to_box(data["masks"])
        keep = ~is_box_near_crop_edge(data["boxes"], crop_box, [0, 0, orig_w, orig_h])
        if not torch.all(keep):
            data.filter(keep)
        data["masks"] = uncrop_masks(data["masks"], crop_box, orig_h, orig_w)
        data["rles"] = [mask_to_rle(m.cpu().numpy()) for m in data["masks"]]
        del data["masks"]
        return data

    def refine_with_m2m(self, points, point_labels, low_res_masks, points_per_batch):
        new_masks, new_ious = [], []
        for cur_points, cur_labels, low_res_mask in batch_iterator(points_per_batch, points, point_labels, low_res_masks):
            best_masks, best_ious, _ = self.predictor._predict(
                cur_points[:, None, :], cur_labels[:, None], mask_input=low_res_mask[:, None, :],
                multimask_output=False, return_logits=True)
            new_masks.append(best_masks)
            new_ious.append(best_ious)
        return torch.cat(new_masks, dim=0), torch.cat(new_ious, dim=0)
