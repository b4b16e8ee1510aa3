"""Twin of REF saber/filters/masks.py:9-183 (R13): turning classifier probabilities into mask lists — class arg-max,
confidence-weighted consensus map + 2-D connected components, semantic merge, label stacks. These run once per slice
on a few dozen masks; they are host (numpy / scipy) logic exactly as in the reference, fed by the device classifier
(saber_b200.classifier.Predictor.batch_predict)."""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def apply_classifier(image, masks, classifier, desired_class: int = None, min_mask_area: int = 100, batchsize: int = 32):
    """REF :9-21. NB the reference's caller passes (target_class, batchsize) positionally, so `min_mask_area` receives
    the batch size 32 there; the twin of that caller reproduces it."""
    sam2_masks = np.array([m["segmentation"].astype(np.uint8) for m in masks])
    predictions = classifier.batch_predict(image, sam2_masks, batchsize)
    return convert_predictions_to_masks(predictions, masks, desired_class, min_mask_area)


def convert_predictions_to_masks(predictions, masks, desired_class: int = None, min_mask_area: int = 100):
    if isinstance(masks, np.ndarray):
        masks = masks_to_list(masks)
    predicted_classes = np.argmax(predictions, axis=1)
    if desired_class > 0 and desired_class is not None:
        confidence_scores = predictions[:, desired_class]
        target = [i for i, p in enumerate(predicted_classes) if p == desired_class]
        masks = [masks[i] for i in target]
        confidence_scores = confidence_scores[target]
        if len(masks) > 0:
            masks = _consensus_based_resolution(masks[0]["segmentation"].shape, masks, confidence_scores)
            masks = [m for m in masks if m["area"] >= min_mask_area]
            masks = sorted(masks, key=lambda x: x["area"], reverse=False)
        return masks
    if len(masks) == 0:
        return np.array([])
    return _semantic_segmentation(masks, predictions)


def _consensus_based_resolution(image_shape, masks, confidences):
    h, w = image_shape
    confidence_map = np.zeros((h, w), dtype=np.float32)
    overlap_count = np.zeros((h, w), dtype=np.int32)
    for md, conf in zip(masks, confidences):
        seg = md["segmentation"]
        confidence_map += seg * conf
        overlap_count += seg
    with np.errstate(divide="ignore", invalid="ignore"):
        avg_confidence = np.nan_to_num(np.divide(confidence_map, overlap_count))
    labeled, n = ndimage.label(overlap_count > 0)
    out = []
    for label in range(1, n + 1):
        comp = labeled == label
        conf = np.mean(avg_confidence[comp])
        ys, xs = np.where(comp)
        y0, y1, x0, x1 = np.min(ys), np.max(ys), np.min(xs), np.max(xs)
        out.append({"segmentation": comp, "area": int(np.sum(comp)),
                    "bbox": [int(x0), int(y0), int(x1 - x0), int(y1 - y0)], "predicted_iou": float(conf),
                    "point_coords": [[int((x0 + x1) / 2), int((y0 + y1) / 2)]], "stability_score": float(conf),
                    "crop_box": [int(x0), int(y0), int(x1), int(y1)]})
    return out


def _semantic_segmentation(masks, predictions):
    predicted_classes = np.argmax(predictions, axis=1)
    max_class = predictions.shape[1]
    out = [{"segmentation": np.zeros(masks[0]["segmentation"].shape, dtype=np.uint8), "area": 0, "label": ii}
           for ii in range(1, max_class)]
    for ii in range(len(masks)):
        c = predicted_classes[ii]
        if c > 0:
            out[c - 1]["segmentation"] = np.logical_or(out[c - 1]["segmentation"], masks[ii]["segmentation"]).astype(bool)
            out[c - 1]["area"] += masks[ii]["area"]
    return out


def masks_to_array(mask_list):
    if not isinstance(mask_list, list):
        return None
    nx, ny = mask_list[0]["segmentation"].shape
    dtype = np.uint8 if len(mask_list) < 256 else (np.uint16 if len(mask_list) < 65536 else np.uint32)
    masks = np.zeros([len(mask_list), nx, ny], dtype=dtype)
    for j, m in enumerate(mask_list):
        masks[j] = m["segmentation"].astype(dtype) * (j + 1)
    return masks


def masks_to_list(masks):
    if isinstance(masks, list):
        return masks
    return [{"segmentation": masks == val, "area": np.sum((masks == val) > 0)} for val in np.unique(masks)]


# ---- REF saber/filters/masks.py:230-309 + gaussian.py:76-138 (R14) on the device ------------------------------------
def _estimate_feature_size_3d(volume_count: int, scale: float = 0.075) -> float:
    return scale * 2 * ((3 * volume_count) / (4 * np.pi)) ** (1 / 3)


def fast_3d_gaussian_smoothing(volume, scale: float = 0.075, deviceID=None):
    """Per label: sigma from the label's voxel count, separable zero-padded Gaussian (x, then y, then z — the axis order
    the reference's conv3d kernels actually have), `> 0.5`, write the label (later labels overwrite) -> uint8 volume.
    numpy in -> numpy out; CUDA tensor in -> CUDA tensor out."""
    import torch
    from .. import ops
    is_numpy = isinstance(volume, np.ndarray)
    dev = f"cuda:{deviceID}" if isinstance(deviceID, int) else (deviceID or "cuda")
    vol = torch.from_numpy(np.ascontiguousarray(volume)).to(dev) if is_numpy else volume.contiguous()
    if vol.dim() != 3:
        raise ValueError(f"Expected 3D input, got {vol.dim()}D")
    if vol.dtype in (torch.int16, torch.int32, torch.uint8, torch.uint16, torch.uint32):
        pass
    else:
        raise TypeError(f"label volume dtype {vol.dtype} not supported")
    labels = [int(v) for v in torch.unique(vol).cpu().tolist() if int(v) != 0]
    result = torch.zeros(vol.shape, dtype=torch.uint8, device=vol.device)
    for label in labels:
        m, count = ops.label_equals(vol, label & 0xFFFFFFFF)
        sigma = _estimate_feature_size_3d(int(count.item()), scale)
        ks = int(2 * 3 * sigma + 1)
        ks = ks + 1 if ks % 2 == 0 else ks
        k = torch.exp(-torch.arange(-(ks // 2), ks // 2 + 1, dtype=torch.float32) ** 2 / (2 * sigma ** 2))
        k = (k / k.sum()).to(vol.device).contiguous()
        for axis in (2, 1, 0):
            m = ops.corr1d_zero(m, k, axis)
        ops.threshold_label_(m, 0.5, label, result)
    return result.cpu().numpy() if is_numpy else result
