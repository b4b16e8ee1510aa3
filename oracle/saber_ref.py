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


# =================================================================================================
# 3-D propagation path and the neighbouring filters (R13, R14, R17)
# =================================================================================================

# ---- REF saber/utils/preprocessing.py:39-66 ---------------------------------------------------
def project_tomogram(vol: np.ndarray, zSlice=None, deltaZ=None) -> np.ndarray:
    if zSlice is not None:
        if deltaZ is not None:
            z0 = int(max(zSlice - deltaZ, 0))
            z1 = int(min(zSlice + deltaZ, vol.shape[0]))
            return np.mean(vol[z0:z1, ], axis=0)
        return vol[zSlice, ]
    return np.mean(vol, axis=0)


# ---- REF saber/adapters/preprocessing.py:72-76 (R1) -------------------------------------------
def normalize_tomogram(tomogram: np.ndarray) -> np.ndarray:
    tomogram = (tomogram - tomogram.min()) / (tomogram.max() - tomogram.min())
    return tomogram * 2 - 1


# ---- skimage.transform.resize restated on scipy (skimage is not installed) -
def skimage_resize(image: np.ndarray, output_shape, order=None, anti_aliasing=None) -> np.ndarray:
    """skimage.transform.resize(image, output_shape, order=…, mode='reflect', anti_aliasing=…), 2-D, following
    skimage/transform/_warps.py: bool input -> order 0 and no anti-aliasing; anti-aliasing = Gaussian with
    sigma = max(0, (in/out - 1)/2), ndimage mode 'mirror'; zoom with grid_mode=True, mode 'mirror'; clip to the input
    range. Called at REF saber/adapters/preprocessing.py:21 (anti_aliasing=True) and sam2/predictor.py:296 (order=0)."""
    image = np.asarray(image)
    in_shape = image.shape
    out_shape = tuple(int(s) for s in output_shape)
    if image.dtype == bool:
        order = 0 if order is None else order
        anti_aliasing = False if anti_aliasing is None else anti_aliasing
        work = image
    else:
        order = 1 if order is None else order
        if image.dtype == np.float16:
            work = image.astype(np.float32)
        elif image.dtype in (np.float32, np.float64):
            work = image
        else:
            work = image.astype(np.float64)
        if anti_aliasing is None:
            anti_aliasing = any(o < i for o, i in zip(out_shape, in_shape))
    factors = np.divide(in_shape, out_shape)
    filtered = work
    if anti_aliasing and work.dtype != bool:
        sigma = np.maximum(0, (factors - 1) / 2)
        if np.any(sigma > 0):
            filtered = ndi.gaussian_filter(work, sigma, cval=0, mode="mirror")
    zoom = [1 / f for f in factors]
    if work.dtype == bool:
        out = ndi.zoom(filtered.astype(np.uint8), zoom, order=0, mode="mirror", cval=0, grid_mode=True).astype(bool)
        return out
    out = ndi.zoom(filtered, zoom, order=order, mode="mirror", cval=0, grid_mode=True)
    if order > 0:  # _clip_warp_output
        out = np.clip(out, work.min(), work.max())
    return out


# ---- REF saber/adapters/preprocessing.py:16-70 (R2; light_modality False) ----------------------
def load_grayscale_image_array(img_array: np.ndarray, image_size: int):
    """-> (images float32 [Z,3,S,S] = 2*resize(slice)-1 repeated over 3 channels, video_height, video_width).
    NB video_height/video_width are read from the *resized* image (REF :24), i.e. both equal image_size."""
    Z = img_array.shape[0]
    images = np.zeros((Z, 3, image_size, image_size), dtype=np.float32)
    for n in range(Z):
        img = skimage_resize(img_array[n], (image_size, image_size), anti_aliasing=True)
        images[n] = np.repeat(img[None, ...], axis=0, repeats=3).astype(np.float32)
    images = 2 * images - 1
    return images, image_size, image_size


# ---- REF saber/filters/gaussian.py:7-74 (R3) --------------------------------------------------
def make_gaussian_kernel(sigma: float) -> np.ndarray:
    import torch
    ks = round(sigma * 3)
    ks = max(ks, 3)
    ks += 1 - ks % 2
    ts = torch.linspace(-ks / 2, ks / 2, ks)
    gauss = torch.exp(-(ts / sigma) ** 2 / 2)
    return (gauss / gauss.sum()).numpy()


def gaussian_smoothing(vol: np.ndarray, sigma: float, dim: int = 0) -> np.ndarray:
    """1-D cross-correlation with the kernel above along `dim`, zero padding (F.conv1d(padding=ks//2)), float32."""
    k = make_gaussian_kernel(sigma).astype(np.float32)
    x = np.moveaxis(np.asarray(vol, dtype=np.float32), dim, -1)
    pad = k.size // 2
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)])
    out = np.zeros_like(x)
    for t in range(k.size):
        out += k[t] * xp[..., t:t + x.shape[-1]]
    return np.moveaxis(out, -1, dim)


# ---- REF saber/filters/estimate_thickness.py:7-112 (R8) ---------------------------------------
def _quadratic(x, a, b, c, d):
    return d * np.maximum(a * (x - b) ** 2 + c, 0)


def _gaussian(x, a, b, c):
    with np.errstate(over="ignore"):
        return a * np.exp(-(x - b) ** 2 / (2 * c ** 2))


def _r2(data, func, params):
    x = np.arange(len(data))
    y = func(x, *params)
    ss_res = np.sum((data - y) ** 2)
    ss_tot = np.sum((data - np.mean(data)) ** 2)
    return 0 if ss_tot == 0 else 1 - ss_res / ss_tot


def fit_organelle_boundaries(frame_scores: np.ndarray) -> np.ndarray:
    from scipy.optimize import curve_fit
    nF, nM = frame_scores.shape
    out = np.zeros((nF, nM))
    for ii in range(nM):
        data = frame_scores[:, ii].copy()
        data = np.maximum(data, 0)
        data -= np.mean(data[-15:-5])
        data = np.maximum(data, 0)
        x = np.arange(len(data), dtype=np.float32)
        try:
            x_max = np.argmax(data[1:-1])
            p1, _ = curve_fit(_quadratic, x, data, p0=[-1e-3, x_max, 1, np.max(data) / 2],
                              bounds=([-np.inf, 0, 0, 0], [0, nF, 10, 10]))
            r2q = _r2(data, _quadratic, p1)
        except Exception:
            r2q = 0
        try:
            x_max = np.argmax(data[1:-1])
            p2, _ = curve_fit(_gaussian, x, data, p0=[np.max(data), x_max, 3e-1],
                              bounds=((0, 0, 0), (np.inf, nF, nF * 0.25 / 2.355)))
            r2g = _r2(data, _gaussian, p2)
        except Exception:
            r2g = 0
        if r2q == 0 and r2g == 0:
            out[:, ii] = 0
        elif r2q > r2g:
            out[:, ii] = _quadratic(x, *p1)
        else:
            out[:, ii] = _gaussian(x, *p2)
    return out


# ---- REF saber/adapters/sam2/predictor.py:232-348 (R7) ----------------------------------------
def normalize_masks(masks) -> list:
    """REF :208-230 for the list-of-arrays / list-of-dicts forms SABER's segmenters pass."""
    if masks is None:
        return []
    if isinstance(masks, np.ndarray) and masks.ndim >= 3:
        return [np.squeeze(masks[i]).astype(np.float32) for i in range(masks.shape[0])]
    out = []
    for m in masks:
        if isinstance(m, dict):
            m = m["segmentation"]
        out.append(np.squeeze(np.asarray(m)).astype(np.float32))
    return out


def segment_volume(predictor, state, start_frame_idx: int, masks, vol_shape, max_frame_num_to_track=None,
                   min_presence_score: float = 0.5):
    """SAM2Adapter.segment_volume on an (oracle or B200) video predictor exposing add_new_mask / propagate_in_video /
    sam_mask_decoder.register_forward_hook. Returns (vol_masks uint16 [Z,H,W], frame_scores [Z,nMasks],
    frame_metrics). Keeps the reference's hook timing: `_current_frame` is updated *after* a frame is yielded, so the
    scores produced while computing frame f are filed under the previously yielded frame."""
    Z, H, W = vol_shape
    mask_list = normalize_masks(masks)
    for obj_id, mask in enumerate(mask_list, start=1):
        if np.max(mask) == 0:
            continue
        predictor.add_new_mask(inference_state=state, frame_idx=start_frame_idx, obj_id=obj_id, mask=mask)
    current = {"frame": None}
    captured = {}

    def _hook(module, inputs, output):
        logits = output[3].detach().cpu().float().numpy()
        captured.setdefault(current["frame"], []).append(logits)

    handle = predictor.sam_mask_decoder.register_forward_hook(_hook)
    vol_masks = np.zeros((Z, H, W), dtype=np.uint16)

    def _apply(frame_idx, obj_ids, mask_logits):
        for i, obj_id in enumerate(obj_ids):
            m = (mask_logits[i] > 0.0).cpu().numpy()
            m = np.squeeze(m).astype(bool)
            if m.shape != (H, W):
                m = skimage_resize(m, (H, W), order=0, anti_aliasing=False)
            vol_masks[frame_idx] = np.where(m, int(obj_id), vol_masks[frame_idx])

    for frame_idx, obj_ids, logits in predictor.propagate_in_video(
            state, start_frame_idx=start_frame_idx, max_frame_num_to_track=max_frame_num_to_track, reverse=False):
        current["frame"] = frame_idx
        _apply(frame_idx, obj_ids, logits)
    for frame_idx, obj_ids, logits in predictor.propagate_in_video(
            state, start_frame_idx=start_frame_idx, max_frame_num_to_track=max_frame_num_to_track, reverse=True):
        current["frame"] = frame_idx
        if not vol_masks[frame_idx].any():
            _apply(frame_idx, obj_ids, logits)
    handle.remove()
    nM = len(mask_list)
    frame_scores = np.zeros([Z, nM])
    metrics = {}
    if nM > 0:
        for fidx, scores in captured.items():
            if fidx is None:
                continue
            vals = np.concatenate([s.flatten() for s in scores])
            n = min(len(vals), nM)
            frame_scores[fidx, :n] = vals[:n]
        bounds = fit_organelle_boundaries(frame_scores)
        for fidx in range(Z):
            metrics[fidx] = {}
            for mi in range(nM):
                ps = float(bounds[fidx, mi])
                metrics[fidx][mi + 1] = {"presence_score": ps}
                if ps < min_presence_score:
                    vol_masks[fidx][vol_masks[fidx] == mi + 1] = 0
    return vol_masks.astype(np.uint16), frame_scores, metrics


# ---- REF saber/filters/masks.py:61-121 (R13) --------------------------------------------------
def consensus_based_resolution(image_shape, masks, confidences):
    h, w = image_shape
    cmap = np.zeros((h, w), dtype=np.float32)
    cnt = np.zeros((h, w), dtype=np.int32)
    for md, conf in zip(masks, confidences):
        seg = md["segmentation"]
        cmap += seg * conf
        cnt += seg
    with np.errstate(divide="ignore", invalid="ignore"):
        avg = np.nan_to_num(np.divide(cmap, cnt))
    lab, n = ndi.label(cnt > 0)
    out = []
    for lb in range(1, n + 1):
        comp = lab == lb
        conf = np.mean(avg[comp])
        ys, xs = np.where(comp)
        y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
        out.append({"segmentation": comp, "area": int(comp.sum()),
                    "bbox": [int(x0), int(y0), int(x1 - x0), int(y1 - y0)], "predicted_iou": float(conf),
                    "point_coords": [[int((x0 + x1) / 2), int((y0 + y1) / 2)]], "stability_score": float(conf),
                    "crop_box": [int(x0), int(y0), int(x1), int(y1)]})
    return out


def convert_predictions_to_masks(predictions, masks, desired_class=None, min_mask_area=100):
    """REF saber/filters/masks.py:23-59 (list-of-dict input)."""
    pred_cls = np.argmax(predictions, axis=1)
    if desired_class > 0 and desired_class is not None:
        conf = predictions[:, desired_class]
        idx = [i for i, p in enumerate(pred_cls) if p == desired_class]
        masks = [masks[i] for i in idx]
        conf = conf[idx]
        if len(masks) > 0:
            masks = consensus_based_resolution(masks[0]["segmentation"].shape, masks, conf)
            masks = [m for m in masks if m["area"] >= min_mask_area]
            masks = sorted(masks, key=lambda m: m["area"], reverse=False)
        return masks
    if len(masks) == 0:
        return np.array([])
    n_cls = predictions.shape[1]
    out = [{"segmentation": np.zeros(masks[0]["segmentation"].shape, dtype=np.uint8), "area": 0, "label": ii}
           for ii in range(1, n_cls)]
    for ii in range(len(masks)):
        c = pred_cls[ii]
        if c > 0:
            out[c - 1]["segmentation"] = np.logical_or(out[c - 1]["segmentation"], masks[ii]["segmentation"]).astype(bool)
            out[c - 1]["area"] += masks[ii]["area"]
    return out


def masks_to_array(mask_list):
    """REF saber/filters/masks.py:157-183."""
    nx, ny = mask_list[0]["segmentation"].shape
    dtype = np.uint8 if len(mask_list) < 256 else (np.uint16 if len(mask_list) < 65536 else np.uint32)
    out = np.zeros([len(mask_list), nx, ny], dtype=dtype)
    for j, m in enumerate(mask_list):
        out[j] = m["segmentation"].astype(dtype) * (j + 1)
    return out


# ---- REF saber/filters/masks.py:230-309 + gaussian.py:76-138 (R14) -----------------------------
def estimate_feature_size_3d(binary_volume: np.ndarray, scale: float = 0.075) -> float:
    volume = np.sum(binary_volume)
    return scale * 2 * ((3 * volume) / (4 * np.pi)) ** (1 / 3)


def gaussian_smoothing_3d(volume: np.ndarray, sigma: float) -> np.ndarray:
    """Separable zero-padded Gaussian; pass order as the reference labels them (last axis, middle axis, first axis)."""
    import torch
    ks = int(2 * 3 * sigma + 1)
    ks = ks + 1 if ks % 2 == 0 else ks
    k = torch.exp(-torch.arange(-(ks // 2), ks // 2 + 1, dtype=torch.float32) ** 2 / (2 * sigma ** 2))
    k = (k / k.sum()).numpy()
    y = volume.astype(np.float32)
    for axis in (2, 1, 0):
        y = ndi.correlate1d(y, k, axis=axis, mode="constant", cval=0.0).astype(np.float32)
    return y


def fast_3d_gaussian_smoothing(volume: np.ndarray, scale: float = 0.075) -> np.ndarray:
    labels = np.unique(volume)
    labels = labels[labels != 0]
    result = np.zeros_like(volume, dtype=np.uint8)
    for lb in labels:
        m = volume == lb
        sm = gaussian_smoothing_3d(m, estimate_feature_size_3d(m, scale))
        result[sm > 0.5] = lb
    return result


# ---- REF saber/analysis/refine_membranes.py:100-117,274-333 (R17) ------------------------------
def ball_kernel(radius: int) -> np.ndarray:
    r = np.arange(2 * radius + 1) - radius
    z, y, x = np.meshgrid(r, r, r, indexing="ij")
    return (x ** 2 + y ** 2 + z ** 2 <= radius ** 2)


def binary_erosion_ball(image: np.ndarray, radius: int) -> np.ndarray:
    """`conv3d(zero-padded image, ball) >= sum(ball)`: every voxel under the ball must be set (zero padding)."""
    if image.sum() == 0:
        return image.astype(np.float32)
    return ndi.binary_erosion(image > 0, structure=ball_kernel(radius), border_value=0).astype(np.float32)


def binary_dilation_ball(image: np.ndarray, radius: int) -> np.ndarray:
    if image.sum() == 0:
        return image.astype(np.float32)
    return ndi.binary_dilation(image > 0, structure=ball_kernel(radius), border_value=0).astype(np.float32)


def morphological_opening_ball(image: np.ndarray, radius: int) -> np.ndarray:
    if image.sum() == 0:
        return image.astype(np.float32)
    return binary_dilation_ball(binary_erosion_ball((image > 0).astype(np.float32), radius), radius)
