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


# ---- REF saber/analysis/refine_membranes.py:120-548 (§8f row 3: the membrane-refinement workflow) ------------------------
def _keep_components(binary: np.ndarray, min_size: int) -> np.ndarray:
    """REF :136-158 / :202-222: components (scipy default = 6-connectivity) with at least min_size voxels."""
    labels, n = ndi.label(binary)
    if n == 0:
        return np.zeros_like(binary, dtype=bool)
    counts = np.bincount(labels.ravel())
    keep = counts >= min_size
    keep[0] = False
    return keep[labels]


def _largest_component(binary: np.ndarray) -> np.ndarray:
    """REF :224-249: the component with the most voxels, the lowest label among equals (np.argmax)."""
    labels, n = ndi.label(binary)
    if n == 0:
        return np.zeros_like(binary, dtype=bool)
    counts = np.bincount(labels.ravel())[1:]
    return labels == (int(np.argmax(counts)) + 1)


def refine_membranes(organelle_seg: np.ndarray, membrane_seg: np.ndarray, ball_size: int = 3, min_membrane_area: int = 10000,
                     edge_trim_z: int = 5, edge_trim_xy: int = 3, min_roi_relative_size: float = 0.15,
                     keep_surface_membranes: bool = False):
    """OrganelleMembraneFilter.run (REF :445-548) + _process_organelle_batch (REF :335-443) on binary numpy volumes.
    Returns (organelles [n,Z,Y,X], membranes [n,Z,Y,X]) in the organelle dtype, or two zero [Z,Y,X] volumes when
    nothing survives (REF :473-480, :515-522). Output values are label + 1 (the even / odd relabelling of REF
    :484-485, 431-435 divided back by two at :529-530)."""
    Z, Y, X = organelle_seg.shape
    empty = (np.zeros_like(organelle_seg), np.zeros_like(organelle_seg))
    # REF :120-135 — `[t:-t]` is empty for t == 0 and the assignment is skipped unless t < size // 2
    trimmed = np.zeros((Z, Y, X), bool)
    if 0 < edge_trim_z < Z // 2 and 0 < edge_trim_xy < Y // 2 and edge_trim_xy < X // 2:
        zt, t = edge_trim_z, edge_trim_xy
        trimmed[zt:-zt, t:-t, t:-t] = membrane_seg[zt:-zt, t:-t, t:-t] != 0
    membrane = _keep_components(trimmed, min_membrane_area)                       # REF :459-462
    present = membrane.any(axis=(1, 2))                                           # REF :466
    filtered = organelle_seg * present[:, None, None].astype(organelle_seg.dtype)  # REF :467
    labels = np.unique(filtered)
    labels = labels[labels > 0]
    if len(labels) == 0:
        return empty
    pad = ball_size // 2
    out_o, out_m = [], []
    for lab in labels:
        org = filtered == lab
        idx = np.argwhere(org)                                                    # REF :251-272
        mins, maxs = idx.min(0), idx.max(0) + 1
        if ((maxs - mins) < min_roi_relative_size * np.array([Z, Y, X])).any():
            continue
        mins = np.maximum(mins - pad, 0)
        maxs = np.minimum(maxs + pad, [Z, Y, X])
        sl = tuple(slice(int(a), int(b)) for a, b in zip(mins, maxs))
        org_roi, mem_roi = org[sl], membrane[sl]
        roi_shape = (maxs - mins).astype(np.float32)
        if np.float32(roi_shape.max()) / np.float32(roi_shape.min()) > 3.0:       # REF :366-376
            dilate_size, morph_ball = 1, max(1, ball_size // 2)
        else:
            dilate_size, morph_ball = 2, ball_size
        ball_d = ball_kernel(dilate_size)
        enhanced = ndi.binary_dilation(mem_roi, structure=ball_d) & ndi.binary_dilation(org_roi, structure=ball_d)  # :379-384
        cleaned = enhanced & _keep_components(enhanced, 100)                      # REF :394-395
        if keep_surface_membranes and cleaned.any():                              # REF :160-199
            boundary = org_roi & ~ndi.binary_erosion(org_roi, structure=np.ones((3, 3, 3)))
            lab_m, n_m = ndi.label(cleaned)
            size = np.bincount(lab_m.ravel(), minlength=n_m + 1)
            over = np.bincount(lab_m[boundary].ravel(), minlength=n_m + 1)
            keep = np.zeros(n_m + 1, bool)
            keep[1:] = over[1:] / np.maximum(size[1:], 1) > 0.1
            cleaned = keep[lab_m]
        if not cleaned.any():
            continue
        # REF :405-410 — organelle voxels carry the even label (>= 4), so `organelle - membrane` is non-zero wherever
        # EITHER is set: the "combined" mask the opening sees is the union, not the difference.
        comb = org_roi | cleaned
        opened = ndi.binary_dilation(ndi.binary_erosion(comb, structure=ball_kernel(morph_ball)), structure=ball_kernel(morph_ball))
        if not opened.any():
            opened = comb                                                         # REF :416-418
        comb_out = _largest_component(opened)                                     # REF :425
        org_out = _largest_component(org_roi & comb_out)                          # REF :428-429
        mem_out = cleaned & comb_out
        mem_out = mem_out & _keep_components(mem_out, 50)                         # REF :432-433
        full_o = np.zeros_like(organelle_seg)
        full_m = np.zeros_like(organelle_seg)
        full_o[sl][org_out] = lab + 1
        full_m[sl][mem_out] = lab + 1
        out_o.append(full_o)
        out_m.append(full_m)
    if not out_o:
        return empty
    return np.stack(out_o), np.stack(out_m)


def convert_to_3d_labels(masks_4d: np.ndarray) -> np.ndarray:
    """REF :548-573: later instances overwrite earlier ones."""
    if masks_4d.ndim == 3:
        return masks_4d
    out = np.zeros(masks_4d.shape[1:], masks_4d.dtype)
    for m in masks_4d:
        out[m > 0] = m[m > 0]
    return out


# ---- REF saber/filters/downsample.py:67-129,153-204 and saber/filters/tomograms.py:67-184 (§8f row 2) --------------------
def _crop_window(n_in: int, n_new: int):
    n_new = n_new - (n_new % 2)                                  # REF downsample.py:117-119 / :183-184
    return (n_in - n_new) // 2 + (n_in % 2), n_new                # REF :122-127 / :190-191


def fourier_rescale_3d(volume: np.ndarray, input_voxel_size, output_voxel_size) -> np.ndarray:
    """FourierRescale3D.run (REF downsample.py:67-129) in float64 numpy; returns float32 like the reference."""
    vin = (input_voxel_size,) * 3 if np.isscalar(input_voxel_size) else input_voxel_size
    vout = (output_voxel_size,) * 3 if np.isscalar(output_voxel_size) else output_voxel_size
    wins = [_crop_window(n, int(round(n * i / o))) for n, i, o in zip(volume.shape, vin, vout)]
    spec = np.fft.fftshift(np.fft.fftn(volume.astype(np.float64), norm="ortho"))
    spec = spec[tuple(slice(s, s + m) for s, m in wins)]
    return np.fft.ifftn(np.fft.ifftshift(spec), norm="ortho").real.astype(np.float32)


def fourier_rescale_2d(image: np.ndarray, scale_factor: float) -> np.ndarray:
    """FourierRescale2D._rescale (REF downsample.py:153-204): modulus of the inverse transform, default norms."""
    h, w = image.shape
    wins = [_crop_window(h, int(h / scale_factor)), _crop_window(w, int(w / scale_factor))]
    spec = np.fft.fftshift(np.fft.fft2(image.astype(np.float64)))
    spec = spec[tuple(slice(s, s + m) for s, m in wins)]
    return np.abs(np.fft.ifft2(np.fft.ifftshift(spec))).astype(np.float32)


def cosine_filter(sz, apix: float, lp=0, lpd=0, hp=0, hpd=0) -> np.ndarray:
    """Filter3D.cosine_filter / construct_filter (REF tomograms.py:67-137), fp32 like the reference."""
    D, H, W = sz
    to_pix = lambda ang: max(sz) / (ang / apix)
    zz, yy, xx = np.meshgrid(np.arange(D, dtype=np.float32) - D // 2, np.arange(H, dtype=np.float32) - H // 2,
                             np.arange(W, dtype=np.float32) - W // 2, indexing="ij")
    r = np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)

    def edge(freq, decay, highpass):
        if freq == 0 and decay == 0:
            return np.ones_like(r)
        m = (r < np.float32(freq)).astype(np.float32)
        if decay != 0:
            half = decay / 2.0
            sel = (r > np.float32(freq - half)) & (r < np.float32(freq + half))
            m[sel] = np.float32(0.5) + np.float32(0.5) * np.cos(np.float32(np.pi) * (r[sel] - np.float32(freq - half)) / np.float32(decay))
        return 1 - m if highpass else m

    return edge(to_pix(lp) if lp > 0 else 0, lpd, False) * edge(to_pix(hp) if hp > 0 else 0, hpd, True)


def filter3d_apply(volume: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """Filter3D.apply (REF tomograms.py:172-194)."""
    spec = np.fft.fftshift(np.fft.fftn(volume.astype(np.float64))) * filt
    return np.fft.ifftn(np.fft.ifftshift(spec)).real.astype(np.float32)
