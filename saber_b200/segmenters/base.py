"""B200 twin of REF saber/segmenters/base.py: ``saber2D`` (:18-232) and ``saber3D`` (:234-280).

Same constructor arguments, attributes and method names. ``segment_image`` returns the reference's list of
mask dicts; ``segment_image_device`` is the resident variant (CUDA slice in, ordered packed masks out) that
``propagationSegmenter.slice_by_slice`` uses so that only the final label volume is copied to the host.
The expert-classifier branch of ``_apply_classifier`` (REF :170-174) calls ``saber_b200.filters.masks.apply_classifier``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..adapters.base import AdapterConfig, SAM2AdapterConfig, cfgAMG, get_adapter
from . import utils

_I32 = torch.int32


class saber2D:
    def __init__(self, deviceID: int = 0, cfg: Optional[AdapterConfig] = None, amg_cfg: Optional[cfgAMG] = None,
                 min_mask_area: int = 50, window_size: int = 256, overlap_ratio: float = 0.25):
        if cfg is None and amg_cfg is None:
            raise ValueError("Either Provide an AdapterConfig or AMG Config!")
        if cfg is None:
            cfg = SAM2AdapterConfig(cfg=amg_cfg.sam2_cfg, amg_cfg=amg_cfg, min_mask_area=min_mask_area)
        self.min_mask_area = min_mask_area
        self.window_size = window_size
        self.overlap_ratio = overlap_ratio
        if not torch.cuda.is_available():
            raise RuntimeError("saber_b200 segmenters need a B200 (sm_100a) device; there is no CPU path")
        self.device = torch.device(f"cuda:{deviceID}")
        self.deviceID = deviceID
        _classifier = getattr(cfg, "classifier", None)
        self.classifier = _classifier
        self.batchsize = None if _classifier is None else 32
        self.adapter_cfg = cfg
        self.adapter = get_adapter(cfg, self.device)
        self.image = None
        self.save_button = False
        self.remove_repeating_masks = True

    def segment(self, image, target_class=None, text=None, threshold=0.5, display=False, use_sliding_window=False):
        return self.segment_image(image, display=display, use_sliding_window=use_sliding_window, text_prompt=text,
                                  threshold=threshold, target_class=target_class)

    @torch.inference_mode()
    def segment_image(self, image: np.ndarray, display: bool = True, use_sliding_window: bool = False,
                      text_prompt: Optional[str] = None, threshold: Optional[float] = 0.5,
                      target_class: Optional[int] = 1):
        self.target_class = target_class
        if use_sliding_window:
            windows = self.get_sliding_windows(image.shape)
            all_masks = []
            for (y1, x1, y2, x2) in windows:
                window_image = image[y1:y2, x1:x2]
                window_masks = self.adapter.segment_image_2d(window_image, text_prompt=text_prompt, threshold=threshold)
                curr = []
                for mask in window_masks:
                    if mask["area"] < self.min_mask_area:
                        continue
                    mask["offset"] = (y1, x1)
                    mask["bbox"] = self._to_global_bbox(mask["bbox"], y1, x1)
                    curr.append(mask)
                all_masks.extend(self._apply_classifier(window_image, curr))
            self.masks = self.rasterize_masks(image, all_masks)
        else:
            self.masks = self.adapter.segment_image_2d(image, text_prompt=text_prompt, threshold=threshold)
            self.masks = self._apply_classifier(image, self.masks)
        self.image = image
        return self.masks

    def _apply_classifier(self, image, masks):
        masks = [m for m in masks if m["area"] >= self.min_mask_area]
        if self.remove_repeating_masks:
            masks = utils.remove_duplicate_masks(masks, device=self.device)
        if self.classifier is None:
            return sorted(masks, key=lambda m: m["area"], reverse=False)
        from ..filters import masks as filters
        gray = image[:, :, 0] if image.ndim == 3 else image
        # REF :170-174 passes (target_class, batchsize) positionally: batchsize lands in `min_mask_area`
        return filters.apply_classifier(gray, masks, self.classifier, self.target_class, self.batchsize)

    @torch.inference_mode()
    def segment_image_device(self, image: torch.Tensor):
        """Resident twin of segment_image (no sliding window, no classifier): returns (DeviceMasks, order) where
        ``order`` lists DeviceMasks rows in the reference's final list order (area filter -> duplicate removal ->
        stable sort by area)."""
        if self.classifier is not None:
            raise RuntimeError("segment_image_device is the classifier-free resident path; with an expert classifier use "
                               "segment_image (propagationSegmenter.label_slices_device does)")
        dm, recs = self.adapter.segment_image_2d_device(image)
        recs = [r for r in recs if r["area"] >= self.min_mask_area]
        if self.remove_repeating_masks and len(recs) > 1:
            idx = torch.tensor([r["index"] for r in recs], dtype=_I32, device=dm.bits.device)
            m = len(recs)
            keep = utils.remove_duplicate_indices(ops.gather_rows(dm.bits, idx, m), ops.gather_rows(dm.bbox, idx, m),
                                                  ops.gather_rows(dm.area, idx, m),
                                                  [r["stability_score"] for r in recs], dm.hw[1])
            recs = [recs[i] for i in keep]
        recs = sorted(recs, key=lambda r: r["area"], reverse=False)
        return dm, recs

    def get_sliding_windows(self, image_shape: Tuple[int, int]) -> List[Tuple[int, int, int, int]]:
        h, w = image_shape[:2]
        stride = int(self.window_size * (1 - self.overlap_ratio))
        windows = []
        for y in range(0, h, stride):
            for x in range(0, w, stride):
                y2, x2 = min(y + self.window_size, h), min(x + self.window_size, w)
                if (y2 - y) < self.window_size // 2 or (x2 - x) < self.window_size // 2:
                    continue
                windows.append((y, x, y2, x2))
        return windows

    def _to_global_bbox(self, local_bbox, y0, x0):
        x, y, w, h = local_bbox
        return [x + x0, y + y0, w, h]

    def rasterize_masks(self, image, masks):
        H, W = image.shape[:2]
        disp = []
        for m in masks:
            y0, x0 = m["offset"]
            seg = m["segmentation"]
            h, w = seg.shape
            full = np.zeros((H, W), dtype=bool)
            y1, x1 = max(0, y0), max(0, x0)
            y2, x2 = min(H, y0 + h), min(W, x0 + w)
            sy1, sx1 = y1 - y0, x1 - x0
            full[y1:y2, x1:x2] = seg[sy1:sy1 + (y2 - y1), sx1:sx1 + (x2 - x1)]
            m2 = dict(m)
            m2["segmentation"] = full
            disp.append(m2)
        return disp


class saber3D(saber2D):
    def __init__(self, deviceID: int = 0, cfg: AdapterConfig = None, amg_cfg: cfgAMG = None, min_mask_area: int = 50):
        super().__init__(deviceID=deviceID, cfg=cfg, amg_cfg=amg_cfg, min_mask_area=min_mask_area)
        self.video_predictor = self.adapter
        self._vol_loaded = False
        self.min_logits = 0.5
        self.confidence_debug = False
        self.nframes = None
        self.filter_threshold = 0.5

    def propagate(self, mask_shape, target_class: Optional[int] = 1):
        mask_arrays = [m["segmentation"] for m in self.masks] if isinstance(self.masks[0], dict) else self.masks
        vol_masks = self.video_predictor.segment_volume(
            start_frame_idx=self.ann_frame_idx, masks=mask_arrays, vol_shape=tuple(int(v) for v in mask_shape),
            max_frame_num_to_track=self.nframes, min_presence_score=self.filter_threshold)
        self.video_predictor.reset_state()
        return vol_masks
