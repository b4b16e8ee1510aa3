"""Micrograph wrapper of `saber2D` — REF saber/segmenters/micro.py:8-60: same constructor, the 1280-pixel "consider
down-sampling" notice, and `segment` delegating to `saber2D.segment` (single inference or sliding window)."""
from __future__ import annotations

from typing import Optional

import torch

from ..adapters.base import AdapterConfig, cfgAMG
from .base import saber2D


class cryoMicroSegmenter(saber2D):
    def __init__(self, deviceID: int = 0, cfg: Optional[AdapterConfig] = None, amg_cfg: Optional[cfgAMG] = None,
                 min_mask_area: int = 50, window_size: int = 256, overlap_ratio: float = 0.25):
        super().__init__(cfg=cfg, amg_cfg=amg_cfg, deviceID=deviceID, min_mask_area=min_mask_area, window_size=window_size,
                         overlap_ratio=overlap_ratio)
        self.max_pixels = 1280  # largest side before the notice (REF :25)

    @torch.inference_mode()
    def segment(self, image0, target_class: Optional[int] = None, text: Optional[str] = None, display: bool = True,
                threshold: Optional[float] = 0.5, use_sliding_window: bool = False):
        self.image0 = image0
        nx, ny = image0.shape
        if (nx > self.max_pixels or ny > self.max_pixels) and not use_sliding_window:
            print(f"Image is Larger than {self.max_pixels} pixels in at least one dimension.\nCurrent Size: ({nx}, {ny})")
            print("Consider Downsampling or Using Sliding Window Inference.")
        return super().segment(image0, target_class=target_class, text=text, threshold=threshold, display=display,
                               use_sliding_window=use_sliding_window)
