"""B200 twin of REF saber/segmenters/tomo.py: ``tomoSegmenter`` (:14-160) and ``multiDepthTomoSegmenter`` (:163-254).

segment_slab: z-Gaussian (sigma 5) -> min-max -> z-slab mean around zSlice -> 2-D AMG; segment_vol: slab masks seed the
z-axis memory propagation (adapter.segment_volume). The tomogram is uploaded once and stays on the device; the smoothed /
normalised volume, the slab projection, the frame features and the label volume never visit the host (the reference
round-trips through numpy between every stage)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from ..adapters.base import AdapterConfig, cfgAMG
from ..filters import gaussian as gauss
from . import utils
from .base import saber3D


class tomoSegmenter(saber3D):
    def __init__(self, deviceID: int = 0, cfg: Optional[AdapterConfig] = None, amg_cfg: Optional[cfgAMG] = None,
                 min_mask_area: int = 50):
        super().__init__(deviceID=deviceID, cfg=cfg, amg_cfg=amg_cfg, min_mask_area=min_mask_area)
        self.filter_threshold = 0.5
        self._vol_src = None

    def _load(self, vol) -> torch.Tensor:
        """REF :44-46: gaussian_smoothing(vol, 5, dim=0) then preprocessing.normalize -> resident fp32 volume."""
        if self._vol_src is not vol:
            v = (vol if isinstance(vol, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vol, dtype=np.float32)))
            v = v.to(self.device, dtype=torch.float32).contiguous()
            v = gauss.gaussian_smoothing(v, 5, dim=0)
            self.vol = ops.minmax_affine(v, ops.minmax(v), 1e-8, 1.0, 0.0)
            self._vol_src = vol
            self._vol_loaded = False
        return self.vol

    @torch.inference_mode()
    def segment_slab(self, vol, slab_thickness: int = 10, zSlice: Optional[int] = None, display: bool = True,
                     text: Optional[str] = None, target_class: Optional[int] = 1):
        v = self._load(vol)
        if zSlice is None:
            zSlice = int(v.shape[0] // 2)
        z0 = int(max(zSlice - slab_thickness, 0))
        z1 = int(min(zSlice + slab_thickness, v.shape[0]))
        self.image0 = ops.mean_z(v, z0, z1)  # project_tomogram(vol, zSlice, slab_thickness)
        self.segment_image(self.image0, display=False, text_prompt=text, target_class=target_class)
        return self.masks

    def segment(self, vol, thickness: int = 10, zSlice: int = None, text=None, target_class=1, save_run=None,
                display: bool = False):
        return self.segment_vol(vol, thickness, zSlice, text, target_class, save_run, display)

    @torch.inference_mode()
    def segment_vol(self, vol, thickness: int, zSlice: int = None, text=None, target_class=1, save_run=None,
                    display: bool = False):
        self.is_tomogram_mode = True
        self.segment_slab(vol, thickness, zSlice, display=False, text=text, target_class=target_class)
        if len(self.masks) == 0:
            return None
        if not self._vol_loaded:
            self.video_predictor.set_volume(self.vol)
            self._vol_loaded = True
        nx = self.vol.shape[0]
        ny, nz = self.masks[0]["segmentation"].shape[:2]
        self.ann_frame_idx = zSlice if zSlice is not None else nx // 2
        return self.propagate((nx, ny, nz))


class multiDepthTomoSegmenter(tomoSegmenter):
    def __init__(self, deviceID: int = 0, cfg: Optional[AdapterConfig] = None, amg_cfg: Optional[cfgAMG] = None,
                 target_class: int = 1, min_mask_area: int = 100, min_rel_box_size: float = 0.025):
        self.min_rel_box_size = min_rel_box_size
        self.target_class = target_class
        super().__init__(deviceID=deviceID, cfg=cfg, amg_cfg=amg_cfg, min_mask_area=min_mask_area)
        if target_class < 1:
            raise ValueError("Multi-Depth Tomogram Segmenter only supports Single-Class Segmentation currently.")

    def segment(self, vol, thickness: int, num_slabs: int = 3, delta_z: int = 30, save_run=None, display=False):
        self.show_segments = display
        if self.target_class > 0 or self.classifier is None:
            return self.single_segment(vol, thickness, num_slabs, delta_z)
        print("Multiclass Segmentation is not implemented yet")

    @torch.inference_mode()
    def single_segment(self, vol, thickness, num_slabs, delta_z):
        """REF :206-254: slabs at centre + (i - n//2) * delta_z, binarised propagation results merged with max, then
        26-connected component separation."""
        depth = vol.shape[0]
        center = depth // 2
        combined = np.zeros(tuple(vol.shape), dtype=np.uint16)
        for i in range(num_slabs):
            slab_center = int(center + (i - num_slabs // 2) * delta_z)
            if slab_center < 0 or slab_center >= depth:
                continue
            masks3d = self.segment_vol(vol, thickness, zSlice=slab_center, display=False)
            if masks3d is None:
                continue
            np.maximum(combined, (masks3d > 0).astype(np.uint16), out=combined)
        return utils.separate_masks(combined, device=self.device)
