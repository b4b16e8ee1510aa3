"""B200 twin of REF saber/adapters/sam2/{predictor,automask,amg}.py — the SAM2 adapter SABER's segmenters
drive. Same class / method names and return types; the arithmetic underneath is the sm_100a kernel path.

* ``build_amg``                         REF saber/adapters/sam2/automask.py:49-86
* ``FilteredSAM2MaskGenerator``         REF saber/adapters/sam2/amg.py:139-201 (min/max area + relative box filters)
* ``SAM2Adapter.segment_image_2d``      REF saber/adapters/sam2/predictor.py:48-70
* ``SAM2Adapter.segment_image_2d_device`` resident variant (CUDA slice in, DeviceMasks out) used by the
  slice-wise segmenter so masks never leave the GPU.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from ..sam2.automatic_mask_generator import DeviceMasks, SAM2AutomaticMaskGenerator
from ..sam2.build_sam import build_sam2
from ..utils import preprocessing as prep
from .base import BaseAdapter, SAM2AdapterConfig, cfgAMG

_CFG_TO_ARCH = {"tiny": "tiny", "small": "small", "base": "base_plus", "large": "large"}


def filter_masks_by_area(anns: List[Dict[str, Any]], min_area: Optional[int] = None,
                         max_area: Optional[int] = None) -> List[Dict[str, Any]]:
    if min_area is None and max_area is None:
        return anns
    out = []
    for ann in anns:
        area = ann.get("area", 0)
        if (min_area is None or area >= min_area) and (max_area is None or area <= max_area):
            out.append(ann)
    return out


def filter_masks_by_relative_box_size(anns, max_rel_box_size=None, min_rel_box_size=None, image_height=None,
                                      image_width=None):
    if max_rel_box_size is None and min_rel_box_size is None:
        return anns
    if image_height is None or image_width is None:
        raise ValueError("image_height and image_width must be provided for relative size filtering")
    out = []
    for ann in anns:
        bbox = ann.get("bbox", None)
        if bbox is None:
            continue
        _, _, w, h = bbox
        rw, rh = w / image_width, h / image_height
        ok = True
        if max_rel_box_size is not None:
            ok = ok and rw < max_rel_box_size and rh < max_rel_box_size
        if min_rel_box_size is not None:
            ok = ok and rw > min_rel_box_size and rh > min_rel_box_size
        if ok:
            out.append(ann)
    return out


class FilteredSAM2MaskGenerator:
    def __init__(self, base_generator, min_rel_box_size=None, max_rel_box_size=None, min_area_filter=None,
                 max_area_filter=None):
        self.base_generator = base_generator
        self.max_rel_box_size = max_rel_box_size
        self.min_rel_box_size = min_rel_box_size
        self.min_area_filter = min_area_filter
        self.max_area_filter = max_area_filter

    def _filter(self, anns, hw):
        anns = filter_masks_by_relative_box_size(anns, self.max_rel_box_size, self.min_rel_box_size, hw[0], hw[1])
        return filter_masks_by_area(anns, self.min_area_filter, self.max_area_filter)

    def generate(self, image) -> List[Dict[str, Any]]:
        return self._filter(self.base_generator.generate(image), image.shape[:2])

    def generate_device(self, image):
        """Resident variant: (DeviceMasks, kept record list); each record carries ``index`` into the DeviceMasks."""
        dm = self.base_generator.generate_device(image)
        recs = self.base_generator.records(dm)
        for i, r in enumerate(recs):
            r["index"] = i
        return dm, self._filter(recs, dm.hw)

    def set_filters(self, min_rel_box_size=None, max_rel_box_size=None, min_area_filter=None):
        if min_rel_box_size is not None:
            self.min_rel_box_size = min_rel_box_size
        if max_rel_box_size is not None:
            self.max_rel_box_size = max_rel_box_size
        if min_area_filter is not None:
            self.min_area_filter = min_area_filter

    def __getattr__(self, name):
        return getattr(self.base_generator, name)


def build_amg(amg_params: Dict[str, Any], min_mask_area: int, device="cuda", checkpoint: Optional[str] = None,
              seed: int = 0, model=None):
    """REF saber/adapters/sam2/automask.py:49-86: build_sam2(apply_postprocessing=True) + AMG + area filter."""
    if model is None:
        model = build_sam2(_CFG_TO_ARCH[amg_params["sam2_cfg"]], checkpoint, device=device,
                           apply_postprocessing=True, seed=seed)
        model.eval()
    gen = SAM2AutomaticMaskGenerator(
        model=model, points_per_side=amg_params["npoints"], points_per_batch=amg_params["points_per_batch"],
        pred_iou_thresh=amg_params["pred_iou_thresh"], stability_score_thresh=amg_params["stability_score_thresh"],
        stability_score_offset=amg_params["stability_score_offset"], crop_n_layers=amg_params["crop_n_layers"],
        box_nms_thresh=amg_params["box_nms_thresh"],
        crop_n_points_downscale_factor=amg_params["crop_n_points_downscale_factor"],
        use_m2m=amg_params["use_m2m"], multimask_output=amg_params["multimask_output"])
    return FilteredSAM2MaskGenerator(base_generator=gen, min_area_filter=min_mask_area)


class SAM2Adapter(BaseAdapter):
    def __init__(self, config: SAM2AdapterConfig, device: str = "cuda"):
        if config.num_maskmem > 7:
            raise ValueError("num_maskmem must be less than 7")
        self._config = config
        self.device = torch.device(device)
        self.frame_metrics: Dict[int, Dict[int, Dict[str, Any]]] = {}
        self._vol_shape = None
        self.inference_state = None
        self._current_frame = None
        self._mask_generator = None
        self.predictor = None  # video predictor: built on first volume call

    # ------------------------------------------------------------------ 2-D segmentation
    def _amg(self):
        if self._mask_generator is None:
            if self._config.amg_cfg is not None:
                amg_dict = self._config.amg_cfg.dict()
            else:
                amg_dict = cfgAMG(sam2_cfg=self._config.cfg).dict()
            self._mask_generator = build_amg(amg_dict, self._config.min_mask_area, device=self.device,
                                             checkpoint=self._config.checkpoint, seed=self._config.seed)
        return self._mask_generator

    @torch.inference_mode()
    def segment_image_2d(self, image: np.ndarray, text_prompt: str = None, threshold: float = None):
        out_rgb = True if image.ndim == 2 else False
        if image.ndim != 2:
            raise NotImplementedError("saber_b200: segment_image_2d takes a grayscale (H,W) slice")
        x = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(self.device)
        x = prep.prepare(x, to_rgb=out_rgb)
        return self._amg().generate(x)

    @torch.inference_mode()
    def segment_image_2d_device(self, image: torch.Tensor):
        """CUDA (H,W) slice -> (DeviceMasks, filtered record list); nothing but the records leaves the GPU."""
        assert image.is_cuda and image.dim() == 2
        x = prep.prepare_device(image)
        return self._amg().generate_device(x)

    # ------------------------------------------------------------------ 3-D (memory propagation) — next §8 row
    def _no3d(self, *a, **k):
        raise NotImplementedError("saber_b200: z-axis memory propagation is not built yet (SURVEY §8a U6-U9)")

    set_volume = add_new_mask = add_new_points_or_box = propagate_in_video = segment_volume = _no3d

    def reset_state(self, inference_state=None) -> None:
        self.inference_state = None
        self.frame_metrics = {}
