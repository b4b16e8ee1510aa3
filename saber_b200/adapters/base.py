"""B200 twin of REF saber/adapters/base.py (:7-33 SAM2AdapterConfig, :48-89 BaseAdapter, :92-97
get_adapter) and REF saber/adapters/sam2/amg.py:4-37 (cfgAMG) — same field names, defaults and
validation errors, so configuration code written for the reference is accepted unchanged."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Iterator, List, Literal, Optional, Tuple

import numpy as np
from pydantic import BaseModel, ConfigDict, Field, field_validator, model_validator


class cfgAMG(BaseModel):
    npoints: int = Field(gt=0, default=32)
    points_per_batch: int = Field(gt=0, default=64)
    pred_iou_thresh: float = Field(gt=0, le=1.0, default=0.7)
    stability_score_thresh: float = Field(ge=0, le=1.0, default=0.92)
    stability_score_offset: float = Field(default=0.7)
    crop_n_layers: int = Field(ge=0, default=2)
    box_nms_thresh: float = Field(gt=0, le=1.0, default=0.7)
    crop_n_points_downscale_factor: int = Field(gt=0, default=2)
    use_m2m: bool = Field(default=True)
    multimask_output: bool = Field(default=True)
    sam2_cfg: str = Field(default="small")

    @field_validator("sam2_cfg")
    @classmethod
    def validate_sam2_cfg(cls, v: str) -> str:
        valid = ["tiny", "small", "base", "large"]
        if v not in valid:
            raise ValueError(f"sam2_cfg must be one of {valid}, got {v}")
        return v

    def dict(self, *args: Any, **kwargs: Any) -> Dict[str, Any]:
        return self.model_dump(*args, **kwargs)

    def to_dict(self, *args: Any, **kwargs: Any) -> Dict[str, Any]:
        return self.dict(*args, **kwargs)


class SAM2AdapterConfig(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)

    model_type: Literal["sam2"] = "sam2"
    cfg: str = Field("small", description="tiny / small / base / large")
    checkpoint: Optional[str] = None
    num_maskmem: int = 2
    light_modality: bool = False
    amg_cfg: Optional[Any] = None
    min_mask_area: int = 50
    classifier: Optional[Any] = None
    # not in the reference: without a checkpoint file the adapter raises unless random initialisation (seed `seed`) of the
    # named architecture is requested explicitly (synthetic benchmarks / parity tests; saber_b200.pretrained_weights)
    allow_random_init: bool = False
    seed: int = 0

    @model_validator(mode="after")
    def _derive_from_classifier(self) -> "SAM2AdapterConfig":
        if self.classifier is not None and self.amg_cfg is None:
            amg_params = self.classifier.config["amg_params"]
            self.cfg = amg_params.get("sam2_cfg", self.cfg)
            self.amg_cfg = cfgAMG(**amg_params)
        return self

    @field_validator("cfg")
    @classmethod
    def _check_cfg(cls, v):
        if v not in {"tiny", "small", "base", "large"}:
            raise ValueError(f"cfg must be one of tiny/small/base/large, got '{v}'")
        return v


AdapterConfig = SAM2AdapterConfig


class BaseAdapter(ABC):
    frame_metrics: Dict[int, Dict[int, Dict[str, Any]]]

    @abstractmethod
    def segment_image_2d(self, image: np.ndarray, text_prompt: Optional[str] = None) -> List[Dict[str, Any]]: ...

    @abstractmethod
    def set_volume(self, tomogram: np.ndarray, offload_video_to_cpu: bool = False) -> None: ...

    @abstractmethod
    def add_new_mask(self, frame_idx: int, obj_id: int, mask: np.ndarray, inference_state=None) -> Tuple: ...

    @abstractmethod
    def add_new_points_or_box(self, frame_idx: int, obj_id: int, inference_state=None, **kwargs) -> Tuple: ...

    @abstractmethod
    def propagate_in_video(self, start_frame_idx, max_frame_num_to_track=None, reverse=False,
                           inference_state=None) -> Iterator: ...

    @abstractmethod
    def segment_volume(self, start_frame_idx: int, masks=None, vol_shape=None, max_frame_num_to_track=None,
                       min_presence_score: float = 0.5, inference_state=None) -> np.ndarray: ...

    @abstractmethod
    def reset_state(self, inference_state=None) -> None: ...


def get_adapter(config: AdapterConfig, device: str = "cuda") -> BaseAdapter:
    if config.model_type == "sam2":
        from .sam2 import SAM2Adapter
        return SAM2Adapter(config, device)
    raise ValueError(f"saber_b200 implements the SAM2 adapter only (got model_type={config.model_type!r})")
