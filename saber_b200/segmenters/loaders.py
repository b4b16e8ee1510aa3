"""Per-GPU model loaders for SABER's GPUPool (`init_fn`, REF saber/segmenters/loaders.py:9-65): each returns the `models`
dict the workers receive (`extract_sam2_candidates`, `segment_micrograph_core`, `segment_tomogram_core`). Checkpoints are
resolved by the adapters (`saber_b200.pretrained_weights`); nothing is random-initialised unless the configuration opts in."""
from __future__ import annotations

import torch

from ..adapters.base import SAM2AdapterConfig, cfgAMG
from ..classifier.predictor import get_predictor
from .micro import cryoMicroSegmenter
from .tomo import multiDepthTomoSegmenter, tomoSegmenter


def micrograph_workflow(gpu_id: int, cfg: cfgAMG, model_weights: str, model_config: str, target_class: int):
    torch.cuda.set_device(gpu_id)
    predictor = get_predictor(model_weights, model_config, gpu_id)
    adapter_cfg = SAM2AdapterConfig(classifier=predictor, amg_cfg=cfg)
    return {"segmenter": cryoMicroSegmenter(cfg=adapter_cfg, deviceID=gpu_id), "target_class": target_class}


def tomogram_workflow(gpu_id: int, model_weights: str, model_config: str, target_class: int, num_slabs: int):
    torch.cuda.set_device(gpu_id)
    predictor = get_predictor(model_weights, model_config, gpu_id)
    cfg_obj = SAM2AdapterConfig(classifier=predictor)
    if num_slabs > 1:
        segmenter = multiDepthTomoSegmenter(cfg=cfg_obj, deviceID=gpu_id, target_class=target_class)
    else:
        segmenter = tomoSegmenter(cfg=cfg_obj, deviceID=gpu_id)
    return {"predictor": predictor, "segmenter": segmenter, "target_class": target_class}


def base_microsegmenter(gpu_id: int, cfg: cfgAMG):
    torch.cuda.set_device(gpu_id)
    return {"segmenter": cryoMicroSegmenter(amg_cfg=cfg, deviceID=gpu_id)}


def base_tomosegmenter(gpu_id: int, cfg: cfgAMG):
    torch.cuda.set_device(gpu_id)
    return {"segmenter": tomoSegmenter(amg_cfg=cfg, deviceID=gpu_id)}
