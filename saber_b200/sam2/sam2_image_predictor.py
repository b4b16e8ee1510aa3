"""Drop-in for ``sam2.sam2_image_predictor.SAM2ImagePredictor`` on the B200 kernels.

Surface used by the reference: ``SAM2ImagePredictor(model)``, ``.model``, ``.set_image_batch(list of
(H,W,3) float32 arrays)``, ``._features["image_embed"]`` ``[B,256,64,64]`` and
``._features["high_res_feats"]`` (REF saber/classifier/models/SAM2.py:46-51,145-151), plus the
``set_image`` / ``_predict`` pair the automatic mask generator drives (upstream
sam2/automatic_mask_generator.py). Restates sam2/sam2_image_predictor.py + sam2/utils/transforms.py
(SURVEY §8a U4): float input is not rescaled, Resize(1024^2) is bilinear with antialias, ImageNet
mean/std, ``no_mem_embed`` is added to the lowest-resolution feature.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import ops
from . import arch

_F32 = torch.float32
ImageLike = Union[np.ndarray, torch.Tensor]


class _Features(dict):
    """``_features`` dict: token-major device features are authoritative; the NCHW views the reference
    reads (``image_embed``, ``high_res_feats``) are materialised on first access."""

    def __init__(self, feat, s0, s1, batch):
        super().__init__()
        self.tok = {"embed": feat, "s0": s0, "s1": s1}
        self.batch = batch

    def __missing__(self, key):
        B = self.batch
        if key == "image_embed":
            v = ops.nhwc_to_nchw(self.tok["embed"], B, 4096, _F32).view(B, 256, 64, 64)
        elif key == "high_res_feats":
            v = [ops.nhwc_to_nchw(self.tok["s0"], B, 65536, _F32).view(B, 32, 256, 256),
                 ops.nhwc_to_nchw(self.tok["s1"], B, 16384, _F32).view(B, 64, 128, 128)]
        else:
            raise KeyError(key)
        self[key] = v
        return v


class SAM2ImagePredictor:
    def __init__(self, sam_model, mask_threshold: float = 0.0, max_hole_area: float = 0.0,
                 max_sprinkle_area: float = 0.0, **kwargs):
        self.model = sam_model
        self.mask_threshold = mask_threshold
        self.resolution = sam_model.image_size
        self._bb_feat_sizes = [(256, 256), (128, 128), (64, 64)]
        self.max_encode_batch = int(os.environ.get("SB_ENCODE_BATCH", "24"))  # crops per encoder pass (all 21 AMG crops at once)
        self.reset_predictor()

    @property
    def device(self):
        return self.model.device

    def reset_predictor(self) -> None:
        self._is_image_set = False
        self._features: Optional[_Features] = None
        self._orig_hw: Optional[List[tuple]] = None
        self._is_batch = False

    # ------------------------------------------------------------------
    def _to_device_image(self, image: ImageLike) -> torch.Tensor:
        """(H,W,3) / (H,W) uint8 or float -> contiguous fp32 CUDA tensor (uint8 is divided by 255 as ToTensor does)."""
        if isinstance(image, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(image))
        else:
            t = image
        if t.dtype == torch.uint8:
            t = t.to(self.device, non_blocking=True).float() / 255.0  # host-format conversion, not model arithmetic
        else:
            t = t.to(self.device, dtype=_F32, non_blocking=True)
        assert t.dim() in (2, 3) and (t.dim() == 2 or t.shape[2] == 3), f"expected HxW or HxWx3 image, got {tuple(t.shape)}"
        return t.contiguous()

    @torch.no_grad()
    def encode_crops(self, image: torch.Tensor, crops: torch.Tensor) -> _Features:
        """Encode ``crops`` (int32 [n,4] xyxy, device) of one device image; returns token-major features."""
        self.model._require_gpu()
        n = crops.shape[0]
        feats, s0s, s1s = [], [], []
        for k0 in range(0, n, self.max_encode_batch):
            kb = min(self.max_encode_batch, n - k0)
            x = ops.resize_normalize(image, crops[k0:k0 + kb].contiguous(), self.resolution)
            out = self.model.encoder.forward(x)
            del x
            feats.append(ops.add_cast(out["feat"], self.model.no_mem_embed_vec, _F32))
            s0s.append(out["s0"])
            s1s.append(out["s1"])
        cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs, 0))
        return _Features(cat(feats), cat(s0s), cat(s1s), n)

    @torch.no_grad()
    def set_image(self, image: ImageLike) -> None:
        self.reset_predictor()
        img = self._to_device_image(image)
        H, W = img.shape[:2]
        self._orig_hw = [(H, W)]
        crops = torch.tensor([[0, 0, W, H]], dtype=torch.int32, device=self.device)
        self._features = self.encode_crops(img, crops)
        self._is_image_set = True

    @torch.no_grad()
    def set_image_batch(self, image_list: Sequence[ImageLike]) -> None:
        self.reset_predictor()
        assert isinstance(image_list, (list, tuple)) and len(image_list) > 0
        self._orig_hw = []
        parts = []
        for im in image_list:
            img = self._to_device_image(im)
            H, W = img.shape[:2]
            self._orig_hw.append((H, W))
            crops = torch.tensor([[0, 0, W, H]], dtype=torch.int32, device=self.device)
            parts.append(self.encode_crops(img, crops))
        if len(parts) == 1:
            self._features = parts[0]
        else:
            self._features = _Features(torch.cat([p.tok["embed"] for p in parts], 0),
                                       torch.cat([p.tok["s0"] for p in parts], 0),
                                       torch.cat([p.tok["s1"] for p in parts], 0), len(parts))
        self._is_image_set = True
        self._is_batch = True

    def get_image_embedding(self) -> torch.Tensor:
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self._features["image_embed"]

    # ------------------------------------------------------------------
    @torch.no_grad()
    def _predict(self, point_coords: torch.Tensor, point_labels: torch.Tensor, boxes=None,
                 mask_input: Optional[torch.Tensor] = None, multimask_output: bool = True,
                 return_logits: bool = False, img_idx: int = -1):
        """point_coords [B,Np,2] fp32 (model-input pixels), point_labels [B,Np] int. Returns
        (masks, iou_predictions, low_res_masks) shaped like upstream: masks are the low-res logits
        bilinearly up-sampled to the original size (via torch-free kernels), thresholded unless return_logits."""
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        if boxes is not None:
            raise NotImplementedError("box prompts are not on SABER's SAM2 path")
        dec = self.model.decoder
        B = self._features.batch
        k = img_idx % B
        tok = self._features.tok
        emb = tok["embed"][k * 4096:(k + 1) * 4096]
        s0 = tok["s0"][k * 65536:(k + 1) * 65536]
        s1 = tok["s1"][k * 16384:(k + 1) * 16384]
        tokens = dec.prompt_tokens(point_coords.to(self.device, _F32).contiguous(),
                                   point_labels.to(self.device, torch.int32).contiguous())
        mi = None
        if mask_input is not None:
            mi = mask_input.to(self.device, _F32).reshape(-1, 256, 256).contiguous()
        out = dec.forward(emb, s0, s1, tokens, mi, multimask_output=multimask_output)
        P = tokens.shape[0]
        if multimask_output:
            low = out["masks"][:, 1:4]
            ious = out["ious"][:, 1:4]
        elif out.get("sel_idx") is not None:
            idx = out["sel_idx"].long()
            low = out["masks"][torch.arange(P, device=self.device), idx].unsqueeze(1)
            ious = out["sel_iou"].unsqueeze(1)
        else:
            low = out["masks"][:, 0:1]
            ious = out["ious"][:, 0:1]
        H, W = self._orig_hw[k]
        masks = ops.upsample_bilinear(low.contiguous(), H, W)
        low = low.clamp(-32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks, ious, low
