"""B200 twin of SABER's expert classifier: same class / method names as

  * ``Predictor.{predict, batch_predict, preprocess, apply_crops}``  REF saber/classifier/models/predictor.py:9-234
  * ``SAM2Classifier.{forward, apply_mask_to_features}``              REF saber/classifier/models/SAM2.py:21-197
  * ``get_predictor``                                                 REF saber/classifier/models/common.py:24-50

B200 design: the slice is normalised once on the device; bounding boxes of all candidate masks come from one kernel
and one small D2H copy (the crop geometry is integer host logic, as in the reference); crops, the Hiera encoder (batched,
not one image at a time), ROI/RONI masking and the conv head run device-resident — the reference round-trips every batch
through numpy (REF SAM2.py:136). Eval-mode BatchNorm is folded into the conv weights at load; convs are GEMMs on
token-major activations (3x3 via im2col); dropout is identity in eval mode.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import ops
from ..sam2.build_sam import build_sam2
from ..sam2.sam2_image_predictor import SAM2ImagePredictor

_BF16, _F32, _I32, _U8 = torch.bfloat16, torch.float32, torch.int32, torch.uint8
_CFG_TO_ARCH = {"tiny": "tiny", "small": "small", "base": "base_plus", "large": "large"}


def _fold_bn(w: torch.Tensor, b: torch.Tensor, sd: Dict[str, torch.Tensor], bn: str, eps: float = 1e-5):
    g, beta = sd[bn + ".weight"].float(), sd[bn + ".bias"].float()
    mu, var = sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float()
    s = g / torch.sqrt(var + eps)
    return w * s.view(-1, *([1] * (w.dim() - 1))), (b - mu) * s + beta


class SAM2Classifier:
    """Mask classifier on SAM2 image embeddings. ``head_sd``: state-dict with the reference's keys
    (``projection.{0,1,2,4,5,6,9,10,11}.*``, ``classifier.{0,1,2,4}.*``)."""

    def __init__(self, num_classes: int, backbone_type: str = "large", hidden_dims: int = 256, fuse_features: bool = False,
                 deviceID: int = 0, head_sd: Optional[Dict[str, torch.Tensor]] = None, sam_model=None, seed: int = 0,
                 sam2_checkpoint: Optional[str] = None, allow_random_init: bool = False):
        """The classifier checkpoint only carries ``projection.*`` / ``classifier.*``; the frozen backbone comes from
        the pretrained SAM2.1 checkpoint (REF SAM2.py:27-51 resolves it with pretrained_weights.get_sam2_checkpoint):
        ``sam2_checkpoint`` or the resolver of saber_b200.pretrained_weights; a missing file raises (a trained head on
        a random backbone would classify noise) unless ``allow_random_init`` is set (synthetic parity tests)."""
        if fuse_features:
            raise NotImplementedError("fuse_features=True is commented out in the reference's forward (REF SAM2.py:154-159)")
        self.name = self.__class__.__name__
        self.input_mode = "separate"
        self.num_classes = num_classes
        self.device = torch.device(f"cuda:{deviceID}")
        if sam_model is None:
            sam_model = build_sam2(_CFG_TO_ARCH.get(backbone_type, backbone_type), sam2_checkpoint, device=self.device,
                                   apply_postprocessing=True, seed=seed, allow_random_init=allow_random_init)
        self.backbone = SAM2ImagePredictor(sam_model)
        if head_sd is None:
            raise ValueError("saber_b200 SAM2Classifier needs the trained head weights (head_sd)")
        self._load_head({k: v.detach().float().cpu() for k, v in head_sd.items()})

    def _load_head(self, sd):
        dev = self.device
        w16 = lambda t: t.to(dev, _BF16).contiguous()
        f32 = lambda t: t.to(dev, _F32).contiguous()
        w, b = _fold_bn(sd["projection.0.weight"].reshape(-1, 512), sd["projection.0.bias"], sd, "projection.1")
        self.c1 = (w16(w), f32(b), float(sd["projection.2.weight"].reshape(-1)[0]))
        w = sd["projection.4.weight"]  # [Cout, Cin, 3, 3] -> [Cout, (ky, kx, ci)]
        w, b = _fold_bn(w, sd["projection.4.bias"], sd, "projection.5")
        self.c2 = (w16(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)), f32(b), float(sd["projection.6.weight"].reshape(-1)[0]))
        w = sd["projection.9.weight"]
        w, b = _fold_bn(w, sd["projection.9.bias"], sd, "projection.10")
        self.c3 = (w16(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)), f32(b), float(sd["projection.11.weight"].reshape(-1)[0]))
        self.fc1 = (w16(sd["classifier.0.weight"]), f32(sd["classifier.0.bias"]))
        self.ln = (f32(sd["classifier.1.weight"]), f32(sd["classifier.1.bias"]))
        self.fc_slope = float(sd["classifier.2.weight"].reshape(-1)[0])
        self.fc2 = (w16(sd["classifier.4.weight"]), f32(sd["classifier.4.bias"]))

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, S, S] fp32 CUDA (grayscale crops, replicated to RGB by the transform) -> image_embed token-major
        [B*4096, 256] fp32 (SAM2ImagePredictor.set_image_batch + _features["image_embed"], batched)."""
        B, S, _ = x.shape
        tall = x.reshape(B * S, S).contiguous()
        crops = torch.tensor([[0, b * S, S, (b + 1) * S] for b in range(B)], dtype=_I32, device=x.device)
        return self.backbone.encode_crops(tall, crops).tok["embed"]

    def apply_mask_to_features(self, feat_tok: torch.Tensor, mask_u8: torch.Tensor) -> torch.Tensor:
        return ops.mask_features(feat_tok, mask_u8, mask_u8.shape[0], 64)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """x [B,1,S,S] fp32, mask [B,1,S,S] (0/1) -> logits [B, num_classes] fp32 (CUDA)."""
        B = x.shape[0]
        xs = x[:, 0].to(self.device, _F32).contiguous()
        m = (mask[:, 0] != 0).to(self.device, _U8).contiguous()
        feat = self.encode(xs)
        f = self.apply_mask_to_features(feat, m)  # [B*4096, 512] bf16
        w, b, a = self.c1
        f = ops.prelu(ops.gemm(f, w, b, out_dtype=_F32), a)  # 1x1 conv + BN + PReLU  -> [B*4096,256]
        w, b, a = self.c2
        f = ops.prelu(ops.gemm(ops.im2col_3x3s1(f.view(B, 64, 64, 256)), w, b, out_dtype=_F32), a)
        f = ops.maxpool2x2(f, B, 64, 64)  # [B*1024, 256]
        w, b, a = self.c3
        f = ops.prelu(ops.gemm(ops.im2col_3x3s1(f.view(B, 32, 32, 256)), w, b, out_dtype=_F32), a)
        f = ops.maxpool2x2(f, B, 32, 32)  # [B*256, 128]
        v = ops.mean_tokens(f, B)  # adaptive_avg_pool2d -> [B,128]
        h = ops.gemm(ops.add_cast(v, None, _BF16), *self.fc1, out_dtype=_F32)
        h = ops.layernorm(h, self.ln[0], self.ln[1], 1e-5, _F32)
        h = ops.prelu(h, self.fc_slope)
        return ops.gemm(h, *self.fc2, out_dtype=_F32)

    __call__ = forward


class Predictor:
    def __init__(self, model_config=None, model_weights=None, min_area: int = 250, deviceID: int = 0, model=None,
                 num_classes: Optional[int] = None, sam2_checkpoint: Optional[str] = None,
                 allow_random_init: bool = False):
        """Reference signature ``Predictor(model_config.yaml, model_weights.pth, deviceID=)``; alternatively pass a built
        ``model`` (and ``num_classes``) directly. As in the reference the backbone is ALWAYS hiera-large:
        ``get_classifier_model('SAM2', ...)`` drops ``model_size`` (REF classifier/models/common.py:15-17, SAM2.py:27), so
        the head was trained on large embeddings whatever ``amg_params.sam2_cfg`` says."""
        self.min_area = min_area
        self.device = torch.device(f"cuda:{deviceID}")
        if model is None:
            import yaml
            with open(model_config, "r") as f:
                self.config = yaml.safe_load(f)
            ck = torch.load(model_weights, map_location="cpu", weights_only=True)
            sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
            model = SAM2Classifier(self.config["model"]["num_classes"], "large", deviceID=deviceID, head_sd=sd,
                                   sam2_checkpoint=sam2_checkpoint, allow_random_init=allow_random_init)
        else:
            self.config = {"model": {"num_classes": num_classes if num_classes is not None else model.num_classes}}
        self.model = model
        self.output_size = 320

    # ---- REF predictor.py:208-234 + RandMaskCrop.py:44-180 ---------------------------------------------------------
    def apply_crops(self, image: torch.Tensor, masks: torch.Tensor):
        """image [H,W] fp32 CUDA (already normalised), masks uint8 [N,H,W] CUDA -> (crops [N,S,S] fp32, masks [N,S,S]
        uint8, areas int32 [N]). One D2H copy of the N bounding boxes; the crop geometry is the reference's integer
        arithmetic."""
        H, W = image.shape
        bbox = ops.mask_bbox(masks).cpu().numpy()
        geom = np.zeros((masks.shape[0], 4), dtype=np.int32)
        for n, (y0, y1, x0, x1) in enumerate(bbox):
            full = (0, 0, H, W)
            if y0 < 0:
                geom[n] = full
                continue
            bh, bw = max(1, int(y1 - y0)), max(1, int(x1 - x0))
            if (bh / H) >= 0.9 and (bw / W) >= 0.9:
                geom[n] = full
                continue
            ch, cw = int(bh * (1 + 1.5)), int(bw * (1 + 1.5))
            cy, cx = (int(y0) + int(y1)) // 2, (int(x0) + int(x1)) // 2
            top, left = cy - ch // 2, cx - cw // 2
            top = max(0, min(top, H - ch))
            left = max(0, min(left, W - cw))
            geom[n] = (top, left, min(ch, H), min(cw, W))
        g = torch.from_numpy(geom).to(image.device)
        return ops.crop_resize(image, masks, g, self.output_size)

    def preprocess(self, images: torch.Tensor, masks: torch.Tensor, areas: torch.Tensor):
        """REF :62-113: drop crops whose (cropped) mask area < min_area."""
        valid = (areas.cpu().numpy() >= self.min_area).nonzero()[0].tolist()
        if not valid:
            return None, None, []
        idx = torch.tensor(valid, dtype=torch.long, device=images.device)
        return images[idx].contiguous(), masks[idx].contiguous(), valid

    @torch.inference_mode()
    def predict(self, image, masks) -> np.ndarray:
        nc = self.config["model"]["num_classes"]
        img = (image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)))
        img = img.to(self.device, _F32).contiguous()
        m = (masks if isinstance(masks, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(masks)))
        m = (m.to(self.device) != 0).to(_U8).contiguous()
        n = m.shape[0]
        img = ops.standardize(img)  # monai NormalizeIntensity
        ci, cm, areas = self.apply_crops(img, m)
        xi, xm, valid = self.preprocess(ci, cm, areas)
        if not valid:
            return np.zeros((n, nc), dtype=np.float32)
        logits = self.model(xi[:, None], xm[:, None])
        probs = ops.softmax_rows(logits.contiguous()).cpu().numpy()
        full = np.zeros((n, probs.shape[1]), dtype=np.float32)
        full[valid] = probs
        return full

    @torch.inference_mode()
    def batch_predict(self, image, masks, batch_size: int = 32) -> np.ndarray:
        nc = self.config["model"]["num_classes"]
        total = masks.shape[0]
        out = np.zeros((total, nc), dtype=np.float32)
        for s in range(0, total, batch_size):
            e = min(s + batch_size, total)
            out[s:e] = self.predict(image, masks[s:e])
        return out


def get_predictor(model_weights, model_config, deviceID: int = 0, sam2_checkpoint: Optional[str] = None,
                  allow_random_init: bool = False):
    """REF saber/classifier/models/common.py:24-50."""
    if model_weights is None or model_config is None:
        return None
    if not os.path.exists(model_weights):
        raise FileNotFoundError(f"Model weights file {model_weights} does not exist.")
    if not os.path.exists(model_config):
        raise FileNotFoundError(f"Model config file {model_config} does not exist.")
    return Predictor(model_config, model_weights, deviceID=deviceID, sam2_checkpoint=sam2_checkpoint,
                     allow_random_init=allow_random_init)
