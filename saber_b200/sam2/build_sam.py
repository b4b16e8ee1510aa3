"""Drop-in for ``sam2.build_sam`` (the seam SABER binds: REF saber/adapters/sam2/automask.py:55-62,
REF saber/adapters/sam2/predictor.py:4,24-26, REF saber/classifier/models/SAM2.py:14,45).

``build_sam2`` / ``build_sam2_video_predictor`` return objects backed by the B200 kernels. A real upstream
``sam2.1_hiera_*.pt`` loads by name (``ckpt_path``, or resolved by ``saber_b200.pretrained_weights`` as the reference
resolves its own, REF saber/pretrained_weights.py:174-202). Without a checkpoint construction RAISES unless random
initialisation of the named architecture (deterministic, ``seed``) is requested explicitly — ``state_dict=``,
``allow_random_init=True`` or ``SABER_B200_ALLOW_RANDOM_INIT=1`` (synthetic benchmarks / parity tests).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import ops
from . import arch
from .decoder import MaskDecoder
from .encoder import HieraEncoder


class SAM2Model(nn.Module):
    """Owner of the SAM2.1 weights (upstream names) and of the B200 executors built from them."""

    def __init__(self, cfg: str, state_dict: Dict[str, torch.Tensor], device="cuda",
                 dynamic_multimask_via_stability: bool = False, num_maskmem: int = 7):
        super().__init__()
        self.cfg = arch.resolve(cfg)
        self.image_size = arch.IMAGE_SIZE
        self.hidden_dim = arch.HIDDEN
        self.mem_dim = arch.MEM_DIM
        self.num_maskmem = num_maskmem
        self.directly_add_no_mem_embed = True
        self.dynamic_multimask_via_stability = dynamic_multimask_via_stability
        expected = arch.param_shapes(self.cfg, num_maskmem)
        missing = [k for k in expected if k not in state_dict]
        if missing:
            raise RuntimeError(f"SAM2 state-dict is missing {len(missing)} tensors, e.g. {missing[:4]}")
        self._names = {}
        for k, (shape, _) in expected.items():
            t = state_dict[k]
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"SAM2 state-dict tensor {k} has shape {tuple(t.shape)}, expected {shape}")
            pname = k.replace(".", "__")
            self._names[pname] = k
            self.register_parameter(pname, nn.Parameter(t.detach().clone().float(), requires_grad=False))
        self._device = None
        self.encoder: Optional[HieraEncoder] = None
        self.decoder: Optional[MaskDecoder] = None
        self.to(device)

    # -- nn.Module plumbing ---------------------------------------------------------------
    def upstream_state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, p).detach() for p, k in self._names.items()}

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        dev = next(self.parameters()).device
        if dev.type == "cuda" and dev != self._device:
            self._build_executors(dev)
        elif dev.type != "cuda":
            self.encoder = self.decoder = None
            self._device = dev
        return out

    def _build_executors(self, dev):
        ops.require_b200()
        sd = {k: v.cpu() for k, v in self.upstream_state_dict().items()}  # weight prep runs on the host
        with torch.cuda.device(dev):
            self.encoder = HieraEncoder(sd, self.cfg, dev)
            self.decoder = MaskDecoder(sd, dev, self.dynamic_multimask_via_stability)
            self.no_mem_embed_vec = sd["no_mem_embed"].reshape(-1).to(dev, torch.float32).contiguous()
        self._device = dev

    @property
    def device(self):
        return next(self.parameters()).device

    def _require_gpu(self):
        if self.encoder is None:
            raise RuntimeError("saber_b200 SAM2 model is not on a CUDA device; there is no CPU path — "
                               "move it with .to('cuda') on a B200")

    # -- image side -----------------------------------------------------------------------
    @torch.no_grad()
    def forward_image(self, img_batch: torch.Tensor):
        """img_batch [B,3,1024,1024] fp32 CUDA -> token-major features (see HieraEncoder.forward)."""
        self._require_gpu()
        return self.encoder.forward(img_batch)


def _load_state_dict(cfg: str, ckpt_path: Optional[str], seed: int, num_maskmem: int = 7,
                     allow_random_init: bool = False):
    from .. import pretrained_weights as pw
    if ckpt_path is None:
        ckpt_path = pw.find_sam2_checkpoint(cfg)
    if ckpt_path is None:
        if not pw.random_init_allowed(allow_random_init):
            raise FileNotFoundError("saber_b200: " + pw.missing_message(cfg))
        return arch.random_state_dict(cfg, seed=seed, num_maskmem=num_maskmem)
    ck = torch.load(ckpt_path, map_location="cpu", weights_only=True)
    return ck["model"] if "model" in ck else ck


def build_sam2(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=None,
               apply_postprocessing=True, seed: int = 0, state_dict=None, allow_random_init: bool = False,
               **kwargs) -> SAM2Model:
    """Same call shape as upstream ``sam2.build_sam.build_sam2``. ``apply_postprocessing`` switches on
    dynamic multimask via stability (delta 0.05, thresh 0.98) in the mask decoder, as upstream does."""
    cfg = arch.resolve(config_file)
    sd = state_dict if state_dict is not None else _load_state_dict(cfg, ckpt_path, seed,
                                                                    allow_random_init=allow_random_init)
    model = SAM2Model(cfg, sd, device=device, dynamic_multimask_via_stability=bool(apply_postprocessing))
    if mode == "eval":
        model.eval()
    return model


def build_sam2_video_predictor(*args, **kwargs):
    """``sam2.build_sam.build_sam2_video_predictor`` (REF saber/adapters/sam2/predictor.py:4,24-26); implemented in
    ``sam2_video_predictor.py`` (imported lazily: that module subclasses SAM2Model from this one)."""
    from .sam2_video_predictor import build_sam2_video_predictor as _b
    return _b(*args, **kwargs)
