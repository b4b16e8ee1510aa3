"""B200 executor of the SAM2.1 image encoder (Hiera trunk + FPN neck + conv_s0/conv_s1).

Layout: activations are token-major ``[B*H*W, C]`` (NHWC) so every Linear / 1x1 conv is a plain
row-major GEMM for the tcgen05 kernel and windows never need partition copies. The residual stream
is fp32, GEMM operands bf16, accumulation fp32. Arithmetic restated from upstream
sam2/modeling/backbones/{hieradet,image_encoder}.py; all math runs in our kernels
(``saber_b200.ops``) — torch only allocates buffers.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from .. import ops
from . import arch

_BF16, _F32 = torch.bfloat16, torch.float32


class HieraEncoder:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: str, device):
        self.cfg = arch.resolve(cfg)
        self.device = torch.device(device)
        h = arch.HIERA[self.cfg]
        self.specs, self.stage_ends = arch.block_specs(self.cfg)
        dev = self.device
        t = "image_encoder.trunk."

        def w16(name):  # bf16 operand of the tcgen05 GEMM (fp32 in the validation mode: ops.weight)
            return ops.weight(sd[name], dev)

        def f32(name):
            return sd[name].to(dev, _F32).contiguous()

        # patch embedding as an im2col GEMM: weight [E, 3*49] zero-padded to K=160
        E = h["embed_dim"]
        pw = sd[t + "patch_embed.proj.weight"].reshape(E, 147)
        self.kp = 160
        self.patch_w = torch.zeros((E, self.kp), dtype=ops.act_dtype(), device=dev)
        self.patch_w[:, :147] = pw.to(dev, ops.act_dtype())
        self.patch_b = f32(t + "patch_embed.proj.bias")
        # positional embedding for the fixed 256x256 token grid (weights-only, done once at load):
        # bicubic background + tiled window embedding (hieradet.Hiera._get_pos_embed)
        T = arch.IMAGE_SIZE // 4
        pe = F.interpolate(sd[t + "pos_embed"].float(), size=(T, T), mode="bicubic")
        we = sd[t + "pos_embed_window"].float()
        pe = pe + we.tile([x // y for x, y in zip(pe.shape, we.shape)])
        self.pos_embed = pe.permute(0, 2, 3, 1).reshape(T * T, E).to(dev, _F32).contiguous()
        self.T = T

        self.blocks: List[dict] = []
        for i, s in enumerate(self.specs):
            b = f"{t}blocks.{i}."
            blk = dict(spec=s, n1w=f32(b + "norm1.weight"), n1b=f32(b + "norm1.bias"),
                       qkv_w=w16(b + "attn.qkv.weight"), qkv_b=f32(b + "attn.qkv.bias"),
                       proj_w=w16(b + "attn.proj.weight"), proj_b=f32(b + "attn.proj.bias"),
                       n2w=f32(b + "norm2.weight"), n2b=f32(b + "norm2.bias"),
                       fc1_w=w16(b + "mlp.layers.0.weight"), fc1_b=f32(b + "mlp.layers.0.bias"),
                       fc2_w=w16(b + "mlp.layers.1.weight"), fc2_b=f32(b + "mlp.layers.1.bias"))
            if s["dim"] != s["dim_out"]:
                blk["skip_w"] = w16(b + "proj.weight")
                blk["skip_b"] = f32(b + "proj.bias")
            self.blocks.append(blk)

        n = "image_encoder.neck.convs."
        self.neck_w = [ops.weight(sd[f"{n}{i}.conv.weight"].reshape(arch.HIDDEN, -1), dev) for i in range(4)]
        self.neck_b = [f32(f"{n}{i}.conv.bias") for i in range(4)]
        md = "sam_mask_decoder."
        self.s0_w = ops.weight(sd[md + "conv_s0.weight"].reshape(32, 256), dev)
        self.s0_b = f32(md + "conv_s0.bias")
        self.s1_w = ops.weight(sd[md + "conv_s1.weight"].reshape(64, 256), dev)
        self.s1_b = f32(md + "conv_s1.bias")

    # ------------------------------------------------------------------
    def forward(self, img: torch.Tensor) -> Dict[str, torch.Tensor]:
        """img: [B, 3, 1024, 1024] fp32 CUDA (already normalised).

        Returns token-major fp32 features: ``feat`` [B*4096, 256] (neck level 2, i.e. upstream
        ``backbone_fpn[2]`` / vision_features, *before* no_mem_embed), ``s1`` [B*16384, 64] and
        ``s0`` [B*65536, 32] (conv_s1 / conv_s0 applied, as in SAM2Base.forward_image).
        """
        assert img.is_cuda and img.dtype == _F32 and img.shape[1:] == (3, arch.IMAGE_SIZE, arch.IMAGE_SIZE)
        B = img.shape[0]
        T = self.T
        cols = ops.im2col_k7s4(img.contiguous(), self.kp)
        x = ops.gemm(cols, self.patch_w, self.patch_b, residual=self.pos_embed, res_mod=T * T, out_dtype=_F32)
        del cols
        H = W = T
        stage_out = []
        for i, blk in enumerate(self.blocks):
            x, H, W = self._block(x, blk, B, H, W)
            if i in self.stage_ends:
                stage_out.append((x, H, W))
        # ---- FPN neck (top-down only into level 2) + high-res projections
        (x0, H0, W0), (x1, H1, W1), (x2, H2, W2), (x3, H3, W3) = stage_out
        lat3 = ops.gemm(ops.add_cast(x3, None, _BF16), self.neck_w[0], self.neck_b[0], out_dtype=_F32)
        feat = ops.gemm(ops.add_cast(x2, None, _BF16), self.neck_w[1], self.neck_b[1], out_dtype=_F32)
        ops.add_upsample2x_(feat, lat3, B, H2, W2)
        lat1 = ops.gemm(ops.add_cast(x1, None, _BF16), self.neck_w[2], self.neck_b[2], out_dtype=_BF16)
        s1 = ops.gemm(lat1, self.s1_w, self.s1_b, out_dtype=_F32)
        lat0 = ops.gemm(ops.add_cast(x0, None, _BF16), self.neck_w[3], self.neck_b[3], out_dtype=_BF16)
        s0 = ops.gemm(lat0, self.s0_w, self.s0_b, out_dtype=_F32)
        return {"feat": feat, "s1": s1, "s0": s0, "B": B}

    def _block(self, x, blk, B, H, W):
        s = blk["spec"]
        xn = ops.layernorm(x, blk["n1w"], blk["n1b"], 1e-6, _BF16)
        pool = 2 if s["q_pool"] else 1
        if "skip_w" in blk:
            sc = ops.gemm(xn, blk["skip_w"], blk["skip_b"], out_dtype=_F32)
            shortcut = ops.maxpool2x2(sc, B, H, W) if pool == 2 else sc
        else:
            shortcut = x
        qkv = ops.gemm(xn, blk["qkv_w"], blk["qkv_b"], out_dtype=_BF16)
        att = ops.window_attention(qkv, blk["qkv_b"], B, H, W, s["heads"], s["window"], pool)
        Ho, Wo = H // pool, W // pool
        x = ops.gemm(att, blk["proj_w"], blk["proj_b"], residual=shortcut, out_dtype=_F32)
        xn = ops.layernorm(x, blk["n2w"], blk["n2b"], 1e-6, _BF16)
        hdn = ops.gemm(xn, blk["fc1_w"], blk["fc1_b"], act=ops.ACT_GELU, out_dtype=_BF16)
        x = ops.gemm(hdn, blk["fc2_w"], blk["fc2_b"], residual=x, out_dtype=_F32)
        return x, Ho, Wo
