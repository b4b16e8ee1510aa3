"""B200 executors of SAM2's memory attention (4 RoPE transformer layers that condition the current frame on the memory
bank) and memory encoder (mask down-sampler + CXBlock fuser), batched over the objects tracked in one frame.

Token-major layout: current-frame stream ``[B*4096, 256]`` (fp32 residual), memory bank ``[B*Nk, 64]`` bf16 where
``Nk = 4096 * n_spatial_memories + 4 * n_object_pointers``. Folded at load time (results equal upstream up to fp
re-association):
  * the cross-attention key positional term: ``k_proj(memory + pos) = memory @ Wk^T + (pos @ Wk^T + bk)``; the second
    term only depends on weights and on which temporal slot a memory frame occupies, so it is the K GEMM's broadcast
    residual;
  * ``+ 0.1 * curr_pos`` (sine encoding of the 64x64 grid) is a constant added once;
  * CXBlock's layer scale ``gamma`` is folded into ``pwconv2``;
  * layer 0's self-attention and cross-attention queries do not depend on the object and are computed once per frame.
Restates sam2/modeling/memory_attention.py, memory_encoder.py, sam/transformer.py::RoPEAttention and
position_encoding.py; reached from REF saber/adapters/sam2/predictor.py:196-202.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import ops

_BF16, _F32 = torch.bfloat16, torch.float32
NT = 4096  # 64 x 64 tokens


def sine_pe_2d(num_pos_feats: int, size: int = 64, temperature: float = 10000.0) -> torch.Tensor:
    """PositionEmbeddingSine(num_pos_feats, normalize=True)(x) for an [*, *, size, size] map -> token-major
    [size*size, num_pos_feats] fp32 (weights-free constant, evaluated once on the host)."""
    npf = num_pos_feats // 2
    scale = 2 * math.pi
    y = torch.arange(1, size + 1, dtype=torch.float32).view(-1, 1).repeat(1, size)
    x = torch.arange(1, size + 1, dtype=torch.float32).view(1, -1).repeat(size, 1)
    eps = 1e-6
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(size * size, num_pos_feats).contiguous()


def axial_rope_table(dim: int = 256, size: int = 64, theta: float = 10000.0) -> torch.Tensor:
    """compute_axial_cis(dim, size, size) as [size*size, dim/2, 2] (cos, sin) fp32."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: (dim // 4)].float() / dim))
    t = torch.arange(size * size, dtype=torch.float32)
    tx = (t % size).float()
    ty = torch.div(t, size, rounding_mode="floor").float()
    fx = torch.outer(tx, freqs)
    fy = torch.outer(ty, freqs)
    ang = torch.cat([fx, fy], dim=-1)  # [ntok, dim/2]
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


def sine_pe_1d(pos: torch.Tensor, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """sam2_utils.get_1d_sine_pe."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    e = pos.unsqueeze(-1) / dim_t
    return torch.cat([e.sin(), e.cos()], dim=-1)


class MemoryAttention:
    def __init__(self, sd: Dict[str, torch.Tensor], device):
        dev = self.device = torch.device(device)

        def w16(t):  # bf16 GEMM operand (fp32 in the validation mode)
            return ops.weight(t, dev)

        def f32(t):
            return t.to(dev, _F32).contiguous()

        self.sd_cpu = {k: v.float() for k, v in sd.items() if k.startswith("memory_attention.")}
        self.layers = []
        for l in range(4):
            b = f"memory_attention.layers.{l}."
            g = lambda n: sd[b + n].float()
            L = dict(
                n1w=f32(g("norm1.weight")), n1b=f32(g("norm1.bias")), n2w=f32(g("norm2.weight")), n2b=f32(g("norm2.bias")),
                n3w=f32(g("norm3.weight")), n3b=f32(g("norm3.bias")),
                sa_qk_w=w16(torch.cat([g("self_attn.q_proj.weight"), g("self_attn.k_proj.weight")], 0)),
                sa_qk_b=f32(torch.cat([g("self_attn.q_proj.bias"), g("self_attn.k_proj.bias")], 0)),
                sa_v_w=w16(g("self_attn.v_proj.weight")), sa_v_b=f32(g("self_attn.v_proj.bias")),
                sa_o_w=w16(g("self_attn.out_proj.weight")), sa_o_b=f32(g("self_attn.out_proj.bias")),
                ca_q_w=w16(g("cross_attn_image.q_proj.weight")), ca_q_b=f32(g("cross_attn_image.q_proj.bias")),
                ca_k_w=w16(g("cross_attn_image.k_proj.weight")),
                ca_v_w=w16(g("cross_attn_image.v_proj.weight")), ca_v_b=f32(g("cross_attn_image.v_proj.bias")),
                ca_o_w=w16(g("cross_attn_image.out_proj.weight")), ca_o_b=f32(g("cross_attn_image.out_proj.bias")),
                l1_w=w16(g("linear1.weight")), l1_b=f32(g("linear1.bias")),
                l2_w=w16(g("linear2.weight")), l2_b=f32(g("linear2.bias")),
            )
            self.layers.append(L)
        self.nw, self.nb = f32(sd["memory_attention.norm.weight"]), f32(sd["memory_attention.norm.bias"])
        self.rope = axial_rope_table(256, 64).to(dev)
        self.pos01 = (0.1 * sine_pe_2d(256, 64)).to(dev).contiguous()  # 0.1 * curr_pos (pos_enc_at_input)

    def key_pos_term(self, pos: torch.Tensor) -> List[torch.Tensor]:
        """pos [n, 64] fp32 (host) -> per layer [n, 256] fp32 device: pos @ Wk^T + bk (weights-only constant)."""
        out = []
        for l in range(4):
            b = f"memory_attention.layers.{l}.cross_attn_image.k_proj."
            out.append((pos.float() @ self.sd_cpu[b + "weight"].t() + self.sd_cpu[b + "bias"]).to(self.device).contiguous())
        return out

    def forward(self, curr: torch.Tensor, memory: torch.Tensor, pos_k: Sequence[torch.Tensor], n_ptr_tokens: int,
                B: int) -> torch.Tensor:
        """curr [4096,256] fp32: this frame's (unconditioned) vision features, shared by the B objects.
        memory [B*Nk, 64] bf16; pos_k[l] [Nk,256] fp32: key positional term of layer l. Returns [B*4096,256] fp32."""
        assert curr.shape == (NT, 256) and memory.dtype == _BF16 and memory.shape[0] % B == 0
        if ops.VALIDATE_FP32:  # the bank stores bf16 (as upstream): widening it is exact
            memory = ops.add_cast(memory.contiguous(), None, _F32)
        Nk = memory.shape[0] // B
        n_rope = Nk - n_ptr_tokens
        tgt = ops.add_cast(curr, self.pos01, _F32)  # [4096,256], shared by all objects until the first cross-attention
        nb = 1  # batch entries of tgt
        for l, L in enumerate(self.layers):
            # ---- RoPE self-attention (1 head x 256)
            t2 = ops.layernorm(tgt, L["n1w"], L["n1b"], 1e-5, _BF16)
            qk = ops.gemm(t2, L["sa_qk_w"], L["sa_qk_b"], out_dtype=_F32)
            v = ops.gemm(t2, L["sa_v_w"], L["sa_v_b"])
            q = ops.rope_apply(qk[:, 0:256], self.rope, NT)
            k = ops.rope_apply(qk[:, 256:512], self.rope, NT)
            a = ops.attention(q, k, v, nb, 1, NT, NT)
            tgt = ops.gemm(a, L["sa_o_w"], L["sa_o_b"], residual=tgt, out_dtype=_F32)
            # ---- RoPE cross-attention to the memory bank (keys carry pos, object-pointer tokens are not rotated)
            t2 = ops.layernorm(tgt, L["n2w"], L["n2b"], 1e-5, _BF16)
            q = ops.rope_apply(ops.gemm(t2, L["ca_q_w"], L["ca_q_b"], out_dtype=_F32), self.rope, NT)
            kf = ops.gemm(memory, L["ca_k_w"], None, residual=pos_k[l], res_mod=Nk, out_dtype=_F32)
            k = ops.rope_apply(kf, self.rope, Nk, n_rope=n_rope)
            del kf
            v = ops.gemm(memory, L["ca_v_w"], L["ca_v_b"])
            a = ops.attention(q, k, v, B, 1, NT, Nk, q_shared=(nb == 1 and B > 1))
            tgt = ops.gemm(a, L["ca_o_w"], L["ca_o_b"], residual=tgt, res_mod=(NT if nb == 1 and B > 1 else 0),
                           out_dtype=_F32)
            nb = B
            # ---- MLP
            t2 = ops.layernorm(tgt, L["n3w"], L["n3b"], 1e-5, _BF16)
            h = ops.gemm(t2, L["l1_w"], L["l1_b"], act=ops.ACT_RELU)
            tgt = ops.gemm(h, L["l2_w"], L["l2_b"], residual=tgt, out_dtype=_F32)
        return ops.layernorm(tgt, self.nw, self.nb, 1e-5, _F32)


class MemoryEncoder:
    def __init__(self, sd: Dict[str, torch.Tensor], device):
        dev = self.device = torch.device(device)

        def w16(t):
            return t.to(dev, _BF16).contiguous()

        def f32(t):
            return t.float().to(dev).contiguous()

        me = "memory_encoder."
        self.stages = []
        for k in range(3):
            p = f"{me}mask_downsampler.encoder."
            self.stages.append((f32(sd[p + f"{3 * k}.weight"]), f32(sd[p + f"{3 * k}.bias"]),
                                f32(sd[p + f"{3 * k + 1}.weight"]), f32(sd[p + f"{3 * k + 1}.bias"])))
        p = f"{me}mask_downsampler.encoder."
        self.c4_w = w16(sd[p + "9.weight"].float().permute(0, 2, 3, 1).reshape(256, 9 * 64))  # column = (ky*3+kx)*64 + ci
        self.c4_b = f32(sd[p + "9.bias"])
        self.ln4_w, self.ln4_b = f32(sd[p + "10.weight"]), f32(sd[p + "10.bias"])
        self.c5_w = w16(sd[p + "12.weight"].reshape(256, 256))
        self.c5_b = f32(sd[p + "12.bias"])
        self.pp_w = w16(sd[me + "pix_feat_proj.weight"].reshape(256, 256))
        self.pp_b = f32(sd[me + "pix_feat_proj.bias"])
        self.cx = []
        for l in range(2):
            b = f"{me}fuser.layers.{l}."
            gamma = sd[b + "gamma"].float()
            self.cx.append(dict(
                dw_w=f32(sd[b + "dwconv.weight"].reshape(256, 49)), dw_b=f32(sd[b + "dwconv.bias"]),
                n_w=f32(sd[b + "norm.weight"]), n_b=f32(sd[b + "norm.bias"]),
                p1_w=w16(sd[b + "pwconv1.weight"]), p1_b=f32(sd[b + "pwconv1.bias"]),
                p2_w=w16(gamma[:, None] * sd[b + "pwconv2.weight"].float()), p2_b=f32(gamma * sd[b + "pwconv2.bias"].float())))
        self.out_w = w16(sd[me + "out_proj.weight"].reshape(64, 256))
        self.out_b = f32(sd[me + "out_proj.bias"])
        self.no_obj_embed_spatial = f32(sd["no_obj_embed_spatial"].reshape(-1))
        self.pos = sine_pe_2d(64, 64)  # maskmem_pos_enc, token-major [4096,64] (host; consumed by key_pos_term)

    def project_pix(self, pix_feat: torch.Tensor) -> torch.Tensor:
        """pix_feat_proj of one frame's raw vision features [4096,256] fp32 (shared by all objects of the frame)."""
        return ops.gemm(ops.add_cast(pix_feat, None, _BF16), self.pp_w, self.pp_b, out_dtype=_F32)

    def forward(self, pix_proj: torch.Tensor, low_res_masks: torch.Tensor, obj_scores: torch.Tensor,
                binarize: bool) -> torch.Tensor:
        """pix_proj [4096,256] fp32 (project_pix), low_res_masks [B,256,256] fp32 mask logits, obj_scores [B] fp32.
        SAM2Base._encode_new_memory: bilinear x4 -> sigmoid (or > 0) * 20 - 10 -> memory encoder -> occlusion embedding
        -> bf16. Returns maskmem_features [B*4096, 64] bf16 token-major."""
        B = low_res_masks.shape[0]
        hi = ops.upsample_bilinear(low_res_masks.contiguous(), 1024, 1024)  # [B,1024,1024]
        x = hi.view(B, 1024, 1024, 1)
        for k, (w, b, g, be) in enumerate(self.stages):
            x = ops.conv3x3s2_ln_gelu(x, k, w, b, g, be, 1e-6, in_xf=(2 if binarize else 1) if k == 0 else 0)
        del hi
        cols = ops.im2col_3x3s2(x)  # [B*4096, 576]
        h = ops.gemm(cols, self.c4_w, self.c4_b, out_dtype=_F32)
        h = ops.layernorm(h, self.ln4_w, self.ln4_b, 1e-6, _BF16, act=ops.ACT_GELU)
        x = ops.gemm(h, self.c5_w, self.c5_b, residual=pix_proj, res_mod=NT, out_dtype=_F32)  # + pix_feat_proj(pix)
        for c in self.cx:
            t = ops.dwconv7_ln(x, B, 64, 64, c["dw_w"], c["dw_b"], c["n_w"], c["n_b"], 1e-6)
            t = ops.gemm(t, c["p1_w"], c["p1_b"], act=ops.ACT_GELU)
            x = ops.gemm(t, c["p2_w"], c["p2_b"], residual=x, out_dtype=_F32)
        o = ops.gemm(ops.add_cast(x, None, _BF16), self.out_w, self.out_b, out_dtype=_F32)
        return ops.add_vec_cond(o, obj_scores.reshape(-1).contiguous(), self.no_obj_embed_spatial, B)
