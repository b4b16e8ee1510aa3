"""ORACLE (test infrastructure only). Restatement of upstream sam2 memory attention (RoPE) and memory
encoder: sam2/modeling/{memory_attention,memory_encoder}.py, sam2/modeling/sam/transformer.py
(RoPEAttention) and sam2/modeling/position_encoding.py (axial RoPE). Reached from
REF saber/adapters/sam2/predictor.py:196-202 (propagate_in_video). SURVEY §8a U7/U8.
Pinned against HF transformers Sam2VideoModel modules in tests/test_oracle_vs_hf.py.
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modeling import Attention, LayerNorm2d


def init_t_xy(end_x, end_y):
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    t_x = (t % end_x).float()
    t_y = torch.div(t, end_x, rounding_mode="floor").float()
    return t_x, t_y


def compute_axial_cis(dim, end_x, end_y, theta=10000.0):
    freqs_x = 1.0 / (theta ** (torch.arange(0, dim, 4)[: (dim // 4)].float() / dim))
    freqs_y = 1.0 / (theta ** (torch.arange(0, dim, 4)[: (dim // 4)].float() / dim))
    t_x, t_y = init_t_xy(end_x, end_y)
    freqs_x = torch.outer(t_x, freqs_x)
    freqs_y = torch.outer(t_y, freqs_y)
    freqs_cis_x = torch.polar(torch.ones_like(freqs_x), freqs_x)
    freqs_cis_y = torch.polar(torch.ones_like(freqs_y), freqs_y)
    return torch.cat([freqs_cis_x, freqs_cis_y], dim=-1)


def reshape_for_broadcast(freqs_cis, x):
    ndim = x.ndim
    assert freqs_cis.shape == (x.shape[-2], x.shape[-1])
    shape = [d if i >= ndim - 2 else 1 for i, d in enumerate(x.shape)]
    return freqs_cis.view(*shape)


def apply_rotary_enc(xq, xk, freqs_cis, repeat_freqs_k=False):
    xq_ = torch.view_as_complex(xq.float().reshape(*xq.shape[:-1], -1, 2))
    xk_ = torch.view_as_complex(xk.float().reshape(*xk.shape[:-1], -1, 2)) if xk.shape[-2] != 0 else None
    freqs_cis = reshape_for_broadcast(freqs_cis, xq_)
    xq_out = torch.view_as_real(xq_ * freqs_cis).flatten(3)
    if xk_ is None:
        return xq_out.type_as(xq), xk
    if repeat_freqs_k:
        r = xk_.shape[-2] // xq_.shape[-2]
        freqs_cis = freqs_cis.repeat(*([1] * (freqs_cis.ndim - 2)), r, 1)
    xk_out = torch.view_as_real(xk_ * freqs_cis).flatten(3)
    return xq_out.type_as(xq), xk_out.type_as(xk)


class RoPEAttention(Attention):
    def __init__(self, *args, rope_theta=10000.0, rope_k_repeat=False, feat_sizes=(64, 64), **kwargs):
        super().__init__(*args, **kwargs)
        self.rope_theta = rope_theta
        self.freqs_cis = compute_axial_cis(self.internal_dim // self.num_heads, feat_sizes[0], feat_sizes[1], rope_theta)
        self.rope_k_repeat = rope_k_repeat

    def forward(self, q, k, v, num_k_exclude_rope=0):
        q, k, v = self.q_proj(q), self.k_proj(k), self.v_proj(v)
        q = self._separate_heads(q, self.num_heads)
        k = self._separate_heads(k, self.num_heads)
        v = self._separate_heads(v, self.num_heads)
        w = h = math.sqrt(q.shape[-2])
        self.freqs_cis = self.freqs_cis.to(q.device)
        if self.freqs_cis.shape[0] != q.shape[-2]:
            self.freqs_cis = compute_axial_cis(self.internal_dim // self.num_heads, int(w), int(h), self.rope_theta).to(q.device)
        if q.shape[-2] != k.shape[-2]:
            assert self.rope_k_repeat
        num_k_rope = k.size(-2) - num_k_exclude_rope
        k = k.clone()
        q, k[:, :, :num_k_rope] = apply_rotary_enc(q, k[:, :, :num_k_rope], freqs_cis=self.freqs_cis,
                                                   repeat_freqs_k=self.rope_k_repeat)
        out = F.scaled_dot_product_attention(q, k, v)
        return self.out_proj(self._recombine_heads(out))


class MemoryAttentionLayer(nn.Module):
    def __init__(self, d_model=256, dim_feedforward=2048, pos_enc_at_attn=False,
                 pos_enc_at_cross_attn_keys=True, pos_enc_at_cross_attn_queries=False):
        super().__init__()
        self.d_model = d_model
        self.self_attn = RoPEAttention(embedding_dim=256, num_heads=1, downsample_rate=1, rope_theta=10000.0,
                                       feat_sizes=(64, 64))
        self.cross_attn_image = RoPEAttention(embedding_dim=256, num_heads=1, downsample_rate=1, rope_theta=10000.0,
                                              feat_sizes=(64, 64), rope_k_repeat=True, kv_in_dim=64)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.pos_enc_at_attn = pos_enc_at_attn
        self.pos_enc_at_cross_attn_queries = pos_enc_at_cross_attn_queries
        self.pos_enc_at_cross_attn_keys = pos_enc_at_cross_attn_keys

    def forward(self, tgt, memory, pos=None, query_pos=None, num_k_exclude_rope=0):
        tgt2 = self.norm1(tgt)
        q = k = tgt2 + query_pos if self.pos_enc_at_attn else tgt2
        tgt = tgt + self.self_attn(q, k, v=tgt2)
        tgt2 = self.norm2(tgt)
        tgt2 = self.cross_attn_image(
            q=tgt2 + query_pos if self.pos_enc_at_cross_attn_queries else tgt2,
            k=memory + pos if self.pos_enc_at_cross_attn_keys else memory, v=memory,
            num_k_exclude_rope=num_k_exclude_rope)
        tgt = tgt + tgt2
        tgt2 = self.norm3(tgt)
        tgt = tgt + self.linear2(F.relu(self.linear1(tgt2)))
        return tgt


class MemoryAttention(nn.Module):
    def __init__(self, d_model=256, pos_enc_at_input=True, num_layers=4):
        super().__init__()
        self.d_model = d_model
        self.layers = nn.ModuleList(MemoryAttentionLayer(d_model) for _ in range(num_layers))
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model)
        self.pos_enc_at_input = pos_enc_at_input

    def forward(self, curr, memory, curr_pos=None, memory_pos=None, num_obj_ptr_tokens=0):
        if isinstance(curr, list):
            assert len(curr) == 1 and len(curr_pos) == 1
            curr, curr_pos = curr[0], curr_pos[0]
        output = curr
        if self.pos_enc_at_input and curr_pos is not None:
            output = output + 0.1 * curr_pos
        output = output.transpose(0, 1)
        curr_pos = curr_pos.transpose(0, 1)
        memory = memory.transpose(0, 1)
        memory_pos = memory_pos.transpose(0, 1)
        for layer in self.layers:
            output = layer(tgt=output, memory=memory, pos=memory_pos, query_pos=curr_pos,
                           num_k_exclude_rope=num_obj_ptr_tokens)
        return self.norm(output).transpose(0, 1)


class MaskDownSampler(nn.Module):
    def __init__(self, embed_dim=256, kernel_size=4, stride=4, padding=0, total_stride=16, activation=nn.GELU):
        super().__init__()
        num_layers = int(math.log2(total_stride) // math.log2(stride))
        assert stride ** num_layers == total_stride
        self.encoder = nn.Sequential()
        cin, cout = 1, 1
        for _ in range(num_layers):
            cout = cin * (stride ** 2)
            self.encoder.append(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=padding))
            self.encoder.append(LayerNorm2d(cout))
            self.encoder.append(activation())
            cin = cout
        self.encoder.append(nn.Conv2d(cout, embed_dim, kernel_size=1))

    def forward(self, x):
        return self.encoder(x)


class CXBlock(nn.Module):
    def __init__(self, dim, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=kernel_size, padding=padding, groups=dim if use_dwconv else 1)
        self.norm = LayerNorm2d(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True)
                      if layer_scale_init_value > 0 else None)

    def forward(self, x):
        inp = x
        x = self.norm(self.dwconv(x)).permute(0, 2, 3, 1)
        x = self.pwconv2(self.act(self.pwconv1(x)))
        if self.gamma is not None:
            x = self.gamma * x
        return inp + x.permute(0, 3, 1, 2)


class Fuser(nn.Module):
    def __init__(self, layer, num_layers):
        super().__init__()
        self.proj = nn.Identity()
        self.layers = nn.ModuleList(copy.deepcopy(layer) for _ in range(num_layers))

    def forward(self, x):
        x = self.proj(x)
        for layer in self.layers:
            x = layer(x)
        return x


class MemoryEncoder(nn.Module):
    def __init__(self, out_dim, mask_downsampler, fuser, position_encoding, in_dim=256):
        super().__init__()
        self.mask_downsampler = mask_downsampler
        self.pix_feat_proj = nn.Conv2d(in_dim, in_dim, kernel_size=1)
        self.fuser = fuser
        self.position_encoding = position_encoding
        self.out_proj = nn.Identity()
        if out_dim != in_dim:
            self.out_proj = nn.Conv2d(in_dim, out_dim, kernel_size=1)

    def forward(self, pix_feat, masks, skip_mask_sigmoid=False):
        if not skip_mask_sigmoid:
            masks = torch.sigmoid(masks)
        masks = self.mask_downsampler(masks)
        x = self.pix_feat_proj(pix_feat) + masks
        x = self.out_proj(self.fuser(x))
        pos = self.position_encoding(x).to(x.dtype)
        return {"vision_features": x, "vision_pos_enc": [pos]}
