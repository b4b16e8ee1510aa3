"""SAM2.1 architecture tables and deterministic random initialisation (no network: REF
saber/pretrained_weights.py:20-65 downloads checkpoints; here ``ckpt_path=None`` means random-init
weights of the *named* architecture, upstream checkpoint names).

Config names follow REF saber/pretrained_weights.py:183-202 (``configs/sam2.1/sam2.1_hiera_{t,s,b+,l}.yaml``)
and SABER's short names tiny/small/base/large (REF saber/adapters/base.py:28-33).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

HIERA = {
    "tiny": dict(embed_dim=96, num_heads=1, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9),
                 bkg_size=(7, 7), window_spec=(8, 4, 14, 7)),
    "small": dict(embed_dim=96, num_heads=1, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13),
                  bkg_size=(7, 7), window_spec=(8, 4, 14, 7)),
    "base_plus": dict(embed_dim=112, num_heads=2, stages=(2, 3, 16, 3), global_att_blocks=(12, 16, 20),
                      bkg_size=(14, 14), window_spec=(8, 4, 14, 7)),
    "large": dict(embed_dim=144, num_heads=2, stages=(2, 6, 36, 4), global_att_blocks=(23, 33, 43),
                  bkg_size=(7, 7), window_spec=(8, 4, 16, 8)),
}
ALIASES = {
    "t": "tiny", "s": "small", "b+": "base_plus", "base": "base_plus", "l": "large",
    "configs/sam2.1/sam2.1_hiera_t.yaml": "tiny", "configs/sam2.1/sam2.1_hiera_s.yaml": "small",
    "configs/sam2.1/sam2.1_hiera_b+.yaml": "base_plus", "configs/sam2.1/sam2.1_hiera_l.yaml": "large",
    "sam2.1_hiera_t.yaml": "tiny", "sam2.1_hiera_s.yaml": "small", "sam2.1_hiera_b+.yaml": "base_plus",
    "sam2.1_hiera_l.yaml": "large",
}

IMAGE_SIZE = 1024
HIDDEN = 256
MEM_DIM = 64


def resolve(name: str) -> str:
    name = ALIASES.get(name, name)
    if name not in HIERA:
        raise ValueError(f"unknown SAM2.1 config {name!r} (expected one of {sorted(HIERA)} or a sam2.1 yaml name)")
    return name


def block_specs(cfg: str):
    """Per-block (dim, dim_out, heads, window_size, q_pool) following Hiera's stage logic: the first
    block of a stage still uses the previous stage's window size and pools queries 2x2."""
    h = HIERA[resolve(cfg)]
    stages = h["stages"]
    depth = sum(stages)
    stage_ends = [sum(stages[:i]) - 1 for i in range(1, len(stages) + 1)]
    q_pool_blocks = [x + 1 for x in stage_ends[:-1]][:3]
    embed_dim, num_heads, cur_stage = h["embed_dim"], h["num_heads"], 1
    specs = []
    for i in range(depth):
        dim_out = embed_dim
        window = h["window_spec"][cur_stage - 1]
        if i in h["global_att_blocks"]:
            window = 0
        if i - 1 in stage_ends:
            dim_out = embed_dim * 2
            num_heads = num_heads * 2
            cur_stage += 1
        specs.append(dict(dim=embed_dim, dim_out=dim_out, heads=num_heads, window=window,
                          q_pool=(i in q_pool_blocks)))
        embed_dim = dim_out
    return specs, stage_ends


def param_shapes(cfg: str, num_maskmem: int = 7) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, init kind). Kinds: linear_w(fan_in) / bias(fan_in) / norm_w / norm_b / embed /
    small / gauss / gamma."""
    cfg = resolve(cfg)
    h = HIERA[cfg]
    P: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def lin(name, out_f, in_f):
        P[name + ".weight"] = ((out_f, in_f), f"w:{in_f}")
        P[name + ".bias"] = ((out_f,), f"b:{in_f}")

    def conv(name, out_c, in_c, kh, kw, groups=1):
        fan = (in_c // groups) * kh * kw
        P[name + ".weight"] = ((out_c, in_c // groups, kh, kw), f"w:{fan}")
        P[name + ".bias"] = ((out_c,), f"b:{fan}")

    def convT(name, in_c, out_c, kh, kw):
        fan = out_c * kh * kw  # torch's fan_in convention for ConvTranspose2d weight [in, out, kh, kw]
        P[name + ".weight"] = ((in_c, out_c, kh, kw), f"w:{fan}")
        P[name + ".bias"] = ((out_c,), f"b:{fan}")

    def norm(name, c):
        P[name + ".weight"] = ((c,), "norm_w")
        P[name + ".bias"] = ((c,), "norm_b")

    def mlp(name, i, hdim, o, n):
        dims = [i] + [hdim] * (n - 1) + [o]
        for k in range(n):
            lin(f"{name}.layers.{k}", dims[k + 1], dims[k])

    def attn(name, emb, internal, kv_in=None):
        kv_in = kv_in or emb
        lin(name + ".q_proj", internal, emb)
        lin(name + ".k_proj", internal, kv_in)
        lin(name + ".v_proj", internal, kv_in)
        lin(name + ".out_proj", emb, internal)

    # ---- top-level parameters
    P["maskmem_tpos_enc"] = ((num_maskmem, 1, 1, MEM_DIM), "small")
    P["no_mem_embed"] = ((1, 1, HIDDEN), "small")
    P["no_mem_pos_enc"] = ((1, 1, HIDDEN), "small")
    P["no_obj_ptr"] = ((1, HIDDEN), "small")
    P["no_obj_embed_spatial"] = ((1, MEM_DIM), "small")
    # ---- trunk
    t = "image_encoder.trunk."
    E = h["embed_dim"]
    P[t + "pos_embed"] = ((1, E, *h["bkg_size"]), "small")
    P[t + "pos_embed_window"] = ((1, E, h["window_spec"][0], h["window_spec"][0]), "small")
    conv(t + "patch_embed.proj", E, 3, 7, 7)
    specs, stage_ends = block_specs(cfg)
    for i, s in enumerate(specs):
        b = f"{t}blocks.{i}."
        norm(b + "norm1", s["dim"])
        lin(b + "attn.qkv", 3 * s["dim_out"], s["dim"])
        lin(b + "attn.proj", s["dim_out"], s["dim_out"])
        norm(b + "norm2", s["dim_out"])
        lin(b + "mlp.layers.0", 4 * s["dim_out"], s["dim_out"])
        lin(b + "mlp.layers.1", s["dim_out"], 4 * s["dim_out"])
        if s["dim"] != s["dim_out"]:
            lin(b + "proj", s["dim_out"], s["dim"])
    chans = [specs[e]["dim_out"] for e in stage_ends[::-1]]
    for i, c in enumerate(chans):
        conv(f"image_encoder.neck.convs.{i}.conv", HIDDEN, c, 1, 1)
    # ---- memory attention
    for l in range(4):
        b = f"memory_attention.layers.{l}."
        attn(b + "self_attn", HIDDEN, HIDDEN)
        attn(b + "cross_attn_image", HIDDEN, HIDDEN, kv_in=MEM_DIM)
        lin(b + "linear1", 2048, HIDDEN)
        lin(b + "linear2", HIDDEN, 2048)
        for k in (1, 2, 3):
            norm(b + f"norm{k}", HIDDEN)
    norm("memory_attention.norm", HIDDEN)
    # ---- memory encoder
    cin = 1
    for k in range(4):
        cout = cin * 4
        conv(f"memory_encoder.mask_downsampler.encoder.{3 * k}", cout, cin, 3, 3)
        norm(f"memory_encoder.mask_downsampler.encoder.{3 * k + 1}", cout)
        cin = cout
    conv("memory_encoder.mask_downsampler.encoder.12", HIDDEN, cin, 1, 1)
    conv("memory_encoder.pix_feat_proj", HIDDEN, HIDDEN, 1, 1)
    for l in range(2):
        b = f"memory_encoder.fuser.layers.{l}."
        P[b + "gamma"] = ((HIDDEN,), "gamma")
        conv(b + "dwconv", HIDDEN, HIDDEN, 7, 7, groups=HIDDEN)
        norm(b + "norm", HIDDEN)
        lin(b + "pwconv1", 4 * HIDDEN, HIDDEN)
        lin(b + "pwconv2", HIDDEN, 4 * HIDDEN)
    conv("memory_encoder.out_proj", MEM_DIM, HIDDEN, 1, 1)
    # ---- prompt encoder
    pe = "sam_prompt_encoder."
    P[pe + "pe_layer.positional_encoding_gaussian_matrix"] = ((2, HIDDEN // 2), "gauss")
    for i in range(4):
        P[pe + f"point_embeddings.{i}.weight"] = ((1, HIDDEN), "embed")
    P[pe + "not_a_point_embed.weight"] = ((1, HIDDEN), "embed")
    conv(pe + "mask_downscaling.0", 4, 1, 2, 2)
    norm(pe + "mask_downscaling.1", 4)
    conv(pe + "mask_downscaling.3", 16, 4, 2, 2)
    norm(pe + "mask_downscaling.4", 16)
    conv(pe + "mask_downscaling.6", HIDDEN, 16, 1, 1)
    P[pe + "no_mask_embed.weight"] = ((1, HIDDEN), "embed")
    # ---- mask decoder
    md = "sam_mask_decoder."
    for l in range(2):
        b = f"{md}transformer.layers.{l}."
        attn(b + "self_attn", HIDDEN, HIDDEN)
        norm(b + "norm1", HIDDEN)
        attn(b + "cross_attn_token_to_image", HIDDEN, HIDDEN // 2)
        norm(b + "norm2", HIDDEN)
        lin(b + "mlp.layers.0", 2048, HIDDEN)
        lin(b + "mlp.layers.1", HIDDEN, 2048)
        norm(b + "norm3", HIDDEN)
        norm(b + "norm4", HIDDEN)
        attn(b + "cross_attn_image_to_token", HIDDEN, HIDDEN // 2)
    attn(md + "transformer.final_attn_token_to_image", HIDDEN, HIDDEN // 2)
    norm(md + "transformer.norm_final_attn", HIDDEN)
    P[md + "iou_token.weight"] = ((1, HIDDEN), "embed")
    P[md + "mask_tokens.weight"] = ((4, HIDDEN), "embed")
    P[md + "obj_score_token.weight"] = ((1, HIDDEN), "embed")
    convT(md + "output_upscaling.0", HIDDEN, HIDDEN // 4, 2, 2)
    norm(md + "output_upscaling.1", HIDDEN // 4)
    convT(md + "output_upscaling.3", HIDDEN // 4, HIDDEN // 8, 2, 2)
    conv(md + "conv_s0", HIDDEN // 8, HIDDEN, 1, 1)
    conv(md + "conv_s1", HIDDEN // 4, HIDDEN, 1, 1)
    for i in range(4):
        mlp(md + f"output_hypernetworks_mlps.{i}", HIDDEN, HIDDEN, HIDDEN // 8, 3)
    mlp(md + "iou_prediction_head", HIDDEN, HIDDEN, 4, 3)
    mlp(md + "pred_obj_score_head", HIDDEN, HIDDEN, 1, 3)
    # ---- pointers
    mlp("obj_ptr_proj", HIDDEN, HIDDEN, HIDDEN, 3)
    lin("obj_ptr_tpos_proj", MEM_DIM, HIDDEN)
    conv("mask_downsample", 1, 1, 4, 4)
    return P


def random_state_dict(cfg: str, seed: int = 0, num_maskmem: int = 7) -> Dict[str, torch.Tensor]:
    """Deterministic random init (CPU generator, fp32). Linear/conv follow torch's default bound
    1/sqrt(fan_in); parameters whose upstream init is degenerate (zeros / ones / 1e-6) get small
    random values instead so every term of the arithmetic is exercised by the parity tests."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for name, (shape, kind) in param_shapes(cfg, num_maskmem).items():
        if kind.startswith("w:") or kind.startswith("b:"):
            bound = 1.0 / math.sqrt(int(kind[2:]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "norm_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "norm_b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "embed":
            t = torch.randn(shape, generator=g)
        elif kind == "small":
            t = 0.02 * torch.randn(shape, generator=g)
        elif kind == "gauss":
            t = torch.randn(shape, generator=g)
        elif kind == "gamma":
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            raise ValueError(kind)
        sd[name] = t
    return sd
