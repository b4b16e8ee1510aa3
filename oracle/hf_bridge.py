"""ORACLE support (test infrastructure only): parameter-name map between the independent HF
``transformers`` 5.5.0 SAM2 implementation (present in this image) and upstream ``sam2`` checkpoint
names. Used to pin the oracle restatement numerically: identical weights are
loaded into HF ``Sam2Model`` / ``Sam2VideoModel`` and into ``oracle.sam2_ref`` and outputs compared.
"""
from __future__ import annotations

import re
from typing import Dict

import torch

_MLP_HEADS = ("output_hypernetworks_mlps", "iou_prediction_head", "pred_obj_score_head", "obj_ptr_proj",
              "object_pointer_proj")


def _map_mlp_head(rest: str) -> str:
    # HF: proj_in / layers.K / proj_out  ->  upstream: layers.0 / layers.K+1 / layers.last
    m = re.match(r"(.*)\.proj_in\.(weight|bias)$", rest)
    if m:
        return f"{m.group(1)}.layers.0.{m.group(2)}"
    m = re.match(r"(.*)\.layers\.(\d+)\.(weight|bias)$", rest)
    if m:
        return f"{m.group(1)}.layers.{int(m.group(2)) + 1}.{m.group(3)}"
    m = re.match(r"(.*)\.proj_out\.(weight|bias)$", rest)
    if m:
        return f"{m.group(1)}.layers.2.{m.group(2)}"  # all heads on this path have 3 layers
    return rest


def hf_to_upstream(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Translate an HF Sam2Model / Sam2VideoModel state-dict into upstream sam2 names."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in hf_sd.items():
        v = v.detach().clone()
        if k in ("no_memory_embedding",):
            out["no_mem_embed"] = v
            continue
        if k == "no_memory_positional_encoding":
            out["no_mem_pos_enc"] = v
            continue
        if k == "memory_temporal_positional_encoding":
            out["maskmem_tpos_enc"] = v
            continue
        if k == "no_object_pointer":
            out["no_obj_ptr"] = v
            continue
        if k == "occlusion_spatial_embedding_parameter":
            out["no_obj_embed_spatial"] = v
            continue
        if k.startswith("shared_image_embedding."):
            continue  # duplicate of prompt_encoder.shared_embedding
        if k == "prompt_encoder.shared_embedding.positional_embedding":
            out["sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"] = v
            continue
        if k == "prompt_encoder.point_embed.weight":
            for i in range(4):
                out[f"sam_prompt_encoder.point_embeddings.{i}.weight"] = v[i:i + 1].clone()
            continue
        if k.startswith("vision_encoder.backbone."):
            r = k[len("vision_encoder.backbone."):]
            r = r.replace("patch_embed.projection", "patch_embed.proj")
            r = r.replace("layer_norm1", "norm1").replace("layer_norm2", "norm2")
            r = r.replace("mlp.proj_in", "mlp.layers.0").replace("mlp.proj_out", "mlp.layers.1")
            out["image_encoder.trunk." + r] = v
            continue
        if k.startswith("vision_encoder.neck.convs."):
            m = re.match(r"vision_encoder\.neck\.convs\.(\d+)\.(weight|bias)", k)
            out[f"image_encoder.neck.convs.{m.group(1)}.conv.{m.group(2)}"] = v
            continue
        if k.startswith("prompt_encoder."):
            r = k[len("prompt_encoder."):]
            r = (r.replace("mask_embed.conv1", "mask_downscaling.0").replace("mask_embed.layer_norm1", "mask_downscaling.1")
                  .replace("mask_embed.conv2", "mask_downscaling.3").replace("mask_embed.layer_norm2", "mask_downscaling.4")
                  .replace("mask_embed.conv3", "mask_downscaling.6"))
            out["sam_prompt_encoder." + r] = v
            continue
        if k.startswith("mask_decoder."):
            r = k[len("mask_decoder."):]
            if r.startswith("transformer."):
                r = r.replace(".o_proj.", ".out_proj.")
                r = re.sub(r"layer_norm(\d)", r"norm\1", r)
                r = r.replace("layer_norm_final_attn", "norm_final_attn")
                r = r.replace("mlp.proj_in", "mlp.layers.0").replace("mlp.proj_out", "mlp.layers.1")
            elif r.startswith(_MLP_HEADS):
                r = _map_mlp_head(r)
            else:
                r = (r.replace("upscale_conv1", "output_upscaling.0").replace("upscale_layer_norm", "output_upscaling.1")
                      .replace("upscale_conv2", "output_upscaling.3"))
            out["sam_mask_decoder." + r] = v
            continue
        if k.startswith("memory_attention."):
            r = k.replace(".o_proj.", ".out_proj.")
            r = re.sub(r"layer_norm(\d)", r"norm\1", r)
            r = r.replace("memory_attention.layer_norm.", "memory_attention.norm.")
            out[r] = v
            continue
        if k.startswith("memory_encoder."):
            r = k[len("memory_encoder."):]
            m = re.match(r"mask_downsampler\.layers\.(\d+)\.(conv|layer_norm)\.(weight|bias)", r)
            if m:
                idx = int(m.group(1)) * 3 + (0 if m.group(2) == "conv" else 1)
                out[f"memory_encoder.mask_downsampler.encoder.{idx}.{m.group(3)}"] = v
                continue
            m = re.match(r"mask_downsampler\.final_conv\.(weight|bias)", r)
            if m:
                out[f"memory_encoder.mask_downsampler.encoder.12.{m.group(1)}"] = v
                continue
            r = (r.replace("feature_projection", "pix_feat_proj").replace("memory_fuser", "fuser")
                  .replace("depthwise_conv", "dwconv").replace("layer_norm", "norm")
                  .replace("pointwise_conv1", "pwconv1").replace("pointwise_conv2", "pwconv2")
                  .replace(".scale", ".gamma"))
            if r.startswith("projection."):
                r = "out_proj." + r[len("projection."):]
            out["memory_encoder." + r] = v
            continue
        if k.startswith("mask_downsample."):
            out[k] = v
            continue
        if k.startswith("object_pointer_proj."):
            out[_map_mlp_head(k).replace("object_pointer_proj", "obj_ptr_proj")] = v
            continue
        if k.startswith("temporal_positional_encoding_projection_layer."):
            out[k.replace("temporal_positional_encoding_projection_layer", "obj_ptr_tpos_proj")] = v
            continue
        raise KeyError(f"unmapped HF parameter {k}")
    return out


def hf_image_config(cfg: str):
    """HF Sam2Config for one of the four SAM2.1 Hiera sizes."""
    from transformers import Sam2Config, Sam2HieraDetConfig, Sam2VisionConfig

    table = {
        "tiny": dict(hidden_size=96, num_attention_heads=1, blocks_per_stage=[1, 2, 7, 2],
                     embed_dim_per_stage=[96, 192, 384, 768], num_attention_heads_per_stage=[1, 2, 4, 8],
                     window_size_per_stage=[8, 4, 14, 7], global_attention_blocks=[5, 7, 9],
                     window_positional_embedding_background_size=[7, 7]),
        "small": dict(hidden_size=96, num_attention_heads=1, blocks_per_stage=[1, 2, 11, 2],
                      embed_dim_per_stage=[96, 192, 384, 768], num_attention_heads_per_stage=[1, 2, 4, 8],
                      window_size_per_stage=[8, 4, 14, 7], global_attention_blocks=[7, 10, 13],
                      window_positional_embedding_background_size=[7, 7]),
        "base_plus": dict(hidden_size=112, num_attention_heads=2, blocks_per_stage=[2, 3, 16, 3],
                          embed_dim_per_stage=[112, 224, 448, 896], num_attention_heads_per_stage=[2, 4, 8, 16],
                          window_size_per_stage=[8, 4, 14, 7], global_attention_blocks=[12, 16, 20],
                          window_positional_embedding_background_size=[14, 14]),
        "large": dict(hidden_size=144, num_attention_heads=2, blocks_per_stage=[2, 6, 36, 4],
                      embed_dim_per_stage=[144, 288, 576, 1152], num_attention_heads_per_stage=[2, 4, 8, 16],
                      window_size_per_stage=[8, 4, 16, 8], global_attention_blocks=[23, 33, 43],
                      window_positional_embedding_background_size=[7, 7]),
    }
    t = table[cfg]
    backbone = Sam2HieraDetConfig(**t)
    vision = Sam2VisionConfig(backbone_config=backbone, backbone_channel_list=list(reversed(t["embed_dim_per_stage"])))
    return Sam2Config(vision_config=vision)
