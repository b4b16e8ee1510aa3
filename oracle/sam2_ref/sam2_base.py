"""ORACLE (test infrastructure only). Restatement of upstream sam2/modeling/sam2_base.py (SAM2Base)
with the SAM2.1 flag set listed in SURVEY §8c, plus sam2/build_sam.py's architecture table for the
four ``configs/sam2.1/sam2.1_hiera_{t,s,b+,l}.yaml`` files that REF saber/pretrained_weights.py:183-202
selects. Reached from REF saber/adapters/sam2/predictor.py:24-26 and automask.py:62.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .memory import CXBlock, Fuser, MaskDownSampler, MemoryAttention, MemoryEncoder
from .modeling import (MLP, FpnNeck, Hiera, ImageEncoder, MaskDecoder, PositionEmbeddingSine, PromptEncoder,
                       TwoWayTransformer)

NO_OBJ_SCORE = -1024.0

HIERA_CFGS = {
    "tiny": dict(embed_dim=96, num_heads=1, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9),
                 window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 14, 7)),
    "small": dict(embed_dim=96, num_heads=1, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13),
                  window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 14, 7)),
    "base_plus": dict(embed_dim=112, num_heads=2, stages=(2, 3, 16, 3), global_att_blocks=(12, 16, 20),
                      window_pos_embed_bkg_spatial_size=(14, 14), window_spec=(8, 4, 14, 7)),
    "large": dict(embed_dim=144, num_heads=2, stages=(2, 6, 36, 4), global_att_blocks=(23, 33, 43),
                  window_pos_embed_bkg_spatial_size=(7, 7), window_spec=(8, 4, 16, 8)),
}
CFG_ALIASES = {
    "t": "tiny", "s": "small", "b+": "base_plus", "base": "base_plus", "l": "large",
    "configs/sam2.1/sam2.1_hiera_t.yaml": "tiny", "configs/sam2.1/sam2.1_hiera_s.yaml": "small",
    "configs/sam2.1/sam2.1_hiera_b+.yaml": "base_plus", "configs/sam2.1/sam2.1_hiera_l.yaml": "large",
}


def resolve_cfg(name: str) -> str:
    name = CFG_ALIASES.get(name, name)
    if name not in HIERA_CFGS:
        raise ValueError(f"unknown SAM2.1 config {name!r}")
    return name


def get_1d_sine_pe(pos_inds, dim, temperature=10000):
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32, device=pos_inds.device)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    pos_embed = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([pos_embed.sin(), pos_embed.cos()], dim=-1)


def select_closest_cond_frames(frame_idx, cond_frame_outputs, max_cond_frame_num):
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    raise NotImplementedError("max_cond_frames_in_attn != -1 is not on SABER's path")


class SAM2Base(nn.Module):
    def __init__(self, cfg: str = "large", num_maskmem: int = 7, image_size: int = 1024,
                 dynamic_multimask_via_stability: bool = False, binarize_mask_from_pts_for_mem_enc: bool = False,
                 fill_hole_area: int = 0):
        super().__init__()
        cfg = resolve_cfg(cfg)
        h = HIERA_CFGS[cfg]
        trunk = Hiera(**h)
        neck = FpnNeck(position_encoding=PositionEmbeddingSine(num_pos_feats=256, normalize=True, temperature=10000),
                       d_model=256, backbone_channel_list=list(trunk.channel_list), fpn_top_down_levels=[2, 3],
                       fpn_interp_model="nearest")
        self.image_encoder = ImageEncoder(trunk=trunk, neck=neck, scalp=1)
        self.use_high_res_features_in_sam = True
        self.num_feature_levels = 3
        self.use_obj_ptrs_in_encoder = True
        self.max_obj_ptrs_in_encoder = 16
        self.mask_downsample = nn.Conv2d(1, 1, kernel_size=4, stride=4)
        self.add_tpos_enc_to_obj_ptrs = True
        self.proj_tpos_enc_in_obj_ptrs = True
        self.use_signed_tpos_enc_to_obj_ptrs = True
        self.only_obj_ptrs_in_the_past_for_eval = True
        self.memory_attention = MemoryAttention(d_model=256, pos_enc_at_input=True, num_layers=4)
        self.hidden_dim = 256
        self.memory_encoder = MemoryEncoder(
            out_dim=64, position_encoding=PositionEmbeddingSine(num_pos_feats=64, normalize=True, temperature=10000),
            mask_downsampler=MaskDownSampler(kernel_size=3, stride=2, padding=1),
            fuser=Fuser(CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True), num_layers=2))
        self.mem_dim = 64
        self.num_maskmem = num_maskmem
        self.maskmem_tpos_enc = nn.Parameter(torch.zeros(num_maskmem, 1, 1, self.mem_dim))
        nn.init.trunc_normal_(self.maskmem_tpos_enc, std=0.02)
        self.no_mem_embed = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        self.no_mem_pos_enc = nn.Parameter(torch.zeros(1, 1, self.hidden_dim))
        nn.init.trunc_normal_(self.no_mem_embed, std=0.02)
        nn.init.trunc_normal_(self.no_mem_pos_enc, std=0.02)
        self.directly_add_no_mem_embed = True
        self.sigmoid_scale_for_mem_enc = 20.0
        self.sigmoid_bias_for_mem_enc = -10.0
        self.binarize_mask_from_pts_for_mem_enc = binarize_mask_from_pts_for_mem_enc
        self.non_overlap_masks_for_mem_enc = False
        self.memory_temporal_stride_for_eval = 1
        self.use_mask_input_as_output_without_sam = True
        self.multimask_output_in_sam = True
        self.multimask_min_pt_num = 0
        self.multimask_max_pt_num = 1
        self.multimask_output_for_tracking = True
        self.use_multimask_token_for_obj_ptr = True
        self.iou_prediction_use_sigmoid = True
        self.image_size = image_size
        self.backbone_stride = 16
        self.pred_obj_scores = True
        self.pred_obj_scores_mlp = True
        self.fixed_no_obj_ptr = True
        self.soft_no_obj_ptr = False
        self.no_obj_ptr = nn.Parameter(torch.zeros(1, self.hidden_dim))
        nn.init.trunc_normal_(self.no_obj_ptr, std=0.02)
        self.use_mlp_for_obj_ptr_proj = True
        self.no_obj_embed_spatial = nn.Parameter(torch.zeros(1, self.mem_dim))
        nn.init.trunc_normal_(self.no_obj_embed_spatial, std=0.02)
        self.max_cond_frames_in_attn = -1
        self.fill_hole_area = fill_hole_area
        # SAM heads
        self.sam_prompt_embed_dim = self.hidden_dim
        self.sam_image_embedding_size = self.image_size // self.backbone_stride
        self.sam_prompt_encoder = PromptEncoder(
            embed_dim=256, image_embedding_size=(self.sam_image_embedding_size,) * 2,
            input_image_size=(self.image_size, self.image_size), mask_in_chans=16)
        self.sam_mask_decoder = MaskDecoder(
            num_multimask_outputs=3,
            transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
            transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256, use_high_res_features=True,
            iou_prediction_use_sigmoid=True, pred_obj_scores=True, pred_obj_scores_mlp=True,
            use_multimask_token_for_obj_ptr=True,
            dynamic_multimask_via_stability=dynamic_multimask_via_stability,
            dynamic_multimask_stability_delta=0.05, dynamic_multimask_stability_thresh=0.98)
        self.obj_ptr_proj = MLP(self.hidden_dim, self.hidden_dim, self.hidden_dim, 3)
        self.obj_ptr_tpos_proj = nn.Linear(self.hidden_dim, self.mem_dim)

    @property
    def device(self):
        return next(self.parameters()).device

    # ---------------- image side ----------------
    def forward_image(self, img_batch):
        backbone_out = self.image_encoder(img_batch)
        backbone_out["backbone_fpn"][0] = self.sam_mask_decoder.conv_s0(backbone_out["backbone_fpn"][0])
        backbone_out["backbone_fpn"][1] = self.sam_mask_decoder.conv_s1(backbone_out["backbone_fpn"][1])
        return backbone_out

    def _prepare_backbone_features(self, backbone_out):
        backbone_out = backbone_out.copy()
        feature_maps = backbone_out["backbone_fpn"][-self.num_feature_levels:]
        vision_pos_embeds = backbone_out["vision_pos_enc"][-self.num_feature_levels:]
        feat_sizes = [(x.shape[-2], x.shape[-1]) for x in vision_pos_embeds]
        vision_feats = [x.flatten(2).permute(2, 0, 1) for x in feature_maps]
        vision_pos_embeds = [x.flatten(2).permute(2, 0, 1) for x in vision_pos_embeds]
        return backbone_out, vision_feats, vision_pos_embeds, feat_sizes

    # ---------------- SAM heads ----------------
    def _forward_sam_heads(self, backbone_features, point_inputs=None, mask_inputs=None, high_res_features=None,
                           multimask_output=False):
        B = backbone_features.size(0)
        device = backbone_features.device
        if point_inputs is not None:
            sam_point_coords = point_inputs["point_coords"]
            sam_point_labels = point_inputs["point_labels"]
        else:
            sam_point_coords = torch.zeros(B, 1, 2, device=device)
            sam_point_labels = -torch.ones(B, 1, dtype=torch.int32, device=device)
        if mask_inputs is not None:
            if mask_inputs.shape[-2:] != self.sam_prompt_encoder.mask_input_size:
                sam_mask_prompt = F.interpolate(mask_inputs.float(), size=self.sam_prompt_encoder.mask_input_size,
                                                align_corners=False, mode="bilinear", antialias=True)
            else:
                sam_mask_prompt = mask_inputs
        else:
            sam_mask_prompt = None
        sparse, dense = self.sam_prompt_encoder(points=(sam_point_coords, sam_point_labels), boxes=None,
                                                masks=sam_mask_prompt)
        low_res_multimasks, ious, sam_output_tokens, object_score_logits = self.sam_mask_decoder(
            image_embeddings=backbone_features, image_pe=self.sam_prompt_encoder.get_dense_pe(),
            sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=multimask_output,
            repeat_image=False, high_res_features=high_res_features)
        is_obj_appearing = object_score_logits > 0
        low_res_multimasks = torch.where(is_obj_appearing[:, None, None], low_res_multimasks, NO_OBJ_SCORE)
        low_res_multimasks = low_res_multimasks.float()
        high_res_multimasks = F.interpolate(low_res_multimasks, size=(self.image_size, self.image_size),
                                            mode="bilinear", align_corners=False)
        sam_output_token = sam_output_tokens[:, 0]
        if multimask_output:
            best = torch.argmax(ious, dim=-1)
            bi = torch.arange(B, device=device)
            low_res_masks = low_res_multimasks[bi, best].unsqueeze(1)
            high_res_masks = high_res_multimasks[bi, best].unsqueeze(1)
            if sam_output_tokens.size(1) > 1:
                sam_output_token = sam_output_tokens[bi, best]
        else:
            low_res_masks, high_res_masks = low_res_multimasks, high_res_multimasks
        obj_ptr = self.obj_ptr_proj(sam_output_token)
        lam = is_obj_appearing.float()
        obj_ptr = lam * obj_ptr
        obj_ptr = obj_ptr + (1 - lam) * self.no_obj_ptr
        return (low_res_multimasks, high_res_multimasks, ious, low_res_masks, high_res_masks, obj_ptr,
                object_score_logits)

    def _use_mask_as_output(self, backbone_features, high_res_features, mask_inputs):
        out_scale, out_bias = 20.0, -10.0
        mask_inputs_float = mask_inputs.float()
        high_res_masks = mask_inputs_float * out_scale + out_bias
        low_res_masks = F.interpolate(high_res_masks, size=(high_res_masks.size(-2) // 4, high_res_masks.size(-1) // 4),
                                      align_corners=False, mode="bilinear", antialias=True)
        ious = mask_inputs.new_ones(mask_inputs.size(0), 1).float()
        _, _, _, _, _, obj_ptr, _ = self._forward_sam_heads(
            backbone_features=backbone_features, mask_inputs=self.mask_downsample(mask_inputs_float),
            high_res_features=high_res_features)
        is_obj_appearing = torch.any(mask_inputs.flatten(1).float() > 0.0, dim=1)[..., None]
        lam = is_obj_appearing.float()
        object_score_logits = out_scale * lam + out_bias
        obj_ptr = lam * obj_ptr
        obj_ptr = obj_ptr + (1 - lam) * self.no_obj_ptr
        return low_res_masks, high_res_masks, ious, low_res_masks, high_res_masks, obj_ptr, object_score_logits

    # ---------------- memory ----------------
    def _prepare_memory_conditioned_features(self, frame_idx, is_init_cond_frame, current_vision_feats,
                                             current_vision_pos_embeds, feat_sizes, output_dict, num_frames,
                                             track_in_reverse=False):
        B = current_vision_feats[-1].size(1)
        C = self.hidden_dim
        H, W = feat_sizes[-1]
        device = current_vision_feats[-1].device
        if self.num_maskmem == 0:
            return current_vision_feats[-1].permute(1, 2, 0).view(B, C, H, W)
        num_obj_ptr_tokens = 0
        tpos_sign_mul = -1 if track_in_reverse else 1
        if is_init_cond_frame:
            pix = current_vision_feats[-1] + self.no_mem_embed
            return pix.permute(1, 2, 0).view(B, C, H, W)
        to_cat_memory, to_cat_memory_pos_embed = [], []
        assert len(output_dict["cond_frame_outputs"]) > 0
        cond_outputs = output_dict["cond_frame_outputs"]
        selected_cond_outputs, unselected_cond_outputs = select_closest_cond_frames(
            frame_idx, cond_outputs, self.max_cond_frames_in_attn)
        t_pos_and_prevs = [(0, out) for out in selected_cond_outputs.values()]
        stride = self.memory_temporal_stride_for_eval
        for t_pos in range(1, self.num_maskmem):
            t_rel = self.num_maskmem - t_pos
            if t_rel == 1:
                prev_frame_idx = frame_idx - t_rel if not track_in_reverse else frame_idx + t_rel
            else:
                if not track_in_reverse:
                    prev_frame_idx = ((frame_idx - 2) // stride) * stride
                    prev_frame_idx = prev_frame_idx - (t_rel - 2) * stride
                else:
                    prev_frame_idx = -(-(frame_idx + 2) // stride) * stride
                    prev_frame_idx = prev_frame_idx + (t_rel - 2) * stride
            out = output_dict["non_cond_frame_outputs"].get(prev_frame_idx, None)
            if out is None:
                out = unselected_cond_outputs.get(prev_frame_idx, None)
            t_pos_and_prevs.append((t_pos, out))
        for t_pos, prev in t_pos_and_prevs:
            if prev is None:
                continue
            feats = prev["maskmem_features"].to(device)
            to_cat_memory.append(feats.flatten(2).permute(2, 0, 1))
            maskmem_enc = prev["maskmem_pos_enc"][-1].to(device)
            maskmem_enc = maskmem_enc.flatten(2).permute(2, 0, 1)
            maskmem_enc = maskmem_enc + self.maskmem_tpos_enc[self.num_maskmem - t_pos - 1]
            to_cat_memory_pos_embed.append(maskmem_enc)
        if self.use_obj_ptrs_in_encoder:
            max_obj_ptrs_in_encoder = min(num_frames, self.max_obj_ptrs_in_encoder)
            ptr_cond_outputs = {t: out for t, out in selected_cond_outputs.items()
                                if (t >= frame_idx if track_in_reverse else t <= frame_idx)}
            pos_and_ptrs = [((frame_idx - t) * tpos_sign_mul, out["obj_ptr"]) for t, out in ptr_cond_outputs.items()]
            for t_diff in range(1, max_obj_ptrs_in_encoder):
                t = frame_idx + t_diff if track_in_reverse else frame_idx - t_diff
                if t < 0 or (num_frames is not None and t >= num_frames):
                    break
                out = output_dict["non_cond_frame_outputs"].get(t, unselected_cond_outputs.get(t, None))
                if out is not None:
                    pos_and_ptrs.append((t_diff, out["obj_ptr"]))
            if len(pos_and_ptrs) > 0:
                pos_list, ptrs_list = zip(*pos_and_ptrs)
                obj_ptrs = torch.stack(ptrs_list, dim=0)
                t_diff_max = max_obj_ptrs_in_encoder - 1
                obj_pos = torch.tensor(pos_list, device=device)
                obj_pos = get_1d_sine_pe(obj_pos / t_diff_max, dim=C)
                obj_pos = self.obj_ptr_tpos_proj(obj_pos)
                obj_pos = obj_pos.unsqueeze(1).expand(-1, B, self.mem_dim)
                if self.mem_dim < C:
                    obj_ptrs = obj_ptrs.reshape(-1, B, C // self.mem_dim, self.mem_dim)
                    obj_ptrs = obj_ptrs.permute(0, 2, 1, 3).flatten(0, 1)
                    obj_pos = obj_pos.repeat_interleave(C // self.mem_dim, dim=0)
                to_cat_memory.append(obj_ptrs)
                to_cat_memory_pos_embed.append(obj_pos)
                num_obj_ptr_tokens = obj_ptrs.shape[0]
        # maskmem_features are stored in bf16; upstream reaches fp32 here through torch.cat's type promotion with the fp32
        # object pointers (or runs under bf16 autocast). Without pointer tokens (a conditioning frame only in the
        # "future") nothing promotes: make the cast explicit, the values are the same.
        memory = torch.cat(to_cat_memory, dim=0).float()
        memory_pos_embed = torch.cat(to_cat_memory_pos_embed, dim=0)
        pix = self.memory_attention(curr=current_vision_feats[-1:], curr_pos=current_vision_pos_embeds[-1:],
                                    memory=memory, memory_pos=memory_pos_embed, num_obj_ptr_tokens=num_obj_ptr_tokens)
        return pix.permute(1, 2, 0).view(B, C, H, W)

    def _encode_new_memory(self, current_vision_feats, feat_sizes, pred_masks_high_res, object_score_logits,
                           is_mask_from_pts):
        B = current_vision_feats[-1].size(1)
        C = self.hidden_dim
        H, W = feat_sizes[-1]
        pix_feat = current_vision_feats[-1].permute(1, 2, 0).view(B, C, H, W)
        binarize = self.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts
        if binarize:
            mask_for_mem = (pred_masks_high_res > 0).float()
        else:
            mask_for_mem = torch.sigmoid(pred_masks_high_res)
        mask_for_mem = mask_for_mem * self.sigmoid_scale_for_mem_enc
        mask_for_mem = mask_for_mem + self.sigmoid_bias_for_mem_enc
        out = self.memory_encoder(pix_feat, mask_for_mem, skip_mask_sigmoid=True)
        maskmem_features = out["vision_features"]
        maskmem_pos_enc = out["vision_pos_enc"]
        is_obj_appearing = (object_score_logits > 0).float()
        maskmem_features = maskmem_features + (1 - is_obj_appearing[..., None, None]) * \
            self.no_obj_embed_spatial[..., None, None].expand(*maskmem_features.shape)
        return maskmem_features, maskmem_pos_enc

    def _use_multimask(self, is_init_cond_frame, point_inputs):
        num_pts = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        return (self.multimask_output_in_sam and (is_init_cond_frame or self.multimask_output_for_tracking)
                and (self.multimask_min_pt_num <= num_pts <= self.multimask_max_pt_num))

    def track_step(self, frame_idx, is_init_cond_frame, current_vision_feats, current_vision_pos_embeds, feat_sizes,
                   point_inputs, mask_inputs, output_dict, num_frames, track_in_reverse=False, run_mem_encoder=True,
                   prev_sam_mask_logits=None):
        current_out = {"point_inputs": point_inputs, "mask_inputs": mask_inputs}
        high_res_features = [x.permute(1, 2, 0).view(x.size(1), x.size(2), *s)
                             for x, s in zip(current_vision_feats[:-1], feat_sizes[:-1])]
        if mask_inputs is not None and self.use_mask_input_as_output_without_sam:
            pix_feat = current_vision_feats[-1].permute(1, 2, 0).view(-1, self.hidden_dim, *feat_sizes[-1])
            sam_outputs = self._use_mask_as_output(pix_feat, high_res_features, mask_inputs)
        else:
            pix_feat = self._prepare_memory_conditioned_features(
                frame_idx=frame_idx, is_init_cond_frame=is_init_cond_frame,
                current_vision_feats=current_vision_feats[-1:], current_vision_pos_embeds=current_vision_pos_embeds[-1:],
                feat_sizes=feat_sizes[-1:], output_dict=output_dict, num_frames=num_frames,
                track_in_reverse=track_in_reverse)
            if prev_sam_mask_logits is not None:
                assert point_inputs is not None and mask_inputs is None
                mask_inputs = prev_sam_mask_logits
            multimask_output = self._use_multimask(is_init_cond_frame, point_inputs)
            sam_outputs = self._forward_sam_heads(backbone_features=pix_feat, point_inputs=point_inputs,
                                                  mask_inputs=mask_inputs, high_res_features=high_res_features,
                                                  multimask_output=multimask_output)
        _, _, _, low_res_masks, high_res_masks, obj_ptr, object_score_logits = sam_outputs
        current_out["pred_masks"] = low_res_masks
        current_out["pred_masks_high_res"] = high_res_masks
        current_out["obj_ptr"] = obj_ptr
        current_out["object_score_logits"] = object_score_logits
        if run_mem_encoder and self.num_maskmem > 0:
            maskmem_features, maskmem_pos_enc = self._encode_new_memory(
                current_vision_feats=current_vision_feats, feat_sizes=feat_sizes, pred_masks_high_res=high_res_masks,
                object_score_logits=object_score_logits, is_mask_from_pts=(point_inputs is not None))
            current_out["maskmem_features"] = maskmem_features
            current_out["maskmem_pos_enc"] = maskmem_pos_enc
        else:
            current_out["maskmem_features"] = None
            current_out["maskmem_pos_enc"] = None
        return current_out


def build_sam2(config_file="large", ckpt_path=None, device="cpu", mode="eval", apply_postprocessing=True,
               state_dict=None, **kw):
    """Oracle twin of sam2.build_sam.build_sam2 (apply_postprocessing -> dynamic multimask via stability)."""
    model = SAM2Base(cfg=config_file, dynamic_multimask_via_stability=bool(apply_postprocessing), **kw)
    _load(model, ckpt_path, state_dict)
    return model.to(device).eval()


def _load(model, ckpt_path, state_dict):
    if state_dict is None and ckpt_path is not None:
        state_dict = torch.load(ckpt_path, map_location="cpu", weights_only=True)["model"]
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        if missing or unexpected:
            raise RuntimeError(f"state-dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
