"""ORACLE (test infrastructure only). Restatement of upstream sam2/sam2_image_predictor.py and
sam2/utils/transforms.py (SAM2Transforms) as used by REF saber/classifier/models/SAM2.py:145-151
(set_image_batch, _features) and by the automatic mask generator (REF saber/adapters/sam2/automask.py:66).
SURVEY §8a U4.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class SAM2Transforms:
    """ToTensor (float HWC input is NOT rescaled) -> Resize(1024^2, bilinear, antialias) -> ImageNet normalise."""

    def __init__(self, resolution, mask_threshold, max_hole_area=0.0, max_sprinkle_area=0.0):
        self.resolution = resolution
        self.mask_threshold = mask_threshold
        self.max_hole_area = max_hole_area
        self.max_sprinkle_area = max_sprinkle_area
        self.mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
        self.std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)

    def __call__(self, x: np.ndarray) -> torch.Tensor:
        if x.dtype == np.uint8:
            t = torch.from_numpy(np.ascontiguousarray(x)).permute(2, 0, 1).float() / 255.0
        else:
            t = torch.from_numpy(np.ascontiguousarray(x)).permute(2, 0, 1).float()
        t = F.interpolate(t[None], size=(self.resolution, self.resolution), mode="bilinear", align_corners=False,
                          antialias=True)[0]
        return (t - self.mean) / self.std

    def forward_batch(self, img_list):
        return torch.stack([self(img) for img in img_list], dim=0)

    def transform_coords(self, coords, normalize=False, orig_hw=None):
        if normalize:
            assert orig_hw is not None
            h, w = orig_hw
            coords = coords.clone()
            coords[..., 0] = coords[..., 0] / w
            coords[..., 1] = coords[..., 1] / h
        return coords * self.resolution

    def postprocess_masks(self, masks, orig_hw):
        # max_hole_area / max_sprinkle_area are 0 on SABER's AMG path (min_mask_region_area = 0)
        masks = masks.float()
        return F.interpolate(masks, orig_hw, mode="bilinear", align_corners=False)


class SAM2ImagePredictor:
    def __init__(self, sam_model, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0):
        self.model = sam_model
        self._transforms = SAM2Transforms(self.model.image_size, mask_threshold, max_hole_area, max_sprinkle_area)
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False
        self.mask_threshold = mask_threshold
        self._bb_feat_sizes = [(256, 256), (128, 128), (64, 64)]

    @property
    def device(self):
        return self.model.device

    @torch.no_grad()
    def set_image(self, image: np.ndarray):
        self.reset_predictor()
        assert isinstance(image, np.ndarray)
        self._orig_hw = [image.shape[:2]]
        input_image = self._transforms(image)[None].to(self.device)
        self._set_features(input_image, 1)
        self._is_image_set = True

    @torch.no_grad()
    def set_image_batch(self, image_list: List[np.ndarray]):
        self.reset_predictor()
        self._orig_hw = [im.shape[:2] for im in image_list]
        img_batch = self._transforms.forward_batch(image_list).to(self.device)
        self._set_features(img_batch, len(image_list))
        self._is_image_set = True
        self._is_batch = True

    def _set_features(self, img_batch, batch_size):
        backbone_out = self.model.forward_image(img_batch)
        _, vision_feats, _, _ = self.model._prepare_backbone_features(backbone_out)
        if self.model.directly_add_no_mem_embed:
            vision_feats[-1] = vision_feats[-1] + self.model.no_mem_embed
        feats = [feat.permute(1, 2, 0).view(batch_size, -1, *fs)
                 for feat, fs in zip(vision_feats[::-1], self._bb_feat_sizes[::-1])][::-1]
        self._features = {"image_embed": feats[-1], "high_res_feats": feats[:-1]}

    @torch.no_grad()
    def _predict(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                 return_logits=False, img_idx=-1):
        assert self._is_image_set
        concat_points = (point_coords, point_labels) if point_coords is not None else None
        assert boxes is None, "box prompts are not on SABER's path"
        sparse, dense = self.model.sam_prompt_encoder(points=concat_points, boxes=None, masks=mask_input)
        batched_mode = concat_points is not None and concat_points[0].shape[0] > 1
        high_res = [lvl[img_idx].unsqueeze(0) for lvl in self._features["high_res_feats"]]
        low_res_masks, iou_predictions, _, _ = self.model.sam_mask_decoder(
            image_embeddings=self._features["image_embed"][img_idx].unsqueeze(0),
            image_pe=self.model.sam_prompt_encoder.get_dense_pe(),
            sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
            multimask_output=multimask_output, repeat_image=batched_mode, high_res_features=high_res)
        masks = self._transforms.postprocess_masks(low_res_masks, self._orig_hw[img_idx])
        low_res_masks = torch.clamp(low_res_masks, -32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks, iou_predictions, low_res_masks

    def reset_predictor(self):
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False
