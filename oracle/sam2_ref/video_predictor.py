"""ORACLE (test infrastructure only). Restatement of upstream sam2/sam2_video_predictor.py (SAM2VideoPredictor,
sam2 >= 1.1.0: independent per-object inference) and sam2/utils/misc.py::fill_holes_in_mask_scores, limited to the
surface REF saber/adapters/sam2/predictor.py uses:

  build_sam2_video_predictor (REF :24-26), maskmem_tpos_enc / num_maskmem re-assignment (REF :31-34),
  _get_image_feature (REF :114,152), add_new_mask (REF :164-169), propagate_in_video (+ preflight) (REF :196-202),
  reset_state (REF :358-361), sam_mask_decoder.register_forward_hook (REF :284).

The inference-state dict is *built by SABER* (REF :130-150), not by init_state; this class consumes those keys.
SURVEY §8a U6, Appendix A4. Upstream hydra overrides applied by build_sam2_video_predictor:
binarize_mask_from_pts_for_mem_enc=true, fill_hole_area=8, dynamic multimask via stability (delta 0.05, thresh 0.98).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from scipy import ndimage as ndi

from .sam2_base import NO_OBJ_SCORE, SAM2Base, _load


def fill_holes_in_mask_scores(mask: torch.Tensor, max_area: int) -> torch.Tensor:
    """sam2/utils/misc.py: background (score <= 0) connected components (8-connectivity, as sam2/csrc/
    connected_components.cu) with area <= max_area are set to +0.1. mask: [N,1,H,W] fp32."""
    assert max_area > 0
    out = mask.clone()
    eight = np.ones((3, 3), dtype=bool)
    m = mask.detach().cpu().numpy()
    for n in range(m.shape[0]):
        bg = m[n, 0] <= 0
        lab, k = ndi.label(bg, structure=eight)
        if k == 0:
            continue
        areas = np.bincount(lab.ravel())
        hole = (lab > 0) & (areas[lab] <= max_area)
        out[n, 0][torch.from_numpy(hole)] = 0.1
    return out


class SAM2VideoPredictor(SAM2Base):
    def __init__(self, cfg="large", fill_hole_area=0, non_overlap_masks=False, clear_non_cond_mem_around_input=False,
                 add_all_frames_to_correct_as_cond=False, **kw):
        super().__init__(cfg=cfg, fill_hole_area=fill_hole_area, **kw)
        self.non_overlap_masks = non_overlap_masks
        self.clear_non_cond_mem_around_input = clear_non_cond_mem_around_input
        self.add_all_frames_to_correct_as_cond = add_all_frames_to_correct_as_cond

    # ---- object bookkeeping ---------------------------------------------------------------
    def _obj_id_to_idx(self, state, obj_id):
        idx = state["obj_id_to_idx"].get(obj_id, None)
        if idx is not None:
            return idx
        idx = len(state["obj_id_to_idx"])
        state["obj_id_to_idx"][obj_id] = idx
        state["obj_idx_to_id"][idx] = obj_id
        state["obj_ids"] = list(state["obj_id_to_idx"])
        state["point_inputs_per_obj"][idx] = {}
        state["mask_inputs_per_obj"][idx] = {}
        state["output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        state["temp_output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        state["frames_tracked_per_obj"][idx] = {}
        return idx

    def _get_obj_num(self, state):
        return len(state["obj_idx_to_id"])

    # ---- features -------------------------------------------------------------------------
    def _get_image_feature(self, state, frame_idx, batch_size):
        image, backbone_out = state["cached_features"].get(frame_idx, (None, None))
        if backbone_out is None:
            image = state["images"][frame_idx].to(state["device"]).float().unsqueeze(0)
            backbone_out = self.forward_image(image)
            state["cached_features"] = {frame_idx: (image, backbone_out)}
        expanded_image = image.expand(batch_size, -1, -1, -1)
        exp = {"backbone_fpn": [f.expand(batch_size, -1, -1, -1) for f in backbone_out["backbone_fpn"]],
               "vision_pos_enc": [p.expand(batch_size, -1, -1, -1) for p in backbone_out["vision_pos_enc"]]}
        return (expanded_image,) + self._prepare_backbone_features(exp)

    # ---- prompting ------------------------------------------------------------------------
    @torch.inference_mode()
    def add_new_mask(self, inference_state, frame_idx, obj_id, mask):
        state = inference_state
        obj_idx = self._obj_id_to_idx(state, obj_id)
        if not isinstance(mask, torch.Tensor):
            mask = torch.tensor(mask, dtype=torch.bool)
        assert mask.dim() == 2
        mh, mw = mask.shape
        mi = mask[None, None].float().to(state["device"])
        if mh != self.image_size or mw != self.image_size:
            mi = F.interpolate(mi, size=(self.image_size, self.image_size), align_corners=False, mode="bilinear",
                               antialias=True)
            mi = (mi >= 0.5).float()
        state["mask_inputs_per_obj"][obj_idx][frame_idx] = mi
        state["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        tracked = state["frames_tracked_per_obj"][obj_idx]
        is_init_cond_frame = frame_idx not in tracked
        reverse = False if is_init_cond_frame else tracked[frame_idx]["reverse"]
        obj_out = state["output_dict_per_obj"][obj_idx]
        obj_tmp = state["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        current_out, _ = self._run_single_frame_inference(
            state, obj_out, frame_idx, 1, is_init_cond_frame, None, mi, reverse, run_mem_encoder=False)
        obj_tmp[key][frame_idx] = current_out
        cons = self._consolidate_temp_output_across_obj(state, frame_idx, is_cond, consolidate_at_video_res=True)
        _, video_res = self._get_orig_video_res_output(state, cons["pred_masks_video_res"])
        return frame_idx, state["obj_ids"], video_res

    @torch.inference_mode()
    def add_new_points_or_box(self, inference_state, frame_idx, obj_id, points=None, labels=None, clear_old_points=True,
                              normalize_coords=True, box=None):
        """sam2_video_predictor.SAM2VideoPredictor.add_new_points_or_box (forwarded by REF saber/adapters/sam2/
        predictor.py:171-180): a box is two extra points with labels 2, 3 placed before the clicks; coordinates are
        normalised by the video size and scaled to the model input; a previous prediction on the frame becomes the
        dense mask prompt (clamped to +-32)."""
        state = inference_state
        obj_idx = self._obj_id_to_idx(state, obj_id)
        point_inputs_per_frame = state["point_inputs_per_obj"][obj_idx]
        mask_inputs_per_frame = state["mask_inputs_per_obj"][obj_idx]
        if (points is not None) != (labels is not None):
            raise ValueError("points and labels must be provided together")
        if points is None and box is None:
            raise ValueError("at least one of points or box must be provided as input")
        if points is None:
            points = torch.zeros(0, 2, dtype=torch.float32)
        elif not isinstance(points, torch.Tensor):
            points = torch.tensor(points, dtype=torch.float32)
        if labels is None:
            labels = torch.zeros(0, dtype=torch.int32)
        elif not isinstance(labels, torch.Tensor):
            labels = torch.tensor(labels, dtype=torch.int32)
        if points.dim() == 2:
            points = points.unsqueeze(0)
        if labels.dim() == 1:
            labels = labels.unsqueeze(0)
        if box is not None:
            if not clear_old_points:
                raise ValueError("cannot add box without clearing old points, since box prompt must be provided before "
                                 "any point prompt (please use clear_old_points=True instead)")
            if not isinstance(box, torch.Tensor):
                box = torch.tensor(box, dtype=torch.float32, device=points.device)
            box_coords = box.reshape(1, 2, 2)
            box_labels = torch.tensor([2, 3], dtype=torch.int32, device=labels.device).reshape(1, 2)
            points = torch.cat([box_coords, points], dim=1)
            labels = torch.cat([box_labels, labels], dim=1)
        if normalize_coords:
            points = points / torch.tensor([state["video_width"], state["video_height"]]).to(points.device)
        points = points * self.image_size
        points = points.to(state["device"])
        labels = labels.to(state["device"])
        old = None if clear_old_points else point_inputs_per_frame.get(frame_idx, None)
        if old is None:
            point_inputs = {"point_coords": points, "point_labels": labels}
        else:
            point_inputs = {"point_coords": torch.cat([old["point_coords"], points], dim=1),
                            "point_labels": torch.cat([old["point_labels"], labels], dim=1)}
        point_inputs_per_frame[frame_idx] = point_inputs
        mask_inputs_per_frame.pop(frame_idx, None)
        tracked = state["frames_tracked_per_obj"][obj_idx]
        is_init_cond_frame = frame_idx not in tracked
        reverse = False if is_init_cond_frame else tracked[frame_idx]["reverse"]
        obj_out = state["output_dict_per_obj"][obj_idx]
        obj_tmp = state["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        prev_sam_mask_logits = None
        prev_out = obj_tmp[key].get(frame_idx)
        if prev_out is None:
            prev_out = obj_out["cond_frame_outputs"].get(frame_idx)
            if prev_out is None:
                prev_out = obj_out["non_cond_frame_outputs"].get(frame_idx)
        if prev_out is not None and prev_out["pred_masks"] is not None:
            prev_sam_mask_logits = torch.clamp(prev_out["pred_masks"].to(state["device"]), -32.0, 32.0)
        current_out, _ = self._run_single_frame_inference(
            state, obj_out, frame_idx, 1, is_init_cond_frame, point_inputs, None, reverse, run_mem_encoder=False,
            prev_sam_mask_logits=prev_sam_mask_logits)
        obj_tmp[key][frame_idx] = current_out
        cons = self._consolidate_temp_output_across_obj(state, frame_idx, is_cond, consolidate_at_video_res=True)
        _, video_res = self._get_orig_video_res_output(state, cons["pred_masks_video_res"])
        return frame_idx, state["obj_ids"], video_res

    @torch.inference_mode()
    def clear_all_prompts_in_frame(self, inference_state, frame_idx, obj_id, need_output=True):
        """Remove every input of one object on one frame; a conditioning output of that frame is downgraded to a
        non-conditioning one (REF saber/adapters/sam2/predictor.py:358-361 forwards here)."""
        state = inference_state
        obj_idx = self._obj_id_to_idx(state, obj_id)
        state["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        state["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        tmp = state["temp_output_dict_per_obj"]
        tmp[obj_idx]["cond_frame_outputs"].pop(frame_idx, None)
        tmp[obj_idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        obj_out = state["output_dict_per_obj"][obj_idx]
        out = obj_out["cond_frame_outputs"].pop(frame_idx, None)
        if out is not None:
            obj_out["non_cond_frame_outputs"][frame_idx] = out
            state["frames_tracked_per_obj"][obj_idx].pop(frame_idx, None)
        if not need_output:
            return None
        is_cond = any(frame_idx in d["cond_frame_outputs"] for d in tmp.values())
        cons = self._consolidate_temp_output_across_obj(state, frame_idx, is_cond, consolidate_at_video_res=True)
        _, video_res = self._get_orig_video_res_output(state, cons["pred_masks_video_res"])
        return frame_idx, state["obj_ids"], video_res

    @torch.inference_mode()
    def remove_object(self, inference_state, obj_id, strict=False, need_output=True):
        """Drop one object from the state and re-index the per-object containers (REF :363-366 forwards here)."""
        state = inference_state
        old_idx = state["obj_id_to_idx"].get(obj_id, None)
        updated_frames = []
        if old_idx is None:
            if not strict:
                return state["obj_ids"], updated_frames
            raise RuntimeError(f"Cannot remove object id {obj_id} as it doesn't exist. "
                               f"All existing object ids: {state['obj_ids']}.")
        if len(state["obj_id_to_idx"]) == 1:
            self.reset_state(state)
            return state["obj_ids"], updated_frames
        input_frames = set(state["point_inputs_per_obj"][old_idx]) | set(state["mask_inputs_per_obj"][old_idx])
        for frame_idx in input_frames:
            self.clear_all_prompts_in_frame(state, frame_idx, obj_id, need_output=False)
        old_obj_ids = state["obj_ids"]
        old_inds = list(range(len(old_obj_ids)))
        remain = [k for k in old_inds if k != old_idx]
        new_obj_ids = [old_obj_ids[k] for k in remain]
        new_inds = list(range(len(new_obj_ids)))
        old_to_new = dict(zip(remain, new_inds))
        state["obj_id_to_idx"] = dict(zip(new_obj_ids, new_inds))
        state["obj_idx_to_id"] = dict(zip(new_inds, new_obj_ids))
        state["obj_ids"] = new_obj_ids
        for name in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj",
                     "frames_tracked_per_obj"):
            container = state[name]
            kept = []
            for k in old_inds:
                v = container.pop(k)
                if k in old_to_new:
                    kept.append((old_to_new[k], v))
            container.update(kept)
        if need_output:
            tmp = state["temp_output_dict_per_obj"]
            for frame_idx in input_frames:
                is_cond = any(frame_idx in d["cond_frame_outputs"] for d in tmp.values())
                cons = self._consolidate_temp_output_across_obj(state, frame_idx, is_cond, consolidate_at_video_res=True)
                _, video_res = self._get_orig_video_res_output(state, cons["pred_masks_video_res"])
                updated_frames.append((frame_idx, video_res))
        return state["obj_ids"], updated_frames

    def _consolidate_temp_output_across_obj(self, state, frame_idx, is_cond, consolidate_at_video_res=False):
        B = self._get_obj_num(state)
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if consolidate_at_video_res:
            Hc, Wc, mk = state["video_height"], state["video_width"], "pred_masks_video_res"
        else:
            Hc = Wc = self.image_size // 4
            mk = "pred_masks"
        cons = {mk: torch.full((B, 1, Hc, Wc), NO_OBJ_SCORE, dtype=torch.float32, device=state["storage_device"])}
        for i in range(B):
            out = state["temp_output_dict_per_obj"][i][key].get(frame_idx, None)
            if out is None:
                out = state["output_dict_per_obj"][i]["cond_frame_outputs"].get(frame_idx, None)
            if out is None:
                out = state["output_dict_per_obj"][i]["non_cond_frame_outputs"].get(frame_idx, None)
            if out is None:
                continue
            om = out["pred_masks"]
            if om.shape[-2:] == cons[mk].shape[-2:]:
                cons[mk][i:i + 1] = om
            else:
                cons[mk][i:i + 1] = F.interpolate(om, size=cons[mk].shape[-2:], mode="bilinear", align_corners=False)
        return cons

    def _get_orig_video_res_output(self, state, any_res_masks):
        vh, vw = state["video_height"], state["video_width"]
        any_res_masks = any_res_masks.to(state["device"])
        if any_res_masks.shape[-2:] == (vh, vw):
            video_res = any_res_masks
        else:
            video_res = F.interpolate(any_res_masks, size=(vh, vw), mode="bilinear", align_corners=False)
        assert not self.non_overlap_masks
        return any_res_masks, video_res

    # ---- propagation ----------------------------------------------------------------------
    @torch.inference_mode()
    def propagate_in_video_preflight(self, state):
        B = self._get_obj_num(state)
        if B == 0:
            raise RuntimeError("No input points or masks are provided for any object; please add inputs first.")
        for i in range(B):
            obj_out = state["output_dict_per_obj"][i]
            obj_tmp = state["temp_output_dict_per_obj"][i]
            for is_cond in (False, True):
                key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
                for frame_idx, out in obj_tmp[key].items():
                    if out["maskmem_features"] is None:
                        hi = F.interpolate(out["pred_masks"].to(state["device"]), size=(self.image_size, self.image_size),
                                           mode="bilinear", align_corners=False)
                        out["maskmem_features"], out["maskmem_pos_enc"] = self._run_memory_encoder(
                            state, frame_idx, 1, hi, out["object_score_logits"], is_mask_from_pts=True)
                    obj_out[key][frame_idx] = out
                obj_tmp[key].clear()
            if len(obj_out["cond_frame_outputs"]) == 0:
                raise RuntimeError(f"No input points or masks are provided for object id {state['obj_idx_to_id'][i]}; "
                                   "please add inputs first.")
            for frame_idx in obj_out["cond_frame_outputs"]:
                obj_out["non_cond_frame_outputs"].pop(frame_idx, None)

    @torch.inference_mode()
    def propagate_in_video(self, inference_state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        state = inference_state
        self.propagate_in_video_preflight(state)
        obj_ids = state["obj_ids"]
        num_frames = state["num_frames"]
        B = self._get_obj_num(state)
        if start_frame_idx is None:
            start_frame_idx = min(t for d in state["output_dict_per_obj"].values() for t in d["cond_frame_outputs"])
        if max_frame_num_to_track is None:
            max_frame_num_to_track = num_frames
        if reverse:
            end = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end - 1, -1) if start_frame_idx > 0 else []
        else:
            end = min(start_frame_idx + max_frame_num_to_track, num_frames - 1)
            order = range(start_frame_idx, end + 1)
        for frame_idx in order:
            per_obj = [None] * B
            for i in range(B):
                obj_out = state["output_dict_per_obj"][i]
                if frame_idx in obj_out["cond_frame_outputs"]:
                    pred = obj_out["cond_frame_outputs"][frame_idx]["pred_masks"].to(state["device"])
                else:
                    cur, pred = self._run_single_frame_inference(state, obj_out, frame_idx, 1, False, None, None, reverse,
                                                                 run_mem_encoder=True)
                    obj_out["non_cond_frame_outputs"][frame_idx] = cur
                state["frames_tracked_per_obj"][i][frame_idx] = {"reverse": reverse}
                per_obj[i] = pred
            allp = torch.cat(per_obj, dim=0) if len(per_obj) > 1 else per_obj[0]
            _, video_res = self._get_orig_video_res_output(state, allp)
            yield frame_idx, obj_ids, video_res

    def _run_single_frame_inference(self, state, output_dict, frame_idx, batch_size, is_init_cond_frame, point_inputs,
                                    mask_inputs, reverse, run_mem_encoder, prev_sam_mask_logits=None):
        _, _, feats, pos, sizes = self._get_image_feature(state, frame_idx, batch_size)
        assert point_inputs is None or mask_inputs is None
        cur = self.track_step(frame_idx=frame_idx, is_init_cond_frame=is_init_cond_frame, current_vision_feats=feats,
                              current_vision_pos_embeds=pos, feat_sizes=sizes, point_inputs=point_inputs,
                              mask_inputs=mask_inputs, output_dict=output_dict, num_frames=state["num_frames"],
                              track_in_reverse=reverse, run_mem_encoder=run_mem_encoder,
                              prev_sam_mask_logits=prev_sam_mask_logits)
        sdev = state["storage_device"]
        mm = cur["maskmem_features"]
        if mm is not None:
            mm = mm.to(torch.bfloat16).to(sdev)
        pred_gpu = cur["pred_masks"]
        if self.fill_hole_area > 0:
            pred_gpu = fill_holes_in_mask_scores(pred_gpu, self.fill_hole_area)
        compact = {"maskmem_features": mm, "maskmem_pos_enc": self._get_maskmem_pos_enc(state, cur),
                   "pred_masks": pred_gpu.to(sdev), "obj_ptr": cur["obj_ptr"],
                   "object_score_logits": cur["object_score_logits"]}
        return compact, pred_gpu

    def _run_memory_encoder(self, state, frame_idx, batch_size, high_res_masks, object_score_logits, is_mask_from_pts):
        _, _, feats, _, sizes = self._get_image_feature(state, frame_idx, batch_size)
        mm, mpos = self._encode_new_memory(current_vision_feats=feats, feat_sizes=sizes,
                                           pred_masks_high_res=high_res_masks, object_score_logits=object_score_logits,
                                           is_mask_from_pts=is_mask_from_pts)
        mm = mm.to(torch.bfloat16).to(state["storage_device"])
        return mm, self._get_maskmem_pos_enc(state, {"maskmem_pos_enc": mpos})

    def _get_maskmem_pos_enc(self, state, cur):
        consts = state["constants"]
        o = cur["maskmem_pos_enc"]
        if o is None:
            return None
        if "maskmem_pos_enc" not in consts:
            assert isinstance(o, list)
            consts["maskmem_pos_enc"] = [x[0:1].clone() for x in o]
        B = o[0].size(0)
        return [x.expand(B, -1, -1, -1) for x in consts["maskmem_pos_enc"]]

    # ---- state ----------------------------------------------------------------------------
    @torch.inference_mode()
    def reset_state(self, inference_state):
        s = inference_state
        for k in ("obj_id_to_idx", "obj_idx_to_id", "obj_ids", "point_inputs_per_obj", "mask_inputs_per_obj",
                  "output_dict_per_obj", "temp_output_dict_per_obj", "frames_tracked_per_obj"):
            s[k].clear()


def build_sam2_video_predictor(config_file="large", ckpt_path=None, device="cpu", mode="eval", apply_postprocessing=True,
                               vos_optimized=False, state_dict=None, **kw):
    """Oracle twin of sam2.build_sam.build_sam2_video_predictor (hydra overrides listed in the module docstring)."""
    model = SAM2VideoPredictor(cfg=config_file, fill_hole_area=8, binarize_mask_from_pts_for_mem_enc=True,
                               dynamic_multimask_via_stability=bool(apply_postprocessing), **kw)
    _load(model, ckpt_path, state_dict)
    return model.to(device).eval()


def empty_inference_state(images: torch.Tensor, video_height: int, video_width: int, device) -> dict:
    """The state dict SABER builds (REF saber/adapters/sam2/predictor.py:130-150), for tests."""
    dev = torch.device(device)
    s = {"images": images, "num_frames": len(images), "offload_video_to_cpu": False, "offload_state_to_cpu": False,
         "video_height": video_height, "video_width": video_width, "device": dev, "storage_device": dev,
         "point_inputs_per_obj": {}, "mask_inputs_per_obj": {}, "cached_features": {}, "constants": {},
         "obj_id_to_idx": OrderedDict(), "obj_idx_to_id": OrderedDict(), "obj_ids": [], "output_dict_per_obj": {},
         "temp_output_dict_per_obj": {}, "frames_tracked_per_obj": {}}
    return s
