"""Shared synthetic generators for the tests (network-independent logits, masks, images)."""
import numpy as np


def synth_logits(n, seed, S=256, edge_frac=0.25):
    """[n,S,S] fp32 low-res logit planes: soft blobs of varying steepness (=> a spread of stability scores) plus
    low-amplitude noise; about edge_frac of them hug the plane border (near-crop-edge filter)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    planes = np.empty((n, S, S), np.float32)
    for i in range(n):
        if rng.uniform() < edge_frac:
            cy, cx = rng.choice([2.0, S - 3.0]), rng.uniform(0, S)
        else:
            cy, cx = rng.uniform(0.2 * S, 0.8 * S), rng.uniform(0.2 * S, 0.8 * S)
        ry, rx = rng.uniform(0.04 * S, 0.18 * S), rng.uniform(0.04 * S, 0.18 * S)
        amp = rng.choice([6.0, 20.0, 60.0, 120.0])
        d = np.sqrt(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2)
        planes[i] = amp * (1 - d) + 0.15 * rng.normal(size=(S, S))
        if i % 11 == 10:
            planes[i] = -5.0 - np.abs(planes[i])  # an empty mask
    return planes


def synth_boxes(n, seed, extent=900):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, extent, (n, 2))
    wh = rng.uniform(5, 300, (n, 2))
    boxes = np.round(np.concatenate([xy, xy + wh], 1)).astype(np.int32)
    scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)
    return boxes, scores


def oracle_amg_from_captures(caps, hw, *, points_per_side, crop_n_layers, crop_n_points_downscale_factor,
                             pred_iou_thresh, stability_score_thresh, stability_score_offset, box_nms_thresh,
                             multimask_output, mask_threshold=0.0, crop_nms_thresh=0.7, crop_overlap_ratio=512 / 1500):
    """Upstream AMG post-decoder pipeline evaluated by the oracle on captured low-res logits (the 'given
    identical logits' checker): mask_post -> per-crop NMS -> cross-crop NMS -> output dicts."""
    import torch
    from oracle import amg_post_ref as R
    from oracle.sam2_ref import amg as up
    crop_boxes, layer_idxs = up.generate_crop_boxes(hw, crop_n_layers, crop_overlap_ratio)
    grids = up.build_all_layer_point_grids(points_per_side, crop_n_layers, crop_n_points_downscale_factor)
    cpp_total = 3 if multimask_output else 1
    recs_all = []
    for k, (cb, layer) in enumerate(zip(crop_boxes, layer_idxs)):
        x0, y0, x1, y1 = cb
        pts64 = grids[layer] * np.array([y1 - y0, x1 - x0])[None, ::-1]
        pts = torch.as_tensor(pts64, dtype=torch.float32)
        pts_full = (pts + torch.tensor([[x0, y0]])).numpy()
        mine = sorted([c for c in caps if c["crop"] == k], key=lambda c: c["base"])
        crop_base = mine[0]["base"]
        recs = []
        for c in mine:
            n, cpp = c["n"], c["cpp"]
            prompt = np.arange(n) // cpp
            token = c["sel"][prompt] if c["sel"] is not None else (1 + np.arange(n) % 3 if cpp == 3 else np.zeros(n, int))
            planes = c["planes"] if c["planes"].ndim == 3 else c["planes"][prompt, token]
            ious = c["ious4"][prompt, token]
            r = R.mask_post(planes, ious, cb, hw, pred_iou_thresh, mask_threshold, stability_score_offset,
                            stability_score_thresh, only_iou_survivors=True)
            row_of = {int(s_): j for j, s_ in enumerate(r["rows"])} if "rows" in r else None
            for i in np.flatnonzero(r["keep"]):
                slot = c["base"] + i
                seg = r["masks"][i] if row_of is None else r["masks_rows"][row_of[int(i)]]
                recs.append(dict(segmentation=seg, area=int(r["area"][i]), bbox_xyxy=r["bbox"][i],
                                 predicted_iou=float(ious[i]), stability_score=float(r["stability"][i]),
                                 point_coords=[pts_full[(slot - crop_base) // cpp_total].tolist()],
                                 crop_box=[x0, y0, x1 - x0, y1 - y0], crop_xyxy=cb))
        if recs:
            boxes = np.stack([r["bbox_xyxy"] for r in recs]).astype(np.float32)
            keep = R.nms(boxes, np.array([r["predicted_iou"] for r in recs], np.float32), box_nms_thresh)
            recs = [recs[i] for i in keep]
        recs_all.extend(recs)
    if len(crop_boxes) > 1 and recs_all:
        cbx = torch.tensor([r["crop_xyxy"] for r in recs_all]).float()
        scores = (1 / ((cbx[:, 2] - cbx[:, 0]) * (cbx[:, 3] - cbx[:, 1]))).numpy()
        boxes = np.stack([r["bbox_xyxy"] for r in recs_all]).astype(np.float32)
        keep = R.nms(boxes, scores, crop_nms_thresh)
        recs_all = [recs_all[i] for i in keep]
    out = []
    for r in recs_all:
        bx0, by0, bx1, by1 = (int(v) for v in r["bbox_xyxy"])
        out.append(dict(segmentation=r["segmentation"], area=r["area"], bbox=[bx0, by0, bx1 - bx0, by1 - by0],
                        predicted_iou=r["predicted_iou"], point_coords=r["point_coords"],
                        stability_score=r["stability_score"], crop_box=r["crop_box"]))
    return out


def assert_mask_lists_equal(got, want):
    assert len(got) == len(want), (len(got), len(want))
    for i, (g, w) in enumerate(zip(got, want)):
        for key in ("area", "bbox", "crop_box", "point_coords"):
            assert g[key] == w[key], (i, key, g[key], w[key])
        assert np.float32(g["predicted_iou"]) == np.float32(w["predicted_iou"]), (i, g["predicted_iou"], w["predicted_iou"])
        gs, wsb = np.float32(g["stability_score"]), np.float32(w["stability_score"])
        assert gs == wsb or (np.isnan(gs) and np.isnan(wsb)), (i, gs, wsb)
        np.testing.assert_array_equal(g["segmentation"], w["segmentation"], err_msg=f"mask {i}")
