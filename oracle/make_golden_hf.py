"""Generates tests/golden/hf_decoder.npz and tests/golden/hf_video_tracking.npz from the independent HF ``transformers``
5.5.0 SAM2 implementation (present in the build image; NOT the reference, but the only other SAM2 arithmetic available
offline — SURVEY 8c). They pin the oracle restatement of upstream ``sam2`` for the rows round 1 left unpinned:

  hf_decoder.npz        — prompt encoder + mask decoder (U2, U3): point prompts (1 and 2 points, positive / negative /
                          box-corner labels), dense mask prompts, multimask and single-mask output with the dynamic
                          multimask-via-stability selection, IoU head, object-score head, output tokens.
  hf_video_tracking.npz — the video state machine (U6-U9): two objects seeded with masks on one frame, memory encoded,
                          tracked forwards and backwards with num_maskmem = 2; per-frame low-res mask logits and object
                          scores. HF differs from upstream in mechanics (inference-session object, batched memory
                          encoding, no hole filling) but not in arithmetic, so the oracle runs with fill_hole_area = 0.

Weights: ``torch.manual_seed(0)`` random init of the HF modules, translated to upstream names by oracle/hf_bridge.py; the
fixtures store only the seeds — the tests re-create the HF-initialised weights (transformers is part of the image).
Run in the build container:  python -m oracle.make_golden_hf
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def decoder_inputs(seed: int = 41):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(1, 256, 64, 64, generator=g) * 0.5
    s0 = torch.randn(1, 32, 256, 256, generator=g) * 0.5
    s1 = torch.randn(1, 64, 128, 128, generator=g) * 0.5
    # prompts: [P, n_points, 2] model-input pixel coordinates
    pts1 = torch.tensor([[[100.0, 200.0]], [[512.5, 512.5]], [[900.0, 30.0]], [[17.0, 1000.0]]])
    lab1 = torch.tensor([[1], [1], [0], [1]], dtype=torch.int32)
    pts2 = torch.tensor([[[100.0, 200.0], [300.0, 420.0]], [[600.0, 610.0], [50.0, 60.0]], [[200.0, 100.0], [800.0, 900.0]]])
    lab2 = torch.tensor([[1, 0], [1, 1], [2, 3]], dtype=torch.int32)  # last: a box given as its two corners
    mask_in = torch.randn(2, 1, 256, 256, generator=g) * 6
    mpts = torch.tensor([[[400.0, 400.0]], [[700.0, 300.0]]])
    mlab = torch.tensor([[1], [1]], dtype=torch.int32)
    return emb, s0, s1, (pts1, lab1), (pts2, lab2), (mask_in, mpts, mlab)


def tie_shared_pe(hf):
    """Upstream has ONE random-Fourier matrix (sam_prompt_encoder.pe_layer) for point prompts and the dense image PE; HF
    keeps two parameters that checkpoints fill with the same values but random init does not."""
    with torch.no_grad():
        hf.shared_image_embedding.positional_embedding.copy_(hf.prompt_encoder.shared_embedding.positional_embedding)


def make_decoder_golden():
    from transformers import Sam2Model

    from oracle.hf_bridge import hf_image_config, hf_to_upstream
    torch.manual_seed(0)
    hf = Sam2Model(hf_image_config("tiny")).eval()
    tie_shared_pe(hf)
    sd = hf_to_upstream(hf.state_dict())
    emb, s0, s1, (pts1, lab1), (pts2, lab2), (mask_in, mpts, mlab) = decoder_inputs()
    out = {}
    with torch.no_grad():
        pe = hf.get_image_wide_positional_embeddings()
        out["dense_pe_sub"] = pe[0, ::8, ::4, ::4].numpy().copy()

        def run(points, labels, masks, multimask, tag):
            P = points.shape[0]
            if masks is None:
                sparse, dense = hf.prompt_encoder(input_points=points[None], input_labels=labels[None], input_boxes=None,
                                                  input_masks=None)
                e, hi = emb, [s0, s1]
            else:  # HF's dense prompt is per batch entry: one prompt per entry
                sparse, dense = hf.prompt_encoder(input_points=points[:, None], input_labels=labels[:, None],
                                                  input_boxes=None, input_masks=masks)
                e, hi = emb.expand(P, -1, -1, -1), [s0.expand(P, -1, -1, -1), s1.expand(P, -1, -1, -1)]
            low, iou, tokens, obj = hf.mask_decoder(
                image_embeddings=e, image_positional_embeddings=pe.repeat(e.shape[0], 1, 1, 1),
                sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=multimask,
                high_resolution_features=hi)
            out[f"{tag}_sparse"] = sparse.reshape(P, -1, 256).numpy().copy()
            out[f"{tag}_low_sub"] = low.reshape(P, -1, 256, 256)[:, :, ::4, ::4].numpy().copy()
            out[f"{tag}_low_stats"] = np.array([low.mean().item(), low.std().item(), low.abs().max().item()])
            out[f"{tag}_iou"] = iou.reshape(P, -1).numpy().copy()
            out[f"{tag}_obj"] = obj.reshape(P, -1).numpy().copy()
            out[f"{tag}_tokens"] = tokens.reshape(P, -1, 256).numpy().copy()

        run(pts1, lab1, None, True, "p1_multi")
        run(pts1, lab1, None, False, "p1_single")
        run(pts2, lab2, None, True, "p2_multi")
        run(pts2, lab2, None, False, "p2_single")
        run(mpts, mlab, mask_in, False, "mask_single")
        run(mpts, mlab, mask_in, True, "mask_multi")
    np.savez_compressed(os.path.join(GOLD, "hf_decoder.npz"), weight_seed=0, input_seed=41, **out)
    print("hf_decoder.npz:", {k: v.shape for k, v in out.items() if k.endswith("_iou")})


VIDEO = dict(frames=4, seed_frame=1, fwd=2, weight_seed=0, input_seed=43)


def video_inputs():
    g = torch.Generator().manual_seed(VIDEO["input_seed"])
    T = VIDEO["frames"]
    base = torch.randn(1, 1, 64, 64, generator=g)
    frames = []
    for t in range(T):
        f = torch.nn.functional.interpolate(base + 0.3 * t * torch.randn(1, 1, 64, 64, generator=g), size=(1024, 1024),
                                            mode="bilinear", align_corners=False)
        frames.append(f[0].expand(3, -1, -1).clone())
    video = torch.stack(frames)  # [T,3,1024,1024]
    yy, xx = torch.meshgrid(torch.arange(1024.0), torch.arange(1024.0), indexing="ij")
    m1 = (((yy - 400) / 180) ** 2 + ((xx - 500) / 250) ** 2 <= 1).float()
    m2 = (((yy - 750) / 120) ** 2 + ((xx - 300) / 140) ** 2 <= 1).float()
    return video, [m1, m2]


def make_video_golden():
    from transformers import Sam2VideoConfig, Sam2VideoModel
    from transformers.models.sam2_video.modeling_sam2_video import Sam2VideoInferenceSession

    torch.manual_seed(VIDEO["weight_seed"])
    hf = Sam2VideoModel(Sam2VideoConfig(num_maskmem=2)).eval()
    tie_shared_pe(hf)
    video, masks = video_inputs()
    sess = Sam2VideoInferenceSession(video=video, video_height=1024, video_width=1024, dtype=torch.float32)
    k = VIDEO["seed_frame"]
    for obj_id, m in enumerate(masks, start=1):
        oi = sess.obj_id_to_idx(obj_id)
        sess.add_mask_inputs(oi, k, m[None, None])
        sess.obj_with_new_inputs.append(obj_id)
    out = {}
    with torch.no_grad():
        o = hf(sess, frame_idx=k)
        out[f"cond_f{k}"] = o.pred_masks.reshape(2, -1, o.pred_masks.shape[-1])[:, ::4, ::4].numpy().copy()
        out[f"cond_obj_f{k}"] = o.object_score_logits.reshape(-1).numpy().copy()
        for o in hf.propagate_in_video_iterator(sess, start_frame_idx=k, max_frame_num_to_track=VIDEO["fwd"], reverse=False):
            pm = o.pred_masks
            out[f"fwd_f{o.frame_idx}"] = pm.reshape(2, pm.shape[-2], pm.shape[-1])[:, ::4, ::4].numpy().copy()
            out[f"fwd_obj_f{o.frame_idx}"] = o.object_score_logits.reshape(-1).numpy().copy()
        for o in hf.propagate_in_video_iterator(sess, start_frame_idx=k, max_frame_num_to_track=1, reverse=True):
            pm = o.pred_masks
            out[f"bwd_f{o.frame_idx}"] = pm.reshape(2, pm.shape[-2], pm.shape[-1])[:, ::4, ::4].numpy().copy()
            out[f"bwd_obj_f{o.frame_idx}"] = o.object_score_logits.reshape(-1).numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "hf_video_tracking.npz"), **{k2: np.asarray(v) for k2, v in VIDEO.items()}, **out)
    print("hf_video_tracking.npz:", {k2: v.shape for k2, v in out.items()})


def main():
    os.makedirs(GOLD, exist_ok=True)
    make_decoder_golden()
    make_video_golden()


if __name__ == "__main__":
    main()
