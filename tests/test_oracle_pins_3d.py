"""Pins the 3-D-path restatements of oracle/saber_ref.py against golden vectors produced by the REFERENCE's own code
(oracle/make_golden_3d.py: saber.filters.gaussian / estimate_thickness / masks, saber.analysis.refine_membranes run
unmodified with their unused third-party imports stubbed), and oracle/sam2_ref/memory.py against the independent HF
transformers Sam2VideoModel. skimage.transform.resize has no reference-run pin (skimage is absent): its restatement is
checked against torch's independent bilinear kernel in the interior and on the documented order-0 rule."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from scipy import ndimage as ndi

from oracle import saber_ref
from saber_b200 import synth


def test_gaussian_z_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "saber3d_gaussian_z.npz"))
    vol = synth.make_tomogram(tuple(g["shape"]), seed=int(g["seed"]), n_ellipsoids=5).numpy()
    out = saber_ref.gaussian_smoothing(vol, 5, dim=0)
    assert out.shape == vol.shape and out.dtype == np.float32
    np.testing.assert_allclose(out, g["out"], rtol=0, atol=2e-6)
    k = saber_ref.make_gaussian_kernel(5)
    assert k.size == 15 and abs(k.sum() - 1) < 1e-6


def test_fast_gauss3d_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "saber3d_fast_gauss3d.npz"))
    out = saber_ref.fast_3d_gaussian_smoothing(g["labels_in"], scale=0.075)
    assert out.dtype == np.uint8
    assert len(np.unique(g["labels_in"])) > 2
    np.testing.assert_array_equal(out, g["out"])


def test_fit_boundaries_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "saber3d_fit_boundaries.npz"))
    out = saber_ref.fit_organelle_boundaries(g["frame_scores"].copy())
    np.testing.assert_allclose(out, g["out"], rtol=1e-9, atol=1e-12)
    assert out[:, 0].max() > 1 and np.all(out[:, 3] == 0)


def test_classifier_mask_conversion_matches_reference_golden(golden_dir):
    from oracle.make_golden import synth_mask_list
    g = np.load(os.path.join(golden_dir, "saber3d_classifier_masks.npz"))
    masks = synth_mask_list((96, 128), 12, seed=24)
    inst = saber_ref.convert_predictions_to_masks(g["preds"], [dict(m) for m in masks], desired_class=1, min_mask_area=32)
    np.testing.assert_array_equal(saber_ref.masks_to_array(inst), g["inst_array"])
    np.testing.assert_allclose([m["predicted_iou"] for m in inst], g["inst_conf"], rtol=1e-6)
    sem = saber_ref.convert_predictions_to_masks(g["preds"], [dict(m) for m in masks], desired_class=0, min_mask_area=32)
    np.testing.assert_array_equal(np.stack([m["segmentation"].astype(np.uint8) for m in sem]), g["sem"])
    np.testing.assert_array_equal([m["area"] for m in sem], g["sem_area"])


def test_morphology_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "saber3d_morphology.npz"))
    roi = g["roi"].astype(np.float32)
    for r in (1, 2, 3):
        np.testing.assert_array_equal(saber_ref.binary_erosion_ball(roi, r).astype(np.uint8), g[f"erode{r}"])
        np.testing.assert_array_equal(saber_ref.binary_dilation_ball(roi, r).astype(np.uint8), g[f"dilate{r}"])
        np.testing.assert_array_equal(saber_ref.morphological_opening_ball(roi, r).astype(np.uint8), g[f"open{r}"])
    assert g["erode1"].sum() > 0 and g["open2"].sum() < g["roi"].sum()


def test_skimage_resize_restatement_properties():
    rng = np.random.default_rng(0)
    img = rng.normal(size=(58, 60)).astype(np.float32)
    up = saber_ref.skimage_resize(img, (64, 64), anti_aliasing=True)  # up-sampling: sigma = 0, zoom only
    assert up.dtype == np.float32 and up.shape == (64, 64)
    ref = F.interpolate(torch.from_numpy(img)[None, None], (64, 64), mode="bilinear", align_corners=False)[0, 0].numpy()
    np.testing.assert_allclose(up[2:-2, 2:-2], ref[2:-2, 2:-2], atol=2e-6)  # interior = pixel-centre bilinear
    assert np.abs(up - ref).max() > 1e-3  # border differs: mirror about the edge pixel centre (Appendix A1)
    same = saber_ref.skimage_resize(img, (58, 60), anti_aliasing=True)
    np.testing.assert_allclose(same, img, atol=1e-6)
    down = saber_ref.skimage_resize(img, (29, 30), anti_aliasing=True)  # sigma = 0.5 Gaussian then zoom
    blur = ndi.gaussian_filter(img, 0.5, mode="mirror")
    np.testing.assert_allclose(down, 0.25 * (blur[0::2, 0::2] + blur[1::2, 0::2] + blur[0::2, 1::2] + blur[1::2, 1::2]),
                               atol=1e-5)
    m = rng.random((64, 64)) > 0.5
    near = saber_ref.skimage_resize(m, (58, 60), order=0, anti_aliasing=False)
    iy = np.floor((np.arange(58) + 0.5) * 64 / 58).astype(int)
    ix = np.floor((np.arange(60) + 0.5) * 64 / 60).astype(int)
    assert near.dtype == bool
    np.testing.assert_array_equal(near, m[np.ix_(iy, ix)])


def test_load_grayscale_and_normalize():
    vol = synth.make_tomogram((3, 58, 60), seed=4, n_ellipsoids=2).numpy()
    nv = saber_ref.normalize_tomogram(vol)
    assert nv.min() == -1 and nv.max() == 1
    images, vh, vw = saber_ref.load_grayscale_image_array(nv, 64)
    assert images.shape == (3, 3, 64, 64) and (vh, vw) == (64, 64)
    assert images.min() >= -3 - 1e-6 and images.max() <= 1 + 1e-6
    np.testing.assert_array_equal(images[:, 0], images[:, 2])


def test_oracle_memory_modules_match_hf_golden(golden_dir):
    """oracle MemoryAttention / MemoryEncoder with HF's weights reproduce HF's outputs (seeded like make_golden_3d)."""
    from transformers import Sam2VideoConfig, Sam2VideoModel

    from oracle.hf_bridge import hf_to_upstream
    from oracle.sam2_ref.sam2_base import SAM2Base
    g = np.load(os.path.join(golden_dir, "hf_video_memory.npz"))
    torch.manual_seed(int(g["weight_seed"]))
    hf = Sam2VideoModel(Sam2VideoConfig(num_maskmem=2))
    sd = hf_to_upstream(hf.state_dict())
    del hf
    m = SAM2Base("tiny", num_maskmem=2)
    sub = {k: v for k, v in sd.items() if k.startswith(("memory_attention.", "memory_encoder."))}
    missing, unexpected = m.load_state_dict(sub, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith(("memory_attention.", "memory_encoder."))]
    m.eval()
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    curr = torch.randn(4096, 1, 256, generator=gen) * 0.5
    curr_pos = torch.randn(4096, 1, 256, generator=gen) * 0.5
    n_ptr = int(g["n_ptr"])
    memory = torch.randn(2 * 4096 + n_ptr, 1, 64, generator=gen) * 0.5
    memory_pos = torch.randn(2 * 4096 + n_ptr, 1, 64, generator=gen) * 0.5
    with torch.no_grad():
        att = m.memory_attention(curr=curr, curr_pos=curr_pos, memory=memory, memory_pos=memory_pos,
                                 num_obj_ptr_tokens=n_ptr)
        pix = torch.randn(1, 256, 64, 64, generator=gen) * 0.5
        msk = torch.randn(1, 1, 1024, 1024, generator=gen) * 4
        out = m.memory_encoder(pix, torch.sigmoid(msk) * 20 - 10, skip_mask_sigmoid=True)
    att = att.reshape(4096, 256)
    np.testing.assert_allclose(att[::16, ::4].numpy(), g["att_sub"], atol=2e-4, rtol=1e-4)
    assert abs(float(att.std()) - float(g["att_std"])) < 1e-4
    np.testing.assert_allclose(out["vision_features"][0, ::2, ::4, ::4].numpy(), g["mm_sub"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(out["vision_pos_enc"][0][0, ::2, ::4, ::4].numpy(), g["mpos_sub"], atol=2e-6)


def test_oracle_video_state_machine_matches_hf_tracking_golden(golden_dir):
    """U6-U9 pin of the whole tracking recurrence: two objects seeded with masks on frame 1 of a 4-frame video, memory
    encoded (binarised prompt masks), tracked forwards (frames 1-3) and backwards (1-0) with num_maskmem = 2 — per
    frame the low-res mask logits and object scores of oracle.sam2_ref.video_predictor vs HF Sam2VideoModel on the
    same HF-initialised weights (oracle/make_golden_hf.py). This exercises what the module-level pins cannot: which
    stored outputs feed which memory slot, the temporal position rows, the object-pointer selection (incl. forward
    pointers picked up by the reverse pass), mask-as-output conditioning, bf16 memory storage. fp32 recurrence with
    bf16-stored memories on both sides: 5e-3 abs on logits of magnitude ~1e1."""
    from transformers import Sam2VideoConfig, Sam2VideoModel

    from oracle.hf_bridge import hf_to_upstream
    from oracle.make_golden_hf import tie_shared_pe, video_inputs
    from oracle.sam2_ref.video_predictor import SAM2VideoPredictor, empty_inference_state
    g = np.load(os.path.join(golden_dir, "hf_video_tracking.npz"))
    torch.manual_seed(int(g["weight_seed"]))
    hf = Sam2VideoModel(Sam2VideoConfig(num_maskmem=2)).eval()
    tie_shared_pe(hf)
    sd = hf_to_upstream(hf.state_dict())
    orc = SAM2VideoPredictor(cfg="tiny", num_maskmem=2, fill_hole_area=0, binarize_mask_from_pts_for_mem_enc=True,
                             dynamic_multimask_via_stability=True).eval()
    missing, unexpected = orc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected[:5]
    assert not [m for m in missing if not m.startswith("image_encoder.trunk.pos_embed")], missing[:5]
    video, masks = video_inputs()
    st = empty_inference_state(video, 1024, 1024, "cpu")
    k = int(g["seed_frame"])
    for obj_id, m in enumerate(masks, start=1):
        orc.add_new_mask(st, k, obj_id, m.bool())

    def low(frame, key):
        return torch.cat([st["output_dict_per_obj"][i][key][frame]["pred_masks"] for i in range(2)])[:, 0]

    def obj(frame, key):
        return torch.cat([st["output_dict_per_obj"][i][key][frame]["object_score_logits"].reshape(-1) for i in range(2)])

    def check(tag, frame, key):
        want, got = g[f"{tag}_f{frame}"], low(frame, key)[:, ::4, ::4].numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-3 * max(1.0, np.abs(want).max() / 10), err_msg=f"{tag} f{frame}")
        np.testing.assert_allclose(obj(frame, key).numpy(), g[f"{tag}_obj_f{frame}"], rtol=0, atol=5e-3, err_msg=f"{tag} obj f{frame}")

    for f, ids, logits in orc.propagate_in_video(st, start_frame_idx=k, max_frame_num_to_track=int(g["fwd"]), reverse=False):
        assert list(ids) == [1, 2] and logits.shape == (2, 1, 1024, 1024)
    check("cond", k, "cond_frame_outputs")
    check("fwd", k, "cond_frame_outputs")
    for f in (k + 1, k + 2):
        check("fwd", f, "non_cond_frame_outputs")
    for f, ids, logits in orc.propagate_in_video(st, start_frame_idx=k, max_frame_num_to_track=1, reverse=True):
        pass
    check("bwd", k, "cond_frame_outputs")
    check("bwd", k - 1, "non_cond_frame_outputs")
