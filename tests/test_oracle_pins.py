"""Pins the oracle: restatements in oracle/ are checked against (a) golden vectors produced by the
reference's own importable functions (oracle/make_golden.py), (b) torch / torchvision CPU kernels, and
(c) the independent HF transformers SAM2 implementation."""
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import amg_post_ref, saber_ref
from saber_b200 import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_prepare_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "saber_prepare.npz"))
    img = synth.make_tomogram((1, 600, 640), seed=int(g["seed"]), n_ellipsoids=12)[0].numpy()
    assert sha(img) == str(g["in_sha"]), "synthetic input generator drifted"
    out = saber_ref.prepare(img, to_rgb=True)
    assert out.dtype == np.float32 and out.shape == (600, 640, 3)
    assert sha(out) == str(g["out_sha"])
    np.testing.assert_array_equal(out[::8, ::8, 0], g["out_sub"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_separate_masks_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"saber_separate_masks_{name}.npz"))
    vol = synth.make_label_volume(tuple(g["shape"]), seed=int(g["seed"]), n_ellipsoids=int(g["n"]),
                                  speckle=float(g["speckle"])).numpy().view(np.uint16)
    assert sha(vol) == str(g["in_sha"])
    lab = saber_ref.separate_masks(vol, min_mask_area=int(g["min_mask_area"]))
    assert lab.dtype == np.uint32
    np.testing.assert_array_equal(lab, g["labels"])


def test_remove_duplicates_matches_reference_golden(golden_dir):
    from oracle.make_golden import synth_mask_list
    g = np.load(os.path.join(golden_dir, "saber_remove_duplicates.npz"))
    masks = synth_mask_list(tuple(g["hw"]), int(g["n"]), seed=int(g["seed"]))
    np.testing.assert_array_equal(np.array([m["area"] for m in masks]), g["areas"])
    kept = saber_ref.remove_duplicate_masks(masks)
    idx = [next(i for i, m in enumerate(masks) if m is k) for k in kept]
    np.testing.assert_array_equal(np.array(idx), g["kept"])
    assert len(idx) < len(masks), "fixture should contain duplicates"


@pytest.mark.parametrize("hw", [(1024, 1024), (683, 684), (342, 341), (200, 333), (256, 256)])
def test_bilinear_restatement_is_bitwise_torch_cpu(hw):
    rng = np.random.default_rng(hw[0])
    p = (rng.normal(size=(3, 256, 256)) * 4).astype(np.float32)
    a = amg_post_ref.upsample_bilinear(p, hw)
    b = F.interpolate(torch.from_numpy(p)[None], hw, mode="bilinear", align_corners=False)[0].numpy()
    np.testing.assert_array_equal(a, b)


def test_nms_restatement_matches_torchvision():
    import torchvision
    rng = np.random.default_rng(2)
    for n in (1, 7, 300, 1500):
        xy = rng.uniform(0, 900, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 300, (n, 2)).astype(np.float32)
        boxes = np.round(np.concatenate([xy, xy + wh], 1))
        scores = np.round(rng.uniform(0, 1, n), 2).astype(np.float32)  # many ties
        k1 = amg_post_ref.nms(boxes, scores, 0.7)
        k2 = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.7).numpy()
        np.testing.assert_array_equal(k1, k2)


def test_mask_post_restatement_matches_upstream_sequence():
    """amg_post_ref.mask_post == the upstream torch op sequence restated in oracle/sam2_ref/amg.py."""
    from oracle.sam2_ref import amg as up
    from util import synth_logits
    rng = np.random.default_rng(4)
    n, crop, hw = 24, (100, 40, 612, 400), (480, 640)
    x0, y0, x1, y1 = crop
    planes = synth_logits(n, seed=4)
    ious = rng.uniform(0.5, 1, n).astype(np.float32)
    r = amg_post_ref.mask_post(planes, ious, crop, hw, 0.7, 0.0, 0.7, 0.92)
    m = F.interpolate(torch.from_numpy(planes)[:, None], (y1 - y0, x1 - x0), mode="bilinear", align_corners=False)[:, 0]
    stab = up.calculate_stability_score(m, 0.0, 0.7)
    np.testing.assert_array_equal(r["stability"], stab.numpy())
    binm = m > 0.0
    boxes = up.batched_mask_to_box(binm)
    near = up.is_box_near_crop_edge(boxes, list(crop), [0, 0, hw[1], hw[0]])
    keep = (torch.from_numpy(ious) > 0.7) & (stab >= 0.92) & ~near
    np.testing.assert_array_equal(r["keep"], keep.numpy())
    np.testing.assert_array_equal(r["bbox"], up.uncrop_boxes_xyxy(boxes, list(crop)).numpy())
    np.testing.assert_array_equal(r["masks"], up.uncrop_masks(binm, list(crop), hw[0], hw[1]).numpy())
    assert r["keep"].any() and not r["keep"].all()


def test_oracle_encoder_matches_hf_golden(golden_dir):
    """oracle.sam2_ref Hiera+FPN (tiny) vs the HF transformers output frozen in tests/golden (same weights)."""
    from transformers import Sam2Model
    from oracle.hf_bridge import hf_image_config, hf_to_upstream
    from oracle.sam2_ref.sam2_base import SAM2Base
    g = np.load(os.path.join(golden_dir, "hf_tiny_encoder.npz"))
    torch.manual_seed(int(g["weight_seed"]))
    hf = Sam2Model(hf_image_config("tiny")).eval()
    sd = hf_to_upstream(hf.state_dict())
    orc = SAM2Base("tiny").eval()
    missing, unexpected = orc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected[:5]
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    x = torch.randn(1, 3, 1024, 1024, generator=gen)
    with torch.no_grad():
        bo = orc.forward_image(x)
        _, vf, _, _ = orc._prepare_backbone_features(bo)
    emb = vf[2][:, 0]
    s1 = vf[1][:, 0]
    np.testing.assert_allclose(emb[::16, ::4].numpy(), g["embed_sub"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(s1[::64, ::4].numpy(), g["s1_sub"], rtol=0, atol=2e-5)


@pytest.fixture(scope="module")
def hf_tiny_oracle():
    """HF-initialised (seed 0) tiny weights loaded into the oracle's SAM2Base (dynamic multimask as upstream's
    apply_postprocessing builds it)."""
    from transformers import Sam2Model
    from oracle.hf_bridge import hf_image_config, hf_to_upstream
    from oracle.sam2_ref.sam2_base import SAM2Base
    from oracle.make_golden_hf import tie_shared_pe
    torch.manual_seed(0)
    hf = Sam2Model(hf_image_config("tiny")).eval()
    tie_shared_pe(hf)
    orc = SAM2Base("tiny", dynamic_multimask_via_stability=True).eval()
    missing, unexpected = orc.load_state_dict(hf_to_upstream(hf.state_dict()), strict=False)
    assert not unexpected, unexpected[:5]
    return orc


@pytest.mark.parametrize("tag,n_pts,use_mask,multi", [("p1_multi", 1, False, True), ("p1_single", 1, False, False),
                                                     ("p2_multi", 2, False, True), ("p2_single", 2, False, False),
                                                     ("mask_single", 1, True, False), ("mask_multi", 1, True, True)])
def test_oracle_prompt_encoder_and_mask_decoder_match_hf_golden(golden_dir, hf_tiny_oracle, tag, n_pts, use_mask, multi):
    """U2 / U3 pin: prompt encoder (points with labels 0-3 and the pad point, dense mask prompt, dense PE) and mask
    decoder (two-way transformer, up-scaling with high-res skips, hyper-networks, IoU head, object-score head, multimask
    selection incl. dynamic multimask via stability) of oracle.sam2_ref vs the independent HF implementation on the same
    weights and inputs (oracle/make_golden_hf.py). fp32 on both sides: 1e-4 on logits of magnitude ~1e1."""
    from oracle.make_golden_hf import decoder_inputs
    g = np.load(os.path.join(golden_dir, "hf_decoder.npz"))
    orc = hf_tiny_oracle
    emb, s0, s1, p1, p2, (mask_in, mpts, mlab) = decoder_inputs(int(g["input_seed"]))
    pts, lab = (mpts, mlab) if use_mask else (p1 if n_pts == 1 else p2)
    P = pts.shape[0]
    with torch.no_grad():
        pe = orc.sam_prompt_encoder.get_dense_pe()
        np.testing.assert_allclose(pe[0, ::8, ::4, ::4].numpy(), g["dense_pe_sub"], rtol=0, atol=1e-5)
        sparse, dense = orc.sam_prompt_encoder(points=(pts, lab), boxes=None, masks=(mask_in if use_mask else None))
        np.testing.assert_allclose(sparse.numpy(), g[f"{tag}_sparse"], rtol=0, atol=1e-5)
        low, iou, tokens, obj = orc.sam_mask_decoder(
            image_embeddings=emb.expand(P, -1, -1, -1), image_pe=pe, sparse_prompt_embeddings=sparse,
            dense_prompt_embeddings=dense, multimask_output=multi, repeat_image=False,
            high_res_features=[s0.expand(P, -1, -1, -1), s1.expand(P, -1, -1, -1)])
    scale = float(g[f"{tag}_low_stats"][2])
    np.testing.assert_allclose(low[:, :, ::4, ::4].numpy(), g[f"{tag}_low_sub"], rtol=0, atol=1e-4 * max(1.0, scale))
    np.testing.assert_allclose(iou.numpy(), g[f"{tag}_iou"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(obj.numpy(), g[f"{tag}_obj"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(tokens.numpy(), g[f"{tag}_tokens"], rtol=0, atol=1e-4)
