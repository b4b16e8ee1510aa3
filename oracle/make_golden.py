"""Generates tests/golden/*.npz by running the REFERENCE's own importable functions
(/root/reference/saber/utils/preprocessing.py, /root/reference/saber/segmenters/utils.py) and the independent HF
transformers SAM2 implementation on seeded inputs. Run in the build container only (the GPU box has no
/root/reference): ``python -m oracle.make_golden``. The fixtures pin oracle/saber_ref.py and oracle/sam2_ref/*.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from saber_b200 import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_mask_list(hw, n, seed):
    """Overlapping blobs with engineered near-duplicates (for remove_duplicate_masks)."""
    rng = np.random.default_rng(seed)
    H, W = hw
    yy, xx = np.mgrid[0:H, 0:W]
    masks = []
    for k in range(n):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        ry, rx = rng.uniform(6, 40), rng.uniform(6, 40)
        seg = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        masks.append(seg)
        if k % 3 == 0:  # a near-duplicate: same blob grown/shrunk by a sliver
            grow = rng.uniform(0.97, 1.03)
            masks.append(((yy - cy) / (ry * grow)) ** 2 + ((xx - cx) / (rx * grow)) ** 2 <= 1)
    out = []
    for i, seg in enumerate(masks):
        out.append({"segmentation": seg, "area": int(seg.sum()), "stability_score": float(np.round(rng.uniform(0.9, 1.0), 3))})
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    sys.path.insert(0, "/root/reference")
    from saber.segmenters import utils as ref_utils  # the reference's own code
    from saber.utils import preprocessing as ref_prep

    # ---- prepare (R4)
    img = synth.make_tomogram((1, 600, 640), seed=3, n_ellipsoids=12)[0].numpy()
    out = ref_prep.prepare(img, to_rgb=True)
    np.savez_compressed(os.path.join(GOLD, "saber_prepare.npz"), seed=3, shape=np.array([600, 640]),
                        out_sha=sha(out), out_sub=out[::8, ::8, 0].copy(), in_sha=sha(img))
    # ---- separate_masks (R11)
    for name, shape, seed, n, speckle, mma in [("a", (40, 96, 112), 5, 30, 0.002, 5), ("b", (24, 64, 64), 6, 12, 0.0, 100)]:
        vol = synth.make_label_volume(shape, seed=seed, n_ellipsoids=n, speckle=speckle).numpy().view(np.uint16)
        lab = ref_utils.separate_masks(vol, min_mask_area=mma)
        np.savez_compressed(os.path.join(GOLD, f"saber_separate_masks_{name}.npz"), shape=np.array(shape), seed=seed,
                            n=n, speckle=speckle, min_mask_area=mma, in_sha=sha(vol), labels=lab)
    # ---- remove_duplicate_masks (R10)
    masks = synth_mask_list((160, 200), 24, seed=9)
    kept = ref_utils.remove_duplicate_masks([dict(m) for m in masks])
    kept_idx = [next(i for i, m in enumerate(masks) if m["segmentation"] is k["segmentation"]) for k in kept]
    np.savez_compressed(os.path.join(GOLD, "saber_remove_duplicates.npz"), seed=9, hw=np.array([160, 200]), n=24,
                        kept=np.array(kept_idx), areas=np.array([m["area"] for m in masks]))
    # ---- HF SAM2 (independent implementation) on the tiny architecture: encoder + decoder vectors
    from oracle.hf_bridge import hf_image_config, hf_to_upstream
    from transformers import Sam2Model
    torch.manual_seed(0)
    hf = Sam2Model(hf_image_config("tiny")).eval()
    sd = hf_to_upstream(hf.state_dict())
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3, 1024, 1024, generator=g)
    with torch.no_grad():
        feats = hf.get_image_features(x)
    fpn = feats["fpn_hidden_states"] if isinstance(feats, dict) else feats[0]
    emb = fpn[-1][:, 0]  # [4096,256] token-major, before no_mem_embed
    s1 = fpn[1][:, 0]    # [16384,64] (conv_s1 applied)
    np.savez_compressed(os.path.join(GOLD, "hf_tiny_encoder.npz"), weight_seed=0, input_seed=11,
                        embed_sub=emb[::16, ::4].numpy().copy(), s1_sub=s1[::64, ::4].numpy().copy(),
                        embed_mean=float(emb.mean()), embed_std=float(emb.std()))
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
