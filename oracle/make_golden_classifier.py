"""Pins oracle/classifier_ref.py::crop_and_resize_adaptive against the REFERENCE's own function
(/root/reference/saber/classifier/datasets/RandMaskCrop.py, imported with monai stubbed) and writes
tests/golden/saber_classifier_crops.npz. Build container only: ``python -m oracle.make_golden_classifier``."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    from oracle.make_golden import synth_mask_list
    rng = np.random.default_rng(41)
    img = torch.from_numpy(rng.normal(size=(1, 300, 340)).astype(np.float32))
    masks = [m["segmentation"] for m in synth_mask_list((300, 340), 10, seed=42)]
    masks.append(np.zeros((300, 340), bool))
    masks.append(np.ones((300, 340), bool))
    m2 = np.zeros((300, 340), bool)
    m2[5, 7] = True
    masks.append(m2)
    return img, masks


def main():
    from oracle.make_golden_3d import _Stub, install_stubs
    install_stubs()
    import importlib.machinery
    if "monai.transforms" not in sys.modules:
        m = _Stub("monai.transforms")
        m.__spec__ = importlib.machinery.ModuleSpec("monai.transforms", None)
        m.__path__ = []
        sys.modules["monai.transforms"] = m
        setattr(sys.modules["monai"], "transforms", m)

    class MapTransform:  # the reference subclasses it at import time
        def __init__(self, keys, allow_missing_keys=False):
            self.keys = keys

    sys.modules["monai.transforms"].MapTransform = MapTransform
    sys.path.insert(0, "/root/reference")
    from saber.classifier.datasets.RandMaskCrop import crop_and_resize_adaptive as ref_crop
    from oracle import classifier_ref as C
    img, masks = cases()
    oi_all, om_all = [], []
    for m in masks:
        mt = torch.from_numpy(m.astype(np.float32))
        ri, rm = ref_crop(img, mt)
        oi, om = C.crop_and_resize_adaptive(img, mt)
        assert torch.equal(ri, oi) and torch.equal(rm, om), "oracle crop differs from the reference"
        oi_all.append(ri[0, ::8, ::8].numpy())
        om_all.append(rm[0].numpy().astype(np.uint8))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "saber_classifier_crops.npz"), img_sub=np.stack(oi_all),
                        mask=np.packbits(np.stack(om_all), axis=-1))
    print("pinned", len(masks), "crops against the reference")


if __name__ == "__main__":
    main()
