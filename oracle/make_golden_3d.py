"""Generates tests/golden/saber3d_*.npz by running the REFERENCE's own code for the 3-D path rows that import in the
build container once their *unused* third-party imports (mrcfile, matplotlib, skimage, zarr, copick ...) are replaced
by empty stub modules: saber.filters.gaussian (R3/R14), saber.filters.estimate_thickness (R8), saber.filters.masks
(R13/R14), saber.analysis.refine_membranes (R17). Also dumps HF transformers Sam2VideoModel memory-attention /
memory-encoder vectors (independent implementation) that pin oracle/sam2_ref/memory.py (U7/U8).
Run in the build container only: ``python -m oracle.make_golden_3d``.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        m.__spec__ = importlib.machinery.ModuleSpec(m.__name__, None)
        m.__path__ = []
        sys.modules[m.__name__] = m
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        raise RuntimeError(f"stubbed dependency {self.__name__} was called: this row is NOT pinned by the reference")


def install_stubs():
    for name in ("mrcfile", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches",
                 "matplotlib.widgets", "matplotlib.cm", "matplotlib.figure", "skimage", "skimage.transform",
                 "skimage.measure", "skimage.morphology", "zarr", "copick", "copick_utils", "rich_click", "monai",
                 "starfile", "ome_zarr", "napari", "pyqtgraph", "PyQt5", "cv2", "PIL", "tifffile", "nibabel", "h5py",
                 "hyperspy", "hyperspy.api", "rsciio", "imageio", "sam2", "sam2.build_sam"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        m = _Stub(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)


def main():
    os.makedirs(GOLD, exist_ok=True)
    install_stubs()
    sys.path.insert(0, "/root/reference")
    from saber_b200 import synth

    # ---- R3 gaussian_smoothing(vol, 5, dim=0)
    import saber.filters.gaussian as ref_gauss
    vol = synth.make_tomogram((40, 48, 56), seed=21, n_ellipsoids=5).numpy()
    sm = ref_gauss.gaussian_smoothing(vol, 5, dim=0)
    np.savez_compressed(os.path.join(GOLD, "saber3d_gaussian_z.npz"), seed=21, shape=np.array(vol.shape), out=sm)
    # ---- R14 gaussian_smoothing_3d / fast_3d_gaussian_smoothing
    import saber.filters.masks as ref_masks
    lab = synth.make_label_volume((24, 40, 48), seed=22, n_ellipsoids=4, rmin=5.0, rmax=10.0).numpy().astype(np.uint16)
    # make the labels distinct per component so several sigmas are exercised
    from scipy import ndimage as ndi
    lab2, n = ndi.label(lab > 0)
    lab2 = lab2.astype(np.uint16)
    ref_masks.io.get_available_devices = lambda deviceID=None: torch.device("cpu")
    sm3 = ref_masks.fast_3d_gaussian_smoothing(lab2, scale=0.075, deviceID=None)
    np.savez_compressed(os.path.join(GOLD, "saber3d_fast_gauss3d.npz"), labels_in=lab2, out=sm3)
    # ---- R8 fit_organelle_boundaries
    import saber.filters.estimate_thickness as ref_thick
    rng = np.random.default_rng(23)
    Z = 64
    x = np.arange(Z)
    fs = np.stack([8 * np.exp(-(x - 30) ** 2 / (2 * 5.0 ** 2)) + rng.normal(0, 0.2, Z),
                   np.maximum(-0.02 * (x - 20) ** 2 + 6, -3) + rng.normal(0, 0.2, Z),
                   rng.normal(-2, 0.3, Z), np.zeros(Z)], axis=1)
    mb = ref_thick.fit_organelle_boundaries(fs.copy(), plot=False)
    np.savez_compressed(os.path.join(GOLD, "saber3d_fit_boundaries.npz"), frame_scores=fs, out=mb)
    # ---- R13 consensus resolution / convert_predictions_to_masks / masks_to_array
    from oracle.make_golden import synth_mask_list
    masks = synth_mask_list((96, 128), 12, seed=24)
    preds = rng.dirichlet(np.ones(3), size=len(masks)).astype(np.float32)
    inst = ref_masks.convert_predictions_to_masks(preds, [dict(m) for m in masks], desired_class=1, min_mask_area=32)
    arr = ref_masks.masks_to_array(inst) if len(inst) else np.zeros((0, 96, 128), np.uint8)
    sem = ref_masks.convert_predictions_to_masks(preds, [dict(m) for m in masks], desired_class=0, min_mask_area=32)
    np.savez_compressed(os.path.join(GOLD, "saber3d_classifier_masks.npz"), preds=preds, inst_array=arr,
                        inst_conf=np.array([m["predicted_iou"] for m in inst], dtype=np.float64),
                        sem=np.stack([m["segmentation"].astype(np.uint8) for m in sem]),
                        sem_area=np.array([m["area"] for m in sem]))
    # ---- R17 morphology
    import saber.analysis.refine_membranes as ref_rm
    cls = [getattr(ref_rm, n) for n in dir(ref_rm) if isinstance(getattr(ref_rm, n), type) and
           hasattr(getattr(ref_rm, n), "_torch_erosion_3d")][0]
    obj = cls.__new__(cls)
    obj.device = torch.device("cpu")
    roi = (synth.make_label_volume((20, 28, 32), seed=25, n_ellipsoids=5, rmin=3.0, rmax=8.0, speckle=0.01).numpy() > 0)
    roi_t = torch.from_numpy(roi.astype(np.float32))
    out = {}
    for r in (1, 2, 3):
        k = obj._create_ball_kernel(r)
        out[f"erode{r}"] = obj._torch_erosion_3d(roi_t, k).numpy().astype(np.uint8)
        out[f"dilate{r}"] = obj._torch_dilation_3d(roi_t, k).numpy().astype(np.uint8)
        out[f"open{r}"] = obj._morphological_opening_gpu(roi_t, k).numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "saber3d_morphology.npz"), roi=roi.astype(np.uint8), **out)

    # ---- HF Sam2VideoModel memory attention / memory encoder (tiny backbone; memory modules are size-independent)
    from transformers import Sam2VideoConfig, Sam2VideoModel
    from oracle.hf_bridge import hf_to_upstream
    torch.manual_seed(0)
    hf = Sam2VideoModel(Sam2VideoConfig(num_maskmem=2)).eval()
    sd = hf_to_upstream(hf.state_dict())
    g = torch.Generator().manual_seed(31)
    curr = torch.randn(4096, 1, 256, generator=g) * 0.5
    curr_pos = torch.randn(4096, 1, 256, generator=g) * 0.5
    n_ptr = 8
    memory = torch.randn(2 * 4096 + n_ptr, 1, 64, generator=g) * 0.5
    memory_pos = torch.randn(2 * 4096 + n_ptr, 1, 64, generator=g) * 0.5
    with torch.no_grad():
        out_att = hf.memory_attention(current_vision_features=curr, current_vision_position_embeddings=curr_pos,
                                      memory=memory, memory_posision_embeddings=memory_pos, num_object_pointer_tokens=n_ptr)
        pix = torch.randn(1, 256, 64, 64, generator=g) * 0.5
        msk = torch.randn(1, 1, 1024, 1024, generator=g) * 4
        mm, mpos = hf.memory_encoder(pix, torch.sigmoid(msk) * 20 - 10)
    np.savez_compressed(os.path.join(GOLD, "hf_video_memory.npz"), weight_seed=0, input_seed=31, n_ptr=n_ptr,
                        att_sub=out_att.reshape(4096, 256)[::16, ::4].numpy().copy(),
                        att_mean=float(out_att.mean()), att_std=float(out_att.std()),
                        mm_sub=mm[0, ::2, ::4, ::4].numpy().copy(), mpos_sub=mpos[0, ::2, ::4, ::4].numpy().copy())
    torch.save({k: v for k, v in sd.items() if k.startswith(("memory_attention.", "memory_encoder."))},
               os.path.join(GOLD, "hf_video_memory_weights_keys.pt")) if False else None
    print("3-D golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
