"""B200 twin of REF saber/utils/preprocessing.py (same function names and argument meaning).

``prepare`` / ``contrast`` / ``normalize`` run the box-filter, contrast and min-max kernels of
``csrc/preprocess.cu`` on the device; numpy in -> numpy out keeps the reference's call shape, CUDA
tensor in -> CUDA tensor out is the resident fast path used by the segmenters.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

_F32 = torch.float32


def _dev(image, device=None):
    if isinstance(image, torch.Tensor):
        return image.to(dtype=_F32).contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device or "cuda"), False


def prepare_device(image: torch.Tensor) -> torch.Tensor:
    """[H,W] fp32 CUDA slice -> [H,W] fp32 in [0,1]: contrast(std_cutoff=3) then normalize (REF :67-81).
    SABER's RGB output is three identical copies of this plane."""
    assert image.is_cuda and image.dim() == 2
    return ops.prepare_slice(image.to(_F32).contiguous(), box=500, cutoff=3.0)


def prepare(image, to_rgb: bool = False, device=None):
    """REF saber/utils/preprocessing.py:67-81."""
    x, was_tensor = _dev(image, device)
    if x.dim() != 2:
        raise NotImplementedError("saber_b200 prepare: only 2-D grayscale slices are on the hot path")
    y = prepare_device(x)
    if to_rgb:
        y = y[..., None].expand(-1, -1, 3).contiguous()
    return y if was_tensor else y.cpu().numpy()


def project_tomogram(vol, zSlice=None, deltaZ=None):
    """REF saber/utils/preprocessing.py:39-65 (mean over a z-range); accepts numpy or CUDA tensors."""
    if zSlice is not None:
        if deltaZ is not None:
            z0 = int(max(zSlice - deltaZ, 0))
            z1 = int(min(zSlice + deltaZ, vol.shape[0]))
            sub = vol[z0:z1]
            return sub.mean(dim=0) if isinstance(sub, torch.Tensor) else np.mean(sub, axis=0)
        return vol[zSlice]
    return vol.mean(dim=0) if isinstance(vol, torch.Tensor) else np.mean(vol, axis=0)
