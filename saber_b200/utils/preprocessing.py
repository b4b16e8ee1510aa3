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


def _channel_axis_box(size: int = 500) -> np.ndarray:
    """scipy.ndimage.uniform_filter(image, size=500) on an (H,W,3) array also filters the CHANNEL axis: a 500-wide box
    over 3 samples with 'reflect' boundary handling is a fixed 3 x 3 mixing matrix (rows sum to 1, every entry is
    166/500 or 167/500 or 168/500)."""
    from scipy.ndimage import uniform_filter1d
    # column j of the filtered identity = response to a unit input in channel j: out[i] = sum_j F[i][j] x[j]
    return uniform_filter1d(np.eye(3, dtype=np.float64), size=size, axis=0, mode="reflect").astype(np.float32)


def prepare_rgb_device(image: torch.Tensor) -> torch.Tensor:
    """(H,W,3) fp32 CUDA image -> (H,W,3) fp32 in [0,1] exactly as the reference treats a 3-D array (REF :4-37,67-81):
    contrast() box-filters all three axes (size 500, reflect), z-scores and clips to +-3; normalize() is a global
    min-max. The channel-axis filter is the constant 3 x 3 mix of _channel_axis_box()."""
    assert image.is_cuda and image.dim() == 3 and image.shape[2] == 3
    return ops.prepare_rgb(image.to(_F32).contiguous(), _channel_axis_box(), box=500, cutoff=3.0)


def prepare(image, to_rgb: bool = False, device=None):
    """REF saber/utils/preprocessing.py:67-81."""
    x, was_tensor = _dev(image, device)
    if x.dim() == 3 and x.shape[2] == 3:
        y = prepare_rgb_device(x)  # `to_rgb` only repeats 2-D inputs (REF :78-80)
        return y if was_tensor else y.cpu().numpy()
    if x.dim() != 2:
        raise ValueError("prepare expects an (H,W) slice or an (H,W,3) image")
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
