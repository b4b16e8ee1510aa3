"""Twin of REF saber/filters/gaussian.py:7-74 (R3): 1-D Gaussian smoothing along z on the device."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def make_gaussian_kernel(sigma):
    ks = round(sigma * 3)
    ks = max(ks, 3)
    ks += 1 - ks % 2
    ts = torch.linspace(-ks / 2, ks / 2, ks)
    gauss = torch.exp(-(ts / sigma) ** 2 / 2)
    return gauss / gauss.sum()


def gaussian_smoothing(input_tensor, sigma, dim=0, device="cuda"):
    """numpy (Z,Y,X) in -> numpy out (the reference's contract); CUDA tensor in -> CUDA tensor out. dim must be 0:
    the only call on the path is `gaussian_smoothing(vol, 5, dim=0)` (REF saber/segmenters/tomo.py:45)."""
    if dim not in (0, -3):
        raise NotImplementedError("saber_b200 gaussian_smoothing: only dim=0 (z) is on the path")
    is_numpy = isinstance(input_tensor, np.ndarray)
    x = (torch.from_numpy(np.ascontiguousarray(input_tensor)).float().to(device) if is_numpy
         else input_tensor.to(dtype=torch.float32).contiguous())
    y = ops.gaussian_z(x, make_gaussian_kernel(sigma).to(x.device, torch.float32).contiguous())
    return y.cpu().numpy() if is_numpy else y
