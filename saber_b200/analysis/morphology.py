"""Twin of the morphology primitives of REF saber/analysis/refine_membranes.py:100-117,274-333 (R17): binary erosion /
dilation / opening with a radius-r ball on an organelle ROI, zero padded. The reference evaluates them as dense fp32
conv3d with (2r+1)^3 taps on 0/1 data and thresholds the sums; the results are integer set operations, computed here
directly on uint8 voxels. Float {0,1} tensors in -> float {0,1} tensors out (the reference's contract)."""
from __future__ import annotations

import torch

from .. import ops


def _u8(image: torch.Tensor) -> torch.Tensor:
    return (image > 0).to(torch.uint8).contiguous()


def torch_erosion_3d(image: torch.Tensor, radius: int) -> torch.Tensor:
    if float(image.sum()) == 0:
        return image
    return ops.morph_ball(_u8(image), radius, 0).to(image.dtype)


def torch_dilation_3d(image: torch.Tensor, radius: int) -> torch.Tensor:
    if float(image.sum()) == 0:
        return image
    return ops.morph_ball(_u8(image), radius, 1).to(image.dtype)


def morphological_opening(image: torch.Tensor, radius: int) -> torch.Tensor:
    if float(image.sum()) == 0:
        return image
    return ops.morph_ball(ops.morph_ball(_u8(image), radius, 0), radius, 1).to(image.dtype)
