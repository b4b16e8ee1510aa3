"""Fourier-crop rescaling on the device — same surface and results as REF saber/filters/downsample.py
(``FourierRescale3D`` :4-129, ``FourierRescale2D`` :131-204).

The reference runs ``fftn -> fftshift -> centre crop -> ifftshift -> ifftn`` through torch.fft, materialising every
intermediate. Here each axis is one line pass of ``csrc/fft.cu``; the shift + crop live in the store index of the forward
passes (so each later pass already works on the cropped extent) and the normalisation and ``.real`` / ``abs`` are fused
into the last inverse pass. Arithmetic is fp32 (complex64), as torch.fft uses for float32 input; float64 input is computed
in fp32 and returned as float64. No CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import ops


def _crop_window(n_in: int, n_new: int):
    """(start, length) of the kept band on the fftshift-ed axis — REF downsample.py:117-127 / :187-194."""
    n_new = n_new - (n_new % 2)
    return (n_in - n_new) // 2 + (n_in % 2), n_new


def _as_f32_cuda(x, device):
    if not torch.cuda.is_available():
        raise RuntimeError("saber_b200 Fourier rescaling needs a CUDA device (no CPU fallback)")
    t = torch.as_tensor(x)
    src_dtype = t.dtype
    return t.to(device=device, dtype=torch.float32).contiguous(), src_dtype


class FourierRescale3D:
    def __init__(self, input_voxel_size, output_voxel_size):
        if isinstance(input_voxel_size, (int, float)):
            input_voxel_size = (input_voxel_size,) * 3
        if isinstance(output_voxel_size, (int, float)):
            output_voxel_size = (output_voxel_size,) * 3
        self.input_voxel_size = input_voxel_size
        self.output_voxel_size = output_voxel_size
        if any(o < i for i, o in zip(input_voxel_size, output_voxel_size)):
            raise ValueError("Output voxel size must be greater than or equal to the input voxel size.")
        if not torch.cuda.is_available():
            raise RuntimeError("saber_b200 FourierRescale3D needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda")

    def calculate_cropping(self, volume):
        """(start_d, start_h, start_w, new_d, new_h, new_w) — REF downsample.py:97-129."""
        dims = volume.shape[-3:]
        news = [int(round(n * i / o)) for n, i, o in zip(dims, self.input_voxel_size, self.output_voxel_size)]
        wins = [_crop_window(n, m) for n, m in zip(dims, news)]
        return (*[w[0] for w in wins], *[w[1] for w in wins])

    def rescale_device(self, volume: torch.Tensor) -> torch.Tensor:
        """float32 [D,H,W] on the device -> float32 [d,h,w] on the device (the resident form of `run`)."""
        sd, sh, sw, nd, nh, nw = self.calculate_cropping(volume)
        D, H, W = volume.shape
        if min(nd, nh, nw) <= 0:
            raise ValueError("the requested voxel size leaves an empty volume")
        spec = ops.fft_lines(volume, 2, crop=(sw, nw))
        spec = ops.fft_lines(spec, 1, crop=(sh, nh))
        spec = ops.fft_lines(spec, 0, crop=(sd, nd))
        spec = ops.fft_lines(spec, 0, inverse=True)
        spec = ops.fft_lines(spec, 1, inverse=True)
        # norm='ortho' on both transforms: 1 / sqrt(D H W) forward, 1 / sqrt(d h w) inverse
        scale = 1.0 / math.sqrt(float(D) * H * W) / math.sqrt(float(nd) * nh * nw)
        return ops.fft_lines(spec, 2, inverse=True, out_mode="real", scale=scale)

    def batched_rescale(self, volume: torch.Tensor) -> torch.Tensor:
        v, _ = _as_f32_cuda(volume, self.device)
        if v.dim() == 3:
            return self.rescale_device(v)
        return torch.stack([self.rescale_device(x.contiguous()) for x in v])

    def single_rescale(self, volume: torch.Tensor) -> torch.Tensor:
        return self.batched_rescale(volume)

    def run(self, volume):
        """REF downsample.py:35-65: numpy in -> numpy out, tensor in -> CPU tensor out; 3-D or batched 4-D."""
        return_numpy = isinstance(volume, np.ndarray)
        v, src_dtype = _as_f32_cuda(volume, self.device)
        if v.dim() not in (3, 4):
            raise ValueError("FourierRescale3D expects a (D,H,W) volume or a (B,D,H,W) batch")
        out = self.batched_rescale(v).cpu()
        if src_dtype == torch.float64:
            out = out.double()
        return out.numpy() if return_numpy else out


class FourierRescale2D:
    @staticmethod
    def run_resolution(image, input_pixsize: float, target_pixsize: float, device=None):
        scale_factor = target_pixsize / input_pixsize
        if target_pixsize <= input_pixsize:
            raise ValueError(f"Target pixel size ({target_pixsize}Å) must be larger than current pixel size ({input_pixsize}Å)")
        return FourierRescale2D._rescale(image, scale_factor, device)

    @staticmethod
    def run(image, scale_factor: float, device=None):
        if scale_factor < 1:
            raise ValueError("Scale factor must be greater than 1")
        return FourierRescale2D._rescale(image, scale_factor, device)

    @staticmethod
    def rescale_device(image: torch.Tensor, scale_factor: float) -> torch.Tensor:
        """float32 [h,w] on the device -> float32 [h',w'] on the device: |ifft2(crop(fft2(image)))| (default fft norms)."""
        h, w = image.shape
        sh, nh = _crop_window(h, int(h / scale_factor))
        sw, nw = _crop_window(w, int(w / scale_factor))
        if min(nh, nw) <= 0:
            raise ValueError("the scale factor leaves an empty image")
        spec = ops.fft_lines(image, 1, crop=(sw, nw))
        spec = ops.fft_lines(spec, 0, crop=(sh, nh))
        spec = ops.fft_lines(spec, 0, inverse=True)
        return ops.fft_lines(spec, 1, inverse=True, out_mode="abs", scale=1.0 / (float(nh) * nw))

    @staticmethod
    def rescale_stack_device(stack: torch.Tensor, scale_factor: float) -> torch.Tensor:
        """float32 [Z,h,w] on the device -> [Z,h',w']: every slice rescaled as `rescale_device` does, in four launches for
        the whole stack (the reference loops `FourierRescale2D.run` over the slices of a movie / FIB stack, REF
        saber/utils/io.py:37-39). The line passes treat the leading axis as a batch."""
        Z, h, w = stack.shape
        sh, nh = _crop_window(h, int(h / scale_factor))
        sw, nw = _crop_window(w, int(w / scale_factor))
        if min(nh, nw) <= 0:
            raise ValueError("the scale factor leaves an empty image")
        spec = ops.fft_lines(stack, 2, crop=(sw, nw))
        spec = ops.fft_lines(spec, 1, crop=(sh, nh))
        spec = ops.fft_lines(spec, 1, inverse=True)
        return ops.fft_lines(spec, 2, inverse=True, out_mode="abs", scale=1.0 / (float(nh) * nw))

    @staticmethod
    def _rescale(image, scale_factor: float, device=None):
        """REF downsample.py:153-204."""
        is_numpy = isinstance(image, np.ndarray)
        img, src_dtype = _as_f32_cuda(image, torch.device("cuda") if device is None else device)
        if img.dim() != 2:
            raise ValueError("FourierRescale2D expects a 2-D image")
        with torch.cuda.device(img.device):
            out = FourierRescale2D.rescale_device(img, scale_factor).cpu()
        if src_dtype == torch.float64:
            out = out.double()
        return out.numpy() if is_numpy else out
