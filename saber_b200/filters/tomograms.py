"""Cosine low-/high-pass of tomograms in Fourier space on the device — same surface and results as
REF saber/filters/tomograms.py (``Filter3D`` :12-184).

The reference builds a D x H x W filter volume at construction and evaluates
``ifftn(ifftshift(fftshift(fftn(data)) * filter)).real``. Here the radial filter is a function of the signed frequency
coordinates evaluated inside the store of the last forward FFT pass (``csrc/fft.cu``), so neither the filter volume nor
the shifted spectra are materialised; ``Filter3D.filter`` builds the volume on demand for inspection. fp32 throughout.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class Filter3D:
    def __init__(self, apix, sz, lp=0, lpd=0, hp=0, hpd=0, device=None):
        self.apix, self.sz = apix, tuple(int(s) for s in sz)
        self.lp, self.lpd, self.hp, self.hpd = lp, lpd, hp, hpd
        self.dtype = torch.float32
        if not torch.cuda.is_available():
            raise RuntimeError("saber_b200 Filter3D needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda") if device is None else torch.device(device)
        if self.lp > self.hp and self.lp > 0 and self.hp > 0:
            raise ValueError("Low-pass cutoff resolution must be less than high-pass cutoff resolution.")
        self.lp_pix = self.angst_to_pix(self.lp) if self.lp > 0 else 0
        self.hp_pix = self.angst_to_pix(self.hp) if self.hp > 0 else 0
        self.lpd_pix, self.hpd_pix = self.lpd, self.hpd
        self._filter = None

    def angst_to_pix(self, ang):
        return max(self.sz) / (ang / self.apix)

    def _bandpass(self):
        """{freq, freq - decay/2, freq + decay/2, decay} for the low- and the high-pass edge, rounded to fp32 the way the
        reference's python scalars are when they meet the fp32 radius tensor (REF tomograms.py:94-137)."""
        out = []
        for freq, decay in ((self.lp_pix, self.lpd_pix), (self.hp_pix, self.hpd_pix)):
            half = decay / 2.0
            out += [freq, freq - half, freq + half, decay]
        return [float(np.float32(v)) for v in out]

    @property
    def filter(self) -> torch.Tensor:
        """The fftshift-ed filter volume (REF cosine_filter :67-92), built on first use."""
        if self._filter is None:
            with torch.cuda.device(self.device):
                self._filter = ops.bandpass_volume(self.sz, self._bandpass())
        return self._filter

    def extract_1d_profile(self, axis="x"):
        """REF tomograms.py:139-170."""
        f = self.filter.cpu().numpy()
        D, H, W = f.shape
        if axis == "x":
            central, freqs = f[D // 2, H // 2, :], np.fft.fftfreq(W, d=self.apix)
        elif axis == "y":
            central, freqs = f[D // 2, :, W // 2], np.fft.fftfreq(H, d=self.apix)
        elif axis == "z":
            central, freqs = f[:, H // 2, W // 2], np.fft.fftfreq(D, d=self.apix)
        else:
            raise ValueError("Axis must be one of 'x', 'y', or 'z'.")
        keep = freqs >= 0
        return freqs[keep][::-1], central[keep]

    def apply(self, data) -> torch.Tensor:
        """REF tomograms.py:172-194: float32 [D,H,W] on the device (the reference also leaves the result there)."""
        v = torch.as_tensor(data).to(device=self.device, dtype=self.dtype).contiguous()
        if tuple(v.shape) != self.sz:
            raise ValueError(f"data of shape {tuple(v.shape)} does not match the filter size {self.sz}")
        D, H, W = self.sz
        with torch.cuda.device(self.device):
            spec = ops.fft_lines(v, 2)
            spec = ops.fft_lines(spec, 1)
            spec = ops.fft_lines(spec, 0, bandpass=self._bandpass())
            spec = ops.fft_lines(spec, 0, inverse=True)
            spec = ops.fft_lines(spec, 1, inverse=True)
            return ops.fft_lines(spec, 2, inverse=True, out_mode="real", scale=1.0 / (float(D) * H * W))
