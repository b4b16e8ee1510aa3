"""Deterministic synthetic inputs (there is no network / dataset here): cryo-ET-like tomograms made of unit
Gaussian noise plus dark soft-edged ellipsoids, and blob label volumes for the integer stages.
Used by tests/, bench.py and __graft_entry__.smoke(); generated with torch so the same code runs on CPU
(tests) and on the GPU (bench, where a 200x1024x1024 volume is built in HBM)."""
from __future__ import annotations

import numpy as np
import torch


def ellipsoid_params(shape, n: int, seed: int, rmin: float = 10.0, rmax: float = 80.0):
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    c = rng.uniform(0, 1, (n, 3)) * np.array([Z, Y, X])
    r = rng.uniform(rmin, rmax, (n, 3))
    r[:, 0] = np.minimum(r[:, 0], max(2.0, Z / 3))
    return c.astype(np.float32), r.astype(np.float32)


def ellipsoid_field(shape, n: int, seed: int, device="cpu", rmin: float = 10.0, rmax: float = 80.0,
                    z_range=None) -> torch.Tensor:
    """min over ellipsoids of the normalised radial distance (1.0 = surface); slices z_range only."""
    Z, Y, X = shape
    c, r = ellipsoid_params(shape, n, seed, rmin, rmax)
    z0, z1 = (0, Z) if z_range is None else z_range
    zz = torch.arange(z0, z1, device=device, dtype=torch.float32)[:, None, None]
    yy = torch.arange(Y, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(X, device=device, dtype=torch.float32)[None, None, :]
    d = torch.full((z1 - z0, Y, X), 1e9, device=device)
    for k in range(n):
        dk = ((zz - c[k, 0]) / r[k, 0]) ** 2 + ((yy - c[k, 1]) / r[k, 1]) ** 2 + ((xx - c[k, 2]) / r[k, 2]) ** 2
        d = torch.minimum(d, dk)
    return d.sqrt()


def make_tomogram(shape, seed: int = 0, n_ellipsoids: int = 40, device="cpu", z_range=None) -> torch.Tensor:
    """fp32 (Z,Y,X) [or the z_range slab]: N(0,1) noise + ellipsoids of intensity -1.5 with ~2-voxel soft edges."""
    Z, Y, X = shape
    z0, z1 = (0, Z) if z_range is None else z_range
    g = torch.Generator(device="cpu").manual_seed(seed)
    parts = []
    for z in range(z0, z1):  # per-slice streams so any slab is reproducible independently
        g.manual_seed(seed * 1_000_003 + z)
        parts.append(torch.randn((Y, X), generator=g))
    noise = torch.stack(parts).to(device)
    d = ellipsoid_field(shape, n_ellipsoids, seed, device, z_range=(z0, z1))
    rbar = 40.0
    return noise - 1.5 * torch.sigmoid((1.0 - d) * (rbar / 2.0))


def make_label_volume(shape, seed: int = 0, n_ellipsoids: int = 40, device="cpu", rmin=4.0, rmax=20.0,
                      speckle: float = 0.0) -> torch.Tensor:
    """uint16-valued (stored int16) instance label volume of overlapping ellipsoids (+ optional speckle noise
    that creates tiny components for the small-component filter)."""
    Z, Y, X = shape
    c, r = ellipsoid_params(shape, n_ellipsoids, seed, rmin, rmax)
    zz = torch.arange(Z, device=device, dtype=torch.float32)[:, None, None]
    yy = torch.arange(Y, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(X, device=device, dtype=torch.float32)[None, None, :]
    lab = torch.zeros(shape, dtype=torch.int16, device=device)
    for k in range(n_ellipsoids):
        inside = (((zz - c[k, 0]) / r[k, 0]) ** 2 + ((yy - c[k, 1]) / r[k, 1]) ** 2 + ((xx - c[k, 2]) / r[k, 2]) ** 2) <= 1
        lab[inside] = k + 1
    if speckle > 0:
        g = torch.Generator(device="cpu").manual_seed(seed + 7)
        sp = (torch.rand(shape, generator=g) < speckle).to(device)
        lab[sp & (lab == 0)] = 1
    return lab
