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


def make_organelle_membrane(shape, seed: int = 0, n_organelles: int = 4, speckle: float = 0.002, gap: float = 0.15,
                            blob: float = 2.0):
    """(organelle labels int32 [Z,Y,X] with values 1..n, membrane uint8 {0,1}): ellipsoidal organelles well inside the
    volume, each wrapped in a ~2-3 voxel shell broken by random gaps, plus speckle in both volumes — the input pair of
    the membrane-refinement workflow (REF saber/analysis/refine_membranes.py:445)."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    zz, yy, xx = np.meshgrid(np.arange(Z, dtype=np.float32), np.arange(Y, dtype=np.float32), np.arange(X, dtype=np.float32),
                             indexing="ij")
    org = np.zeros(shape, np.int32)
    mem = np.zeros(shape, np.uint8)
    blocks = rng.random((Z // 4 + 1, Y // 4 + 1, X // 4 + 1)) > gap  # 4^3 blocks knocked out of the shells
    keep = np.repeat(np.repeat(np.repeat(blocks, 4, 0), 4, 1), 4, 2)[:Z, :Y, :X]
    for k in range(n_organelles):
        c = np.array([rng.uniform(0.3, 0.7) * Z, rng.uniform(0.2, 0.8) * Y, rng.uniform(0.2, 0.8) * X], np.float32)
        r = np.array([rng.uniform(0.12, 0.25) * Z, rng.uniform(0.1, 0.2) * Y, rng.uniform(0.1, 0.2) * X], np.float32)
        if k == n_organelles - 1:
            r[2] *= 0.45  # one thin organelle
        d = np.sqrt(((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2)
        org[d < 1.0] = k + 1
        t = 2.5 / float(r.min())
        mem[(d > 1.0 - t) & (d < 1.0 + 0.5 * t) & keep] = 1
        mem[(zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2 <= blob * blob] = 1  # an internal blob (not on the surface)
    mem[rng.random(shape) < speckle] = 1
    org[(rng.random(shape) < speckle) & (org == 0)] = n_organelles + 1
    return org, mem
