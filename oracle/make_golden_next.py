"""Generates the fixtures of the SURVEY §8f "next" rows by running the REFERENCE's own code in the build container
(``python -m oracle.make_golden_next``):

* tests/golden/saber_refine_membranes.npz — REF saber/analysis/refine_membranes.py OrganelleMembraneFilter.run on CPU
  (scipy connected components + scipy binary_opening) over ``saber_b200.synth.make_organelle_membrane`` volumes;
* tests/golden/saber_fourier.npz — REF saber/filters/downsample.py FourierRescale3D / FourierRescale2D and
  REF saber/filters/tomograms.py Filter3D.apply on CPU (torch.fft).

Inputs are regenerated from their seeds by the tests; only outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

REFINE_CASES = [
    # name, shape, seed, n_organelles, config overrides
    ("a", (40, 72, 80), 31, 4, dict(ball_size=3, min_membrane_area=200, edge_trim_z=2, edge_trim_xy=2)),
    ("b", (80, 96, 96), 34, 2, dict(ball_size=3, min_membrane_area=100, edge_trim_z=2, edge_trim_xy=2,
                                    keep_surface_membranes=True)),  # drops the internal blob
    ("c", (36, 64, 64), 32, 3, dict(ball_size=5, min_membrane_area=100, edge_trim_z=3, edge_trim_xy=3,
                                    min_roi_relative_size=0.1)),
    ("d", (24, 40, 40), 33, 2, dict(ball_size=3, min_membrane_area=100000, edge_trim_z=2, edge_trim_xy=2)),  # no membrane left
    ("e", (24, 40, 40), 33, 2, dict(ball_size=3, min_membrane_area=50, edge_trim_z=0, edge_trim_xy=2)),  # the [0:-0] quirk
]


def refine_golden():
    from oracle.make_golden_3d import install_stubs
    from saber_b200 import synth
    install_stubs()
    sys.path.insert(0, "/root/reference")
    import saber.analysis.refine_membranes as ref
    out = {}
    for name, shape, seed, n_org, over in REFINE_CASES:
        org, mem = synth.make_organelle_membrane(shape, seed, n_org, blob=3.2 if name == "b" else 2.0)
        f = ref.OrganelleMembraneFilter(ref.FilteringConfig(**over))
        f.device = torch.device("cpu")
        res = f.run(org.copy(), mem.copy(), batch_processing=True)
        o, m = np.asarray(res["organelles"]), np.asarray(res["membranes"])
        out[f"{name}_organelles"] = o
        out[f"{name}_membranes"] = m
        if o.ndim == 4:
            out[f"{name}_organelles_3d"] = f.convert_to_3d_labels(o)
            out[f"{name}_membranes_3d"] = f.convert_to_3d_labels(m)
        print(name, o.shape, o.dtype, int((o > 0).sum()), int((m > 0).sum()), np.unique(o), np.unique(m))
    np.savez_compressed(os.path.join(GOLD, "saber_refine_membranes.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "refine" in sys.argv[1:] or len(sys.argv) == 1:
        refine_golden()
    if "fourier" in sys.argv[1:] or len(sys.argv) == 1:
        fourier_golden()


RESCALE3D_CASES = [  # name, shape, seed, input voxel size, output voxel size
    ("r3a", (24, 58, 40), 51, 5.0, 10.0),
    ("r3b", (25, 45, 64), 52, (4.0, 5.0, 5.0), (7.0, 9.0, 12.0)),   # odd sizes, anisotropic
    ("r3c", (20, 29, 48), 53, 3.0, 4.3),
]
RESCALE2D_CASES = [("r2a", (96, 116), 54, 2.0), ("r2b", (75, 128), 55, 3.3), ("r2c", (58, 50), 56, 1.0)]
FILTER_CASES = [  # name, shape, seed, apix, lp, lpd, hp, hpd
    ("fa", (20, 48, 58), 57, 10.0, 60.0, 6.0, 0.0, 0.0),      # low-pass with cosine decay
    ("fb", (24, 40, 40), 58, 10.0, 50.0, 4.0, 400.0, 2.0),    # band-pass
    ("fc", (25, 45, 32), 59, 8.0, 40.0, 0.0, 0.0, 0.0),       # box low-pass, odd sizes
    ("fd", (16, 32, 32), 60, 8.0, 0.0, 0.0, 200.0, 0.0),      # box high-pass
]


def fourier_golden():
    from oracle.make_golden_3d import install_stubs
    from saber_b200 import synth
    install_stubs()
    sys.path.insert(0, "/root/reference")
    import saber.filters.downsample as ref_ds
    import saber.filters.tomograms as ref_tf
    out = {}
    for name, shape, seed, vin, vout in RESCALE3D_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        r = ref_ds.FourierRescale3D(vin, vout)
        r.device = torch.device("cpu")
        out[name] = r.run(vol)
        print(name, shape, "->", out[name].shape, out[name].dtype)
    for name, shape, seed, sf in RESCALE2D_CASES:
        img = synth.make_tomogram((1, *shape), seed=seed, n_ellipsoids=3).numpy()[0]
        out[name] = ref_ds.FourierRescale2D.run(img, sf, device=torch.device("cpu"))
        print(name, shape, "->", out[name].shape, out[name].dtype)
    for name, shape, seed, apix, lp, lpd, hp, hpd in FILTER_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        f = ref_tf.Filter3D(apix, shape, lp=lp, lpd=lpd, hp=hp, hpd=hpd, device=torch.device("cpu"))
        out[name] = f.apply(vol).numpy()
        out[name + "_filter"] = f.filter.numpy()
        print(name, shape, float(f.filter.min()), float(f.filter.max()), float(f.filter.mean()))
    np.savez_compressed(os.path.join(GOLD, "saber_fourier.npz"), **out)


if __name__ == "__main__":
    main()
