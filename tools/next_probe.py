"""Timing probe for the §8f rows on a B200: Fourier passes (GB/s per pass against the HBM roofline), the whole
FourierRescale3D / Filter3D calls at the BASELINE volume size, and the membrane-refinement workflow.
usage: python tools/next_probe.py [fft] [refine]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from saber_b200 import ops, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def fft_probe():
    from saber_b200.filters.downsample import FourierRescale3D
    from saber_b200.filters.tomograms import Filter3D
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6548.0)
    shape = (200, 928, 960)
    vol = synth.make_tomogram(shape, seed=7, n_ellipsoids=30, device="cuda").contiguous()
    n = vol.numel()
    spec = ops.fft_lines(vol, 2)
    rows = []
    for name, fn, nbytes in [
        ("x rows  real->complex n=960", lambda: ops.fft_lines(vol, 2), n * 12),
        ("y cols  complex       n=928", lambda: ops.fft_lines(spec, 1), n * 16),
        ("z cols  complex       n=200", lambda: ops.fft_lines(spec, 0), n * 16),
        ("x rows  complex->real n=960", lambda: ops.fft_lines(spec, 2, inverse=True, out_mode="real"), n * 12),
    ]:
        ms = timed(fn)
        rows.append((name, ms, nbytes / ms / 1e6))
        print(f"{name}: {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s  {nbytes / ms / 1e6 / peak:.3f} of HBM peak", flush=True)
    del spec
    r = FourierRescale3D(10.0, 20.0)
    ms = timed(lambda: r.rescale_device(vol), 3)
    print(f"FourierRescale3D 200x928x960 -> 100x464x480: {ms:.2f} ms")
    f = Filter3D(10.0, shape, lp=60.0, lpd=6.0, hp=2000.0, hpd=2.0)
    ms = timed(lambda: f.apply(vol), 3)
    print(f"Filter3D.apply 200x928x960 (band-pass): {ms:.2f} ms ({6 * n * 16 / ms / 1e6:.0f} GB/s over 6 passes at 16 B/voxel)")
    t0 = time.perf_counter()
    ref = torch.fft.ifftn(torch.fft.fftn(vol)).real
    torch.cuda.synchronize()
    ms = timed(lambda: torch.fft.ifftn(torch.fft.fftn(vol)).real, 3)
    print(f"(library comparison, not on the path) torch.fft fftn+ifftn (cuFFT): {ms:.2f} ms")


def refine_probe():
    from saber_b200.analysis.refine_membranes import FilteringConfig, OrganelleMembraneFilter
    shape = (200, 464, 480)
    org, mem = synth.make_organelle_membrane(shape, 71, 12, blob=3.0)
    o, m = torch.from_numpy(org).cuda(), torch.from_numpy(mem).cuda()
    f = OrganelleMembraneFilter(FilteringConfig(ball_size=3, min_membrane_area=2000))
    res = f.run_device(o, m)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = f.run_device(o, m)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"refine_membranes run_device {shape}, 12 organelles: {dt * 1e3:.1f} ms, organelles kept "
          f"{int(torch.unique(res['organelles']).numel()) - 1}")


def fft_once():
    """Four launches for an ncu capture: x rows (real in), y columns, z columns, x rows inverse (real out)."""
    vol = synth.make_tomogram((200, 928, 960), seed=7, n_ellipsoids=30, device="cuda").contiguous()
    spec = ops.fft_lines(vol, 2)
    a = ops.fft_lines(spec, 1)
    b = ops.fft_lines(spec, 0)
    c = ops.fft_lines(spec, 2, inverse=True, out_mode="real")
    torch.cuda.synchronize()


if __name__ == "__main__":
    what = sys.argv[1:] or ["fft", "refine"]
    if "fft" in what:
        fft_probe()
    if "fftonce" in what:
        fft_once()
    if "refine" in what:
        refine_probe()
