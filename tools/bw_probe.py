"""Bandwidth-kernel probe: CUDA-event time (L2 flushed before every launch) and achieved GB/s of the streaming kernels
of the path at their production shapes. Run under gpurun: python tools/bw_probe.py [ln|vol|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saber_b200 import ops, synth

PEAK = 6547.8


def timeit(fn, flush, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def report(name, nbytes, ms):
    g = nbytes / ms / 1e6
    print(f"{name:58s} {ms * 1e3:9.1f} us {g:8.1f} GB/s  {g / PEAK:.3f} of {PEAK:.0f}")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda"
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    if what in ("ln", "all"):
        for M, C in ((524288, 144), (131072, 288), (32768, 576), (86016, 576), (8192, 1152), (21504, 1152)):
            x = torch.randn(M, C, device=dev)
            g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
            out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: ops.layernorm(x, g, b, 1e-6, torch.bfloat16, out=out), flush)
            report(f"layernorm fp32->bf16 M={M} C={C}", M * C * 6, ms)
    if what in ("vol", "all"):
        v = synth.make_tomogram((64, 928, 960), seed=2, device=dev).contiguous()
        nv = v.numel()
        mm = ops.minmax(v)
        report("minmax", 4 * nv, timeit(lambda: ops.minmax(v), flush))
        report("minmax_affine", 8 * nv, timeit(lambda: ops.minmax_affine(v, mm, 0.0, 2.0, -1.0), flush))
        from saber_b200.filters.gaussian import make_gaussian_kernel
        w15 = make_gaussian_kernel(5).to(dev, torch.float32).contiguous()
        report("gaussian_z 15 taps", 8 * nv, timeit(lambda: ops.gaussian_z(v, w15), flush))
        report("zoom_linear_mirror 928x960 -> 1024^2", 4 * nv + 4 * 64 * 1024 * 1024,
               timeit(lambda: ops.skimage_resize_stack(v, 1024, 2.0, -1.0), flush))
        report("mean_z 20 slices", 4 * 21 * 928 * 960, timeit(lambda: ops.mean_z(v, 22, 42), flush))
        img = v[0].contiguous()
        report("prepare_slice 928x960", 8 * img.numel(), timeit(lambda: ops.prepare_slice(img, 500, 3.0), flush))
    if what in ("post", "all"):
        # mask_post: 192 candidates of a full-frame 1024^2 crop and of a 683x683 second-layer crop, everything kept
        n, S, H, W = 192, 256, 1024, 1024
        planes = (torch.randn(n, 4, S, S, device=dev) * 4).contiguous()
        ious4 = torch.rand(n, 4, device=dev).contiguous()
        keep = torch.empty(n, dtype=torch.uint8, device=dev)
        stab, iou = torch.empty(n, device=dev), torch.empty(n, device=dev)
        bbox = torch.empty(n, 4, dtype=torch.int32, device=dev)
        area = torch.empty(n, dtype=torch.int32, device=dev)
        bits = torch.empty(n, H, W // 32, dtype=torch.int32, device=dev)
        for (hc, wc), (x0, y0) in (((1024, 1024), (0, 0)), ((683, 683), (341, 341))):
            ms = timeit(lambda: ops.amg_mask_post(planes, ious4, None, 1, n, (hc, wc), (x0, y0), (H, W), 0.0, 0.0, 1.0, 0.0,
                                                  keep, stab, iou, bbox, area, bits, 0), flush)
            report(f"mask_post 192 candidates crop {hc}x{wc}", n * (S * S * 4 + H * W // 8), ms)
    if what in ("ccl", "all"):
        from saber_b200.segmenters import utils as sutils
        for shape in ((200, 1024, 1024), (64, 512, 512)):
            vol = synth.make_label_volume(shape, seed=1, n_ellipsoids=300, device=dev, rmin=8.0, rmax=40.0, speckle=0.0005)
            ms = timeit(lambda: sutils.separate_masks_device(vol, min_mask_area=100), flush, reps=3)
            n = vol.numel()
            report(f"separate_masks {shape}: {n / ms / 1e3:.0f} Mvox/s", 6 * n, ms)
            del vol


if __name__ == "__main__":
    main()
