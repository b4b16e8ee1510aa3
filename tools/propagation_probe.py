"""z-axis propagation throughput probe (BASELINE configs[2] semantics on one GPU): hiera-large, synthetic 928x960 tomogram
of Z frames, N seed masks at the middle slice -> SAM2Adapter.set_volume + segment_volume. Run under gpurun:
  python tools/propagation_probe.py [Z] [N_obj]"""
import os, sys, time
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from saber_b200 import ops, synth
from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG
from saber_b200.adapters.sam2 import SAM2Adapter


def ellipse(hw, cy, cx, ry, rx):
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    return (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).astype(np.float32)


def main():
    Z = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    H, W = 928, 960
    ad = SAM2Adapter(SAM2AdapterConfig(cfg="large", amg_cfg=cfgAMG(sam2_cfg="large"), num_maskmem=2, seed=0), device="cuda:0")
    vol = synth.make_tomogram((Z, H, W), seed=3, n_ellipsoids=10, device="cuda:0")
    ad._video()  # build the predictor (weight init + upload) outside the timed region
    ad.set_volume(vol[:2].contiguous())  # warm the kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ad.set_volume(vol)
    torch.cuda.synchronize()
    t_set = time.perf_counter() - t0
    print(f"set_volume (normalize + resize + encode {Z} frames, hiera-L): {t_set * 1e3:.1f} ms = {t_set / Z * 1e3:.2f} ms/frame")
    rng = np.random.default_rng(0)
    for n_obj in [int(a) for a in sys.argv[2:]] or [1, 8]:
        seeds = [ellipse((H, W), rng.uniform(200, 700), rng.uniform(200, 700), rng.uniform(30, 90), rng.uniform(30, 90))
                 for _ in range(n_obj)]
        for rep in range(2):
            torch.cuda.synchronize()
            n0 = ops.launch_count
            t0 = time.perf_counter()
            out = ad.segment_volume(Z // 2, masks=seeds, vol_shape=(Z, H, W), min_presence_score=-1e9)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ad.reset_state()
        tracked = (Z - 1) * n_obj
        print(f"segment_volume N_obj={n_obj}: {dt * 1e3:.1f} ms for {Z} frames ({dt / Z * 1e3:.2f} ms/frame, "
              f"{dt / max(1, tracked) * 1e3:.2f} ms per object-frame, {ops.launch_count - n0} launches); labels present: "
              f"{sorted(set(np.unique(out).tolist()))[:10]}")


if __name__ == "__main__":
    main()
