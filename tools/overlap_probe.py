"""Does slice s+1's encoder (tensor-bound) overlap with slice s's decoder (HBM-bound) when two segmenters run from two
threads of one process on one GPU, each on its own streams? Prints slices/s for 1 worker and for 2 concurrent workers."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")
from saber_b200 import synth  # noqa: E402
from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG  # noqa: E402
from saber_b200.segmenters.propagation import propagationSegmenter  # noqa: E402

SHAPE = (200, 1024, 1024)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def make():
    return propagationSegmenter(deviceID=0, cfg=SAM2AdapterConfig(cfg="large", amg_cfg=cfgAMG(sam2_cfg="large"), min_mask_area=100,
                                                                 allow_random_init=True), min_mask_area=100)


def main():
    torch.cuda.set_device(0)
    segs = [make() for _ in range(K)]
    slabs = [synth.make_tomogram(SHAPE, seed=0, device="cuda", z_range=(z, z + 1)).contiguous() for z in range(K * N)]
    labels = [torch.empty((1,) + SHAPE[1:], dtype=torch.int16, device="cuda") for _ in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    for w in range(K):  # warm (graph capture) one after the other
        with torch.cuda.stream(streams[w]):
            for _ in range(2):
                segs[w].label_slices_device(slabs[w], labels[w])
        torch.cuda.synchronize()

    def run(w, idx):
        torch.cuda.set_device(0)
        with torch.cuda.stream(streams[w]):
            for i in idx:
                segs[w].label_slices_device(slabs[i], labels[w])
        streams[w].synchronize()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(0, range(N))
    torch.cuda.synchronize()
    t1 = time.perf_counter() - t0
    print(f"1 worker : {N} slices in {t1 * 1e3:.1f} ms = {N / t1:.2f} slices/s", flush=True)
    ths = [threading.Thread(target=run, args=(w, range(w * N, (w + 1) * N))) for w in range(K)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    t2 = time.perf_counter() - t0
    print(f"{K} workers: {K * N} slices in {t2 * 1e3:.1f} ms = {K * N / t2:.2f} slices/s ({K * N / t2 / (N / t1):.3f}x)", flush=True)


if __name__ == "__main__":
    main()
