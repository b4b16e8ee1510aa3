"""separate_masks (26-connected CCL + size filter + compact relabel) on the bench's 200x1024x1024 label volume:
event-timed total, or one pass under `ncu --metrics gpu__time_duration.sum` (argument: once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from saber_b200 import synth  # noqa: E402
from saber_b200.segmenters import utils as sutils  # noqa: E402

SHAPE = (200, 1024, 1024)
vol = synth.make_label_volume(SHAPE, seed=1, n_ellipsoids=300, device="cuda", rmin=8.0, rmax=40.0, speckle=0.0005)
print("foreground fraction", float((vol != 0).float().mean()))
if "once" in sys.argv[1:]:
    out = sutils.separate_masks_device(vol, min_mask_area=100)
    torch.cuda.synchronize()
    print("components", int(out.max()))
else:
    for _ in range(2):
        sutils.separate_masks_device(vol, min_mask_area=100)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sutils.separate_masks_device(vol, min_mask_area=100)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("separate_masks ms:", sorted(ts), "components", int(out.max()))
