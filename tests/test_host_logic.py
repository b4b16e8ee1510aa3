"""CPU tests of the host-side logic of the drop-in layer: AMG geometry vs the oracle's restatement of upstream,
config validation mirrored from the reference, greedy duplicate grouping."""
import numpy as np
import pytest
import torch

from oracle.sam2_ref import amg as up
from saber_b200.sam2 import automatic_mask_generator as ours


@pytest.mark.parametrize("hw", [(1024, 1024), (928, 960), (512, 512), (300, 517), (2048, 1536)])
@pytest.mark.parametrize("layers", [0, 1, 2])
def test_crop_boxes_and_point_grids_match_upstream_restatement(hw, layers):
    b1, l1 = ours.generate_crop_boxes(hw, layers, 512 / 1500)
    b2, l2 = up.generate_crop_boxes(hw, layers, 512 / 1500)
    assert b1 == [list(b) for b in b2] and l1 == list(l2)
    g1 = ours.build_all_layer_point_grids(32, layers, 2)
    g2 = up.build_all_layer_point_grids(32, layers, 2)
    assert len(g1) == len(g2)
    for a, b in zip(g1, g2):
        np.testing.assert_array_equal(a, b)
    assert len(b1) == sum(4 ** i for i in range(layers + 1))


def test_default_amg_has_21_crops_and_9216_candidates():
    boxes, layers = ours.generate_crop_boxes((1024, 1024), 2, 512 / 1500)
    grids = ours.build_all_layer_point_grids(32, 2, 2)
    assert len(boxes) == 21
    assert sum(len(grids[l]) for l in layers) == 3072


def test_configs_mirror_reference_validation():
    from pydantic import ValidationError
    from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG
    c = cfgAMG()
    assert (c.npoints, c.points_per_batch, c.pred_iou_thresh, c.stability_score_thresh, c.stability_score_offset,
            c.crop_n_layers, c.box_nms_thresh, c.crop_n_points_downscale_factor, c.use_m2m, c.multimask_output,
            c.sam2_cfg) == (32, 64, 0.7, 0.92, 0.7, 2, 0.7, 2, True, True, "small")
    with pytest.raises(ValidationError):
        cfgAMG(sam2_cfg="huge")
    with pytest.raises(ValidationError):
        cfgAMG(npoints=0)
    a = SAM2AdapterConfig()
    assert (a.cfg, a.num_maskmem, a.min_mask_area, a.model_type) == ("small", 2, 50, "sam2")
    with pytest.raises(ValidationError):
        SAM2AdapterConfig(cfg="giant")


def test_greedy_groups_equal_reference_semantics():
    """_greedy_groups (host part of the GPU duplicate removal) against the oracle on explicit intersections."""
    from oracle import saber_ref
    from oracle.make_golden import synth_mask_list
    from saber_b200.segmenters.utils import _greedy_groups
    for seed in (1, 2, 3, 4):
        masks = synth_mask_list((80, 90), 20, seed=seed)
        m = len(masks)
        seg = np.stack([x["segmentation"] for x in masks]).reshape(m, -1).astype(np.int64)
        area = seg.sum(1)
        inter = seg @ seg.T
        ratio = np.minimum(area[:, None], area[None]) / np.maximum(np.maximum(area[:, None], area[None]), 1)
        inter = np.where(ratio < 0.9, -1, inter)
        keep = _greedy_groups(area, np.array([x["stability_score"] for x in masks]), inter, 0.9)
        want = saber_ref.remove_duplicate_masks(masks)
        assert [id(masks[i]) for i in keep] == [id(x) for x in want]


def test_sliding_windows_match_reference_formula():
    from saber_b200.segmenters.base import saber2D
    s = object.__new__(saber2D)
    s.window_size, s.overlap_ratio = 256, 0.25
    w = s.get_sliding_windows((600, 700))
    assert w[0] == (0, 0, 256, 256) and all((y2 - y1) >= 128 and (x2 - x1) >= 128 for y1, x1, y2, x2 in w)
    assert (384, 576, 600, 700) not in w  # 124-px wide remainder is skipped
    assert s._to_global_bbox([1, 2, 3, 4], 10, 20) == [21, 12, 3, 4]


def test_estimate_thickness_twin_matches_reference_golden(golden_dir):
    """saber_b200.filters.estimate_thickness (host scipy, R8) reproduces the reference-run golden vector."""
    import os
    import numpy as np
    from saber_b200.filters import estimate_thickness
    g = np.load(os.path.join(golden_dir, "saber3d_fit_boundaries.npz"))
    out = estimate_thickness.fit_organelle_boundaries(g["frame_scores"].copy())
    np.testing.assert_allclose(out, g["out"], rtol=1e-9, atol=1e-12)


def test_gaussian_kernel_twin():
    import numpy as np
    from oracle import saber_ref
    from saber_b200.filters.gaussian import make_gaussian_kernel
    np.testing.assert_array_equal(make_gaussian_kernel(5).numpy(), saber_ref.make_gaussian_kernel(5))
