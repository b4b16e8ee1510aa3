"""-m gpu: expert classifier — B200 Predictor / SAM2Classifier vs the oracle restatement of the
reference's predictor, with identical seeded backbone + head weights (hiera-tiny)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from oracle import classifier_ref as C
    from oracle.make_golden_classifier import cases
    from oracle.sam2_ref.sam2_base import build_sam2 as oracle_build
    from saber_b200 import ops
    from saber_b200.classifier import Predictor, SAM2Classifier
    from saber_b200.sam2 import arch
    from saber_b200.sam2.build_sam import build_sam2
    ops.require_b200()
    sd = arch.random_state_dict("tiny", seed=0)
    head = C.random_head_state_dict(3, seed=1)
    orc = C.Predictor(C.SAM2Classifier(oracle_build("tiny", None, device="cpu", state_dict=sd), 3, head), 3)
    sam = build_sam2("tiny", None, device="cuda:0", state_dict=sd)
    ours = Predictor(model=SAM2Classifier(3, "tiny", head_sd=head, sam_model=sam), num_classes=3)
    img, masks = cases()
    return dict(orc=orc, ours=ours, img=img[0].numpy(), masks=np.stack(masks).astype(np.uint8), C=C)


def test_crops_match_oracle(setup):
    """NormalizeIntensity + crop_and_resize_adaptive: image crops to fp32 rounding, masks and areas bit-exact."""
    from saber_b200 import ops
    C, ours = setup["C"], setup["ours"]
    img = torch.from_numpy(setup["img"])
    nimg = C.normalize_intensity(img)
    got = ops.standardize(img.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), nimg.numpy(), atol=2e-6, rtol=0)
    m = torch.from_numpy(setup["masks"])
    ci, cm, areas = ours.apply_crops(nimg.cuda().contiguous(), m.cuda().contiguous())
    wi, wm = setup["orc"].apply_crops(nimg, m.float())
    np.testing.assert_array_equal(cm.cpu().numpy(), wm.numpy().astype(np.uint8))
    np.testing.assert_array_equal(areas.cpu().numpy(), wm.sum(dim=[1, 2]).numpy().astype(np.int32))
    np.testing.assert_allclose(ci.cpu().numpy(), wi.numpy(), atol=1e-6, rtol=0)


def test_batch_predict_matches_oracle(setup):
    got = setup["ours"].batch_predict(setup["img"], setup["masks"], batch_size=8)
    want = setup["orc"].batch_predict(setup["img"], setup["masks"], batch_size=8)
    assert got.shape == want.shape == (setup["masks"].shape[0], 3)
    zero_rows = (want.sum(1) == 0)
    np.testing.assert_array_equal((got.sum(1) == 0), zero_rows)  # the same masks fall under min_area
    assert zero_rows.sum() >= 1 and (~zero_rows).sum() >= 5
    # bf16 backbone + head vs fp32 oracle: 2e-2 on probabilities (north_star tolerance)
    np.testing.assert_allclose(got, want, atol=2e-2, rtol=0)
    np.testing.assert_allclose(got[~zero_rows].sum(1), 1.0, atol=1e-5)


def test_apply_classifier_filter_path(setup):
    """saber2D._apply_classifier's classifier branch (REF segmenters/base.py:159-176 -> filters/masks.py:9-59)."""
    from oracle import saber_ref
    from saber_b200.filters import masks as filters
    masks = [{"segmentation": m.astype(bool), "area": int(m.sum()), "stability_score": 1.0} for m in setup["masks"][:10]]
    preds = setup["ours"].batch_predict(setup["img"], setup["masks"][:10], 32)
    got = filters.apply_classifier(setup["img"], [dict(m) for m in masks], setup["ours"], 1, 32)
    want = saber_ref.convert_predictions_to_masks(preds, [dict(m) for m in masks], 1, 32)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g["segmentation"], w["segmentation"])
        assert g["area"] == w["area"] and g["bbox"] == w["bbox"]
