"""-m gpu: label-volume filters next to the path against golden vectors produced by the
REFERENCE's own functions (oracle/make_golden_3d.py): bit-exact uint8 / binary outputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fast_3d_gaussian_smoothing_matches_reference_golden(golden_dir):
    from saber_b200.filters.masks import fast_3d_gaussian_smoothing
    g = np.load(os.path.join(golden_dir, "saber3d_fast_gauss3d.npz"))
    out = fast_3d_gaussian_smoothing(g["labels_in"], scale=0.075, deviceID=0)
    assert out.dtype == np.uint8 and out.shape == g["out"].shape
    mism = int((out != g["out"]).sum())
    # fp32 separable sums vs the reference's conv3d summation order: voxels whose smoothed value sits within an ulp of
    # the 0.5 threshold may flip; none do on this fixture
    assert mism == 0, mism


def test_ball_morphology_matches_reference_golden(golden_dir):
    from saber_b200.analysis import morphology as M
    g = np.load(os.path.join(golden_dir, "saber3d_morphology.npz"))
    roi = torch.from_numpy(g["roi"].astype(np.float32)).cuda()
    for r in (1, 2, 3):
        np.testing.assert_array_equal(M.torch_erosion_3d(roi, r).cpu().numpy().astype(np.uint8), g[f"erode{r}"])
        np.testing.assert_array_equal(M.torch_dilation_3d(roi, r).cpu().numpy().astype(np.uint8), g[f"dilate{r}"])
        np.testing.assert_array_equal(M.morphological_opening(roi, r).cpu().numpy().astype(np.uint8), g[f"open{r}"])
    z = torch.zeros(4, 5, 6, device="cuda")
    assert M.morphological_opening(z, 2) is z


def test_gaussian_z_and_slab_matches_reference_golden(golden_dir):
    """R3 (z-Gaussian) vs the reference-run golden; normalize + project_tomogram vs the oracle restatement."""
    from oracle import saber_ref
    from saber_b200 import ops, synth
    from saber_b200.filters.gaussian import gaussian_smoothing
    g = np.load(os.path.join(golden_dir, "saber3d_gaussian_z.npz"))
    vol = synth.make_tomogram(tuple(g["shape"]), seed=int(g["seed"]), n_ellipsoids=5).numpy()
    out = gaussian_smoothing(vol, 5, dim=0, device="cuda:0")
    np.testing.assert_allclose(out, g["out"], atol=2e-6, rtol=0)
    v = torch.from_numpy(out).cuda()
    nv = ops.minmax_affine(v, ops.minmax(v), 1e-8, 1.0, 0.0)
    want = saber_ref.normalize(out)
    np.testing.assert_allclose(nv.cpu().numpy(), want, atol=1e-6, rtol=0)
    proj = ops.mean_z(nv, 10, 30)
    np.testing.assert_allclose(proj.cpu().numpy(), saber_ref.project_tomogram(want, 20, 10), atol=1e-6, rtol=0)


@pytest.mark.parametrize("dim", [1, 2, -1])
def test_gaussian_smoothing_other_axes_match_oracle(dim):
    """REF saber/filters/gaussian.py:17-74 smooths along any axis (default -1); only dim 0 is on the path, the others
    run the zero-padded correlation kernel along y / x. fp32, 1e-5 vs the oracle restatement of F.conv1d."""
    from oracle import saber_ref
    from saber_b200 import synth
    from saber_b200.filters.gaussian import gaussian_smoothing
    vol = synth.make_tomogram((20, 37, 45), seed=23, n_ellipsoids=4).numpy()
    out = gaussian_smoothing(vol, 3, dim=dim, device="cuda:0")
    assert isinstance(out, np.ndarray) and out.shape == vol.shape
    np.testing.assert_allclose(out, saber_ref.gaussian_smoothing(vol, 3, dim=dim % 3), atol=1e-5, rtol=0)
    t = gaussian_smoothing(torch.from_numpy(vol).cuda(), 3, dim=dim)
    assert t.is_cuda and torch.equal(t.cpu(), torch.from_numpy(out))


def test_segment_tomogram_core_smooths_and_writes_uint8(golden_dir):
    """SURVEY §8f row 1 (REF saber/entry_points/inference_core.py:10-93): reader -> segmenter.segment -> adaptive Gaussian
    smoothing (scale 0.05) -> uint8 -> writers.segmentation(run, mask, 'saber', name, session_id, voxel_size). In-memory
    stand-ins for the copick run / reader / writer and the segmenter; the written volume must equal the oracle's
    smoothing of the segmenter's label volume, cast to uint8."""
    import numpy as np
    from oracle import saber_ref
    from saber_b200 import synth
    from saber_b200.entry_points.inference_core import segment_tomogram_core
    labels = synth.make_label_volume((24, 64, 72), seed=5, n_ellipsoids=6, rmin=6.0, rmax=14.0).numpy().view(np.uint16)

    class Run:
        name = "run_001"

    class Reader:
        @staticmethod
        def tomogram(run, voxel_size, algorithm=None):
            return np.zeros((24, 64, 72), dtype=np.float32) if run is not None else None

    written = {}

    class Writer:
        @staticmethod
        def segmentation(run, mask, user_id, name=None, session_id=None, voxel_size=None):
            written.update(run=run.name, mask=mask, user=user_id, name=name, session=session_id, voxel=voxel_size)

    class Segmenter:
        inference_state = "something"
        calls = []

        def segment(self, vol, thickness, *a, **kw):
            self.calls.append((a, kw))
            return labels

    seg = Segmenter()
    out = segment_tomogram_core(Run(), 10.0, "wbp", "organelles", "7", 10, 1, 30, False, seg, gpu_id=0, target_class=1,
                                reader=Reader, writer=Writer)
    assert out is None and seg.inference_state is None
    assert seg.calls[0][1] == {"target_class": 1, "save_run": "run_001-7", "display": False}
    want = saber_ref.fast_3d_gaussian_smoothing(labels, scale=0.05).astype(np.uint8)
    assert written["user"] == "saber" and written["name"] == "organelles" and written["session"] == "7"
    assert written["voxel"] == 10.0 and written["mask"].dtype == np.uint8
    assert np.array_equal(written["mask"], want)
    # multi-slab call shape (REF :52-55) and the "no tomogram" / "no segmentation" early returns
    segment_tomogram_core(Run(), 10.0, "wbp", "o", "7", 10, 3, 30, True, seg, reader=Reader, writer=Writer)
    assert seg.calls[1][0] == (3, 30, "run_001-7", True)
    written.clear()

    class Empty(Segmenter):
        def segment(self, *a, **kw):
            return None

    assert segment_tomogram_core(Run(), 10.0, "wbp", "o", "7", 10, 1, 30, False, Empty(), reader=Reader, writer=Writer) is None
    assert not written
