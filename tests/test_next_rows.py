"""SURVEY §8f "next" rows — the membrane-refinement workflow (row 3) and the Fourier-space rescale / band-pass (row 2).
CPU part: the oracle restatements against fixtures produced by the REFERENCE's own code (oracle/make_golden_next.py).
GPU part (-m gpu): the device implementations against the same fixtures and against the oracle on further inputs."""
import os

import numpy as np
import pytest
import torch

REFINE_CASES = [
    ("a", (40, 72, 80), 31, 4, 2.0, dict(ball_size=3, min_membrane_area=200, edge_trim_z=2, edge_trim_xy=2)),
    ("b", (80, 96, 96), 34, 2, 3.2, dict(ball_size=3, min_membrane_area=100, edge_trim_z=2, edge_trim_xy=2,
                                         keep_surface_membranes=True)),
    ("c", (36, 64, 64), 32, 3, 2.0, dict(ball_size=5, min_membrane_area=100, edge_trim_z=3, edge_trim_xy=3,
                                         min_roi_relative_size=0.1)),
    ("d", (24, 40, 40), 33, 2, 2.0, dict(ball_size=3, min_membrane_area=100000, edge_trim_z=2, edge_trim_xy=2)),
    ("e", (24, 40, 40), 33, 2, 2.0, dict(ball_size=3, min_membrane_area=50, edge_trim_z=0, edge_trim_xy=2)),
]


def _refine_inputs(case):
    from saber_b200 import synth
    _name, shape, seed, n_org, blob, _cfg = case
    return synth.make_organelle_membrane(shape, seed, n_org, blob=blob)


@pytest.mark.parametrize("case", REFINE_CASES, ids=[c[0] for c in REFINE_CASES])
def test_oracle_refine_membranes_matches_reference_golden(case, golden_dir):
    from oracle import saber_ref
    g = np.load(os.path.join(golden_dir, "saber_refine_membranes.npz"))
    org, mem = _refine_inputs(case)
    o, m = saber_ref.refine_membranes(org, mem, **case[5])
    np.testing.assert_array_equal(o, g[f"{case[0]}_organelles"])
    np.testing.assert_array_equal(m, g[f"{case[0]}_membranes"])
    if o.ndim == 4:
        np.testing.assert_array_equal(saber_ref.convert_to_3d_labels(o), g[f"{case[0]}_organelles_3d"])
        assert (o > 0).sum() < (org > 0).sum()  # the workflow removed something


def test_refine_golden_exercises_the_surface_filter(golden_dir):
    """case b differs from the same input without keep_surface_membranes (the internal blob goes away)."""
    from oracle import saber_ref
    case = REFINE_CASES[1]
    org, mem = _refine_inputs(case)
    cfg = dict(case[5])
    _, m_on = saber_ref.refine_membranes(org, mem, **cfg)
    cfg["keep_surface_membranes"] = False
    _, m_off = saber_ref.refine_membranes(org, mem, **cfg)
    assert (m_on > 0).sum() < (m_off > 0).sum()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("conn", [6, 26])
def test_ccl3d_connectivity_and_sizes_vs_scipy(conn):
    from scipy import ndimage as ndi
    from saber_b200 import ops
    rng = np.random.default_rng(5 + conn)
    for shape, p, min_vol in [((17, 33, 45), 0.35, 1), ((9, 64, 70), 0.22, 4), ((30, 31, 29), 0.5, 10), ((1, 50, 37), 0.4, 1)]:
        vol = (rng.random(shape) < p).astype(np.uint8)
        structure = np.ones((3, 3, 3)) if conn == 26 else None
        lab, n = ndi.label(vol, structure=structure)
        counts = np.bincount(lab.ravel())
        keep = counts >= min_vol
        keep[0] = False
        remap = np.zeros(n + 1, np.int32)
        remap[keep] = np.arange(1, keep.sum() + 1)
        want = remap[lab]
        labels, count, sizes = ops.ccl3d(torch.from_numpy(vol).cuda(), min_vol=min_vol, conn=conn, with_sizes=True)
        k = int(count.item())
        assert k == int(keep.sum())
        np.testing.assert_array_equal(labels.cpu().numpy(), want)
        np.testing.assert_array_equal(sizes.cpu().numpy()[:k], counts[keep])


@pytest.mark.gpu
def test_refine_kernels_vs_numpy():
    from scipy import ndimage as ndi
    from saber_b200 import ops
    rng = np.random.default_rng(11)
    Z, Y, X = 13, 37, 41
    lab = rng.integers(0, 6, (Z, Y, X)).astype(np.int32) * (rng.random((Z, Y, X)) < 0.3)
    lab[:, :5] = 0
    lab[3] = 0
    for dt in (np.int32, np.int64, np.uint8, np.int16):
        v = torch.from_numpy(lab.astype(dt)).cuda()
        present = torch.from_numpy((np.arange(Z) % 4 != 1).astype(np.uint8)).cuda()
        table = ops.label_bbox(v, present, 7).cpu().numpy()
        assert table[0, 7] == 0
        for k in range(1, 8):
            idx = np.argwhere((lab == k) & (np.arange(Z) % 4 != 1)[:, None, None])
            if len(idx) == 0:
                assert table[k, 6] == 0
                continue
            assert table[k, 6] == len(idx)
            np.testing.assert_array_equal(table[k, :3], idx.min(0))
            np.testing.assert_array_equal(table[k, 3:6], idx.max(0))
        assert ops.label_bbox(v, None, 3).cpu().numpy()[0, 7] != 0  # labels 4, 5 exceed the cap
        roi = (2, 3, 4, 11, 30, 40)
        got = ops.roi_binarize(v, roi, 2, present).cpu().numpy()
        want = ((lab == 2) & (np.arange(Z) % 4 != 1)[:, None, None])[2:11, 3:30, 4:40]
        np.testing.assert_array_equal(got, want.astype(np.uint8))
        np.testing.assert_array_equal(ops.roi_binarize(v, roi, -1).cpu().numpy(), (lab != 0)[2:11, 3:30, 4:40].astype(np.uint8))
        out = torch.zeros_like(v)
        ops.roi_paste(out, roi, torch.from_numpy(want.astype(np.uint8)).cuda(), 9)
        ref = np.zeros_like(lab)
        ref[2:11, 3:30, 4:40][want] = 9
        np.testing.assert_array_equal(out.cpu().numpy(), ref.astype(dt))
        ops.overlay_nonzero(out, v)
        ref[lab > 0] = lab[lab > 0]
        np.testing.assert_array_equal(out.cpu().numpy(), ref.astype(dt))
    a = (rng.random((Z, Y, X)) < 0.5).astype(np.uint8)
    b = (rng.random((Z, Y, X)) < 0.5).astype(np.uint8)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    np.testing.assert_array_equal(ops.mask_logic(ta, tb, "and").cpu().numpy(), a & b)
    np.testing.assert_array_equal(ops.mask_logic(ta, tb, "or").cpu().numpy(), a | b)
    np.testing.assert_array_equal(ops.mask_logic(ta, tb, "andnot").cpu().numpy(), a & (1 - b))
    np.testing.assert_array_equal(ops.z_any(torch.from_numpy((lab != 0).astype(np.uint8)).cuda()).cpu().numpy(),
                                  (lab != 0).any(axis=(1, 2)).astype(np.uint8))
    np.testing.assert_array_equal(ops.morph_cube(ta, 1, 0).cpu().numpy(),
                                  ndi.binary_erosion(a, structure=np.ones((3, 3, 3))).astype(np.uint8))
    for zt, t in [(2, 3), (0, 3), (2, 0), (7, 3), (2, 19)]:
        want = np.zeros((Z, Y, X), np.uint8)
        if 0 < zt < Z // 2 and 0 < t < Y // 2 and t < X // 2:
            want[zt:-zt, t:-t, t:-t] = (lab != 0)[zt:-zt, t:-t, t:-t]
        np.testing.assert_array_equal(ops.trim_binarize(torch.from_numpy(lab).cuda(), zt, t).cpu().numpy(), want)
    # largest component: the first among equals
    vol = np.zeros((4, 8, 8), np.uint8)
    vol[0, 0, :3] = 1
    vol[2, 4, 2:5] = 1
    vol[3, 7, 7] = 1
    labels, count, sizes = ops.ccl3d(torch.from_numpy(vol).cuda(), 1, 6, with_sizes=True)
    big = ops.label_select(labels, sizes, count, largest=True).cpu().numpy()
    want = np.zeros_like(vol)
    want[0, 0, :3] = 1
    np.testing.assert_array_equal(big, want)
    np.testing.assert_array_equal(ops.label_select(labels).cpu().numpy(), vol)
    over = ops.label_keep_ratio(labels, torch.from_numpy((vol * 0 + (np.arange(8) >= 3)[None, None, :]).astype(np.uint8)).cuda(),
                                sizes, 0.1).cpu().numpy()
    want = np.zeros_like(vol)
    want[2, 4, 2:5] = 1   # 2 of 3 voxels at x >= 3
    want[3, 7, 7] = 1
    np.testing.assert_array_equal(over, want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", REFINE_CASES, ids=[c[0] for c in REFINE_CASES])
def test_refine_membranes_matches_reference_golden(case, golden_dir):
    from saber_b200.analysis.refine_membranes import FilteringConfig, OrganelleMembraneFilter
    g = np.load(os.path.join(golden_dir, "saber_refine_membranes.npz"))
    org, mem = _refine_inputs(case)
    f = OrganelleMembraneFilter(FilteringConfig(**case[5]))
    res = f.run(org, mem, batch_processing=True)
    o, m = res["organelles"], res["membranes"]
    if g[f"{case[0]}_organelles"].ndim == 3:  # the reference returns the zero volume as a device tensor
        assert torch.is_tensor(o) and o.shape == org.shape and int(o.abs().sum()) == 0 and int(m.abs().sum()) == 0
        return
    assert isinstance(o, np.ndarray) and o.dtype == org.dtype
    np.testing.assert_array_equal(o, g[f"{case[0]}_organelles"])
    np.testing.assert_array_equal(m, g[f"{case[0]}_membranes"])
    np.testing.assert_array_equal(f.convert_to_3d_labels(o), g[f"{case[0]}_organelles_3d"])
    dev = f.run_device(torch.from_numpy(org), torch.from_numpy(mem))
    np.testing.assert_array_equal(dev["organelles"].cpu().numpy(), g[f"{case[0]}_organelles_3d"])
    np.testing.assert_array_equal(dev["membranes"].cpu().numpy(), g[f"{case[0]}_membranes_3d"])


@pytest.mark.gpu
def test_refine_membranes_vs_oracle_more_inputs():
    from oracle import saber_ref
    from saber_b200 import synth
    from saber_b200.analysis.refine_membranes import FilteringConfig, OrganelleMembraneFilter
    for seed, shape, n_org, cfg, dt in [
        (41, (48, 80, 72), 5, dict(ball_size=3, min_membrane_area=150, edge_trim_z=3, edge_trim_xy=2), np.int64),
        (42, (56, 64, 96), 3, dict(ball_size=7, min_membrane_area=80, edge_trim_z=1, edge_trim_xy=1, min_roi_relative_size=0.05), np.uint8),
        (43, (72, 88, 88), 2, dict(ball_size=3, min_membrane_area=100, edge_trim_z=2, edge_trim_xy=2, keep_surface_membranes=True), np.int16),
    ]:
        org, mem = synth.make_organelle_membrane(shape, seed, n_org, blob=3.2)
        org = org.astype(dt)
        want_o, want_m = saber_ref.refine_membranes(org, mem, **cfg)
        res = OrganelleMembraneFilter(FilteringConfig(**cfg)).run(org, mem.astype(np.float32))
        assert want_o.ndim == 4
        np.testing.assert_array_equal(res["organelles"], want_o)
        np.testing.assert_array_equal(res["membranes"], want_m)


# ---- §8f row 2: Fourier-crop rescale and cosine band-pass -----------------------------------------------------------------
RESCALE3D_CASES = [("r3a", (24, 58, 40), 51, 5.0, 10.0), ("r3b", (25, 45, 64), 52, (4.0, 5.0, 5.0), (7.0, 9.0, 12.0)),
                   ("r3c", (20, 29, 48), 53, 3.0, 4.3)]
RESCALE2D_CASES = [("r2a", (96, 116), 54, 2.0), ("r2b", (75, 128), 55, 3.3), ("r2c", (58, 50), 56, 1.0)]
FILTER_CASES = [("fa", (20, 48, 58), 57, 10.0, 60.0, 6.0, 0.0, 0.0), ("fb", (24, 40, 40), 58, 10.0, 50.0, 4.0, 400.0, 2.0),
                ("fc", (25, 45, 32), 59, 8.0, 40.0, 0.0, 0.0, 0.0), ("fd", (16, 32, 32), 60, 8.0, 0.0, 0.0, 200.0, 0.0)]
FOURIER_TOL = 1e-5  # of the largest magnitude in the expected array: fp32 transforms on both sides


def _close(got, want, tol=FOURIER_TOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    assert err < tol, err


def test_oracle_fourier_matches_reference_golden(golden_dir):
    from oracle import saber_ref
    from saber_b200 import synth
    g = np.load(os.path.join(golden_dir, "saber_fourier.npz"))
    for name, shape, seed, vin, vout in RESCALE3D_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        _close(saber_ref.fourier_rescale_3d(vol, vin, vout), g[name])
    for name, shape, seed, sf in RESCALE2D_CASES:
        img = synth.make_tomogram((1, *shape), seed=seed, n_ellipsoids=3).numpy()[0]
        _close(saber_ref.fourier_rescale_2d(img, sf), g[name])
    for name, shape, seed, apix, lp, lpd, hp, hpd in FILTER_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        filt = saber_ref.cosine_filter(shape, apix, lp, lpd, hp, hpd)
        _close(filt, g[name + "_filter"], 1e-6)
        _close(saber_ref.filter3d_apply(vol, filt), g[name])


@pytest.mark.gpu
def test_fft_lines_vs_numpy():
    from saber_b200 import ops
    rng = np.random.default_rng(3)
    for shape in [(6, 20, 29), (3, 58, 12), (5, 7, 64), (4, 200, 45), (2, 9, 928), (3, 77, 91), (2, 1, 960), (1, 13, 1)]:
        x = rng.normal(size=shape).astype(np.float32)
        t = torch.from_numpy(x).cuda()
        for axis in (0, 1, 2):
            got = ops.fft_lines(t, axis).cpu().numpy()
            want = np.fft.fft(x.astype(np.float64), axis=axis)
            assert np.abs(got - want).max() / np.abs(want).max() < 2e-6, (shape, axis)
            c = torch.from_numpy(want.astype(np.complex64)).cuda()
            back = ops.fft_lines(c, axis, inverse=True, out_mode="real", scale=1.0 / shape[axis]).cpu().numpy()
            assert np.abs(back - x).max() < 5e-6 * np.abs(x).max(), (shape, axis)
    img = rng.normal(size=(33, 40)).astype(np.float32)
    got = ops.fft_lines(torch.from_numpy(img).cuda(), 1, crop=(11, 18)).cpu().numpy()
    want = np.fft.ifftshift(np.fft.fftshift(np.fft.fft(img.astype(np.float64), axis=1), axes=1)[:, 11:29], axes=1)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6


@pytest.mark.gpu
def test_fourier_rescale_matches_reference_golden(golden_dir):
    from saber_b200 import synth
    from saber_b200.filters.downsample import FourierRescale2D, FourierRescale3D
    g = np.load(os.path.join(golden_dir, "saber_fourier.npz"))
    for name, shape, seed, vin, vout in RESCALE3D_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        r = FourierRescale3D(vin, vout)
        out = r.run(vol)
        assert isinstance(out, np.ndarray) and out.dtype == np.float32
        _close(out, g[name])
        batched = r.run(torch.from_numpy(np.stack([vol, 2 * vol])))
        assert torch.is_tensor(batched) and batched.device.type == "cpu"
        _close(batched[1].numpy(), 2 * g[name])
    with pytest.raises(ValueError):
        FourierRescale3D(10.0, 5.0)
    for name, shape, seed, sf in RESCALE2D_CASES:
        img = synth.make_tomogram((1, *shape), seed=seed, n_ellipsoids=3).numpy()[0]
        _close(FourierRescale2D.run(img, sf), g[name])
    with pytest.raises(ValueError):
        FourierRescale2D.run(np.zeros((8, 8), np.float32), 0.5)
    _close(FourierRescale2D.run_resolution(synth.make_tomogram((1, 96, 116), seed=54, n_ellipsoids=3).numpy()[0], 1.5, 3.0), g["r2a"])
    stack = synth.make_tomogram((5, 75, 128), seed=61, n_ellipsoids=6, device="cuda").contiguous()
    got = FourierRescale2D.rescale_stack_device(stack, 3.3)
    for z in range(5):
        assert torch.equal(got[z], FourierRescale2D.rescale_device(stack[z].contiguous(), 3.3))


@pytest.mark.gpu
def test_filter3d_matches_reference_golden(golden_dir):
    from saber_b200 import synth
    from saber_b200.filters.tomograms import Filter3D
    g = np.load(os.path.join(golden_dir, "saber_fourier.npz"))
    for name, shape, seed, apix, lp, lpd, hp, hpd in FILTER_CASES:
        vol = synth.make_tomogram(shape, seed=seed, n_ellipsoids=4).numpy()
        f = Filter3D(apix, shape, lp=lp, lpd=lpd, hp=hp, hpd=hpd)
        _close(f.filter.cpu().numpy(), g[name + "_filter"], 1e-6)
        out = f.apply(vol)
        assert out.is_cuda and out.dtype == torch.float32
        _close(out.cpu().numpy(), g[name])
    with pytest.raises(ValueError):
        Filter3D(10.0, (8, 8, 8), lp=100.0, hp=50.0)


@pytest.mark.gpu
def test_fourier_full_size_properties():
    """BASELINE-size volume (200 x 928 x 960): identity voxel size round-trips the volume; a 2x rescale preserves the mean
    (the DC term) and halves every extent; an all-pass Filter3D is the identity."""
    from saber_b200 import synth
    from saber_b200.filters.downsample import FourierRescale3D
    from saber_b200.filters.tomograms import Filter3D
    shape = (200, 928, 960)
    vol = synth.make_tomogram(shape, seed=7, n_ellipsoids=30, device="cuda").contiguous()
    same = FourierRescale3D(10.0, 10.0).rescale_device(vol)
    assert same.shape == vol.shape
    assert float((same - vol).abs().max()) < 2e-5 * float(vol.abs().max())
    half = FourierRescale3D(10.0, 20.0).rescale_device(vol)
    assert tuple(half.shape) == (100, 464, 480)
    # ortho norms: mean(out) = mean(in) * sqrt(N_in / N_out)
    assert abs(float(half.double().mean()) - float(vol.double().mean()) * (8 ** 0.5)) < 1e-4
    ident = Filter3D(10.0, shape).apply(vol)
    assert float((ident - vol).abs().max()) < 2e-5 * float(vol.abs().max())


# ---- §8f row 4: training-data store layout and the prep3d / prep2d workers ---------------------------------------------------
def _read_zarr_v2(path):
    """Minimal zarr-v2 reader (json + zlib / numcodecs) used only to check what the writer produced."""
    import json
    import zlib
    meta = json.load(open(os.path.join(path, ".zarray")))
    assert meta["zarr_format"] == 2 and meta["dimension_separator"] == "/" and meta["order"] == "C"
    comp = meta["compressor"]
    if comp["id"] == "zlib":
        decode = zlib.decompress
    else:
        import numcodecs
        decode = numcodecs.get_codec(comp).decode
    shape, chunks = meta["shape"], meta["chunks"]
    out = np.zeros(shape, np.dtype(meta["dtype"]))
    if len(shape) >= 3:
        for i in range(shape[0]):
            p = os.path.join(path, str(i), *["0"] * (len(shape) - 1))
            out[i] = np.frombuffer(decode(open(p, "rb").read()), out.dtype).reshape(chunks[1:])
    else:
        p = os.path.join(path, *["0"] * len(shape))
        out[...] = np.frombuffer(decode(open(p, "rb").read()), out.dtype).reshape(shape)
    return out


class _Cfg:
    class amg_cfg:  # noqa: N801
        @staticmethod
        def to_dict():
            return {"npoints": 32, "pred_iou_thresh": np.float32(0.7), "sam2_cfg": "large"}


class _SlabSegmenter:
    adapter_cfg = _Cfg

    def __init__(self):
        self.calls = []

    def segment_slab(self, vol, slab_thickness, display=False, zSlice=None):
        self.calls.append((slab_thickness, zSlice))
        rng = np.random.default_rng(zSlice)
        self.image0 = vol[zSlice].astype(np.float32)
        areas = [50, 10, 30]
        self.masks = []
        for a in areas:
            m = np.zeros(vol.shape[1:], bool)
            m.ravel()[rng.choice(m.size, a, replace=False)] = True
            self.masks.append({"segmentation": m, "area": a})


def test_zarr_layout_and_prep3d_worker(tmp_path, monkeypatch):
    import json
    from saber_b200.classifier.preprocess import tomo_prep
    from saber_b200.utils import zarr_writer
    monkeypatch.setattr(zarr_writer, "_zarr_writer", None)
    out = str(tmp_path / "training.zarr")
    vol = np.random.default_rng(0).normal(size=(30, 24, 28)).astype(np.float32)

    class Run:
        name = "TS_001"

    class Reader:
        @staticmethod
        def tomogram(run, voxel_size, algorithm):
            return vol if algorithm == "wbp" else None

    seg = _SlabSegmenter()
    tomo_prep.extract_sam2_candidates(Run, out, 10.0, "wbp", 4, 3, 0, {"segmenter": seg}, reader=Reader)
    tomo_prep.extract_sam2_candidates(Run, out, 10.0, "missing", 4, 1, 0, {"segmenter": seg}, reader=Reader)  # no tomogram
    Run.name = "TS_002"
    tomo_prep.extract_sam2_candidates(Run, out, 10.0, "wbp", 4, 1, 0, {"segmenter": seg}, reader=Reader)
    w = zarr_writer.get_zarr_writer(out)
    w.finalize()
    # REF tomo_prep.py:61-72: slabs centred on the volume, one thickness apart
    assert seg.calls == [(4, 11), (4, 15), (4, 19), (4, 15)]
    root = json.load(open(os.path.join(out, ".zattrs")))
    assert root["total_runs"] == 4 and root["creation_complete"] is True
    assert root["amg"] == {"npoints": 32, "pred_iou_thresh": pytest.approx(0.7), "sam2_cfg": "large"}
    assert sorted(d for d in os.listdir(out) if not d.startswith(".")) == ["TS_001_1", "TS_001_2", "TS_001_3", "TS_002"]
    for group, z in [("TS_001_1", 11), ("TS_001_3", 19), ("TS_002", 15)]:
        g = os.path.join(out, group)
        assert json.load(open(os.path.join(g, ".zgroup"))) == {"zarr_format": 2}
        np.testing.assert_array_equal(_read_zarr_v2(os.path.join(g, "0")), vol[z])
        masks = _read_zarr_v2(os.path.join(g, "labels", "0"))
        assert masks.dtype == np.uint8 and masks.shape == (3, 24, 28)
        assert [int((masks[j] == j + 1).sum()) for j in range(3)] == [10, 30, 50]  # sorted by area, labelled j + 1
        ms = json.load(open(os.path.join(g, ".zattrs")))["multiscales"][0]
        assert [a["name"] for a in ms["axes"]] == ["y", "x"] and ms["version"] == "0.4" and ms["name"] == "/"
        assert ms["datasets"][0] == {"coordinateTransformations": [{"scale": [1.0, 1.0], "type": "scale"}], "path": "0"}
        ml = json.load(open(os.path.join(g, "labels", ".zattrs")))["multiscales"][0]
        assert [a["name"] for a in ml["axes"]] == ["z", "y", "x"]
        assert ml["datasets"][0]["coordinateTransformations"][0]["scale"] == [1.0, 1.0, 1.0]
    with pytest.raises(ValueError):
        w.write("TS_002", vol[0], np.zeros((1, 24, 28), np.uint8))  # a run is written once
    w.set_dict_attr("amg", {"npoints": 64, "extra": 1}, merge_missing=True)
    assert json.load(open(os.path.join(out, ".zattrs")))["amg"]["npoints"] == 32
    assert json.load(open(os.path.join(out, ".zattrs")))["amg"]["extra"] == 1


@pytest.mark.gpu
def test_segment_micrograph_core_rescales_and_writes(tmp_path, monkeypatch):
    import json
    from saber_b200 import synth
    from saber_b200.entry_points.inference_core import segment_micrograph_core
    from saber_b200.filters.downsample import FourierRescale2D
    from saber_b200.utils import zarr_writer
    monkeypatch.setattr(zarr_writer, "_zarr_writer", None)
    img = synth.make_tomogram((1, 96, 116), seed=54, n_ellipsoids=3).numpy()[0]

    class Seg:
        adapter_cfg = _Cfg

        def segment(self, image, target_class=-1, display=False, use_sliding_window=False):
            self.image = image
            self.args = (target_class, use_sliding_window)
            m = np.zeros(image.shape, bool)
            m[2:10, 3:9] = True
            self.masks = [{"segmentation": m, "area": int(m.sum())}]

    seg = Seg()
    out = str(tmp_path / "micro.zarr")
    segment_micrograph_core("/data/mic_07.mrc", out, None, 4.0, False, True, 0, {"segmenter": seg, "target_class": 2},
                            read_micrograph=lambda f: (img.astype(np.float64), np.float32(2.0)))
    assert seg.args == (2, True)
    want = FourierRescale2D.run(img, 2.0)
    np.testing.assert_array_equal(seg.image, want)
    g = os.path.join(out, "mic_07")
    np.testing.assert_array_equal(_read_zarr_v2(os.path.join(g, "0")), want)
    assert _read_zarr_v2(os.path.join(g, "labels", "0")).shape == (1, 48, 58)
    scale = json.load(open(os.path.join(g, ".zattrs")))["multiscales"][0]["datasets"][0]["coordinateTransformations"][0]["scale"]
    assert scale == [pytest.approx(0.2), pytest.approx(0.2)]


@pytest.mark.gpu
def test_prep2d_chain_loader_rescale_segment_store(tmp_path, monkeypatch, capsys):
    """The prep2d worker end to end on the real pieces: GPUPool loader (`base_microsegmenter`) -> `cryoMicroSegmenter` ->
    Fourier-crop down-sampling -> AMG -> candidate stack -> zarr group (REF micro_prep.py:104-133, inference_core.py:97-153)."""
    import json
    from saber_b200 import synth
    from saber_b200.adapters.base import cfgAMG
    from saber_b200.entry_points.inference_core import segment_micrograph_core
    from saber_b200.filters.downsample import FourierRescale2D
    from saber_b200.segmenters.loaders import base_microsegmenter
    from saber_b200.utils import zarr_writer
    monkeypatch.setattr(zarr_writer, "_zarr_writer", None)
    models = base_microsegmenter(0, cfgAMG(sam2_cfg="tiny", npoints=8, crop_n_layers=0, pred_iou_thresh=0.3,
                                           stability_score_thresh=0.0))
    seg = models["segmenter"]
    assert type(seg).__name__ == "cryoMicroSegmenter" and seg.max_pixels == 1280
    img = synth.make_tomogram((1, 400, 512), seed=63, n_ellipsoids=8).numpy()[0]
    out = str(tmp_path / "prep2d.zarr")
    segment_micrograph_core("/data/grid_03.tif", out, 2.0, None, False, False, 0, models, read_micrograph=lambda f: (img, None))
    small = FourierRescale2D.run(img, 2.0)
    assert small.shape == (200, 256)
    g = os.path.join(out, "grid_03")
    stored = _read_zarr_v2(os.path.join(g, "0"))
    assert stored.shape[:2] == (200, 256)
    masks = _read_zarr_v2(os.path.join(g, "labels", "0"))
    assert masks.ndim == 3 and masks.shape[1:] == (200, 256) and masks.shape[0] == len(seg.masks) > 0
    for j in range(masks.shape[0]):  # labelled stack: plane j holds value j + 1 on its mask
        np.testing.assert_array_equal(masks[j] > 0, np.asarray(seg.masks[j]["segmentation"], bool))
        assert set(np.unique(masks[j])) <= {0, j + 1}
    scale = json.load(open(os.path.join(g, ".zattrs")))["multiscales"][0]["datasets"][0]["coordinateTransformations"][0]["scale"]
    assert scale == [1, 1]  # no pixel size in a tiff: REF :135-138 falls back to 1
    assert json.load(open(os.path.join(out, ".zattrs")))["amg"]["npoints"] == 8
    big = np.random.default_rng(0).normal(size=(1300, 64)).astype(np.float32)
    seg.segment(big, display=False)
    assert "Consider Downsampling" in capsys.readouterr().out


@pytest.mark.gpu
def test_gpupool_contract(tmp_path, monkeypatch):
    """GPUPool (REF saber/utils/parallelization.py): task i -> GPU i % n, models from init_fn once per GPU, result dicts
    sorted by task id, failures reported not raised; then the pool drives the prep3d worker as `prep3d` does
    (REF tomo_prep.py:150-170: init_fn = base_tomosegmenter-like loader, func = extract_sam2_candidates)."""
    import json
    from saber_b200.classifier.preprocess import tomo_prep
    from saber_b200.utils import zarr_writer
    from saber_b200.utils.parallelization import GPUPool, gpu_map
    n = torch.cuda.device_count()
    loads = []

    def init(gpu_id, tag, scale=1):
        loads.append(gpu_id)
        return {"tag": tag, "scale": scale, "dev": gpu_id}

    def work(x, y=0, gpu_id=None, models=None):
        if x == 3:
            raise ValueError("boom")
        assert torch.cuda.current_device() == gpu_id == models["dev"]
        return float((torch.full((4,), float(x), device="cuda") * models["scale"]).sum()) + y

    pool = GPUPool(init_fn=init, init_args=("t",), init_kwargs={"scale": 2}, verbose=False)
    res = pool.execute(work, [0, (1,), ((2,), {"y": 5}), 3, {"x": 4, "y": 1}], task_ids=[10, 11, 12, 13, 14])
    assert loads == list(range(n)) and [r["task_id"] for r in res] == [10, 11, 12, 13, 14]
    assert [r["gpu_id"] for r in res] == [i % n for i in range(5)]
    assert [r["success"] for r in res] == [True, True, True, False, True] and res[3]["error"] == "boom"
    assert [r["result"] for r in res if r["success"]] == [0.0, 8.0, 21.0, 33.0]
    assert pool.execute(work, []) == []
    pool.shutdown()
    with pytest.raises(ValueError):
        GPUPool(approach="multiprocessing")
    assert [r["result"] for r in gpu_map(lambda v, gpu_id=None: v * 2, [1, 2, 3], verbose=False)] == [2, 4, 6]

    # the prep3d wiring with a stand-in slab segmenter
    monkeypatch.setattr(zarr_writer, "_zarr_writer", None)
    out = str(tmp_path / "prep3d.zarr")
    vol = np.random.default_rng(1).normal(size=(20, 24, 28)).astype(np.float32)

    class Reader:
        @staticmethod
        def tomogram(run, voxel_size, algorithm):
            return vol

    class Run:
        def __init__(self, name):
            self.name = name

    runs = [Run(f"TS_{i:02d}") for i in range(5)]
    with GPUPool(init_fn=lambda g: {"segmenter": _SlabSegmenter()}, verbose=False) as p3:
        res = p3.execute(tomo_prep.extract_sam2_candidates,
                         [((r, out, 10.0, "wbp", 4, 1), {"reader": Reader}) for r in runs], task_ids=[r.name for r in runs])
    assert all(r["success"] for r in res), [r.get("error") for r in res]
    zarr_writer.get_zarr_writer(out).finalize()
    assert json.load(open(os.path.join(out, ".zattrs")))["total_runs"] == 5
    assert sorted(d for d in os.listdir(out) if not d.startswith(".")) == [r.name for r in runs]


def test_host_helpers_of_the_next_rows():
    """Pure host logic (no GPU): GPUPool task forms (REF parallelization.py:143-151), the Fourier crop window
    (REF downsample.py:117-127, 183-191) and the band-pass parameter packing (REF tomograms.py:44-52)."""
    from saber_b200.filters.downsample import _crop_window
    from saber_b200.utils.parallelization import _split_task
    assert _split_task({"a": 1}) == ((), {"a": 1})
    assert _split_task(((1, 2), {"k": 3})) == ((1, 2), {"k": 3})
    assert _split_task([1, 2, 3]) == ((1, 2, 3), {})
    assert _split_task((7,)) == ((7,), {})
    assert _split_task("run") == (("run",), {})
    for n in (64, 65, 200, 928, 929):
        for m in (n, n - 1, n // 2, n // 2 + 1, 3, 2):
            start, length = _crop_window(n, m)
            assert length == m - (m % 2) and start == (n - length) // 2 + (n % 2)
            assert 0 <= start and start + length <= n, (n, m)
            # the kept band is centred on the zero frequency of the fftshift-ed axis (index n // 2)
            assert start <= n // 2 < start + length or length == 0
    import types
    f = types.SimpleNamespace(lp_pix=12.5, lpd_pix=4, hp_pix=0, hpd_pix=0)
    from saber_b200.filters.tomograms import Filter3D
    assert Filter3D._bandpass(f) == [12.5, 10.5, 14.5, 4.0, 0.0, 0.0, 0.0, 0.0]


def _stockham_sim(x, inverse=False):
    """Scalar model of csrc/fft.cu's stage arithmetic (same index expressions, same root-table look-ups): radix-4 / 2 first,
    radix-3 / 5 butterflies, and for any other odd prime R the pre-twiddle pass + output PAIRS (t, R - t) built from
    a_r = v_r + v_{R-r}, b_r = v_r - v_{R-r}. Documents the kernel; checked against numpy's FFT below."""
    n = len(x)
    fac, m = [], n
    while m % 4 == 0:
        fac.append(4); m //= 4
    while m % 2 == 0:
        fac.append(2); m //= 2
    p = 3
    while m > 1:
        while m % p == 0:
            fac.append(p); m //= p
        p += 2
    tw = np.exp(-2j * np.pi * np.arange(n) / n)
    sgn = -1.0 if inverse else 1.0
    root = lambda idx: complex(tw[idx].real, tw[idx].imag * sgn)
    cur, Ns = np.asarray(x, complex).copy(), 1
    for R in fac:
        q, tstep = n // R, n // R // Ns
        nxt = np.zeros(n, complex)
        if R in (2, 3, 4, 5):
            W = np.exp(-2j * np.pi * sgn * np.outer(np.arange(R), np.arange(R)) / R)  # the register butterfly
            for j in range(q):
                k = j % Ns
                v = np.array([cur[j + r * q] * root(r * k * tstep) for r in range(R)])
                j0 = (j - k) * R + k
                nxt[j0 + np.arange(R) * Ns] = W @ v
        else:
            h = (R - 1) // 2
            for i in range(n):  # stage twiddles in place
                r, j = divmod(i, q)
                k = j % Ns
                if r and k:
                    cur[i] *= root(r * k * tstep)
            for j in range(q):
                k = j % Ns
                j0 = (j - k) * R + k
                nxt[j0] = sum(cur[j + r * q] for r in range(R))
                for t in range(1, h + 1):
                    A, B, xi = cur[j], 0j, 0
                    for r in range(1, h + 1):
                        xi = (xi + t) % R
                        w = tw[xi * q]  # (cos, -sin) of 2 pi xi / R, unconjugated in both directions
                        u, z = cur[j + r * q], cur[j + (R - r) * q]
                        A += (u + z) * w.real
                        B += (u - z) * w.imag
                    iB = complex(-sgn * B.imag, sgn * B.real)
                    nxt[j0 + t * Ns], nxt[j0 + (R - t) * Ns] = A + iB, A - iB
        cur, Ns = nxt, Ns * R
    return cur


def test_fft_kernel_stage_arithmetic_model():
    rng = np.random.default_rng(0)
    for n in (1, 2, 12, 29, 58, 77, 200, 232, 960 // 8, 105):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        assert np.abs(_stockham_sim(x) - np.fft.fft(x)).max() < 1e-10 * max(1, n)
        assert np.abs(_stockham_sim(x, True) - np.fft.ifft(x) * n).max() < 1e-10 * max(1, n)
    # the store index of a cropping pass: bin (start + (q + m/2) mod m - n/2) mod n == ifftshift(crop(fftshift(.)))
    for n, m in ((40, 18), (33, 18), (29, 14), (16, 16)):
        start = (n - m) // 2 + (n % 2)
        spec = np.arange(n)
        want = np.fft.ifftshift(np.fft.fftshift(spec)[start:start + m])
        got = [(start + (qo + m // 2) % m - n // 2) % n for qo in range(m)]
        assert list(want) == got


def _ccl_model(vol, conn):
    """Sequential model of csrc/ccl3d.cu: parents start at the head of the foreground run inside a 32-voxel segment of the
    row-major index space (runs never cross a row end); the merge pass performs only the unions the run structure does not
    already imply (26: centre of each preceding row first, +-1 diagonals only when the centre is background, nothing but the
    +1 diagonal when the left neighbour is set; 6: up / back unless the left neighbour and ITS up / back are set)."""
    Z, Y, X = vol.shape
    fg = (vol != 0).ravel()
    n = fg.size
    P = np.full(n, -1, np.int64)
    for i in range(n):
        if fg[i]:
            left_ok = i % 32 != 0 and i % X != 0 and fg[i - 1]
            P[i] = P[i - 1] if left_ok else i

    def find(v):
        while P[v] != v:
            v = P[v]
        return v

    def unite(a, b):
        a, b = find(a), find(b)
        if a != b:
            P[max(a, b)] = min(a, b)

    plane = Y * X
    for i in range(n):
        if not fg[i]:
            continue
        x, y, z = i % X, (i // X) % Y, i // plane
        left = x > 0 and fg[i - 1]
        if left and i % 32 == 0:
            unite(i, i - 1)
        if conn == 6:
            if y > 0 and fg[i - X] and not (left and fg[i - X - 1]):
                unite(i, i - X)
            if z > 0 and fg[i - plane] and not (left and fg[i - plane - 1]):
                unite(i, i - plane)
            continue
        rows = ([i - X] if y > 0 else []) + ([i - plane - X] if z > 0 and y > 0 else []) + ([i - plane] if z > 0 else []) + \
               ([i - plane + X] if z > 0 and y + 1 < Y else [])
        for r in rows:
            if fg[r]:
                if not left:
                    unite(i, r)
                continue
            if not left and x > 0 and fg[r - 1]:
                unite(i, r - 1)
            if x + 1 < X and fg[r + 1]:
                unite(i, r + 1)
    roots = np.array([find(i) if fg[i] else -1 for i in range(n)])
    ids = {r: k + 1 for k, r in enumerate(sorted(set(roots[roots >= 0])))}  # raster order of the first voxel
    return np.array([ids[r] if r >= 0 else 0 for r in roots]).reshape(vol.shape)


@pytest.mark.parametrize("conn", [6, 26])
def test_ccl_merge_rules_model_vs_scipy(conn):
    from scipy import ndimage as ndi
    rng = np.random.default_rng(100 + conn)
    structure = np.ones((3, 3, 3)) if conn == 26 else None
    for shape, p in [((3, 5, 37), 0.5), ((4, 6, 33), 0.3), ((2, 9, 70), 0.65), ((5, 4, 8), 0.4), ((1, 7, 64), 0.55)]:
        vol = (rng.random(shape) < p).astype(np.uint8)
        want, _ = ndi.label(vol, structure=structure)
        np.testing.assert_array_equal(_ccl_model(vol, conn), want)
