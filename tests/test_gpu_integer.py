"""-m gpu: integer / indexing kernels must be BIT-EXACT against the oracle (oracle/amg_post_ref.py,
oracle/saber_ref.py) on seeded inputs, on the reference-derived golden fixtures, and — at BASELINE sizes —
through size-independent properties (idempotence, permutation invariance, counts)."""
import os

import numpy as np
import pytest
import torch

from util import synth_boxes, synth_logits

pytestmark = pytest.mark.gpu
I32, U8, F32 = torch.int32, torch.uint8, torch.float32


@pytest.fixture(scope="module")
def ops():
    from saber_b200 import ops as _ops
    _ops.require_b200()
    return _ops


def _unpack(bits, W):
    b = bits.cpu().numpy().view(np.uint32)
    m, H, WW = b.shape
    return np.unpackbits(b.view(np.uint8).reshape(m, H, WW * 4), axis=-1, bitorder="little")[:, :, :W].astype(bool)


@pytest.mark.parametrize("hw,out_hw", [((256, 256), (1024, 1024)), ((256, 256), (683, 684)), ((256, 256), (928, 960)),
                                       ((256, 256), (256, 256)), ((64, 64), (200, 333))])
def test_upsample_bilinear_bitexact(ops, hw, out_hw):
    from oracle import amg_post_ref as R
    rng = np.random.default_rng(7)
    p = (rng.normal(size=(3,) + hw) * 4).astype(np.float32)
    out = ops.upsample_bilinear(torch.from_numpy(p).cuda(), *out_hw).cpu().numpy()
    np.testing.assert_array_equal(out, R.upsample_bilinear(p, out_hw))


@pytest.mark.parametrize("crop,hw,cpp", [((0, 0, 1024, 1024), (1024, 1024), 1), ((341, 0, 1024, 683), (1024, 1024), 3),
                                         ((100, 40, 612, 400), (480, 640), 1), ((0, 0, 517, 300), (300, 517), 3),
                                         ((37, 11, 137, 75), (100, 150), 1)])
def test_amg_mask_post_bitexact(ops, crop, hw, cpp):
    from oracle import amg_post_ref as R
    x0, y0, x1, y1 = crop
    H, W = hw
    P = 12
    rng = np.random.default_rng(11)
    planes4 = synth_logits(P * 4, seed=21).reshape(P, 4, 256, 256)
    ious4 = rng.uniform(0.5, 1.0, (P, 4)).astype(np.float32)
    sel = rng.integers(0, 4, P).astype(np.int32) if cpp == 1 else None
    n = P * cpp
    prompt = np.arange(n) // cpp
    token = sel[prompt] if sel is not None else 1 + np.arange(n) % 3
    ref = R.mask_post(planes4[prompt, token], ious4[prompt, token], crop, hw, 0.7, 0.0, 0.7, 0.92)
    dev = "cuda"
    base, N = 5, n + 9
    keep = torch.zeros(N, dtype=U8, device=dev)
    stab = torch.zeros(N, dtype=F32, device=dev)
    iou = torch.zeros(N, dtype=F32, device=dev)
    bbox = torch.zeros((N, 4), dtype=I32, device=dev)
    area = torch.zeros(N, dtype=I32, device=dev)
    bits = torch.zeros((N, H, (W + 31) // 32), dtype=I32, device=dev)
    ops.amg_mask_post(torch.from_numpy(planes4).cuda(), torch.from_numpy(ious4).cuda(),
                      None if sel is None else torch.from_numpy(sel).cuda(), cpp, n, (y1 - y0, x1 - x0), (x0, y0), hw,
                      0.7, 0.0, 0.7, 0.92, keep, stab, iou, bbox, area, bits, base)
    k = keep[base:base + n].cpu().numpy().astype(bool)
    np.testing.assert_array_equal(k, ref["keep"])
    assert k.any() and not k.all()
    passed_iou = ious4[prompt, token] > np.float32(0.7)
    np.testing.assert_array_equal(stab[base:base + n].cpu().numpy()[passed_iou], ref["stability"][passed_iou])
    np.testing.assert_array_equal(bbox[base:base + n].cpu().numpy()[passed_iou], ref["bbox"][passed_iou])
    np.testing.assert_array_equal(area[base:base + n].cpu().numpy()[passed_iou], ref["area"][passed_iou])
    np.testing.assert_array_equal(_unpack(bits[base:base + n], W)[passed_iou], ref["masks"][passed_iou])
    assert keep[:base].sum().item() == 0 and keep[base + n:].sum().item() == 0
    unp = ops.unpack_bits(bits[base:base + n].contiguous(), None, n, W).cpu().numpy()
    np.testing.assert_array_equal(unp[passed_iou], ref["masks"][passed_iou])


def _run_nms(ops, boxes, scores, cand, thr):
    dev = "cuda"
    n_cap = max(len(cand), 1)
    cb = (n_cap + 63) // 64
    out_list = torch.full((n_cap + 3,), -1, dtype=I32, device=dev)
    cnt = torch.tensor([len(cand), 3], dtype=I32, device=dev)  # out_count starts at 3: appended after existing entries
    ops.nms_dev(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.tensor(cand, dtype=I32, device=dev),
                cnt[0:1], n_cap, thr, torch.empty(n_cap, dtype=I32, device=dev),
                torch.empty(n_cap * cb, dtype=torch.int64, device=dev), out_list, cnt[1:2])
    m = cnt[1].item() - 3
    return out_list[3:3 + m].cpu().numpy()


@pytest.mark.parametrize("n", [1, 7, 64, 65, 300, 3072, 9216])
def test_nms_bitexact(ops, n):
    from oracle import amg_post_ref as R
    boxes, scores = synth_boxes(n, seed=n)
    rng = np.random.default_rng(n)
    cand = np.sort(rng.choice(n, size=max(1, int(0.8 * n)), replace=False)).astype(np.int32)  # a sub-list of the slots
    got = _run_nms(ops, boxes, scores, cand.tolist(), 0.7)
    want = cand[R.nms(boxes[cand].astype(np.float32), scores[cand], 0.7)]
    np.testing.assert_array_equal(got, want)


def test_nms_empty_list(ops):
    boxes, scores = synth_boxes(8, seed=1)
    dev = "cuda"
    out_list = torch.full((8,), -1, dtype=I32, device=dev)
    cnt = torch.tensor([0, 0], dtype=I32, device=dev)
    ops.nms_dev(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.zeros(8, dtype=I32, device=dev),
                cnt[0:1], 8, 0.7, torch.empty(8, dtype=I32, device=dev), torch.empty(8, dtype=torch.int64, device=dev),
                out_list, cnt[1:2])
    assert cnt[1].item() == 0


def test_compact_keep(ops):
    rng = np.random.default_rng(3)
    keep = (rng.uniform(size=5000) < 0.3).astype(np.uint8)
    cand = torch.full((4000,), -1, dtype=I32, device="cuda")
    cnt = torch.zeros(1, dtype=I32, device="cuda")
    ops.compact_keep(torch.from_numpy(keep).cuda(), 100, 3333, cand, cnt)
    want = 100 + np.flatnonzero(keep[100:3433])
    assert cnt.item() == len(want)
    np.testing.assert_array_equal(cand[:len(want)].cpu().numpy(), want)


def test_remove_duplicates_matches_reference_golden(ops, golden_dir):
    from oracle.make_golden import synth_mask_list
    from saber_b200.segmenters import utils as sutils
    g = np.load(os.path.join(golden_dir, "saber_remove_duplicates.npz"))
    masks = synth_mask_list(tuple(g["hw"]), int(g["n"]), seed=int(g["seed"]))
    kept = sutils.remove_duplicate_masks(masks, device="cuda")
    idx = [next(i for i, m in enumerate(masks) if m is k) for k in kept]
    np.testing.assert_array_equal(np.array(idx), g["kept"])


def test_remove_duplicates_vs_oracle_random(ops):
    from oracle import saber_ref
    from oracle.make_golden import synth_mask_list
    from saber_b200.segmenters import utils as sutils
    for seed in (1, 2, 3):
        masks = synth_mask_list((97, 131), 30, seed=seed)
        got = sutils.remove_duplicate_masks(masks, device="cuda")
        want = saber_ref.remove_duplicate_masks(masks)
        assert [id(m) for m in got] == [id(m) for m in want]
    assert sutils.remove_duplicate_masks([], device="cuda") == []


def test_stitch_labels_bitexact(ops):
    from oracle import saber_ref
    from oracle.make_golden import synth_mask_list
    from saber_b200.segmenters import utils as sutils
    masks = synth_mask_list((120, 205), 20, seed=5)
    seg = np.stack([m["segmentation"] for m in masks])
    bits, _, _ = sutils.pack_masks(seg, "cuda")
    order = np.random.default_rng(0).permutation(len(masks)).astype(np.int32)
    out = torch.zeros((120, 205), dtype=torch.int16, device="cuda")
    ops.stitch_labels(bits, torch.from_numpy(order).cuda(), len(order), 205, out=out)
    want = saber_ref.stitch_slice([seg[i] for i in order], (120, 205))
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint16), want)
    out0 = ops.stitch_labels(bits, None, 0, 205)
    assert out0.cpu().numpy().view(np.uint16).max() == 0


@pytest.mark.parametrize("name", ["a", "b"])
def test_separate_masks_matches_reference_golden(ops, golden_dir, name):
    from saber_b200 import synth
    from saber_b200.segmenters import utils as sutils
    g = np.load(os.path.join(golden_dir, f"saber_separate_masks_{name}.npz"))
    vol = synth.make_label_volume(tuple(g["shape"]), seed=int(g["seed"]), n_ellipsoids=int(g["n"]),
                                  speckle=float(g["speckle"])).numpy().view(np.uint16)
    lab = sutils.separate_masks(vol, min_mask_area=int(g["min_mask_area"]), device="cuda")
    assert lab.dtype == np.uint32
    np.testing.assert_array_equal(lab, g["labels"])


@pytest.mark.parametrize("shape,seed,speckle,mma", [((33, 70, 90), 1, 0.01, 1), ((16, 128, 128), 2, 0.004, 3),
                                                    ((5, 17, 1000), 3, 0.02, 0), ((1, 64, 64), 4, 0.05, 0),
                                                    ((64, 64, 1), 5, 0.05, 0)])
def test_separate_masks_vs_oracle_random(ops, shape, seed, speckle, mma):
    from oracle import saber_ref
    from saber_b200 import synth
    from saber_b200.segmenters import utils as sutils
    vol = synth.make_label_volume(shape, seed=seed, n_ellipsoids=25, speckle=speckle).numpy().view(np.uint16)
    np.testing.assert_array_equal(sutils.separate_masks(vol, mma, device="cuda"), saber_ref.separate_masks(vol, mma))
    # dense random noise: worst case for the union-find (long thin components), 26-connectivity percolates
    rng = np.random.default_rng(seed)
    noise = (rng.uniform(size=shape) < 0.12).astype(np.uint8)
    np.testing.assert_array_equal(sutils.separate_masks(noise, 0, device="cuda"), saber_ref.separate_masks(noise, 0))
    empty = np.zeros(shape, np.uint16)
    assert sutils.separate_masks(empty, 1, device="cuda").max() == 0


def test_separate_masks_properties_at_baseline_size(ops):
    """300x928x960 (BASELINE config 3 shape): idempotence (relabelling the labels is the identity), label set is
    exactly 1..K, every ellipsoid centre that survives carries one label, sizes >= min_vol."""
    from saber_b200 import synth
    from saber_b200.segmenters import utils as sutils
    shape = (300, 928, 960)
    vol = synth.make_label_volume(shape, seed=0, n_ellipsoids=40, device="cuda", rmin=10, rmax=80, speckle=0.0005)
    lab = sutils.separate_masks_device(vol, min_mask_area=100)
    K = int(lab.max().item())
    assert K >= 1
    counts = torch.bincount(lab.flatten(), minlength=K + 1)
    assert (counts[1:] >= 1000).all(), "components below min_vol survived"
    lab2 = sutils.separate_masks_device(lab, min_mask_area=100)
    assert torch.equal(lab, lab2), "relabelling must be idempotent"
    assert ((lab > 0) <= (vol != 0)).all()
    first = torch.stack([(lab == k).flatten().nonzero()[0, 0] for k in range(1, min(K, 8) + 1)])
    assert (first[1:] > first[:-1]).all(), "labels must be numbered in raster order of their first voxel"


def test_prepare_rgb_input_matches_reference_semantics():
    """REF saber/adapters/sam2/predictor.py:58-59 passes (H,W,3) images through prepare(to_rgb=False): the reference's
    uniform_filter then also runs along the channel axis. fp32 vs the oracle (scipy on the 3-D array), 1e-4."""
    import torch
    from oracle import saber_ref
    from saber_b200.utils import preprocessing as prep
    rng = np.random.default_rng(12)
    img = (rng.normal(size=(300, 340, 3)) + np.linspace(0, 2, 340)[None, :, None]).astype(np.float32)
    got = prep.prepare(img, to_rgb=False, device="cuda:0")
    want = saber_ref.prepare(img, to_rgb=False)
    assert got.shape == (300, 340, 3) and want.shape == got.shape
    np.testing.assert_allclose(got, want, atol=1e-4, rtol=0)
    t = prep.prepare(torch.from_numpy(img).cuda(), to_rgb=True)  # to_rgb only repeats 2-D inputs (REF :78-80)
    assert t.is_cuda and t.shape == (300, 340, 3)
