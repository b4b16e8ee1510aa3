"""-m gpu: z-axis memory propagation against the oracle.

* float stages: per-frame video-resolution logits / object scores of the B200 video predictor vs the fp32 oracle
  restatement of upstream SAM2VideoPredictor with identical seeded weights and inputs (hiera-tiny, 6 frames);
  tolerance from BASELINE north_star: 2e-2 relative (bf16), stated per assertion;
* integer stages: SAM2Adapter.segment_volume's uint16 label volume must equal, bit for bit, the oracle restatement of
  REF saber/adapters/sam2/predictor.py:232-348 replayed on the logits / scores the GPU produced ("given identical
  logits"): threshold, skimage order-0 resize, stitch precedence, backward-fills-only-empty-slices, hook bookkeeping
  (incl. the reference's off-by-one), presence-score fit and filtering.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def ellipse(hw, cy, cx, ry, rx):
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    return (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).astype(np.float32)


@pytest.fixture(scope="module")
def setup():
    from oracle import saber_ref
    from oracle.sam2_ref.video_predictor import build_sam2_video_predictor as oracle_build, empty_inference_state
    from saber_b200 import ops, synth
    from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG
    from saber_b200.adapters.sam2 import SAM2Adapter
    from saber_b200.sam2 import arch
    ops.require_b200()
    Z, H, W = 6, 96, 120
    vol = synth.make_tomogram((Z, H, W), seed=7, n_ellipsoids=3).numpy()
    cfg = SAM2AdapterConfig(cfg="tiny", amg_cfg=cfgAMG(sam2_cfg="tiny"), num_maskmem=2, seed=0)
    ad = SAM2Adapter(cfg, device="cuda:0")
    sd = arch.random_state_dict("tiny", seed=0)
    # Random-init heads give three near-identical IoU predictions (argmax decided by bf16 noise) and mask logits of
    # ~1e-1: give the mask logits a realistic dynamic range and separate the IoU / object-score outputs so that the
    # discrete choices (best multimask token, object present) are well-conditioned on both sides.
    for k in list(sd):
        if "output_hypernetworks_mlps" in k and ".layers.2." in k:
            sd[k] = sd[k] * 30.0
    sd["sam_mask_decoder.iou_prediction_head.layers.2.bias"] = torch.tensor([0.0, -1.0, 1.0, 0.0])
    sd["sam_mask_decoder.pred_obj_score_head.layers.2.bias"] = torch.tensor([1.5])
    from saber_b200.sam2.sam2_video_predictor import build_sam2_video_predictor
    vp = build_sam2_video_predictor("tiny", None, device="cuda:0", state_dict=sd)
    vp.maskmem_tpos_enc = torch.nn.Parameter(vp.maskmem_tpos_enc[:2], requires_grad=False)  # as SAM2Adapter._video()
    vp.num_maskmem = 2
    ad.predictor = vp
    # ---- oracle: reference preprocessing + upstream state machine, fp32 on the host
    orc = oracle_build("tiny", None, device="cpu", state_dict=sd)
    orc.maskmem_tpos_enc = torch.nn.Parameter(orc.maskmem_tpos_enc[:2])  # REF saber/adapters/sam2/predictor.py:31-34
    orc.num_maskmem = 2
    images, vh, vw = saber_ref.load_grayscale_image_array(saber_ref.normalize_tomogram(vol), 1024)
    ost = empty_inference_state(torch.from_numpy(images), vh, vw, "cpu")
    seeds = [ellipse((H, W), 40, 50, 18, 25), ellipse((H, W), 70, 90, 12, 14), np.zeros((H, W), np.float32)]
    return dict(vol=vol, ad=ad, orc=orc, ost=ost, images=images, seeds=seeds, shape=(Z, H, W), start=2)


def test_volume_preprocessing_matches_reference_restatement(setup):
    """R1 + R2: normalize_tomogram, skimage-style resize of every slice, 2x - 1 (fp32 tolerance 1e-5)."""
    ad = setup["ad"]
    ad.set_volume(setup["vol"])
    img = ad.inference_state["images"]
    assert tuple(img.shape) == (6, 3, 1024, 1024) and ad.inference_state["video_height"] == 1024
    got = img[:, 0].cpu().numpy()
    np.testing.assert_allclose(got, setup["images"][:, 0], atol=1e-5, rtol=0)
    assert torch.equal(img[:, 0], img[:, 2])


def test_propagation_logits_match_oracle(setup):
    """Forward + backward propagation of 2 seeded objects: logits / scores vs the fp32 oracle."""
    ad, orc, ost, start = setup["ad"], setup["orc"], setup["ost"], setup["start"]
    if ad.inference_state is None:
        ad.set_volume(setup["vol"])
    p, st = ad._video(), ad.inference_state
    p.reset_state(st)
    # Hole filling (area <= 8 background components -> +0.1) is discontinuous in the logits: on the noise-like masks of
    # random-init weights a 1e-3 perturbation re-wires which specks count as holes. It is checked bit-exactly on
    # identical inputs (test_gpu_memory_kernels.py::test_fill_holes_bit_exact and the replay test below) and switched
    # off on both sides here so that this test measures the float path.
    p.fill_hole_area, orc.fill_hole_area = 0, 0
    got_scores, want_scores = [], []
    h1 = p.sam_mask_decoder.register_forward_hook(lambda m, i, o: got_scores.append(float(o[3].reshape(-1)[0])))
    h2 = orc.sam_mask_decoder.register_forward_hook(lambda m, i, o: want_scores.append(float(o[3].reshape(-1)[0])))
    for obj_id, mask in enumerate(setup["seeds"][:2], start=1):
        f, ids, vr = p.add_new_mask(st, start, obj_id, mask)
        f2, ids2, vr2 = orc.add_new_mask(ost, start, obj_id, mask)
        assert ids == ids2 and tuple(vr.shape) == tuple(vr2.shape)
        assert rel(vr, vr2) < 1e-4  # the seed frame's scores are +-10 blends: fp32-exact up to resize rounding
    worst = 0.0
    for reverse in (False, True):
        a = list(p.propagate_in_video(st, start_frame_idx=start, reverse=reverse))
        b = list(orc.propagate_in_video(ost, start_frame_idx=start, reverse=reverse))
        assert [x[0] for x in a] == [x[0] for x in b]
        for (fa, ia, la), (fb, ib, lb) in zip(a, b):
            assert ia == ib and tuple(la.shape) == tuple(lb.shape) == (2, 1, 1024, 1024)
            lbc = lb.clamp(-64, 64)
            r = rel(la.clamp(-64, 64), lbc)
            agree = ((la.cpu() > 0) == (lb > 0)).float().mean().item()
            worst = max(worst, r)
            print(f"frame {fa} reverse={reverse}: rel_l2={r:.4f} sign agreement={agree:.5f}")
            assert r < 2e-2, (fa, reverse, r)  # north_star bf16 tolerance on logits
            assert agree > 0.995
    h1.remove()
    h2.remove()
    p.fill_hole_area, orc.fill_hole_area = 8, 8
    assert len(got_scores) == len(want_scores) > 0
    np.testing.assert_allclose(got_scores, want_scores, rtol=2e-2, atol=5e-2)
    print("worst rel", worst)


class _Replay:
    """Plays back recorded (hook, yield) events behind the video-predictor interface the reference adapter drives."""

    class _Dec:
        def __init__(self):
            self.hooks = []

        def register_forward_hook(self, fn):
            self.hooks.append(fn)
            outer = self

            class H:
                def remove(self_inner):
                    outer.hooks.remove(fn)
            return H()

    def __init__(self, passes):
        self.passes = passes  # {reverse: [events]}
        self.sam_mask_decoder = self._Dec()
        self.added = []

    def add_new_mask(self, inference_state, frame_idx, obj_id, mask):
        self.added.append(obj_id)

    def propagate_in_video(self, state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        for ev in self.passes[reverse]:
            if ev[0] == "hook":
                for fn in list(self.sam_mask_decoder.hooks):
                    fn(None, (), (None, None, None, torch.tensor([[ev[1]]])))
            else:
                yield ev[1], ev[2], ev[3]


@pytest.mark.parametrize("out_hw", [(96, 120), (1024, 1024)])
def test_segment_volume_bit_exact_given_identical_logits(setup, out_hw):
    from oracle import saber_ref
    ad, start = setup["ad"], setup["start"]
    Z = setup["shape"][0]
    if ad.inference_state is None:
        ad.set_volume(setup["vol"])
    p = ad._video()
    p.reset_state(ad.inference_state)
    H, W = out_hw
    seeds = setup["seeds"] if out_hw == (96, 120) else [ellipse((1024, 1024), 400, 500, 180, 250),
                                                         ellipse((1024, 1024), 700, 800, 120, 140)]
    # record what the GPU path produces while the adapter runs the reference's control flow
    events = {False: [], True: []}
    current = {"rev": False}
    hook = p.sam_mask_decoder.register_forward_hook(
        lambda m, i, o: events[current["rev"]].append(("hook", float(o[3].reshape(-1)[0]))))
    orig = p.propagate_in_video

    def recording(state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        current["rev"] = reverse
        for f, ids, logits in orig(state, start_frame_idx=start_frame_idx, max_frame_num_to_track=max_frame_num_to_track,
                                   reverse=reverse):
            events[reverse].append(("yield", f, list(ids), logits.detach().cpu().clone()))
            yield f, ids, logits

    p.propagate_in_video = recording
    try:
        got = ad.segment_volume(start, masks=seeds, vol_shape=(Z, H, W), min_presence_score=0.5)
    finally:
        del p.propagate_in_video
        hook.remove()
    assert got.dtype == np.uint16 and got.shape == (Z, H, W)
    replay = _Replay(events)
    want, frame_scores, metrics = saber_ref.segment_volume(replay, None, start, seeds, (Z, H, W), min_presence_score=0.5)
    n_obj = len([s for s in seeds if s.max() > 0])
    assert replay.added == list(range(1, n_obj + 1))  # the all-zero seed is skipped (REF :262-263)
    np.testing.assert_array_equal(ad.frame_scores, frame_scores)
    np.testing.assert_array_equal(got, want)
    assert ad.frame_metrics.keys() == metrics.keys()
    for f in metrics:
        assert ad.frame_metrics[f] == metrics[f]
    # without the presence filter the propagated objects must be visible in the seed slice
    ad.reset_state()
    raw = ad.segment_volume(start, masks=seeds, vol_shape=(Z, H, W), min_presence_score=-1e9)
    ad.reset_state()
    assert set(np.unique(raw[start])) >= set(range(1, n_obj + 1))
