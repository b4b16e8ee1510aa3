"""Stage-by-stage comparison of the B200 SAM2 executors against the fp32 oracle (run under gpurun).

The oracle runs in fp32 on the same GPU (TF32 off) purely as the checker.
usage: python tools/model_diag.py [cfg=tiny] [what=enc,dec,m2m]
"""
import os
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

from oracle.sam2_ref.image_predictor import SAM2ImagePredictor  # noqa: E402
from oracle.sam2_ref.sam2_base import SAM2Base  # noqa: E402
from saber_b200 import ops  # noqa: E402
from saber_b200.sam2 import arch  # noqa: E402
from saber_b200.sam2.build_sam import build_sam2  # noqa: E402


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item(), b.abs().max().item()


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    what = (sys.argv[2] if len(sys.argv) > 2 else "enc,dec,m2m").split(",")
    dev = "cuda"
    sd = arch.random_state_dict(cfg, seed=0)
    orc = SAM2Base(cfg, dynamic_multimask_via_stability=True)
    orc.load_state_dict(sd, strict=True)
    orc = orc.to(dev).eval()
    model = build_sam2(cfg, None, device=dev, state_dict=sd, apply_postprocessing=True)
    torch.manual_seed(1)
    B = 2
    img = torch.randn(B, 3, 1024, 1024, device=dev)
    with torch.no_grad():
        bo = orc.forward_image(img)
        _, vf, _, _ = orc._prepare_backbone_features(bo)
    ref_feat = vf[2].permute(1, 0, 2).reshape(B * 4096, 256)
    ref_s1 = vf[1].permute(1, 0, 2).reshape(B * 16384, 64)
    ref_s0 = vf[0].permute(1, 0, 2).reshape(B * 65536, 32)
    if "enc" in what:
        out = model.forward_image(img)
        torch.cuda.synchronize()
        print(f"[{cfg}] encoder feat  rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out["feat"], ref_feat))
        print(f"[{cfg}] encoder s1    rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out["s1"], ref_s1))
        print(f"[{cfg}] encoder s0    rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out["s0"], ref_s0))
        for bsz in (1, 4):
            x = torch.randn(bsz, 3, 1024, 1024, device=dev)
            for _ in range(2):
                model.forward_image(x)
            torch.cuda.synchronize()
            n0 = ops.launch_count
            t0 = time.time()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(3):
                model.forward_image(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"[{cfg}] encoder B={bsz}: {ms:.2f} ms/batch ({ms / bsz:.2f} ms/img), host {1e3 * (time.time() - t0) / 3:.2f} ms,"
                  f" launches/batch {(ops.launch_count - n0) // 3}")
    if "dec" in what or "m2m" in what:
        # decoder on the oracle's own features (isolates decoder error from encoder error)
        pred = SAM2ImagePredictor(orc)
        pred._orig_hw = [(1024, 1024)]
        pred._set_features(img[:1], 1)
        pred._is_image_set = True
        P = 64
        pts = torch.rand(P, 1, 2, device=dev) * 1024
        labels = torch.ones(P, 1, dtype=torch.int32, device=dev)
        with torch.no_grad():
            _, ious_ref, low_ref = pred._predict(pts, labels, multimask_output=True, return_logits=True)
            _, ious2_ref, low2_ref = pred._predict(pts, labels, mask_input=low_ref[:, 0:1], multimask_output=False,
                                                   return_logits=True)
        emb = pred._features["image_embed"][0].permute(1, 2, 0).reshape(4096, 256).contiguous()
        s0 = pred._features["high_res_feats"][0][0].permute(1, 2, 0).reshape(65536, 32).contiguous()
        s1 = pred._features["high_res_feats"][1][0].permute(1, 2, 0).reshape(16384, 64).contiguous()
        dec = model.decoder
        tokens = dec.prompt_tokens(pts.contiguous(), labels.contiguous())
        with torch.no_grad():
            sp, _ = orc.sam_prompt_encoder(points=(pts, labels), boxes=None, masks=None)
        print(f"[{cfg}] prompt tokens rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(tokens[:, 6:], sp))
        if "dec" in what:
            out = dec.forward(emb, s0, s1, tokens, None, multimask_output=True)
            torch.cuda.synchronize()
            print(f"[{cfg}] decoder masks rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out["masks"][:, 1:], low_ref))
            print(f"[{cfg}] decoder ious  rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out["ious"][:, 1:], ious_ref))
            agree = ((out["masks"][:, 1:] > 0) == (low_ref > 0)).float().mean().item()
            print(f"[{cfg}] decoder mask sign agreement = {agree:.6f}")
            for _ in range(2):
                dec.forward(emb, s0, s1, tokens, None, True)
            torch.cuda.synchronize()
            n0 = ops.launch_count
            t0 = time.time()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                dec.forward(emb, s0, s1, tokens, None, True)
            e1.record()
            torch.cuda.synchronize()
            print(f"[{cfg}] decoder B={P}: {e0.elapsed_time(e1) / 5:.2f} ms, host {1e3 * (time.time() - t0) / 5:.2f} ms,"
                  f" launches {(ops.launch_count - n0) // 5}")
        if "m2m" in what:
            mi = low_ref[:, 0].contiguous()
            out2 = dec.forward(emb, s0, s1, tokens, mi, multimask_output=False)
            torch.cuda.synchronize()
            idx = out2["sel_idx"].long()
            sel = out2["masks"][torch.arange(P, device=dev), idx]
            print(f"[{cfg}] m2m sel idx histogram: {torch.bincount(idx, minlength=4).tolist()}")
            print(f"[{cfg}] m2m masks rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(sel, low2_ref[:, 0]))
            print(f"[{cfg}] m2m ious  rel_l2/max_abs/ref_max = %.4g %.4g %.4g" % rel(out2["sel_iou"], ious2_ref[:, 0]))
            for _ in range(2):
                dec.forward(emb, s0, s1, tokens, mi, False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(5):
                dec.forward(emb, s0, s1, tokens, mi, False)
            e1.record()
            torch.cuda.synchronize()
            print(f"[{cfg}] m2m decoder B={P}: {e0.elapsed_time(e1) / 5:.2f} ms")


if __name__ == "__main__":
    main()
