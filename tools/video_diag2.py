"""End-to-end state comparison of the B200 video predictor vs the oracle on a short synthetic video (gpurun)."""
import os, sys
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import saber_ref
from oracle.sam2_ref.video_predictor import build_sam2_video_predictor as oracle_build, empty_inference_state
from saber_b200 import synth
from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG
from saber_b200.adapters.sam2 import SAM2Adapter
from saber_b200.sam2 import arch
from saber_b200.sam2.sam2_video_predictor import build_sam2_video_predictor


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def ellipse(hw, cy, cx, ry, rx):
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    return (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).astype(np.float32)


def main():
    Z, H, W = 4, 96, 120
    vol = synth.make_tomogram((Z, H, W), seed=7, n_ellipsoids=3).numpy()
    sd = arch.random_state_dict("tiny", seed=0)
    for k in list(sd):
        if "output_hypernetworks_mlps" in k and ".layers.2." in k:
            sd[k] = sd[k] * 30.0
    sd["sam_mask_decoder.iou_prediction_head.layers.2.bias"] = torch.tensor([0.0, -1.0, 1.0, 0.0])
    sd["sam_mask_decoder.pred_obj_score_head.layers.2.bias"] = torch.tensor([1.5])
    ad = SAM2Adapter(SAM2AdapterConfig(cfg="tiny", amg_cfg=cfgAMG(sam2_cfg="tiny"), num_maskmem=2, seed=0), device="cuda:0")
    vp = build_sam2_video_predictor("tiny", None, device="cuda:0", state_dict=sd)
    vp.maskmem_tpos_enc = torch.nn.Parameter(vp.maskmem_tpos_enc[:2], requires_grad=False); vp.num_maskmem = 2
    ad.predictor = vp
    orc = oracle_build("tiny", None, device="cpu", state_dict=sd)
    orc.maskmem_tpos_enc = torch.nn.Parameter(orc.maskmem_tpos_enc[:2]); orc.num_maskmem = 2
    images, vh, vw = saber_ref.load_grayscale_image_array(saber_ref.normalize_tomogram(vol), 1024)
    ost = empty_inference_state(torch.from_numpy(images), vh, vw, "cpu")
    ad.set_volume(vol)
    st = ad.inference_state
    start = 1
    mask = ellipse((H, W), 40, 50, 18, 25)
    vp.add_new_mask(st, start, 1, mask)
    orc.add_new_mask(ost, start, 1, mask)
    from saber_b200 import ops as _ops
    from oracle.sam2_ref.video_predictor import fill_holes_in_mask_scores
    rec = []
    orig_fill = _ops.fill_holes
    _ops.fill_holes = lambda m, a: (rec.append((m.clone(), orig_fill(m, a))) or rec[-1][1])
    a = list(vp.propagate_in_video(st, start_frame_idx=start, reverse=False))
    _ops.fill_holes = orig_fill
    for i, (mi, mo) in enumerate(rec):
        ref = fill_holes_in_mask_scores(mi.cpu()[:, None], 8)[:, 0]
        print("fill", i, "in==out frac", (mi == mo).float().mean().item(), "ours vs oracle fill mismatches", (mo.cpu() != ref).sum().item(),
              "filled ours", (mo == 0.1).sum().item(), "filled oracle", (ref == 0.1).sum().item(), "bg frac", (mi <= 0).float().mean().item())
    b = list(orc.propagate_in_video(ost, start_frame_idx=start, reverse=False))
    go, oo = st["output_dict_per_obj"][0], ost["output_dict_per_obj"][0]
    for i, (mi, mo) in enumerate(rec[1:2]):
        o_pm = oo["non_cond_frame_outputs"][start + 1]["pred_masks"][0, 0]
        print("pre-fill low vs oracle pred_masks (post-fill)", rel(mi[0].clamp(-64, 64), o_pm.clamp(-64, 64)),
              "ours post-fill vs oracle", rel(mo[0], o_pm), "oracle filled px", (o_pm == 0.1).sum().item())
    # raw features
    _, _, vf, vpos, _ = orc._get_image_feature(ost, start, 1)
    print("feat frame start", rel(st["cached_features"][start]["feat"], vf[2][:, 0]))
    for key, f in (("cond_frame_outputs", start), ("non_cond_frame_outputs", start + 1), ("non_cond_frame_outputs", start + 2)):
        g, o = go[key][f], oo[key][f]
        mm_o = o["maskmem_features"].float().flatten(2).permute(0, 2, 1).reshape(-1, 64)
        print(key, f, "maskmem", rel(g["maskmem_features"], mm_o), "pred_masks", rel(g["pred_masks"].clamp(-64, 64), o["pred_masks"].clamp(-64, 64)),
              "obj_ptr", rel(g["obj_ptr"], o["obj_ptr"]), "score", g["object_score_logits"].item(), o["object_score_logits"].item())
    for (fa, _, la), (fb, _, lb) in zip(a, b):
        print("frame", fa, "video_res", rel(la.clamp(-64, 64), lb.clamp(-64, 64)), "agree", ((la.cpu() > 0) == (lb > 0)).float().mean().item())
    # replay frame start+1 by hand with the ORACLE's stored memory to separate error sources
    f = start + 1
    oc = oo["cond_frame_outputs"][start]
    mem = oc["maskmem_features"].float().flatten(2).permute(0, 2, 1).reshape(-1, 64)
    ptr = oc["obj_ptr"].reshape(4, 64)
    memory = torch.cat([mem, ptr]).to(torch.bfloat16).cuda()
    sig = ((0,), (1,))
    parts = [vp._spatial_key_pos(0), vp._ptr_key_pos((1,), Z)]
    pos_k = [torch.cat([p_[l] for p_ in parts], 0).contiguous() for l in range(4)]
    c = vp._frame(st, f)
    pix = vp.mem_attn.forward(c["feat"], memory, pos_k, 4, 1)
    with torch.no_grad():
        _, _, vf, vpos, sizes = orc._get_image_feature(ost, f, 1)
        want = orc._prepare_memory_conditioned_features(frame_idx=f, is_init_cond_frame=False, current_vision_feats=vf[-1:],
                                                        current_vision_pos_embeds=vpos[-1:], feat_sizes=sizes[-1:],
                                                        output_dict={"cond_frame_outputs": {start: oc}, "non_cond_frame_outputs": {}},
                                                        num_frames=Z)
    print("pix_feat_with_mem (oracle memory)", rel(pix, want[0].flatten(1).t()))
    print("feat frame f", rel(c["feat"], vf[2][:, 0]))
    print("s0", rel(c["s0"], vf[0][:, 0]), "s1", rel(c["s1"], vf[1][:, 0]))
    hr = [x.permute(1, 2, 0).view(1, -1, *s_) for x, s_ in zip(vf[:-1], sizes[:-1])]
    with torch.no_grad():
        o = orc._forward_sam_heads(backbone_features=want, high_res_features=hr, multimask_output=True)
    dec = vp.decoder
    tokens = dec.prompt_tokens(torch.zeros(1, 1, 2, device="cuda"), torch.full((1, 1), -1, dtype=torch.int32, device="cuda"))
    for name, emb, s0_, s1_ in (("oracle inputs", want[0].flatten(1).t().contiguous().cuda(), vf[0][:, 0].contiguous().cuda(), vf[1][:, 0].contiguous().cuda()),
                                ("our inputs", pix, c["s0"], c["s1"])):
        out = dec.forward(emb, s0_, s1_, tokens, None, multimask_output=True)
        print(name, "multimasks", rel(out["masks"][:, 1:4], o[0]), "ious", rel(out["ious"][:, 1:4], o[2]), out["ious"].tolist(), o[2].tolist(),
              "per-token", [rel(out["masks"][:, 1 + i], o[0][:, i]) for i in range(3)], "tokens", rel(out["hs"][:, 2:6], o[0].new_zeros(1)) if False else "")


if __name__ == "__main__":
    main()


def more():
    pass
