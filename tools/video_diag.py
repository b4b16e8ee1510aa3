"""Stage-by-stage diagnostics of the z-propagation modules against the oracle (run under gpurun)."""
import os, sys
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.sam2_ref.video_predictor import build_sam2_video_predictor as oracle_build
from saber_b200 import ops
from saber_b200.sam2 import arch
from saber_b200.sam2.sam2_video_predictor import build_sam2_video_predictor
from saber_b200.sam2.memory import sine_pe_2d


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def main():
    torch.manual_seed(0)
    cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    sd = arch.random_state_dict(cfg, seed=0)
    orc = oracle_build(cfg, None, device="cpu", state_dict=sd)
    orc.maskmem_tpos_enc = torch.nn.Parameter(orc.maskmem_tpos_enc[:2]); orc.num_maskmem = 2
    p = build_sam2_video_predictor(cfg, None, device="cuda:0", state_dict=sd)
    p.maskmem_tpos_enc = torch.nn.Parameter(p.maskmem_tpos_enc[:2], requires_grad=False); p.num_maskmem = 2
    g = torch.Generator().manual_seed(1)
    # ---- constants
    x = torch.zeros(1, 256, 64, 64)
    pos256 = orc.image_encoder.neck.position_encoding(x)[0].flatten(1).t()  # [4096,256]
    print("sine256 const", rel(sine_pe_2d(256, 64), pos256))
    pos64 = orc.memory_encoder.position_encoding(torch.zeros(1, 64, 64, 64))[0].flatten(1).t()
    print("sine64 const", rel(p.mem_enc.pos, pos64))
    # ---- memory attention
    curr = torch.randn(4096, 256, generator=g) * 0.5
    n_ptr = 3
    Nk = 2 * 4096 + 4 * n_ptr
    memory = (torch.randn(Nk, 64, generator=g) * 0.5).to(torch.bfloat16).float()
    mpos = torch.cat([pos64 + orc.maskmem_tpos_enc[1].reshape(1, 64), pos64 + orc.maskmem_tpos_enc[0].reshape(1, 64),
                      torch.randn(4 * n_ptr, 64, generator=g) * 0.1]).detach()
    with torch.no_grad():
        want = orc.memory_attention(curr=[curr[:, None]], curr_pos=[pos256[:, None]], memory=memory[:, None],
                                    memory_pos=mpos[:, None], num_obj_ptr_tokens=4 * n_ptr)[:, 0]
        # layer-wise references
        refs = []
        out = (curr + 0.1 * pos256)[None]
        for layer in orc.memory_attention.layers:
            out = layer(tgt=out, memory=memory[None], pos=mpos[None], query_pos=pos256[None], num_k_exclude_rope=4 * n_ptr)
            refs.append(out[0])
    pos_k = p.mem_attn.key_pos_term(mpos)
    got = p.mem_attn.forward(curr.cuda(), memory.cuda().to(torch.bfloat16), pos_k, 4 * n_ptr, 1)
    print("memory attention out", rel(got, want))
    # B=2 batched (shared layer-0) vs B=1
    mem2 = torch.cat([memory, memory.flip(0)]).cuda().to(torch.bfloat16)
    got2 = p.mem_attn.forward(curr.cuda(), mem2, pos_k, 4 * n_ptr, 2)
    print("memory attention B=2 first vs B=1", rel(got2[:4096], got))
    # ---- memory encoder
    pix = torch.randn(4096, 256, generator=g) * 0.5
    low = torch.randn(2, 256, 256, generator=g) * 5
    score = torch.tensor([1.0, -1.0])
    with torch.no_grad():
        hi = torch.nn.functional.interpolate(low[:, None], size=(1024, 1024), mode="bilinear", align_corners=False)
        feats = [None, None, pix[:, None].expand(-1, 2, -1)]
        for binar in (False, True):
            mm, _ = orc._encode_new_memory(feats, [None, None, (64, 64)], hi, score[:, None], is_mask_from_pts=binar)
            wantm = mm.to(torch.bfloat16).flatten(2).permute(0, 2, 1).reshape(-1, 64)
            gotm = p.mem_enc.forward(p.mem_enc.project_pix(pix.cuda()), low.cuda(), score.cuda(), binarize=binar)
            print(f"memory encoder binarize={binar}", rel(gotm, wantm))
    # ---- SAM heads on conditioned features (per-prompt embeddings)
    pixm = torch.randn(2, 256, 64, 64, generator=g) * 0.5
    s0 = torch.randn(1, 32, 256, 256, generator=g) * 0.3
    s1 = torch.randn(1, 64, 128, 128, generator=g) * 0.3
    with torch.no_grad():
        o = orc._forward_sam_heads(backbone_features=pixm, high_res_features=[s0.expand(2, -1, -1, -1), s1.expand(2, -1, -1, -1)],
                                   multimask_output=True)
    dec = p.decoder
    tokens = dec.prompt_tokens(torch.zeros(2, 1, 2, device="cuda"), torch.full((2, 1), -1, dtype=torch.int32, device="cuda"))
    emb = pixm.flatten(2).permute(0, 2, 1).reshape(-1, 256).contiguous().cuda()
    out = dec.forward(emb, s0[0].flatten(1).t().contiguous().cuda(), s1[0].flatten(1).t().contiguous().cuda(), tokens, None,
                      multimask_output=True)
    print("decoder multimasks", rel(out["masks"][:, 1:4], o[0]), "ious", rel(out["ious"][:, 1:4], o[2]),
          "obj", out["obj"].reshape(-1).tolist(), o[6].reshape(-1).tolist())
    low_g, tok, best = ops.track_select(out["masks"], out["ious"], out["obj"].reshape(-1).contiguous(), out["hs"].contiguous(), None, True)
    print("best", best.tolist(), torch.argmax(o[2], -1).tolist(), "low_res", rel(low_g, o[3][:, 0]))
    ptr = p._obj_ptr(tok, out["obj"].reshape(-1).contiguous())
    print("obj_ptr", rel(ptr, o[5]))


if __name__ == "__main__":
    main()
