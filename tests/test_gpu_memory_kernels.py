"""-m gpu: z-axis propagation kernels (csrc/memory.cu, attention head_dim 256) against the oracle restatement of the
upstream modules (oracle/sam2_ref/memory.py, video_predictor.py) and torch fp32, through the C ABI.
Float kernels: bf16 storage -> 2e-2 relative (north_star); integer / index kernels: bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    from saber_b200 import ops as _ops
    _ops.require_b200()
    return _ops


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm().clamp_min(1e-12)).item()


def test_rope_matches_oracle(ops):
    from oracle.sam2_ref.memory import apply_rotary_enc, compute_axial_cis
    torch.manual_seed(0)
    cis = compute_axial_cis(256, 64, 64)
    cs = torch.view_as_real(cis).contiguous().cuda()  # [4096,128,2]
    q = torch.randn(2, 1, 4096, 256)
    n_ptr = 12
    k = torch.randn(2, 1, 2 * 4096 + n_ptr, 256)
    k_ref = k.clone()
    q_ref, k_ref[:, :, :2 * 4096] = apply_rotary_enc(q, k[:, :, :2 * 4096], cis, repeat_freqs_k=True)
    q_out = ops.rope_apply(q.reshape(-1, 256).cuda(), cs, 4096)
    k_out = ops.rope_apply(k.reshape(-1, 256).cuda(), cs, 2 * 4096 + n_ptr, n_rope=2 * 4096)
    assert rel(q_out, q_ref.reshape(-1, 256)) < 5e-3
    assert rel(k_out, k_ref.reshape(-1, 256)) < 5e-3
    # untouched object-pointer rows are plain bf16 casts
    tail = k_out.view(2, -1, 256)[:, 2 * 4096:]
    assert torch.equal(tail.cpu(), k[:, 0, 2 * 4096:].to(BF16))
    # bf16 input path (a strided view of a fused projection)
    fused = torch.randn(4096, 768, device="cuda").to(BF16)
    o2 = ops.rope_apply(fused[:, 256:512], cs, 4096)
    r2, _ = apply_rotary_enc(fused[:, 256:512].float().cpu()[None, None], torch.zeros(1, 1, 0, 256), cis)
    assert rel(o2, r2[0, 0]) < 5e-3


@pytest.mark.parametrize("hd,nq,nk,B", [(256, 4096, 4096, 1), (256, 4096, 8192 + 28, 2), (256, 100, 77, 1)])
def test_attention_hd256(ops, hd, nq, nk, B):
    torch.manual_seed(1)
    q = (torch.randn(B * nq, hd, device="cuda") * 0.3).to(BF16)
    k = (torch.randn(B * nk, hd, device="cuda") * 0.3).to(BF16)
    v = torch.randn(B * nk, hd, device="cuda").to(BF16)
    out = ops.attention(q, k, v, B, 1, nq, nk)
    ref = F.scaled_dot_product_attention(q.float().view(B, 1, nq, hd), k.float().view(B, 1, nk, hd),
                                         v.float().view(B, 1, nk, hd)).reshape(B * nq, hd)
    assert rel(out, ref) < 1e-2


def test_mask_downsampler_stages(ops):
    """conv3x3s2 + LN2d + GELU stages and the im2col of the GEMM stage vs the oracle MaskDownSampler."""
    from oracle.sam2_ref.memory import MaskDownSampler
    torch.manual_seed(2)
    ds = MaskDownSampler(kernel_size=3, stride=2, padding=1).eval()
    for p in ds.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.5)
    x = torch.randn(2, 1, 256, 256) * 4
    with torch.no_grad():
        h = torch.sigmoid(x) * 20 - 10
        refs = []
        for i in range(3):
            h = ds.encoder[3 * i + 2](ds.encoder[3 * i + 1](ds.encoder[3 * i](h)))
            refs.append(h)
        h4 = ds.encoder[9](h)  # conv 64 -> 256 (before LN)
    cur = x.permute(0, 2, 3, 1).contiguous().cuda()
    for i in range(3):
        conv, ln = ds.encoder[3 * i], ds.encoder[3 * i + 1]
        cur = ops.conv3x3s2_ln_gelu(cur, i, conv.weight.detach().cuda().contiguous(), conv.bias.detach().cuda(),
                                    ln.weight.detach().cuda(), ln.bias.detach().cuda(), ln.eps, in_xf=1 if i == 0 else 0)
        assert rel(cur.permute(0, 3, 1, 2), refs[i]) < 1e-2, i
    cols = ops.im2col_3x3s2(cur)
    w = ds.encoder[9].weight.detach().permute(0, 2, 3, 1).reshape(256, 9 * 64).cuda().to(BF16).contiguous()
    out = ops.gemm(cols, w, ds.encoder[9].bias.detach().cuda(), out_dtype=F32)
    assert rel(out.view(2, 16, 16, 256).permute(0, 3, 1, 2), h4) < 2e-2
    # binarised input transform
    xb = ops.conv3x3s2_ln_gelu(x.permute(0, 2, 3, 1).contiguous().cuda(), 0, ds.encoder[0].weight.detach().cuda().contiguous(),
                               ds.encoder[0].bias.detach().cuda(), ds.encoder[1].weight.detach().cuda(),
                               ds.encoder[1].bias.detach().cuda(), 1e-6, in_xf=2)
    with torch.no_grad():
        rb = ds.encoder[2](ds.encoder[1](ds.encoder[0]((x > 0).float() * 20 - 10)))
    assert rel(xb.permute(0, 3, 1, 2), rb) < 1e-2


def test_dwconv7_ln_matches_cxblock_front(ops):
    from oracle.sam2_ref.memory import CXBlock
    torch.manual_seed(3)
    blk = CXBlock(dim=256).eval()
    blk.norm.weight.data.normal_(1, 0.2)
    blk.norm.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 256, 64, 64)
    with torch.no_grad():
        ref = blk.norm(blk.dwconv(x))  # LayerNorm2d over channels, NCHW
    xin = x.permute(0, 2, 3, 1).reshape(-1, 256).contiguous().cuda()
    out = ops.dwconv7_ln(xin, 2, 64, 64, blk.dwconv.weight.detach().reshape(256, 49).cuda().contiguous(),
                         blk.dwconv.bias.detach().cuda(), blk.norm.weight.detach().cuda(), blk.norm.bias.detach().cuda(),
                         blk.norm.eps)
    assert rel(out.view(2, 64, 64, 256).permute(0, 3, 1, 2), ref) < 1e-2


def test_fill_holes_bit_exact(ops):
    from oracle.sam2_ref.video_predictor import fill_holes_in_mask_scores
    rng = np.random.default_rng(4)
    m = rng.normal(0.5, 1.0, size=(6, 1, 256, 256)).astype(np.float32)
    m[0, 0] = np.abs(m[0, 0]) + 0.1  # no background at all
    m[1, 0] = -np.abs(m[1, 0]) - 0.1  # all background (one big component, not a hole)
    # engineered holes: isolated pixels, an 8-pixel diagonal chain (8-connectivity), a 9-pixel blob (too big)
    m[2, 0] = 3.0
    m[2, 0, 10, 10] = -1
    for t in range(8):
        m[2, 0, 50 + t, 60 + t] = -2
    m[2, 0, 100:103, 100:103] = -1
    m[2, 0, 0, 0] = 0.0  # score == 0 counts as background
    t = torch.from_numpy(m)
    ref = fill_holes_in_mask_scores(t, 8)
    out = ops.fill_holes(t[:, 0].contiguous().cuda(), 8)
    np.testing.assert_array_equal(out.cpu().numpy(), ref[:, 0].numpy())
    assert (ref[2, 0] == 0.1).sum().item() == 10 and ref[2, 0, 101, 101] == -1


@pytest.mark.parametrize("HW", [(1024, 1024), (928, 960), (511, 77)])
def test_stitch_objects_bit_exact(ops, HW):
    from oracle import saber_ref
    H, W = HW
    rng = np.random.default_rng(5)
    N, Sv = 3, 1024
    yy, xx = np.mgrid[0:Sv, 0:Sv]
    logits = np.stack([8 - np.hypot(yy - rng.uniform(200, 800), xx - rng.uniform(200, 800)) / rng.uniform(20, 60)
                       for _ in range(N)]).astype(np.float32)
    ids = np.array([1, 2, 5], dtype=np.int32)
    want = np.full((H, W), 7, dtype=np.uint16)
    for i in range(N):
        m = logits[i] > 0
        if m.shape != (H, W):
            m = saber_ref.skimage_resize(m, (H, W), order=0, anti_aliasing=False)
        want = np.where(m, ids[i], want).astype(np.uint16)
    labels = torch.full((H, W), 7, dtype=torch.int16, device="cuda")
    ops.stitch_objects_(torch.from_numpy(logits).cuda(), torch.from_numpy(ids).cuda(), labels)
    np.testing.assert_array_equal(labels.cpu().numpy().view(np.uint16), want)


def test_track_select_and_misc(ops):
    torch.manual_seed(6)
    B, Nt, S = 5, 8, 64
    masks = torch.randn(B, 4, S, S, device="cuda")
    ious = torch.rand(B, 4, device="cuda")
    ious[1, 2] = ious[1, 3] = 0.9  # tie -> first maximum
    ious[1, 1] = 0.1
    obj = torch.tensor([1.0, -2.0, 0.0, 3.0, 0.5], device="cuda")
    hs = torch.randn(B, Nt, 256, device="cuda")
    low, tok, best = ops.track_select(masks, ious, obj, hs, None, True)
    bi = torch.argmax(ious[:, 1:], dim=-1)
    np.testing.assert_array_equal(best.cpu().numpy(), (1 + bi).cpu().numpy())
    want = torch.where((obj > 0)[:, None, None], masks[torch.arange(B), 1 + bi], torch.full_like(masks[:, 0], -1024.0))
    assert torch.equal(low, want)
    assert torch.equal(tok, hs[torch.arange(B), 3 + bi])
    sel = torch.tensor([0, 1, 2, 3, 0], dtype=torch.int32, device="cuda")
    low2, tok2, _ = ops.track_select(masks, ious, obj, hs, sel, False)
    want2 = torch.where((obj > 0)[:, None, None], masks[torch.arange(B), sel.long()], torch.full_like(masks[:, 0], -1024.0))
    assert torch.equal(low2, want2) and torch.equal(tok2, hs[:, 2])
    ptr = torch.randn(B, 256, device="cuda")
    nop = torch.randn(256, device="cuda")
    want_ptr = torch.where((obj > 0)[:, None], ptr, nop[None])
    ops.objptr_mix_(ptr, obj, nop)
    assert torch.equal(ptr, want_ptr)
    x = torch.randn(2 * 4096, 64, device="cuda")
    sc = torch.tensor([1.0, -1.0], device="cuda")
    vec = torch.randn(64, device="cuda")
    out = ops.add_vec_cond(x, sc, vec, 2)
    want = torch.cat([x[:4096], x[4096:] + vec]).to(BF16)
    assert torch.equal(out, want)
    m = torch.rand(3, 128, 128, device="cuda")
    assert torch.equal(ops.threshold_affine(m, 0.5, 20.0, -10.0), (m >= 0.5).float() * 20 - 10)
    conv = torch.nn.Conv2d(1, 1, 4, 4).cuda()
    big = torch.randn(2, 256, 256, device="cuda")
    with torch.no_grad():
        ref = conv(big[:, None])[:, 0]
    got = ops.conv4x4s4(big, conv.weight.detach().reshape(16).contiguous(), conv.bias.detach())
    assert (got - ref).abs().max().item() < 1e-5
    vol = torch.zeros(4, 32, 32, dtype=torch.int16, device="cuda")
    vol[2, 31, 31] = 9
    vol[3, 0, 0] = 4
    np.testing.assert_array_equal(ops.slice_any(vol).cpu().numpy(), [0, 0, 1, 1])
    ops.erase_label_(vol[3], 4)
    assert vol[3].abs().sum().item() == 0 and vol[2, 31, 31].item() == 9
