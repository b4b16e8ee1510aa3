"""Thin Python wrappers over the C-ABI kernels: torch owns device memory and the stream, every
arithmetic op on the hot path is one of our sm_100a kernels. No CPU / eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import lib as _lib

ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
_BF16, _F32 = torch.bfloat16, torch.float32

launch_count = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("saber_b200 ops need CUDA tensors: there is no CPU fallback")


def _count(n=1):
    global launch_count
    launch_count += n


def require_b200() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("saber_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    L = _lib.load()
    _lib.check(L.sb_require_sm100(), "sb_require_sm100")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, res_mod: int = 0, out_dtype=_BF16,
         out: Optional[torch.Tensor] = None, alpha: float = 1.0, force_bn: int = 0) -> torch.Tensor:
    """out[M,N] = act(alpha * a[M,K] @ w[N,K]^T + bias[N]) + residual[M (mod res_mod), N].

    a, w: bf16 row-major (last dim contiguous, pitch multiple of 8). bias: fp32. residual: bf16 or
    fp32 2-D. out: bf16 or fp32.
    """
    _chk_cuda(a, w, bias, residual, out)
    assert a.dtype == _BF16 and w.dtype == _BF16, (a.dtype, w.dtype)
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2, (a.shape, w.shape)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape == (M, N)
    flags = (1 if out.dtype == _F32 else 0)
    ldr = 0
    if residual is not None:
        assert residual.dim() == 2 and residual.stride(1) == 1 and residual.shape[1] == N
        assert residual.dtype in (_BF16, _F32)
        flags |= 2 if residual.dtype == _F32 else 0
        ldr = residual.stride(0)
    if bias is not None:
        assert bias.dtype == _F32 and bias.numel() == N and bias.is_contiguous()
    L = _lib.load()
    rc = L.sb_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                        out.stride(0), M, N, K, _ptr(bias), act, _ptr(residual), ldr, res_mod,
                        flags, alpha, force_bn, _stream())
    _lib.check(rc, "sb_gemm_bf16")
    _count()
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6,
              out_dtype=_BF16, act: int = ACT_NONE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dim of a 2-D row-major tensor (fp32 or bf16 in; bf16 or fp32 out)."""
    _chk_cuda(x, gamma, beta, out)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype in (_BF16, _F32)
    M, Cc = x.shape
    if out is None:
        out = torch.empty((M, Cc), dtype=out_dtype, device=x.device)
    assert gamma.dtype == _F32 and beta.dtype == _F32
    L = _lib.load()
    rc = L.sb_layernorm(x.data_ptr(), x.stride(0), int(x.dtype == _F32), out.data_ptr(), out.stride(0),
                        int(out.dtype == _F32), gamma.data_ptr(), beta.data_ptr(), M, Cc, eps, act,
                        _stream())
    _lib.check(rc, "sb_layernorm")
    _count()
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, batch: int, heads: int, nq: int,
              nk: int, scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              q_shared: bool = False, kv_shared: bool = False) -> torch.Tensor:
    """Plain batched MHA. q [batch*nq, heads*hd], k/v [batch*nk, heads*hd] (bf16, row views allowed).
    With q_shared / kv_shared the operand holds one batch entry ([nq|nk, C]) read by every batch element."""
    _chk_cuda(q, k, v, out)
    assert q.dtype == _BF16 and k.dtype == _BF16 and v.dtype == _BF16
    Cc = q.shape[1]
    hd = Cc // heads
    assert q.shape[0] == (1 if q_shared else batch) * nq
    assert k.shape[0] == (1 if kv_shared else batch) * nk and v.shape[0] == k.shape[0]
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if out is None:
        out = torch.empty((batch * nq, Cc), dtype=_BF16, device=q.device)
    L = _lib.load()
    rc = L.sb_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                        out.data_ptr(), out.stride(0), batch, heads, hd, nq, nk, scale, int(q_shared),
                        int(kv_shared), _stream())
    _lib.check(rc, "sb_attention")
    _count()
    return out


def window_attention(qkv: torch.Tensor, qkv_bias: Optional[torch.Tensor], batch: int, H: int, W: int,
                     heads: int, ws: int, pool: int = 1, scale: Optional[float] = None) -> torch.Tensor:
    """Hiera (windowed / global, optionally q-pooled) attention over a fused qkv [B*H*W, 3*C] buffer."""
    _chk_cuda(qkv, qkv_bias)
    assert qkv.dtype == _BF16 and qkv.is_contiguous() and qkv.shape[0] == batch * H * W
    C3 = qkv.shape[1]
    Cc = C3 // 3
    hd = Cc // heads
    if scale is None:
        scale = 1.0 / math.sqrt(hd)
    if ws <= 0 or ws >= max(H, W):
        ws = max(H, W)
    Ho, Wo = H // pool, W // pool
    out = torch.empty((batch * Ho * Wo, Cc), dtype=_BF16, device=qkv.device)
    L = _lib.load()
    rc = L.sb_window_attention(qkv.data_ptr(), _ptr(qkv_bias), out.data_ptr(), batch, H, W, heads, hd,
                               ws, pool, scale, _stream())
    _lib.check(rc, "sb_window_attention")
    _count()
    return out


def im2col_k7s4(img: torch.Tensor, kp: int) -> torch.Tensor:
    """[B,Cin,S,S] fp32 -> [B*(S/4)^2, kp] bf16 patches for the 7x7 stride-4 pad-3 patch embedding."""
    _chk_cuda(img)
    assert img.dtype == _F32 and img.is_contiguous() and img.dim() == 4 and img.shape[2] == img.shape[3]
    B, Cin, S, _ = img.shape
    cols = torch.empty((B * (S // 4) ** 2, kp), dtype=_BF16, device=img.device)
    L = _lib.load()
    _lib.check(L.sb_im2col_k7s4(img.data_ptr(), cols.data_ptr(), B, Cin, S, kp, _stream()), "sb_im2col_k7s4")
    _count()
    return cols


def maxpool2x2(x: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """Token-major [B*H*W, C] -> [B*(H/2)*(W/2), C]."""
    _chk_cuda(x)
    assert x.is_contiguous() and x.dtype in (_BF16, _F32)
    Cc = x.shape[1]
    out = torch.empty((B * (H // 2) * (W // 2), Cc), dtype=x.dtype, device=x.device)
    L = _lib.load()
    _lib.check(L.sb_maxpool2x2(x.data_ptr(), out.data_ptr(), int(x.dtype == _F32), B, H, W, Cc, _stream()),
               "sb_maxpool2x2")
    _count()
    return out


def add_upsample2x_(dst: torch.Tensor, src: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """dst[B*H*W, C] += nearest-2x(src[B*(H/2)*(W/2), C]) — both fp32 token-major."""
    _chk_cuda(dst, src)
    assert dst.dtype == _F32 and src.dtype == _F32 and dst.is_contiguous() and src.is_contiguous()
    L = _lib.load()
    _lib.check(L.sb_add_upsample2x(dst.data_ptr(), src.data_ptr(), B, H, W, dst.shape[1], _stream()),
               "sb_add_upsample2x")
    _count()
    return dst


def nhwc_to_nchw(x: torch.Tensor, B: int, HW: int, out_dtype=_F32,
                 chan_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B*HW, C] token-major -> [B, C, HW] (optionally adding a per-channel fp32 vector)."""
    _chk_cuda(x, chan_add)
    assert x.is_contiguous() and x.shape[0] == B * HW
    Cc = x.shape[1]
    out = torch.empty((B, Cc, HW), dtype=out_dtype, device=x.device)
    L = _lib.load()
    _lib.check(L.sb_nhwc_to_nchw(x.data_ptr(), int(x.dtype == _F32), out.data_ptr(), int(out_dtype == _F32),
                                 B, HW, Cc, _ptr(chan_add), _stream()), "sb_nhwc_to_nchw")
    _count()
    return out


def nchw_to_nhwc(x: torch.Tensor, out_dtype=_BF16) -> torch.Tensor:
    """[B, C, ...spatial] -> [B*HW, C] token-major."""
    _chk_cuda(x)
    assert x.is_contiguous()
    B, Cc = x.shape[0], x.shape[1]
    HW = x.numel() // (B * Cc)
    out = torch.empty((B * HW, Cc), dtype=out_dtype, device=x.device)
    L = _lib.load()
    _lib.check(L.sb_nchw_to_nhwc(x.data_ptr(), int(x.dtype == _F32), out.data_ptr(), int(out_dtype == _F32),
                                 B, HW, Cc, _stream()), "sb_nchw_to_nhwc")
    _count()
    return out


def add_cast(a: torch.Tensor, b: Optional[torch.Tensor] = None, out_dtype=_BF16) -> torch.Tensor:
    """out = a + b (b fp32, broadcast by flat index modulo b.numel()), converted to out_dtype."""
    _chk_cuda(a, b)
    assert a.is_contiguous() and a.dtype in (_BF16, _F32)
    out = torch.empty(a.shape, dtype=out_dtype, device=a.device)
    b_mod = 0
    if b is not None:
        assert b.dtype == _F32 and b.is_contiguous()
        b_mod = b.numel() if b.numel() != a.numel() else 0
    L = _lib.load()
    _lib.check(L.sb_add_cast(a.data_ptr(), int(a.dtype == _F32), _ptr(b), b_mod, out.data_ptr(),
                             int(out_dtype == _F32), a.numel(), _stream()), "sb_add_cast")
    _count()
    return out


def prompt_tokens(coords: torch.Tensor, labels: torch.Tensor, gauss: torch.Tensor, point_emb: torch.Tensor,
                  not_a_point: torch.Tensor, out_tokens: torch.Tensor, image_size: int, pad: bool = True) -> torch.Tensor:
    """Decoder token matrix [B, 6 + Np (+1 pad), 256] fp32: output tokens + point-prompt embeddings."""
    _chk_cuda(coords, labels, gauss, point_emb, not_a_point, out_tokens)
    assert coords.dtype == _F32 and coords.is_contiguous() and coords.dim() == 3 and coords.shape[2] == 2
    assert labels.dtype == torch.int32 and labels.is_contiguous() and labels.shape == coords.shape[:2]
    B, Np = coords.shape[0], coords.shape[1]
    Nt = 6 + Np + (1 if pad else 0)
    tokens = torch.empty((B, Nt, 256), dtype=_F32, device=coords.device)
    L = _lib.load()
    _lib.check(L.sb_prompt_tokens(coords.data_ptr(), labels.data_ptr(), B, Np, int(pad), gauss.data_ptr(),
                                  point_emb.data_ptr(), not_a_point.data_ptr(), out_tokens.data_ptr(), image_size,
                                  tokens.data_ptr(), _stream()), "sb_prompt_tokens")
    _count()
    return tokens


def mask_downscale(mask: torch.Tensor, w) -> torch.Tensor:
    """[B, S, S] fp32 mask prompt -> [B*(S/4)^2, 16] bf16 (mask_downscaling convs 0..5 fused)."""
    _chk_cuda(mask)
    assert mask.dtype == _F32 and mask.is_contiguous() and mask.dim() == 3
    B, S, _ = mask.shape
    out = torch.empty((B * (S // 4) ** 2, 16), dtype=_BF16, device=mask.device)
    L = _lib.load()
    _lib.check(L.sb_mask_downscale(mask.data_ptr(), B, S, *[t.data_ptr() for t in w], out.data_ptr(), _stream()),
               "sb_mask_downscale")
    _count()
    return out


def upscale1_post(g1: torch.Tensor, feat_s1: torch.Tensor, s1_batch_stride: int, gamma: torch.Tensor,
                  beta: torch.Tensor, B: int, h: int, w: int) -> torch.Tensor:
    _chk_cuda(g1, feat_s1, gamma, beta)
    assert g1.dtype == _BF16 and g1.is_contiguous() and g1.shape == (B * h * w, 256)
    assert feat_s1.dtype == _F32 and feat_s1.is_contiguous()
    u1 = torch.empty((B * 4 * h * w, 64), dtype=_BF16, device=g1.device)
    L = _lib.load()
    _lib.check(L.sb_upscale1_post(g1.data_ptr(), feat_s1.data_ptr(), s1_batch_stride, gamma.data_ptr(),
                                  beta.data_ptr(), B, h, w, u1.data_ptr(), _stream()), "sb_upscale1_post")
    _count()
    return u1


def upscale2_mask(g2: torch.Tensor, feat_s0: torch.Tensor, s0_batch_stride: int, hyper: torch.Tensor, B: int,
                  H1: int, W1: int) -> torch.Tensor:
    _chk_cuda(g2, feat_s0, hyper)
    assert g2.dtype == _BF16 and g2.is_contiguous() and g2.shape == (B * H1 * W1, 128)
    assert feat_s0.dtype == _F32 and feat_s0.is_contiguous()
    assert hyper.dtype == _F32 and hyper.is_contiguous() and hyper.shape == (B, 4, 32)
    masks = torch.empty((B, 4, 2 * H1, 2 * W1), dtype=_F32, device=g2.device)
    L = _lib.load()
    _lib.check(L.sb_upscale2_mask(g2.data_ptr(), feat_s0.data_ptr(), s0_batch_stride, hyper.data_ptr(), B, H1, W1,
                                  masks.data_ptr(), _stream()), "sb_upscale2_mask")
    _count()
    return masks


def select_mask(masks: torch.Tensor, ious: torch.Tensor, delta: float, thresh: float):
    """dynamic_multimask_via_stability: returns (sel_idx int32 [B], sel_iou fp32 [B])."""
    _chk_cuda(masks, ious)
    assert masks.dtype == _F32 and masks.is_contiguous() and masks.dim() == 4 and masks.shape[1] == 4
    assert ious.dtype == _F32 and ious.is_contiguous() and ious.shape == (masks.shape[0], 4)
    B = masks.shape[0]
    idx = torch.empty((B,), dtype=torch.int32, device=masks.device)
    iou = torch.empty((B,), dtype=_F32, device=masks.device)
    L = _lib.load()
    _lib.check(L.sb_select_mask(masks.data_ptr(), ious.data_ptr(), B, masks.shape[2] * masks.shape[3], delta, thresh,
                                idx.data_ptr(), iou.data_ptr(), _stream()), "sb_select_mask")
    _count()
    return idx, iou
