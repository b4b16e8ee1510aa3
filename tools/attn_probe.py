"""GPU probe: Hiera-L stage-3 attention (16 x 16 windows and global) — parity vs an fp32 SDPA reference and CUDA-event
timing of the tcgen05 kernel (hiera_attn_tc.cu) against the mma.sync kernel (SB_HIERA_TC=0 in a second process).

    python tools/attn_probe.py [--crops 8] [--reps 20]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=8)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from saber_b200 import ops
    from test_gpu_kernels import _window_ref

    ops.require_b200()
    BF16 = torch.bfloat16
    heads, hd, H, W = 8, 72, 64, 64
    C = heads * hd
    mode = os.environ.get("SB_HIERA_TC", "1")
    print(f"SB_HIERA_TC={mode}")
    for ws, name in ((16, "window16"), (0, "global")):
        # ---- parity (2 crops)
        torch.manual_seed(7)
        B = 2
        qkv = (torch.randn(B * H * W, 3 * C, device="cuda") * 1.5).to(BF16)
        bias = torch.zeros(3 * C, device="cuda")
        out = ops.window_attention(qkv, bias, B, H, W, heads, ws, 1)
        torch.cuda.synchronize()
        ref = _window_ref(qkv, bias.to(BF16), B, H, W, heads, hd, ws, 1)
        err = (out.float() - ref).abs()
        print(f"{name}: max abs err {err.max().item():.4e}  mean {err.mean().item():.3e}  ref rms {ref.pow(2).mean().sqrt().item():.3f}"
              f"  nan={bool(torch.isnan(out.float()).any())}")
        # per-head / per-column-block error (layout debugging aid)
        e = err.view(B * H * W, heads, hd)
        print("   err by head:", [f"{x:.2e}" for x in e.amax(dim=(0, 2)).tolist()])
        print("   err cols 0-63 / 64-71:", f"{e[..., :64].max().item():.2e}", f"{e[..., 64:].max().item():.2e}")
        # ---- timing (bench shape)
        B = args.crops
        qkv = (torch.randn(B * H * W, 3 * C, device="cuda")).to(BF16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            ops.window_attention(qkv, bias, B, H, W, heads, ws, 1)
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.window_attention(qkv, bias, B, H, W, heads, ws, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        nk = 256 if ws == 16 else H * W
        flops = 4.0 * B * H * W * nk * hd * heads
        byt = B * H * W * (3 * C + C) * 2
        print(f"{name}: B={B} median {med * 1e3:.1f} us  min {ts[0] * 1e3:.1f} us  {flops / med / 1e9:.1f} TFLOP/s  "
              f"{byt / med / 1e6:.0f} GB/s (qkv read + out write)")
        if mode != "0":
            prof_variants(ops, qkv, B, H, W, heads, ws, name, args.reps, flush)


PROF_SLOTS = ["prod wait k_empty", "prod wait v_empty", "mma wait q_full", "mma wait k_full", "mma wait p_full",
              "mma wait v_full", "mma wait o_empty", "smx wait s_full", "smx tmem ld S", "smx max+xchg", "smx item end",
              "smx exp+store P", "epilogue warp", "smx total", "tiles"]


def prof_variants(ops, qkv, B, H, W, heads, ws, name, reps, flush):
    """Ring-depth variants of the tcgen05 kernel: timing, then the per-role clock breakdown of the instrumented build."""
    import ctypes

    from saber_b200 import lib as _lib
    L = _lib.load()
    out = torch.empty((B * H * W, heads * 72), dtype=torch.bfloat16, device="cuda")
    prof = torch.zeros((148, 16), dtype=torch.int64, device="cuda")
    scale = 72 ** -0.5
    st = torch.cuda.current_stream().cuda_stream
    for variant, label in ((0, "P in TMEM, K3 V3, epilogue warps"),):
        ts = []
        for it in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            # the instrumented entry point with a null profile buffer is rejected, so time the PROF build too
            rc = L.sb_hiera_attention_tc_prof(qkv.data_ptr(), out.data_ptr(), B, H, W, heads, ws, scale, prof.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, _lib.last_error()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        pr = prof.cpu().double()
        tiles = pr[:, 14].clamp(min=1)
        per_tile = (pr[:, :14] / tiles[:, None]).mean(0)
        print(f"  {name} variant {variant} ({label}): instrumented median {ts[len(ts) // 2] * 1e3:.1f} us; clocks per tile (mean over CTAs):")
        print("    " + "  ".join(f"{n}={per_tile[i]:.0f}" for i, n in enumerate(PROF_SLOTS[:14])))


if __name__ == "__main__":
    main()
