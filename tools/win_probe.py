"""Timing of the Hiera-L window-attention shapes that still run on the mma.sync kernel (stages 1, 2, 4 and the q-pooled
first blocks), 8 crops, L2 flushed: us per launch and GB/s of (qkv read + out written) against the HBM peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saber_b200 import ops

SHAPES = [("stage1 ws8", 256, 256, 2, 8, 1, 2), ("stage2 first ws8 pool", 256, 256, 4, 8, 2, 1), ("stage2 ws4", 128, 128, 4, 4, 1, 5),
          ("stage3 first ws4 pool", 128, 128, 8, 4, 2, 1), ("stage3 ws16 (tcgen05)", 64, 64, 8, 16, 1, 32),
          ("stage3 global (tcgen05)", 64, 64, 8, 0, 1, 3), ("stage4 first ws16 pool", 64, 64, 16, 16, 2, 1),
          ("stage4 ws8", 32, 32, 16, 8, 1, 3)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    tot = 0.0
    for name, H, W, heads, ws, pool, count in SHAPES:
        C = heads * 72
        qkv = torch.randn(B * H * W, 3 * C, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(3 * C, device="cuda")
        for _ in range(2):
            ops.window_attention(qkv, bias, B, H, W, heads, ws, pool)
        best = 1e9
        for _ in range(5):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            out = ops.window_attention(qkv, bias, B, H, W, heads, ws, pool)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        nbytes = qkv.numel() * 2 + out.numel() * 2
        tot += best * count
        print(f"{name:28s} H={H:3d} heads={heads:2d} ws={ws:2d} pool={pool}: {best * 1e3:8.1f} us  {nbytes / best / 1e6:7.1f} GB/s "
              f"({nbytes / best / 1e6 / 6547.8:.2f} of HBM peak)  x{count} per encoder pass = {best * count:.3f} ms")
    print(f"attention per encoder pass of {B} crops: {tot:.3f} ms")


if __name__ == "__main__":
    main()
