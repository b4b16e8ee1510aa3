import os, sys
sys.path.insert(0, '/root/repo')
import torch
from saber_b200 import ops
SH = [(32768, 2304, 576, 1, 0, 0), (32768, 1728, 576, 0, 0, 0), (32768, 576, 2304, 0, 1, 1), (32768, 576, 576, 0, 1, 1),
      (131072, 1152, 288, 1, 0, 0), (131072, 288, 1152, 0, 1, 1), (131072, 288, 288, 0, 1, 1), (131072, 864, 288, 0, 0, 0),
      (524288, 576, 144, 1, 0, 0), (524288, 144, 576, 0, 1, 1), (524288, 432, 144, 0, 0, 0), (524288, 144, 144, 0, 1, 1),
      (8192, 4608, 1152, 1, 0, 0), (8192, 1152, 4608, 0, 1, 1), (8192, 3456, 1152, 0, 0, 0), (8192, 1152, 1152, 0, 1, 1),
      (786432, 256, 256, 0, 1, 0), (786432, 128, 256, 0, 1, 0)]
flush = torch.zeros(64 * 1024 * 1024, device='cuda')
for (M, N, K, act, res, of32) in SH:
    a = torch.randn(M, K, device='cuda').to(torch.bfloat16)
    w = (torch.randn(N, K, device='cuda') / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device='cuda')
    r = (torch.randn(M, N, device='cuda') if of32 else torch.randn(M, N, device='cuda').to(torch.bfloat16)) if res else None
    out = torch.empty(M, N, device='cuda', dtype=torch.float32 if of32 else torch.bfloat16)
    line = f"M={M:7d} N={N:5d} K={K:5d} act={act} res={res}: "
    for bn in (64, 128, 192, 256, -128, -192, -256):
        if bn == 64 and N > 600: 
            line += "   --  "; continue
        best = 1e9
        for _ in range(4):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ops.gemm(a, w, bias, act, r, 0, out.dtype, out=out, force_bn=bn)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        line += f"{bn:4d}:{2.0*M*N*K/best/1e9:6.0f} "
    print(line, flush=True)
