"""GEMM probe: GPU time per launch (CUDA graph of R launches, CUDA events) for the shapes that dominate the step.
Run under gpurun. SB_GEMM_DBG=1/2/4 bisects the epilogue (skip stores / TMEM loads / math)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saber_b200 import ops

SHAPES = [  # M, N, K, act, res(0 none, 1 f32), out_f32, bn
    (32768, 2304, 576, 1, 0, 0, 0), (32768, 1728, 576, 0, 0, 0, 0), (32768, 576, 2304, 0, 1, 1, 0),
    (32768, 576, 576, 0, 1, 1, 0), (131072, 1152, 288, 1, 0, 0, 0), (524288, 576, 144, 1, 0, 0, 0),
    (524288, 432, 144, 0, 0, 0, 0), (8192, 4608, 1152, 1, 0, 0, 0), (8192, 8192, 8192, 0, 0, 0, 0),
    (128, 256, 64, 0, 0, 0, 256), (128, 64, 64, 0, 0, 0, 64), (512, 256, 256, 0, 0, 0, 0), (512, 2048, 256, 2, 0, 0, 0),
    (786432, 256, 256, 0, 1, 0, 0), (786432, 128, 256, 0, 1, 0, 0),
    (86016, 2304, 576, 1, 0, 0, 0),  # 15: fc1 at the bench's encoder batch (21 crops x 4096 tokens): roofline.traffic
    (86016, 576, 2304, 0, 1, 1, 0),  # 16: fc2 (192-wide tiles)
]
R = 10


def main():
    only = [int(x) for x in sys.argv[1:]]
    torch.manual_seed(0)
    dev = "cuda"
    
    for idx, (M, N, K, act, res, of32, bn) in enumerate(SHAPES):
        if only and idx not in only:
            continue
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        r = torch.randn(M, N, device=dev) if res else None
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if of32 else torch.bfloat16)

        def run():
            ops.gemm(a, w, bias, act, r, 0, out.dtype, out=out, force_bn=bn)

        run(); run()
        torch.cuda.synchronize()
        if only:  # ncu mode: plain launches only
            run(); run()
            torch.cuda.synchronize()
            continue
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g):
                for _ in range(R):
                    run()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / R
        tiles128 = ((M + 127) // 128)
        print(f"M={M:7d} N={N:5d} K={K:5d} act={act} res={res} f32={of32} bn={bn:3d}: {us:9.1f} us  "
              f"{2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  out {M * N * (4 if of32 else 2) / us / 1e3:7.1f} GB/s")
        if os.environ.get("SB_GEMM_PROF", "1") != "0":  # one instrumented launch: where do the warps wait?
            from saber_b200 import lib as _lib
            L = _lib.load()
            cnt = torch.zeros(8, dtype=torch.int64, device=dev)
            L.sb_gemm_set_prof(cnt.data_ptr())
            run()
            torch.cuda.synchronize()
            L.sb_gemm_set_prof(None)
            c = cnt.tolist()
            ctas, tot = max(c[6], 1), max(c[5], 1)
            etiles = max(c[7], 1)
            print(f"      per CTA (clocks, MMA-warp span {tot / ctas:9.0f}): producer waits slot {c[0] / tot:5.1%}  MMA waits operands "
                  f"{c[1] / tot:5.1%}  MMA waits accumulator {c[2] / tot:5.1%} | epilogue set: waits {c[3] / etiles:7.0f} clk/tile, works "
                  f"{c[4] / etiles:7.0f} clk/tile ({etiles} tiles)")
        del a, w, r, out


if __name__ == "__main__":
    main()
