"""One Hiera-L encoder batch (8 x 1024^2 crops) eagerly, for ncu launch lists / CUDA-event timing (run under gpurun)."""
import os, sys
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saber_b200 import ops
from saber_b200.sam2.build_sam import build_sam2


def main():
    once = "--once" in sys.argv
    model = build_sam2("large", None, device="cuda")
    x = torch.randn(8, 3, 1024, 1024, device="cuda")
    model.encoder.forward(x)
    torch.cuda.synchronize()
    if once:
        torch.cuda.profiler.start()
        model.encoder.forward(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n0 = ops.launch_count
    e0.record()
    for _ in range(3):
        model.encoder.forward(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"encoder batch of 8 crops: {e0.elapsed_time(e1) / 3:.3f} ms ({e0.elapsed_time(e1) / 24:.3f} ms / crop), "
          f"{(ops.launch_count - n0) // 3} launches")
    prof = ops.GemmProfiler()
    with prof:
        model.encoder.forward(x)
    r = prof.summary()
    print(f"GEMM kernels: {r['launches']} launches, {r['ms']:.3f} ms, {r['tflops']:.1f} TFLOP/s")
    for tag, n, ms, tf in prof.by_shape(30):
        print(f"{ms:9.3f} ms  n={n:4d}  {tf:8.1f} TFLOP/s  {tag}")


if __name__ == "__main__":
    main()
