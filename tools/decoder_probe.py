"""One AMG prompt batch (64 points -> decoder, m2m refinement of the 192 candidates, post-processing) of hiera-large on
one 1024^2 crop, eager launches (no CUDA graph) so ncu / the GEMM profiler see every kernel. Run under gpurun:
  python tools/decoder_probe.py            # CUDA-event time of the batch + per-GEMM-shape table
  ncu --metrics gpu__time_duration.sum ... python tools/decoder_probe.py --once"""
import os, sys
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saber_b200 import ops, synth
from saber_b200.sam2.build_sam import build_sam2
from saber_b200.sam2.automatic_mask_generator import SAM2AutomaticMaskGenerator


def main():
    once = "--once" in sys.argv
    cfg = "large"
    model = build_sam2(cfg, None, device="cuda")
    gen = SAM2AutomaticMaskGenerator(model, points_per_side=32, points_per_batch=64, pred_iou_thresh=0.7,
                                     stability_score_thresh=0.92, stability_score_offset=0.7, crop_n_layers=0,
                                     box_nms_thresh=0.7, use_m2m=True, multimask_output=True)
    gen.use_cuda_graph = False
    img = synth.make_tomogram((1, 1024, 1024), seed=0, device="cuda")[0].contiguous()
    img = (img - img.min()) / (img.max() - img.min())
    plan = gen._plan((1024, 1024))
    ws = gen._workspace(plan)
    feats = gen.predictor.encode_crops(img, plan.crops_dev)
    tok = feats.tok
    crop = plan.crops[0]

    def batch():
        gen._process_batch(0, crop.in_points[:64], crop.labels[:64], tok["embed"][:4096], tok["s0"][:65536],
                           tok["s1"][:16384], plan, ws, crop.box, 0, None)

    batch()
    torch.cuda.synchronize()
    if once:
        torch.cuda.profiler.start()
        batch()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n0 = ops.launch_count
    e0.record()
    for _ in range(5):
        batch()
    e1.record()
    torch.cuda.synchronize()
    print(f"batch of 64 points (+192 m2m): {e0.elapsed_time(e1) / 5:.3f} ms, {(ops.launch_count - n0) // 5} launches")
    prof = ops.GemmProfiler()
    with prof:
        batch()
    r = prof.summary()
    print(f"GEMM kernels: {r['launches']} launches, {r['ms']:.3f} ms, {r['tflops']:.1f} TFLOP/s")
    for tag, n, ms, tf in prof.by_shape(40):
        print(f"{ms:9.3f} ms  n={n:4d}  {tf:8.1f} TFLOP/s  {tag}")


if __name__ == "__main__":
    main()
