"""z-axis propagation on N GPUs (BASELINE configs[2] semantics): one process per GPU under torchrun.
Phase A: every rank encodes its z-slab of frames once and the cached features are exchanged (NCCL broadcast per slab);
Phase B: tracked objects are sharded over the ranks (object k -> rank k % N); Phase C: the uint16 label volumes are
merged with an element-wise max after the forward and after the backward pass. Prints device-synchronised wall times
(max over ranks) and checks that every rank ends with the same label volume.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/propagation_probe_dist.py [Z] [N_obj ...]"""
import os, sys, time, zlib
os.environ.setdefault("SABER_B200_ALLOW_RANDOM_INIT", "1")  # probes run on synthetic random-init weights
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from saber_b200 import synth
from saber_b200.adapters.base import SAM2AdapterConfig, cfgAMG
from saber_b200.adapters.sam2 import SAM2Adapter


def ellipse(hw, cy, cx, ry, rx):
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    return (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1).astype(np.float32)


def tmax(dt, dev):
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    Z = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    H, W = 928, 960
    ad = SAM2Adapter(SAM2AdapterConfig(cfg="large", amg_cfg=cfgAMG(sam2_cfg="large"), num_maskmem=2, seed=0), device=dev)
    vol = synth.make_tomogram((Z, H, W), seed=3, n_ellipsoids=10, device=dev)
    ad._video()
    ad.set_volume(vol[: 2 * world].contiguous())  # warm the kernels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ad.set_volume(vol)
    torch.cuda.synchronize()
    t_set = tmax(time.perf_counter() - t0, dev)
    if rank == 0:
        print(f"[{world} GPU] set_volume (normalize + resize + encode {Z} frames z-slab sharded + feature exchange): "
              f"{t_set * 1e3:.1f} ms = {t_set / Z * 1e3:.2f} ms/frame", flush=True)
    rng = np.random.default_rng(0)
    for n_obj in [int(a) for a in sys.argv[2:]] or [8]:
        seeds = [ellipse((H, W), rng.uniform(200, 700), rng.uniform(200, 700), rng.uniform(30, 90), rng.uniform(30, 90))
                 for _ in range(n_obj)]
        for rep in range(2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            out = ad.segment_volume(Z // 2, masks=seeds, vol_shape=(Z, H, W), min_presence_score=-1e9)
            torch.cuda.synchronize()
            dt = tmax(time.perf_counter() - t0, dev)
            ad.reset_state()
        crc = zlib.crc32(np.ascontiguousarray(out).tobytes())
        crcs = [None] * world
        if world > 1:
            dist.all_gather_object(crcs, crc)
        else:
            crcs = [crc]
        if rank == 0:
            print(f"[{world} GPU] segment_volume N_obj={n_obj}: {dt * 1e3:.1f} ms for {Z} frames ({dt / Z * 1e3:.2f} ms/frame, "
                  f"{dt / max(1, (Z - 1) * n_obj) * 1e3:.2f} ms per object-frame); label volume crc32 per rank {crcs} "
                  f"({'identical' if len(set(crcs)) == 1 else 'MISMATCH'}); labels present {sorted(set(np.unique(out).tolist()))[:10]}"
                  f"; shard={ad.prop_shard} relay={ad.relay_stats}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
