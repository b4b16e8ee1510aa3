"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: z-slab ranges, GPUPool task rule, label gather and
the sharded slice_by_slice composition (label function and CC are CPU stand-ins from the oracle here; on the GPU box
the same code runs with the CUDA kernels over NCCL)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from saber_b200 import dist as sbdist


def test_zslab_ranges_cover_volume():
    for Z in (1, 7, 200, 300):
        for world in (1, 2, 3, 8):
            r = [sbdist.zslab_range(Z, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == Z
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert sbdist.tasks_for_rank(10, 1, 4) == [1, 5, 9]


def _worker(rank, world, port, Z, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import saber_ref
    from saber_b200 import synth
    vol = synth.make_label_volume((Z, 40, 48), seed=3, n_ellipsoids=14, speckle=0.01)
    z0, z1 = sbdist.zslab_range(Z, rank, world)

    def label_fn(slab):  # stand-in for the per-slice AMG + stitch: the slab's labels are already there
        return slab.clone()

    def separate_fn(full):
        return torch.from_numpy(saber_ref.separate_masks(full.numpy().view(np.uint16), 1).astype(np.int64))

    out = sbdist.slice_by_slice_sharded(label_fn, separate_fn, vol[z0:z1], Z)
    if rank == 0:
        want = saber_ref.separate_masks(vol.numpy().view(np.uint16), 1)
        q.put(bool(np.array_equal(out.numpy(), want.astype(np.int64))))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_slice_by_slice_world2():
    ctx = mp.get_context("spawn")
    for Z in (9, 10):  # uneven and even slabs
        q = ctx.Queue()
        port = 29500 + (os.getpid() + Z) % 2000
        procs = [ctx.Process(target=_worker, args=(r, 2, port, Z, q)) for r in range(2)]
        for p in procs:
            p.start()
        ok = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert ok


def test_merge_captured_scores_reproduces_single_process_order():
    """5 tracked objects on 2 ranks (0,2,4 | 1,3): the merged hook log equals the single-process log, including a frame
    key that collected two call groups (the reference's off-by-one files fwd(start+1) and bwd(start-1) under `start`)."""
    total, world = 5, 2
    single = {None: [10 + k for k in range(total)], 3: [20 + k for k in range(total)] + [30 + k for k in range(total)],
              4: [40 + k for k in range(total)]}
    n_local = [len([k for k in range(total) if k % world == r]) for r in range(world)]
    per_rank = []
    for r in range(world):
        d = {}
        for key, vals in single.items():
            groups = [vals[g * total:(g + 1) * total] for g in range(len(vals) // total)]
            d[key] = [grp[k] for grp in groups for k in range(total) if k % world == r]
        per_rank.append(d)
    assert sbdist.merge_captured_scores(per_rank, n_local) == single
    # a rank without objects (more ranks than objects)
    per_rank2 = [{7: [1.0]}, {}]
    assert sbdist.merge_captured_scores(per_rank2, [1, 0]) == {7: [1.0]}


def _worker_max(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.integers(0, 60000, size=(world, 6, 20, 24)).astype(np.uint16)
    mine = torch.from_numpy(full[rank].view(np.int16).copy())
    out = sbdist.allreduce_max_labels(mine)
    want = full.max(axis=0)
    ok = bool(np.array_equal(out.numpy().view(np.uint16), want))
    # feature exchange: rank r holds the frames of its slab
    Z = 5
    z0, z1 = sbdist.zslab_range(Z, rank, world)
    cached = {f: {"feat": torch.full((4096, 256), float(f)), "s1": torch.full((16384, 64), float(f) + 0.5),
                  "s0": torch.full((65536, 32), float(f) + 0.25)} for f in range(z0, z1)}
    sbdist.exchange_frame_features(cached, Z)
    ok = ok and sorted(cached) == list(range(Z)) and all(
        float(cached[f]["feat"][0, 0]) == f and float(cached[f]["s1"][5, 5]) == f + 0.5 and
        float(cached[f]["s0"][7, 7]) == f + 0.25 for f in range(Z))
    if rank == 0:
        q.put(ok)
    else:
        assert ok
    dist.barrier()
    dist.destroy_process_group()


def test_label_max_merge_and_feature_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_max, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
