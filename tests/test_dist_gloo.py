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


# ---------------------------------------------------------------------------------------------------------------------
# z-slab relay of the propagation: a CPU stand-in predictor with the real state layout and the real dependency
# structure (conditioning frames + the previous frame's memory + the last 15 object pointers; reverse pass sees
# forward-pass pointers) so that any missing / stale / misordered halo entry changes the result.
# ---------------------------------------------------------------------------------------------------------------------
class _FakeVideo:
    num_maskmem = 2
    max_obj_ptrs = 16

    def _obj_id_to_idx(self, st, obj_id):
        if obj_id not in st["obj_id_to_idx"]:
            idx = len(st["obj_id_to_idx"])
            st["obj_id_to_idx"][obj_id] = idx
            st["obj_ids"] = list(st["obj_id_to_idx"])
            st["output_dict_per_obj"][idx] = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        return st["obj_id_to_idx"][obj_id]

    @staticmethod
    def new_state(Z, feats):
        return {"num_frames": Z, "obj_id_to_idx": {}, "obj_ids": [], "output_dict_per_obj": {}, "feats": feats}

    def add_seed(self, st, frame, obj_id):
        idx = self._obj_id_to_idx(st, obj_id)
        g = torch.Generator().manual_seed(1000 + obj_id)
        st["output_dict_per_obj"][idx]["cond_frame_outputs"][frame] = {
            "maskmem_features": torch.randn(4096, 64, generator=g).to(torch.bfloat16), "maskmem_pos_enc": True,
            "pred_masks": None, "obj_ptr": torch.randn(1, 256, generator=g), "object_score_logits": None}

    def propagate_in_video(self, st, start_frame_idx, max_frame_num_to_track, reverse=False):
        Z = st["num_frames"]
        if reverse:
            order = range(start_frame_idx, max(start_frame_idx - max_frame_num_to_track, 0) - 1, -1)
        else:
            order = range(start_frame_idx, min(start_frame_idx + max_frame_num_to_track, Z - 1) + 1)
        for f in order:
            res = []
            for idx in range(len(st["obj_ids"])):
                od = st["output_dict_per_obj"][idx]
                if f in od["cond_frame_outputs"]:
                    res.append(torch.zeros(1))
                    continue
                acc = torch.zeros(256)
                memsum = torch.zeros(64)
                for cf, out in od["cond_frame_outputs"].items():
                    acc = acc + out["obj_ptr"].view(-1) * 0.5
                    memsum = memsum + out["maskmem_features"].float().mean(0)
                prev = f + 1 if reverse else f - 1
                out = od["non_cond_frame_outputs"].get(prev)
                if out is not None:
                    memsum = memsum + 2.0 * out["maskmem_features"].float().mean(0)
                for t in range(1, self.max_obj_ptrs):
                    tt = f + t if reverse else f - t
                    o = od["non_cond_frame_outputs"].get(tt)
                    if o is not None:
                        acc = acc + o["obj_ptr"].view(-1) / (t + 1.0)
                feat = st["feats"][f]  # KeyError if this rank does not own the frame: the relay must never ask for it
                ptr = torch.tanh(acc + feat[:256] + memsum.sum())
                mem = (torch.outer(torch.arange(4096, dtype=torch.float32) % 7 + 1.0, memsum) * 0.01 + ptr[:64]).to(torch.bfloat16)
                od["non_cond_frame_outputs"][f] = {"maskmem_features": mem, "maskmem_pos_enc": True, "pred_masks": None,
                                                   "obj_ptr": ptr.view(1, 256), "object_score_logits": None}
                res.append(ptr.sum().view(1))
            yield f, st["obj_ids"], torch.cat(res).view(-1, 1)


def _relay_reference(Z, start, obj_ids, feats):
    fv = _FakeVideo()
    st = fv.new_state(Z, feats)
    for o in obj_ids:
        fv.add_seed(st, start, o)
    out = {}
    for f, ids, v in fv.propagate_in_video(st, start, Z, False):
        out[(0, f)] = v.clone()
    for f, ids, v in fv.propagate_in_video(st, start, Z, True):
        out[(1, f)] = v.clone()
    return out


def _relay_worker(rank, world, port, Z, start, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    feats_all = {f: torch.randn(300, generator=g) for f in range(Z)}
    z0, z1 = sbdist.zslab_range(Z, rank, world)
    fv = _FakeVideo()
    st = fv.new_state(Z, {f: feats_all[f] for f in range(z0, z1)})  # own slab only
    obj_ids = [1, 2, 5]
    got, keys = {}, []

    def seed_fn():
        for o in obj_ids:
            fv.add_seed(st, start, o)

    stats = sbdist.relay_propagate(fv, st, obj_ids, start, Z, seed_fn,
                                   lambda ps, f, ids, v: got.__setitem__((ps, f), v.clone()),
                                   lambda k, ps: keys.append((k, ps)), torch.device("cpu"))
    allgot = [None] * world
    dist.all_gather_object(allgot, {k: v.tolist() for k, v in got.items()})
    if rank == 0:
        want = _relay_reference(Z, start, obj_ids, feats_all)
        merged = {}
        for d in allgot:
            for k, v in d.items():
                assert k not in merged, "a frame was tracked on two ranks"
                merged[k] = v
        ok = set(merged) == set(want) and all(merged[k] == want[k].tolist() for k in want)
        q.put((ok, stats["bytes_sent"]))
    assert all(z0 <= f < z1 for (_, f) in got), "tracked a frame outside the own slab"
    dist.barrier()
    dist.destroy_process_group()


def test_zslab_relay_reproduces_single_process_chain():
    """relay_propagate over 2 and 3 gloo ranks == the single-process forward + backward chain, exactly, for seeds in the
    middle, next to a slab boundary (the backward pass needs forward pointers owned by the next rank) and on the first
    frame; every rank only ever touches the frames of its own slab."""
    ctx = mp.get_context("spawn")
    for world, Z, start in ((2, 40, 19), (2, 40, 12), (3, 60, 35), (2, 36, 0), (2, 36, 35)):
        q = ctx.Queue()
        port = 29500 + (os.getpid() + 17 * Z + start + world) % 2000
        procs = [ctx.Process(target=_relay_worker, args=(r, world, port, Z, start, q)) for r in range(world)]
        for p in procs:
            p.start()
        ok, sent = q.get(timeout=180)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert ok, (world, Z, start)


def test_halo_pack_roundtrip():
    fv = _FakeVideo()
    st = fv.new_state(30, {f: torch.randn(300) for f in range(30)})
    for o in (1, 2):
        fv.add_seed(st, 10, o)
    list(fv.propagate_in_video(st, 10, 12, False))
    meta, mem, ptr = sbdist.halo_pack(st, 2, list(range(8, 23)), [22])
    objs = sbdist.halo_unpack(meta, mem, ptr)
    st2 = fv.new_state(30, {})
    sbdist.halo_install(fv, st2, [1, 2], objs)
    for i in range(2):
        a, b = st["output_dict_per_obj"][i], st2["output_dict_per_obj"][i]
        assert sorted(b["cond_frame_outputs"]) == [10]
        assert torch.equal(a["cond_frame_outputs"][10]["maskmem_features"], b["cond_frame_outputs"][10]["maskmem_features"])
        assert sorted(b["non_cond_frame_outputs"]) == list(range(11, 23))  # frames 8-10 were never tracked / are cond
        for f in range(11, 23):
            assert torch.equal(a["non_cond_frame_outputs"][f]["obj_ptr"], b["non_cond_frame_outputs"][f]["obj_ptr"])
        assert torch.equal(a["non_cond_frame_outputs"][22]["maskmem_features"], b["non_cond_frame_outputs"][22]["maskmem_features"])
        assert "maskmem_features" not in b["non_cond_frame_outputs"][21]
