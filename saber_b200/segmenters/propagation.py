"""B200 twin of REF saber/segmenters/propagation.py (``propagationSegmenter``, :11-189).

``slice_by_slice`` (REF :164-189) is the slice-wise zero-shot path of BASELINE config 2: per z-slice
prepare -> AMG -> area filter -> duplicate removal -> sort -> label stitch, then 3-D connected components over
the whole label volume. Here every stage runs on the device; the volume is uploaded once and only the uint32
label volume comes back. ``segment`` / ``single_segment`` / ``segment_3d`` (REF :41-130) drive the z-axis memory
propagation of the adapter (``segment_volume``).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .. import ops
from ..adapters.base import AdapterConfig, cfgAMG
from . import utils
from .base import saber3D

_I32 = torch.int32


class propagationSegmenter(saber3D):
    def __init__(self, deviceID: int = 0, cfg: Optional[AdapterConfig] = None, amg_cfg: Optional[cfgAMG] = None,
                 min_mask_area: int = 100, min_rel_box_size: float = 0.025):
        self.min_rel_box_size = min_rel_box_size
        super().__init__(deviceID=deviceID, cfg=cfg, amg_cfg=amg_cfg, min_mask_area=min_mask_area)
        self.ini_depth = 10
        # concurrent slice workers of label_slices_device (1 = the plain serial loop)
        self.slice_workers = max(1, int(os.environ.get("SB_SLICE_WORKERS", "3")))
        self._peers, self._peer_streams, self._pool = [self], [torch.cuda.Stream(device=self.device)], None

    # ------------------------------------------------------------------
    @torch.inference_mode()
    def label_slices_device(self, volume: torch.Tensor, labels: torch.Tensor, z0: int = 0, z1: Optional[int] = None):
        """Slices z0..z1 of a CUDA (Z,Y,X) fp32 volume -> stitched uint16 labels written into labels[z0:z1]
        (the body of the reference loop, REF :177-187). Returns the number of masks kept per slice.

        Slices are independent, and the two halves of a slice load different units: the encoder is tensor-bound, the
        decoder streams per-prompt image features (HBM-bound). With ``slice_workers`` > 1 (default 3, ``SB_SLICE_WORKERS``)
        the slices are dealt round-robin to that many worker threads, each with its own segmenter instance (own
        workspaces and CUDA graphs, same weights) on its own stream, so one slice's encoder overlaps another's decoder:
        5.8 (serial) -> 6.14 (two workers) -> 6.26 (three) slices/s on hiera-L at 1024^2 (profiles/r02zzj, r02zzr). Results are identical to the serial loop."""
        z1 = volume.shape[0] if z1 is None else z1
        nw = min(self.slice_workers, z1 - z0)
        if nw > 1 and self.classifier is None:
            gen = self.adapter._amg().base_generator
            if gen.capture is not None or gen.phase_ms is not None:
                nw = 1  # test hook / phase instrumentation of this generator: keep every slice on it
        if nw > 1 and self.classifier is None:
            return self._label_slices_concurrent(volume, labels, z0, z1, nw)
        return self._label_slices_serial(volume, labels, z0, z1)

    _GEN_TUNABLES = ("pred_iou_thresh", "stability_score_thresh", "stability_score_offset", "mask_threshold", "box_nms_thresh",
                     "crop_nms_thresh", "use_cuda_graph", "m2m_gate", "graph_lanes", "exec_ppb", "m2m_batch")
    _FILTER_TUNABLES = ("min_area_filter", "min_rel_box_size", "max_rel_box_size")

    def _peer_segmenters(self, nw: int):
        """[self, clone, ...]: a clone shares this segmenter's MODEL (weights) and owns its mask generator — workspaces,
        CUDA graphs, streams. Settings changed on this segmenter's generator after construction are mirrored before every
        run (changed thresholds are baked into captured graphs: the clone's graphs are dropped and re-captured)."""
        from ..adapters.sam2 import build_amg
        src = self.adapter
        fgen = src._amg()
        while len(self._peers) < nw:
            peer = propagationSegmenter(deviceID=self.deviceID, cfg=self.adapter_cfg, min_mask_area=self.min_mask_area,
                                        min_rel_box_size=self.min_rel_box_size)
            peer.slice_workers = 1
            amg_dict = src._config.amg_cfg.dict() if src._config.amg_cfg is not None else cfgAMG(sam2_cfg=src._config.cfg).dict()
            peer.adapter._mask_generator = build_amg(amg_dict, src._config.min_mask_area, device=src.device,
                                                     model=fgen.base_generator.predictor.model)
            self._peers.append(peer)
            self._peer_streams.append(torch.cuda.Stream(device=self.device))
        for w, peer in enumerate(self._peers[1:nw], start=1):
            pf = peer.adapter._amg()
            changed = False
            for k in self._GEN_TUNABLES:
                if getattr(pf.base_generator, k) != getattr(fgen.base_generator, k):
                    setattr(pf.base_generator, k, getattr(fgen.base_generator, k))
                    changed = True
            for k in self._FILTER_TUNABLES:
                if hasattr(fgen, k) and getattr(pf, k, None) != getattr(fgen, k):
                    setattr(pf, k, getattr(fgen, k))
            if changed:
                pf.base_generator._graphs.clear()
            peer.min_mask_area, peer.remove_repeating_masks = self.min_mask_area, self.remove_repeating_masks
        return self._peers[:nw]

    def _label_slices_concurrent(self, volume: torch.Tensor, labels: torch.Tensor, z0: int, z1: int, nw: int):
        from concurrent.futures import ThreadPoolExecutor
        peers = self._peer_segmenters(nw)
        counts = [0] * (z1 - z0)
        first = [z0 + w for w in range(nw)]
        # CUDA graphs are captured the first time a worker meets a frame shape; captures are not run concurrently
        # with other threads' work on the same device: a worker's first slice of a new shape is done here, serially
        shape_key = tuple(volume.shape[1:])
        for w, peer in enumerate(peers):
            gen = peer.adapter._amg().base_generator
            warm = (not gen.use_cuda_graph) or any(k[0] == shape_key for k in gen._graphs)
            if not warm and first[w] < z1:
                counts[first[w] - z0] = peer._label_slices_serial(volume, labels, first[w], first[w] + 1)[0]
                first[w] += nw
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)

        def work(w):
            torch.cuda.set_device(self.device)
            stream = self._peer_streams[w]
            with torch.inference_mode(), torch.cuda.stream(stream):
                stream.wait_event(ready)
                for ii in range(first[w], z1, nw):
                    counts[ii - z0] = peers[w]._label_slices_serial(volume, labels, ii, ii + 1)[0]
                done = torch.cuda.Event()
                done.record(stream)
            return done

        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=8, thread_name_prefix="saber_b200_slice")
        for fut in [self._pool.submit(work, w) for w in range(nw)]:
            cur.wait_event(fut.result())
        return counts

    def _label_slices_serial(self, volume: torch.Tensor, labels: torch.Tensor, z0: int, z1: int):
        W = volume.shape[2]
        counts = []
        for ii in range(z0, z1):
            if self.classifier is not None:
                # expert classifier configured: the reference's per-slice list path (segment_image -> _apply_classifier,
                # REF segmenters/base.py:159-176); the class logic is host list code fed by the device classifier
                masks = self.segment_image(volume[ii], display=False)
                lab = np.zeros(tuple(volume.shape[1:]), dtype=np.uint16)
                for idx, m in enumerate(masks):
                    lab[np.asarray(m["segmentation"], dtype=bool)] = idx + 1
                labels[ii].copy_(torch.from_numpy(lab.view(np.int16)).to(labels.device))
                counts.append(len(masks))
                continue
            dm, recs = self.segment_image_device(volume[ii])
            if len(recs) == 0:
                labels[ii].zero_()
                counts.append(0)
                continue
            order = torch.tensor([r["index"] for r in recs], dtype=_I32, device=volume.device)
            ops.stitch_labels(dm.bits, order, len(recs), W, out=labels[ii])
            counts.append(len(recs))
        return counts

    @torch.inference_mode()
    def slice_by_slice_device(self, volume: torch.Tensor) -> torch.Tensor:
        """CUDA (Z,Y,X) fp32 volume -> CUDA int32 (uint32-valued) separated label volume."""
        assert volume.is_cuda and volume.dim() == 3
        labels = torch.empty(volume.shape, dtype=torch.int16, device=volume.device)  # uint16 payload
        self.label_slices_device(volume, labels)
        return utils.separate_masks_device(labels)

    @torch.inference_mode()
    def slice_by_slice(self, volume, text_prompt: str = None) -> np.ndarray:
        """REF :164-189. ``volume``: host (Z,Y,X) array (numpy, or a — preferably pinned — CPU tensor). One H2D copy
        of the volume, one D2H copy of the uint32 label volume."""
        if isinstance(volume, np.ndarray):
            volume = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32))
        vol = volume.to(self.device, dtype=torch.float32, non_blocking=True)
        return self.slice_by_slice_device(vol).cpu().numpy().view(np.uint32)

    slice_by_slice_host = slice_by_slice

    # ------------------------------------------------------------------
    def segment(self, volume, ini_depth, nframes=None, target_class=1, text_prompt=None, display=False):
        self.ini_depth = ini_depth
        self.nframes = nframes
        self.target_class = target_class
        self.display = display
        if self.target_class > 0 or self.classifier is None:
            return self.single_segment(volume, text_prompt=text_prompt)
        return self.multiclass_segment(volume)

    def segment_3d(self, vol, masks, ann_frame_idx: int = None):
        if not self._vol_loaded:
            self.video_predictor.set_volume(vol)
            self._vol_loaded = True
        self.masks = masks
        nx = vol.shape[0]
        ny, nz = self.masks[0].shape[0], self.masks[0].shape[1]
        self.ann_frame_idx = ann_frame_idx if ann_frame_idx is not None else nx // 2
        return self.propagate((nx, ny, nz))

    def single_segment(self, volume, text_prompt=None):
        final_masks = np.zeros(volume.shape, dtype=np.uint16)
        for ii in range(2, volume.shape[0], self.ini_depth):
            masks = self.segment_image(volume[ii], display=False, target_class=self.target_class,
                                       text_prompt=text_prompt)
            if len(masks) == 0:
                continue
            masks3d = self.segment_3d(volume, [m["segmentation"] for m in masks], ann_frame_idx=ii)
            if self.target_class > 0:
                masks3d = (masks3d > 0).astype(np.uint8)
            np.maximum(final_masks, masks3d, out=final_masks)
        return utils.separate_masks(final_masks, device=self.device)

    @torch.inference_mode()
    def multiclass_segment(self, volume):
        """REF :121-161, quirks included: the seed image is ``prepare``d here and again inside the adapter,
        and ``segment_image_2d`` is called with a ``target_class`` keyword that ``SAM2Adapter.segment_image_2d`` does
        not take (REF adapters/sam2/predictor.py:49-54) — with the SAM2 adapter the call raises TypeError exactly as
        the reference does; adapters that accept the keyword run the loop below. Every voxel keeps the class of the
        most confident propagated object."""
        from ..utils import preprocessing
        final_masks = np.zeros(volume.shape, dtype=np.uint16)
        max_confidence = np.zeros(volume.shape, dtype=np.float32)
        for ii in range(2, volume.shape[0], self.ini_depth):
            im = preprocessing.prepare(volume[ii], to_rgb=True)
            raw_masks = self.adapter.segment_image_2d(im, target_class=self.target_class)
            raw_masks = [m for m in raw_masks if m["area"] >= self.min_mask_area]
            if len(raw_masks) == 0:
                continue
            mask_arrays = np.array([m["segmentation"].astype(np.uint8) for m in raw_masks])
            predictions = self.classifier.batch_predict(im[:, :, 0], mask_arrays, self.batchsize)
            predicted_classes = np.argmax(predictions, axis=1)
            valid = predicted_classes > 0
            if not np.any(valid):
                continue
            mask_list = [raw_masks[i]["segmentation"] for i, ok in enumerate(valid) if ok]
            masks3d = self.segment_3d(volume, mask_list, ann_frame_idx=ii)
            for idx, (probs, class_id) in enumerate(zip(predictions[valid], predicted_classes[valid])):
                region = masks3d == (idx + 1)
                if np.any(region):
                    confidence = probs[class_id]
                    update = region & (confidence > max_confidence)
                    final_masks[update] = class_id
                    max_confidence[update] = confidence
        return final_masks
