"""B200 twin of REF saber/entry_points/inference_core.py:10-93 (`segment_tomogram_core`) — SURVEY §8f row 1: the stage
immediately after the hot path on the batch CLI. Reader -> `segmenter.segment(...)` -> `fast_3d_gaussian_smoothing`
(scale 0.05, per-label adaptive sigma; `saber_b200.filters.masks`, CUDA kernels `label_equals` / `corr1d_zero` /
`threshold_label`) -> uint8 -> `writers.segmentation(run, mask, 'saber', name=..., session_id=..., voxel_size=...)`.

The reference imports copick_utils' readers / writers at module level (REF :5); copick is an I/O dependency outside this
path's scope, so here they are resolved lazily — `reader` / `writer` arguments take any object with the
same two callables (the tests inject in-memory stand-ins), and without them copick_utils must be importable.
The label volume stays on the device between the segmenter and the smoothing when the segmenter offers a resident
variant (`segment_device`); exactly one device->host copy (the uint8 result the writer needs) is made.
"""
from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np
import torch

from ..filters import masks as mask_filters


def _copick_io():
    try:
        from copick_utils.io import readers, writers  # noqa: WPS433 (optional dependency of the CLI layer)
    except Exception as e:  # pragma: no cover - exercised only where copick is installed
        raise RuntimeError("segment_tomogram_core needs copick_utils (readers / writers) or explicit `reader` / `writer` "
                           f"arguments: {e}") from e
    return readers, writers


def segment_tomogram_core(run, voxel_size: float, tomogram_algorithm: str, segmentation_name: str,
                          segmentation_session_id: str, slab_thickness: int, num_slabs: int, delta_z: int,
                          display_segmentation: bool, segmenter, gpu_id: int = 0, target_class: int = 1,
                          reader: Optional[Any] = None, writer: Optional[Any] = None):
    """Same arguments, order of operations and return values as REF :10-93 (None in every branch)."""
    logger = logging.getLogger(__name__)
    if reader is None or writer is None:
        readers, writers = _copick_io()
        reader = reader or readers
        writer = writer or writers
    vol = reader.tomogram(run, voxel_size, algorithm=tomogram_algorithm)
    if vol is None:
        logger.info(f"No Tomogram Found for {run.name}")
        return None
    torch.cuda.set_device(gpu_id)
    img_name = run.name + "-" + segmentation_session_id
    if num_slabs > 1:
        segment_mask = segmenter.segment(vol, slab_thickness, num_slabs, delta_z, img_name, display_segmentation)
    else:
        segment_mask = segmenter.segment(vol, slab_thickness, target_class=target_class, save_run=img_name,
                                         display=display_segmentation)
    if segment_mask is None:
        logger.info(f"No Segmentation Found for {run.name}")
        return None
    if not display_segmentation and segment_mask is not None:
        segment_mask = smooth_to_uint8(segment_mask, scale=0.05, gpu_id=gpu_id)
        writer.segmentation(run, segment_mask, "saber", name=segmentation_name, session_id=segmentation_session_id,
                            voxel_size=float(voxel_size))
        logger.info(f"Saved Segmentation for {run.name} as {segmentation_name}")
    del vol
    del segment_mask
    torch.cuda.empty_cache()
    segmenter.inference_state = None
    return


def smooth_to_uint8(segment_mask, scale: float = 0.05, gpu_id: int = 0) -> np.ndarray:
    """REF :68-73: adaptive Gaussian smoothing of the label volume, then `.astype(np.uint8)`. Host or CUDA label volume
    in; the smoothing runs on `cuda:gpu_id` and the uint8 result comes back in one copy."""
    dev = torch.device(f"cuda:{gpu_id}")
    if isinstance(segment_mask, torch.Tensor):
        vol = segment_mask.to(dev)
    else:
        arr = np.ascontiguousarray(segment_mask)
        if arr.dtype == np.uint16:
            arr = arr.view(np.int16)  # torch's 16-bit integer kernels: same bits (labels < 2^15 on this path)
        elif arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        elif arr.dtype in (np.int64, np.uint64):
            arr = arr.astype(np.int32)
        elif arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        vol = torch.from_numpy(arr).to(dev)
    out = mask_filters.fast_3d_gaussian_smoothing(vol, scale=scale, deviceID=gpu_id)
    return out.cpu().numpy().astype(np.uint8)


def segment_micrograph_core(input: str, output: str, scale_factor: float, target_resolution: float, display_image: bool,
                            use_sliding_window: bool, gpu_id, models, read_micrograph=None):
    """REF saber/entry_points/inference_core.py:97-153 — the 2-D worker of `prep2d` / micrograph inference (SURVEY §8f rows
    2 and 4 meet here): read -> optional Fourier-crop down-sampling (`FourierRescale2D`, csrc/fft.cu) -> `segmenter.segment`
    -> labelled candidate stack -> one zarr group named after the file (pixel size stored in nanometres).
    `read_micrograph(path) -> (array, pixel size in Angstroms or None)` is I/O (mrc / tiff / dm4) and is injected; without
    it `saber.utils.io.read_micrograph` must be importable."""
    import os

    from ..filters.downsample import FourierRescale2D
    from ..utils import zarr_writer
    segmenter = models["segmenter"]
    zwriter = zarr_writer.get_zarr_writer(output)
    zwriter.set_dict_attr("amg", segmenter.adapter_cfg.amg_cfg.to_dict())
    torch.cuda.set_device(gpu_id)
    if read_micrograph is None:
        from saber.utils.io import read_micrograph  # noqa: WPS433 (the reference's file readers; optional)
    image, pixel_size = read_micrograph(input)
    image = image.astype(np.float32)
    if target_resolution is not None and target_resolution > pixel_size:
        image = FourierRescale2D.run(image, target_resolution / pixel_size)
    elif scale_factor is not None:
        image = FourierRescale2D.run(image, scale_factor)
    target_class = models.get("target_class", -1)
    segmenter.segment(image, target_class=target_class, display=False, use_sliding_window=use_sliding_window)
    if isinstance(pixel_size, np.ndarray):
        pixel_size = pixel_size.item()
    masks = mask_filters.masks_to_array(segmenter.masks)
    pixel_size = pixel_size / 10 if pixel_size is not None else 1
    out_image = segmenter.image
    if out_image.ndim == 3:
        out_image = out_image[:, :, 0]
    zwriter.write(run_name=os.path.splitext(os.path.basename(input))[0], image=out_image, masks=masks, pixel_size=pixel_size)
