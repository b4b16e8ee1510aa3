"""Training-data preparation from tomograms (`prep3d`) — the caller after the slab segmenter, same behaviour as
REF saber/classifier/preprocess/tomo_prep.py:13-85. The click / SLURM command wrappers are CLI and stay
out of scope; `extract_sam2_candidates` is the per-run worker GPUPool executes.

copick is I/O outside the path: the tomogram reader is injected (`reader`, any object with ``tomogram(run, voxel_size,
algorithm)``) or imported lazily from copick_utils.
"""
from __future__ import annotations

import numpy as np

from ...filters import masks as mask_filters
from ...utils import zarr_writer


def segment(segmenter, vol, slab_thickness, zSlice):
    """REF tomo_prep.py:13-26: SAM2 candidates of one slab, sorted by area (ascending), as a labelled stack."""
    segmenter.segment_slab(vol, slab_thickness, display=False, zSlice=zSlice)
    image0, masks_list = segmenter.image0, segmenter.masks
    masks_list = sorted(masks_list, key=lambda mask: mask["area"], reverse=False)
    return image0, mask_filters.masks_to_array(masks_list)


def extract_sam2_candidates(run, output, voxel_size, tomogram_algorithm: str, slab_thickness: int, multiple_slabs: int,
                            gpu_id, models, reader=None):
    """REF tomo_prep.py:28-85: one zarr group per run (or per slab: ``<run>_<i+1>``) with the slab image and the uint8
    candidate stack; the pixel size is stored in nanometres (copick voxel sizes are Angstroms)."""
    segmenter = models["segmenter"]
    zwriter = zarr_writer.get_zarr_writer(output)
    zwriter.set_dict_attr("amg", segmenter.adapter_cfg.amg_cfg.to_dict())
    if reader is None:
        from copick_utils.io import readers as reader  # noqa: WPS433 (optional dependency of the CLI layer)
    vol = reader.tomogram(run, voxel_size, tomogram_algorithm)
    if vol is None:
        print("No Tomogram Found for Run: ", run.name)
        return
    voxel_size /= 10
    if multiple_slabs > 1:
        center_index = vol.shape[0] // 2
        for i in range(multiple_slabs):
            slab_center = center_index + (i - multiple_slabs // 2) * slab_thickness
            image_seg, masks = segment(segmenter, vol, slab_thickness, zSlice=slab_center)
            zwriter.write(run_name=f"{run.name}_{i + 1}", image=image_seg, masks=masks.astype(np.uint8), pixel_size=voxel_size)
    else:
        image_seg, masks = segment(segmenter, vol, slab_thickness, zSlice=int(vol.shape[0] // 2))
        zwriter.write(run_name=run.name, image=image_seg, masks=masks.astype(np.uint8), pixel_size=voxel_size)
