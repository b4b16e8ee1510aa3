"""Training-data store written by the prep3d / prep2d callers — same layout and surface as
REF saber/utils/zarr_writer.py:22-231: a zarr-v2 directory store with '/' as the dimension separator,
one group per run holding the image as dataset ``0`` and the candidate masks as ``labels/0``, OME-style ``multiscales``
attributes on both groups (voxel size in nanometres), root attributes for the AMG configuration and the final
``total_runs`` / ``creation_complete`` marks.

The reference goes through the ``zarr`` package (absent from this image, and pure I/O next to the path), so the few files
a v2 store consists of (``.zgroup``, ``.zattrs``, ``.zarray`` and chunk files) are written directly. The compressor is the
reference's Blosc-zstd (clevel 2, bit-shuffle) when ``numcodecs`` is importable and zlib level 2 otherwise — both are
declared in ``.zarray``, so any zarr reader opens the store either way. Thread-safe like the reference (one lock around
the run counter and the attribute files; every run writes its own directory).
"""
from __future__ import annotations

import json
import os
import threading
import zlib
from typing import Any, Dict, Mapping, Optional

import numpy as np


def _to_jsonable(obj):
    """REF zarr_writer.py:8-20."""
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (list, tuple)):
        return [_to_jsonable(x) for x in obj]
    if isinstance(obj, Mapping):
        return {str(k): _to_jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (bool, int, float, str)) or obj is None:
        return obj
    return str(obj)


def _compressor():
    try:
        import numcodecs  # noqa: WPS433 (optional)
        codec = numcodecs.Blosc(cname="zstd", clevel=2, shuffle=2)
        return codec.get_config(), codec.encode
    except Exception:
        return {"id": "zlib", "level": 2}, lambda buf: zlib.compress(bytes(buf), 2)


def _write_json(path: str, obj) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(obj, f, indent=4, sort_keys=True)
    os.replace(tmp, path)


class _Group:
    """A zarr-v2 group directory: `.zgroup`, `.zattrs`, child groups and datasets."""

    def __init__(self, path: str, lock: threading.Lock):
        self.path, self._lock = path, lock
        os.makedirs(path, exist_ok=True)
        _write_json(os.path.join(path, ".zgroup"), {"zarr_format": 2})

    # -- attributes
    def _attrs_path(self):
        return os.path.join(self.path, ".zattrs")

    def get_attrs(self) -> dict:
        p = self._attrs_path()
        if not os.path.exists(p):
            return {}
        with open(p) as f:
            return json.load(f)

    def update_attrs(self, items: Mapping[str, Any]) -> None:
        with self._lock:
            attrs = self.get_attrs()
            attrs.update(_to_jsonable(dict(items)))
            _write_json(self._attrs_path(), attrs)

    # -- children
    def create_group(self, name: str) -> "_Group":
        p = os.path.join(self.path, name)
        if os.path.exists(os.path.join(p, ".zgroup")):
            raise ValueError(f"path {name!r} contains a group")  # zarr's ContainsGroupError
        return _Group(p, self._lock)

    def create_dataset(self, name: str, data: np.ndarray) -> None:
        data = np.ascontiguousarray(data)
        p = os.path.join(self.path, name)
        os.makedirs(p, exist_ok=True)
        # one chunk per leading index for stacks (a mask / a slice at a time), the whole array otherwise
        chunks = (1, *data.shape[1:]) if data.ndim >= 3 else tuple(data.shape)
        config, encode = _compressor()
        _write_json(os.path.join(p, ".zarray"), {
            "zarr_format": 2, "shape": list(data.shape), "chunks": list(chunks), "dtype": data.dtype.str, "fill_value": 0,
            "order": "C", "filters": None, "compressor": config, "dimension_separator": "/"})
        if data.ndim >= 3:
            for i in range(data.shape[0]):
                d = os.path.join(p, str(i), *["0"] * (data.ndim - 2))
                os.makedirs(d, exist_ok=True)
                with open(os.path.join(d, "0"), "wb") as f:
                    f.write(encode(data[i].tobytes()))
        elif data.size:
            d = os.path.join(p, *["0"] * (data.ndim - 1)) if data.ndim > 1 else p
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "0"), "wb") as f:
                f.write(encode(data.tobytes()))


class ParallelZarrWriter:
    """REF zarr_writer.py:26-171 (same methods, arguments and return values)."""

    def __init__(self, zarr_path: str):
        self.zarr_path = zarr_path
        self._lock = threading.Lock()
        self._attr_lock = threading.Lock()
        if os.path.isdir(zarr_path):  # mode='w': start from an empty store
            import shutil
            shutil.rmtree(zarr_path)
        self.zroot = _Group(zarr_path, self._attr_lock)
        self._run_counter = 0
        print(f"Initialized zarr store at: {zarr_path}")

    def set_dict_attr(self, key: str, data: Mapping[str, Any], *, merge_missing: bool = False) -> None:
        safe = _to_jsonable(dict(data))
        with self._lock:
            if merge_missing:
                existing = self.zroot.get_attrs().get(key)
                if isinstance(existing, dict):
                    merged = dict(existing)
                    missing = {k: v for k, v in safe.items() if k not in merged}
                    if missing:
                        merged.update(missing)
                        self.zroot.update_attrs({key: merged})
                    return
            self.zroot.update_attrs({key: safe})

    def get_next_run_index(self) -> int:
        with self._lock:
            run_index = self._run_counter
            self._run_counter += 1
            return run_index

    def write(self, run_name: str, image: np.ndarray, masks: np.ndarray, pixel_size: Optional[float] = None,
              metadata: Optional[Dict[str, Any]] = None) -> int:
        if pixel_size is None:
            pixel_size = 1.0
        run_index = self.get_next_run_index()
        try:
            run_group = self.zroot.create_group(run_name)
            if metadata:
                run_group.update_attrs(metadata)
            run_group.create_dataset("0", np.asarray(image))
            add_attributes(run_group, pixel_size)
            labels_group = run_group.create_group("labels")
            labels_group.create_dataset("0", np.asarray(masks))
            add_attributes(labels_group, pixel_size, True)
            return run_index
        except Exception as e:
            print(f"Error writing {run_name} to zarr: {e}")
            raise

    def finalize(self):
        self.zroot.update_attrs({"total_runs": self._run_counter, "creation_complete": True})
        print(f"Zarr file finalized with {self._run_counter} runs")


_zarr_writer = None
_writer_lock = threading.Lock()


def get_zarr_writer(zarr_path: str) -> ParallelZarrWriter:
    """The process-wide writer (REF zarr_writer.py:173-180: the first path wins)."""
    global _zarr_writer
    with _writer_lock:
        if _zarr_writer is None:
            _zarr_writer = ParallelZarrWriter(zarr_path)
        return _zarr_writer


def add_attributes(zarr_group, voxel_size: float = 1.0, is_3d: bool = False, voxel_size_z: float = 1.0) -> None:
    """REF zarr_writer.py:182-231: `multiscales` with nanometre axes and one scale transformation for dataset "0"."""
    names = ["z", "y", "x"] if is_3d else ["y", "x"]
    axes = [{"name": n, "type": "space", "unit": "nanometer"} for n in names]
    scale = [voxel_size_z, voxel_size, voxel_size] if is_3d else [voxel_size, voxel_size]
    zarr_group.update_attrs({"multiscales": [{
        "axes": axes,
        "datasets": [{"coordinateTransformations": [{"scale": scale, "type": "scale"}], "path": "0"}],
        "name": "/", "version": "0.4"}]})
