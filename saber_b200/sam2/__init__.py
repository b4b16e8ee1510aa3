"""`sam2`-API-compatible backend on the B200 kernels (the seam SABER binds).

``install_as_sam2()`` registers this package under the module name ``sam2`` so that the reference's
``from sam2.build_sam import build_sam2`` etc. resolve here without touching SABER's source.
"""
import importlib
import sys

_SUBMODULES = ("build_sam", "automatic_mask_generator", "sam2_image_predictor", "sam2_video_predictor")


def install_as_sam2(force: bool = False) -> None:
    if "sam2" in sys.modules and not force and sys.modules["sam2"].__name__ != __name__:
        raise RuntimeError("a different `sam2` module is already imported; pass force=True to replace it")
    me = sys.modules[__name__]
    sys.modules["sam2"] = me
    for sub in _SUBMODULES:
        sys.modules[f"sam2.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
