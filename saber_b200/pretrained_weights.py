"""Checkpoint resolution for the SAM2.1 weights — the role of REF saber/pretrained_weights.py:174-202
(``get_sam2_checkpoint``: config name -> (hydra config, ``sam2.1_hiera_*.pt`` under the package's ``checkpoints/``; the
reference downloads missing files, REF :20-65).

There is no network here, so nothing is downloaded: a checkpoint is looked up under ``$SABER_B200_CHECKPOINT_DIR``, the
package's ``checkpoints/`` directory and the reference package's own ``checkpoints/`` (when SABER is installed beside
this repo). If none is found the caller must opt in to deterministic random initialisation EXPLICITLY
(``allow_random_init=True`` / ``SABER_B200_ALLOW_RANDOM_INIT=1``: benchmarks and parity tests on synthetic weights);
otherwise model construction raises — a segmenter silently running on random weights returns plausible-looking but
meaningless masks.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

_FILES = {"tiny": "sam2.1_hiera_tiny.pt", "small": "sam2.1_hiera_small.pt", "base_plus": "sam2.1_hiera_base_plus.pt",
          "large": "sam2.1_hiera_large.pt"}
_ALIASES = {"base": "base_plus", "base+": "base_plus", "b+": "base_plus", "t": "tiny", "s": "small", "l": "large"}


def _dirs() -> List[str]:
    out = []
    env = os.environ.get("SABER_B200_CHECKPOINT_DIR")
    if env:
        out.append(env)
    out.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "checkpoints"))
    try:
        import saber  # the reference package, if it is installed next to this one
        out.append(os.path.join(os.path.dirname(saber.__file__), "checkpoints"))
    except Exception:
        pass
    return out


def find_sam2_checkpoint(arch_name: str) -> Optional[str]:
    """Path of the SAM2.1 checkpoint of ``arch_name`` (tiny / small / base_plus / large), or None."""
    name = _ALIASES.get(arch_name, arch_name)
    fn = _FILES.get(name)
    if fn is None:
        raise ValueError(f"Invalid SAM2 Model Config: {arch_name}")
    for d in _dirs():
        p = os.path.join(d, fn)
        if os.path.isfile(p):
            return p
    return None


def random_init_allowed(flag: bool = False) -> bool:
    return bool(flag) or os.environ.get("SABER_B200_ALLOW_RANDOM_INIT", "") == "1"


def get_sam2_checkpoint(sam2_cfg: str) -> Tuple[str, Optional[str]]:
    """REF saber/pretrained_weights.py:174-202: (config name, checkpoint path). The path is None when no file is
    present; ``build_sam2`` then raises unless random initialisation was explicitly allowed."""
    name = _ALIASES.get(sam2_cfg, sam2_cfg)
    if name not in _FILES:
        raise ValueError(f"Invalid SAM2 Model Config: {sam2_cfg}")
    return name, find_sam2_checkpoint(name)


def missing_message(arch_name: str) -> str:
    return (f"no SAM2.1 checkpoint for '{arch_name}' ({_FILES.get(_ALIASES.get(arch_name, arch_name), '?')}) under "
            f"{_dirs()}; place the upstream file there / set SABER_B200_CHECKPOINT_DIR, pass ckpt_path=, or opt in to "
            "random initialisation explicitly (allow_random_init=True or SABER_B200_ALLOW_RANDOM_INIT=1 — synthetic "
            "benchmarks and parity tests only)")
