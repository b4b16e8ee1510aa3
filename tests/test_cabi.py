"""The C-ABI library builds, loads without a GPU and exports exactly what include/saber_b200.h declares."""
import ctypes
import os
import re

from saber_b200 import lib as sblib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "saber_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_every_bound_symbol():
    hdr = set(_header_symbols())
    bound = set(sblib.SIGNATURES) | {"sb_last_error"}
    assert hdr == bound, (sorted(hdr - bound), sorted(bound - hdr))


def test_library_loads_and_exports_header_symbols():
    L = sblib.load()  # builds with nvcc if missing; raises on failure
    for name in _header_symbols():
        assert hasattr(L, name), f"{name} declared in include/saber_b200.h but not exported"
    assert L.sb_version() >= 100
    assert isinstance(sblib.last_error(), str)


def test_header_arity_matches_ctypes_signatures():
    text = open(os.path.join(ROOT, "include", "saber_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, argtypes in sblib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(argtypes), (name, n, len(argtypes))


def test_product_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from saber_b200 import ops
    with pytest.raises(RuntimeError):
        ops.require_b200()
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from saber_b200.segmenters.propagation import propagationSegmenter
    from saber_b200.adapters.base import cfgAMG
    with pytest.raises(RuntimeError):
        propagationSegmenter(amg_cfg=cfgAMG(sam2_cfg="tiny"))


def test_missing_checkpoint_is_an_error_unless_random_init_is_explicit(monkeypatch, tmp_path):
    """a segmenter must not silently run on random weights. Without a checkpoint file construction raises;
    the explicit opt-ins (argument / env var / state_dict) reach the weight initialiser."""
    import pytest
    from saber_b200 import pretrained_weights as pw
    from saber_b200.sam2 import build_sam
    monkeypatch.delenv("SABER_B200_ALLOW_RANDOM_INIT", raising=False)
    monkeypatch.setenv("SABER_B200_CHECKPOINT_DIR", str(tmp_path))
    assert pw.find_sam2_checkpoint("tiny") is None and pw.get_sam2_checkpoint("base") == ("base_plus", None)
    with pytest.raises(FileNotFoundError, match="allow_random_init"):
        build_sam._load_state_dict("tiny", None, 0)
    sd = build_sam._load_state_dict("tiny", None, 0, allow_random_init=True)
    assert "image_encoder.trunk.patch_embed.proj.weight" in sd
    monkeypatch.setenv("SABER_B200_ALLOW_RANDOM_INIT", "1")
    assert pw.random_init_allowed() and build_sam._load_state_dict("tiny", None, 0).keys() == sd.keys()
    # a checkpoint file under the directory is found by name (REF saber/pretrained_weights.py:183-188 file names)
    import torch
    torch.save({"model": {"x": torch.zeros(1)}}, tmp_path / "sam2.1_hiera_tiny.pt")
    monkeypatch.delenv("SABER_B200_ALLOW_RANDOM_INIT")
    assert pw.find_sam2_checkpoint("tiny") == str(tmp_path / "sam2.1_hiera_tiny.pt")
    assert list(build_sam._load_state_dict("tiny", None, 0)) == ["x"]
    with pytest.raises(ValueError):
        pw.get_sam2_checkpoint("giant")


def test_key_split_heuristics_are_host_only_and_consistent():
    """sb_t2i_*_splits are pure host arithmetic (callable without a GPU): the split count must divide the key count into
    whole 64-key tiles (tcgen05 kernel) / 32-key tiles (mma.sync kernel) for every batch size the decoder uses."""
    from saber_b200 import lib

    L = lib.load()
    for batch in (1, 8, 32, 64, 100, 192, 384):
        for nk in (256, 1024, 4096):
            ns = L.sb_t2i_tc_splits(batch, nk)
            assert ns >= 1 and nk % (ns * 64) == 0, (batch, nk, ns)
            ns2 = L.sb_t2i_fold_splits(batch, nk)
            assert ns2 >= 1 and nk % (ns2 * 32) == 0, (batch, nk, ns2)
            assert ns in (1, 2, 4, 8, 16) and ns2 in (1, 2, 4, 8)
