"""CPU tests: the C-ABI library builds/loads here and exports every symbol include/b200pt.h declares; no compute."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "b200pt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from multimodal_llm_pretraining_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_loads_and_exports_everything():
    from multimodal_llm_pretraining_b200 import _lib
    from multimodal_llm_pretraining_b200.csrc import build

    build.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.b200_abi_version() == _lib.ABI_VERSION


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_llm_pretraining_b200 import _lib

    with pytest.raises(_lib.B200Error):
        _lib.lib_for(0)
