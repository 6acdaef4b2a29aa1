"""CPU tests: the C-ABI library builds, loads and exports every symbol the header declares;
without a GPU every compute path fails loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from deadtrees_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "deadtrees_b200.h").read_text()
    return sorted(set(re.findall(r"^\s*(?:int|int64_t)\s+(dt_\w+)\s*\(", text, flags=re.M)))


def test_header_and_binding_agree(lib):
    from deadtrees_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/deadtrees_b200.h but not exported"


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md says, for every exported entry point, which reference code it replaces."""
    doc = (ROOT / "INTEGRATION.md").read_text()
    missing = [s for s in declared_symbols() if s not in doc]
    assert not missing, f"entry points without a row in INTEGRATION.md: {missing}"


def test_version(lib):
    assert lib.dt_version() == 100


def test_conv_desc_layout_matches_header():
    from deadtrees_b200._lib import ConvDesc
    text = (ROOT / "include" / "deadtrees_b200.h").read_text()
    body = text[text.index("typedef struct {"):text.index("} dt_conv_desc;")]
    fields = re.findall(r"\b([A-Za-z_]\w*)\s*(?=[,;])", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert [f for f, _ in ConvDesc._fields_] == fields
    assert ctypes.sizeof(ConvDesc) == 4 * len(fields)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from deadtrees_b200 import _lib
    from deadtrees_b200.utils.data_handling import make_blocks_vectorized
    assert lib.dt_device_check() != 0
    assert "no CPU fallback" in _lib.last_error()
    with pytest.raises(_lib.DeadtreesB200Error):
        make_blocks_vectorized(np.zeros((3, 4, 4), dtype=np.uint8), 2)
    with pytest.raises(_lib.DeadtreesB200Error):
        _lib.require_device()


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under deadtrees_b200/ or scripts/ may reference it."""
    for p in list((ROOT / "deadtrees_b200").rglob("*.py")) + list((ROOT / "deadtrees").rglob("*.py")) + \
            list((ROOT / "scripts").rglob("*.py")):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
