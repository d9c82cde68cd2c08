"""The C-ABI library loads without a GPU and exports every symbol include/b200ani.h declares."""
import ctypes
import re

from conftest import ROOT


def declared_symbols():
    text = (ROOT / "include" / "b200ani.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ani_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from pyani_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 12
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200ani.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in pyani_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms


def test_result_struct_layout():
    from pyani_b200 import _lib
    from pyani_b200.engine import ANIM_DTYPE

    assert ctypes.sizeof(_lib.AnimResult) == 48 == ANIM_DTYPE.itemsize
    for (name, _), fname in zip(_lib.AnimResult._fields_, ANIM_DTYPE.names):
        assert name == fname
        assert getattr(_lib.AnimResult, name).offset == ANIM_DTYPE.fields[fname][1]


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run rather than fall back to the oracle."""
    import pytest
    import torch

    import pyani_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pyani_b200.EngineError):
        pyani_b200.Engine(0)


def test_product_never_imports_oracle():
    for p in (ROOT / "pyani_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, p


def test_algorithm_constants_agree_between_oracle_and_engine():
    """oracle/anim_spec.h and pyani_b200/csrc/anim_params.cuh are independent copies of one specification."""
    def defs(path):
        out = {}
        for line in path.read_text().splitlines():
            m = re.match(r"#define\s+(AN_[A-Z_]+)\s+(.+?)\s*(/\*.*)?$", line)
            if m:
                out[m.group(1)] = m.group(2).strip()
        return out

    a = defs(ROOT / "oracle" / "anim_spec.h")
    b = defs(ROOT / "pyani_b200" / "csrc" / "anim_params.cuh")
    assert a == b and len(a) >= 20
