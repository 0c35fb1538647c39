"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the product package never touches the oracle; host-side logic (shapes, roofline bytes, TP plan)."""
import os
import re
import subprocess

import pytest

import simplellminference_b200 as pkg
from simplellminference_b200 import _lib
from simplellminference_b200.config import PRESETS, BF16, F32, INT8

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sllm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sllm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()   # also asserts every SIGNATURES entry resolves
    syms = declared_symbols()
    assert len(syms) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sllm_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in include/sllm_b200.h but not exported: {missing}"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert lib.sllm_abi_version() == 1


def test_library_is_sm100a_and_uses_bulk_copy():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_no_cpu_fallback_message_without_library(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(FileNotFoundError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "simplellminference_b200")):
        if any(part in dirpath for part in ("build", "lib", "__pycache__")):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\bfrom\s+oracle\b|\bimport\s+oracle\b|oracle/_ref|liboracle|libref_oracle|llama_oracle\.h", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_roofline_bytes_match_survey():
    # SURVEY.md Appendix C / BASELINE.md §4
    assert abs(PRESETS["llama2-7b"].step_bytes(511, BF16, BF16) / 1e9 - 13.484) < 2e-3
    assert abs(PRESETS["stories110M"].step_bytes(255, F32, F32) / 1e9 - 0.457) < 2e-3
    assert abs(PRESETS["stories15M"].step_bytes(127, F32, F32) / 1e9 - 0.0626) < 2e-4
    assert abs(PRESETS["tinyllama-1.1b"].step_bytes(2047, BF16, BF16) / 1e9 - 2.115) < 2e-3
    assert abs(PRESETS["tinyllama-1.1b"].step_bytes(2047, INT8, BF16, 64) / 1e9 - 1.146) < 2e-3
    assert abs(PRESETS["llama2-7b"].n_params() / 1e6 - 6607.3) < 0.1


def test_shape_validation():
    with pytest.raises(ValueError):
        pkg.ModelShape(100, 16, 60, 60, 64, 8, 1, 4, 4)   # heads*head_dim != hidden
