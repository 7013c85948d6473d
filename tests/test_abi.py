"""CPU: the C-ABI library builds, loads and exports every symbol include/radialog_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from radialog_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "radialog_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.lib_path())
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_header():
    names = set(_declared_symbols())
    bound = set(_lib.SIGNATURES)
    assert names <= bound, f"header symbols without a ctypes signature: {sorted(names - bound)}"


def test_version_and_error_string(lib):
    assert lib.rd_version() >= 100
    assert isinstance(lib.rd_last_error(), bytes)


def test_invalid_arguments_fail_loudly(lib):
    # argument validation happens before any CUDA call, so it is testable without a GPU
    st = lib.rd_linear(None, 8, None, 8, None, 8, 0, 0, 0, None, 0, 0, None, 0, None)
    assert st != 0 and b"bad shape" in lib.rd_last_error()
    st = lib.rd_linear(None, 8, None, 8, None, 8, 4, 4, 12, None, 0, 0, None, 0, None)
    assert st != 0 and b"multiples of 8" in lib.rd_last_error()
    st = lib.rd_attention(None, 0, None, None, None, None, None, 1, 1, 1, 64, 16, 0, None)
    assert st != 0 and b"head_dim" in lib.rd_last_error()
    assert lib.rd_llm_create(None, None) != 0


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped SASS (UTCHMMA, UTMALDG, LDTM), not a legacy mma.sync path."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA.16816" not in sass


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle (test infrastructure) anywhere."""
    pkg = os.path.join(ROOT, "radialog_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} imports the oracle"
