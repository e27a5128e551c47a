"""The C-ABI library builds for sm_100a on a CPU box, loads, and exports every
symbol include/splashy_cuda.h declares.  No compute is attempted here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "splashy_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_tables_agree(built_lib):
    from splashy_b200 import _native
    assert declared_symbols() == sorted(_native.SYMBOLS)


def test_every_declared_symbol_is_exported(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), name
    assert built_lib.spl_abi_version() == 1


def test_library_is_sm_100a_only():
    import subprocess
    from splashy_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", str(_native.LIB_PATH)],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_desc_layout_matches_header_defaults(built_lib):
    from splashy_b200 import _native
    d = _native.SplDesc()
    built_lib.spl_desc_default(C.byref(d))
    assert (d.h, d.rho) == (10.0, 999.97)            # Constants.fs:11,15
    assert list(d.gravity) == [0.0, -9.81, 0.0]      # Constants.fs:26
    assert d.world == 1 and d.precond == _native.SPL_PC_JACOBI
    assert d.tol == 1e-6


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly."""
    if built_lib.spl_device_count() > 0:
        pytest.skip("a GPU is present")
    from splashy_b200 import Solver, SplashyError
    with pytest.raises(SplashyError) as e:
        Solver(4, 4, 4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The product path (the package, its CUDA / C++ sources, the C ABI header) must not
    import, include, link or dlopen anything under oracle/."""
    import re
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|dlopen[^\n]*oracle|"
                     r"#include[^\n]*oracle|liboracle|oracle/_ref|oracle/_build")
    files = list((ROOT / "splashy_b200").rglob("*")) + list((ROOT / "include").rglob("*"))
    checked = 0
    for path in files:
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ""} and path.is_file():
            if path.name in {"Makefile"} or path.suffix:
                m = bad.search(path.read_text(errors="ignore"))
                assert m is None, (path, m.group(0))
                checked += 1
    assert checked >= 15
