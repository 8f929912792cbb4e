"""`-m "not gpu"`: the CUDA C-ABI library builds, loads without a GPU and exports every symbol include/eg_b200.h declares;
compute entry points fail loudly (EG_ERR_NO_DEVICE) instead of falling back to a CPU path."""
import ctypes as C
import pathlib
import re

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _lib():
    from elastic_elgamal_b200 import _ffi, build
    return _ffi, _ffi.load(build.build())


def test_exports_every_declared_symbol():
    _ffi, lib = _lib()
    header = (ROOT / "include" / "eg_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # drop comments
    names = set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", header))
    assert len(names) > 30
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes mirror covers the same set
    assert names <= set(_ffi.PROTOTYPES), sorted(names - set(_ffi.PROTOTYPES))


def test_rust_sys_crate_declares_every_symbol():
    """rust/elastic-elgamal-b200-sys is source-only here (no rustc): keep its extern block in step with the header."""
    header = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "eg_b200.h").read_text(), flags=re.S)
    names = set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", header))
    rust = (ROOT / "rust" / "elastic-elgamal-b200-sys" / "src" / "lib.rs").read_text()
    declared = set(re.findall(r"pub fn (eg_[a-z0-9_]+)\(", rust))
    assert names <= declared, sorted(names - declared)


def test_headers_compile_as_c99_and_cxx17(tmp_path):
    """include/eg_b200.h is plain C (what cgo / bindgen / ctypes consume); the .hpp mirror is header-only C++17."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "eg_b200.h"\nint main(void) { return eg_version() == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-fsyntax-only", str(c)], check=True)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "elastic_elgamal_b200.hpp"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), "-fsyntax-only", str(cpp)], check=True)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    _ffi, lib = _lib()
    h = C.c_void_p()
    assert lib.eg_ctx_create(0, C.byref(h)) == _ffi.ERR_NO_DEVICE and not h.value
    assert lib.eg_version().startswith(b"eg_b200")


def test_hostsim_harness_is_not_loadable_as_the_product(monkeypatch):
    """The CPU-compiled test harness identifies itself and the package refuses it unless a test names it explicitly."""
    import pytest
    from hostsim.build_hostsim import build as build_hostsim
    from elastic_elgamal_b200 import _ffi
    hs = build_hostsim()
    assert b"HOSTSIM" in _ffi.load(hs).eg_version()                 # explicit path: allowed (tests only)
    monkeypatch.setenv("EG_B200_LIB", str(hs))
    with pytest.raises(RuntimeError):
        _ffi.load()
    monkeypatch.delenv("EG_B200_LIB")
    assert b"sm_100a" in _ffi.load().eg_version()


def test_package_never_imports_the_oracle():
    for f in (ROOT / "elastic_elgamal_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".inc", ".h", ".hpp") and f.is_file():
            text = f.read_text()
            assert "import oracle" not in text and "liboracle" not in text and "eg_oracle.h" not in text, f
