"""CPU: the C-ABI library builds/loads and exports every symbol include/emu_b200.h declares; argument validation
that needs no GPU behaves (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "emu_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from emu_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_lists_every_symbol():
    from emu_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()


def test_version_and_counter(lib):
    assert b"sm_100a" in lib.emu_version()
    assert lib.emu_launch_count() >= 0


def test_config_struct_layout_matches_header():
    """EmuConfig in _lib.py must have the same field order/size as the C struct (all 4-byte fields)."""
    from emu_b200 import _lib
    src = open(HEADER).read()
    body = re.search(r"typedef struct EmuConfig \{(.*?)\} EmuConfig;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(int|float)\s+", "", decl)
        for n in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", n.strip()))
    assert names == [f[0] for f in _lib.EmuConfig._fields_]
    assert ctypes.sizeof(_lib.EmuConfig) == 4 * (len(names) - 1) + 4 * 8


def test_no_cpu_fallback(lib):
    """Without a GPU the engine refuses to come up (EMU_ERR_CUDA) instead of computing on the host."""
    import torch
    from emu_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.EmuConfig()
    h = ctypes.c_void_p()
    assert lib.emu_engine_create(ctypes.byref(cfg), 0, 1, None, ctypes.byref(h)) == -2
    with pytest.raises(_lib.EmuError):
        _lib.Engine(cfg)


def test_product_never_imports_oracle():
    """The package must not route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "emu_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
