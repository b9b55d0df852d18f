"""The C-ABI library loads and exports every symbol include/dorylus_b200.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

from dorylus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dorylus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dory_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.dory_abi_version() == _lib.DORY_ABI_VERSION


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.DoryChunk) == 32  # 7 x u32 + u8, padded to 4
    assert ctypes.sizeof(_lib.DoryConfig) == 4 * (3 + 9 + 2) + 4 + 4 + 4
    assert ctypes.sizeof(_lib.DoryStats) == 32


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dorylus_b200.engine import DoryError, Engine

    with pytest.raises(DoryError) as ei:
        Engine([8, 4, 2])
    assert ei.value.code == _lib.ENODEV and "no CPU fallback" in str(ei.value)


def test_product_path_never_touches_the_oracle():
    """No file of the shipped package may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "dorylus_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libdoryref" not in text, f


def test_product_path_never_touches_the_host_check_library():
    """tests/hostcheck/ (the engine object on an emulated CUDA runtime) is test infrastructure: no
    shipped file -- package, host driver, header, bench.py, __graft_entry__.py -- may name it."""
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for top in ("dorylus_b200", "host", "include", "tools"):
        for dirpath, _, names in os.walk(os.path.join(ROOT, top)):
            if "_obj" in dirpath or "__pycache__" in dirpath:
                continue
            files += [os.path.join(dirpath, n) for n in names if n.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".sh"))]
    for f in files:
        assert "hostcheck" not in open(f).read(), f
