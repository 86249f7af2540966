"""The C-ABI library loads on a CPU-only box and exports every symbol include/clann_b200.h declares; entry points fail
loudly (never fall back) when there is no device or the arguments are bad. No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "clann_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:clann|CPUFFINN)_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 27, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from clann_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_legacy_symbols_match_reference_header():
    """Same eight names as libpuffinn-ffi/c_binder.h:14-26."""
    legacy = [n for n in declared_functions() if n.startswith("CPUFFINN_")]
    assert legacy == sorted([
        "CPUFFINN_load_from_file", "CPUFFINN_index_create", "CPUFFINN_index_rebuild", "CPUFFINN_index_insert_cosine",
        "CPUFFINN_search_cosine", "CPUFFINN_get_distance_computations", "CPUFFINN_clear_distance_computations",
        "CPUFFINN_save_index"])


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_errors_without_fallback(lib):
    from clann_b200 import _lib
    h = C.c_void_p()
    cfg = _lib.ClannConfig(4, 1.0, 3, 0.9)
    data = np.zeros((10, 4), np.float32)
    # empty dataset -> DataError (index.rs:72-74), before any device work
    assert lib.clann_init_with_config(data.ctypes.data, 0, 4, C.byref(cfg), C.byref(h)) == _lib.ERR_DATA
    assert b"empty dataset" in lib.clann_last_error()
    assert lib.clann_init_with_config(data.ctypes.data, 10, 4, None, C.byref(h)) == _lib.ERR_ARG
    bad = _lib.ClannConfig(0, 1.0, 3, 0.9)
    assert lib.clann_init_with_config(data.ctypes.data, 10, 4, C.byref(bad), C.byref(h)) == _lib.ERR_CONFIG
    if not _has_gpu():
        assert lib.clann_init_with_config(data.ctypes.data, 10, 4, C.byref(cfg), C.byref(h)) == _lib.ERR_CUDA
        assert b"no CPU fallback" in lib.clann_last_error()
    assert lib.clann_build(None) == _lib.ERR_ARG
    assert lib.clann_search(None, None, 0, None, None, None) == _lib.ERR_ARG
    assert lib.clann_state_bytes(None) == 0


def test_legacy_create_rejects_unknown_type(lib, capfd):
    assert not lib.CPUFFINN_index_create(b"euclidean", 8)       # c_binder.cpp:46-49: stderr + NULL
    assert "Unsupported dataset type" in capfd.readouterr().err
    assert not lib.CPUFFINN_load_from_file(b"/nonexistent.h5", b"index_0")
    assert lib.CPUFFINN_search_cosine(None, None, 3, 0.9, 0.0, 4) is None or not lib.CPUFFINN_search_cosine(None, None, 3, 0.9, 0.0, 4)
    lib.CPUFFINN_clear_distance_computations()
    assert lib.CPUFFINN_get_distance_computations() == 0
    from clann_b200 import _lib
    assert lib.clann_puffinn_search(None, None, 1, 0.9, 0.0, 1, None, None, None) == _lib.ERR_ARG


def test_rust_binding_matches_header(lib):
    """bindings/rust/src/sys.rs (the reference-side binding; not compilable here) declares only functions the header
    declares and the library exports, with the header's status codes."""
    header = open(os.path.join(ROOT, "include", "clann_b200.h")).read()
    rust = open(os.path.join(ROOT, "bindings", "rust", "src", "sys.rs")).read()
    fns = re.findall(r"pub fn ((?:clann|CPUFFINN)_[A-Za-z0-9_]+)\s*\(", rust)
    assert len(fns) >= 18
    declared = set(declared_functions())
    assert not [f for f in fns if f not in declared or not hasattr(lib, f)]
    for name, value in re.findall(r"pub const (CLANN_[A-Z_]+): i32 = (-?\d+);", rust):
        m = re.search(r"\b" + name + r"\s*=\s*(-?\d+)", header)
        assert m and int(m.group(1)) == int(value), name
    # argument counts agree with the C prototypes
    csrc = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    for f in fns:
        c_args = re.search(r"\b" + f + r"\s*\(([^)]*)\)", csrc).group(1).strip()
        n_c = 0 if c_args in ("", "void") else c_args.count(",") + 1
        r_args = re.search(r"pub fn " + f + r"\s*\(([^)]*)\)", rust).group(1).strip()
        n_r = 0 if r_args == "" else r_args.count(",") + 1
        assert n_c == n_r, (f, c_args, r_args)
