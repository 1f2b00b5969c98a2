"""The C-ABI library loads and exports every symbol include/moonsr.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

import pytest

from moonsuperresolution_b200 import _lib


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "libmoonsr.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`"


def test_every_declared_symbol_is_exported_and_typed():
    handle = _lib.lib()
    declared = _lib.header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/moonsr.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(declared)


def test_prototype_arity_matches_header():
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes prototype {len(argtypes)}"


def test_version_and_error_string_without_gpu():
    handle = _lib.lib()
    assert handle.msr_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    handle = _lib.lib()
    rc = handle.msr_blend_tile(None, None, None, None, 0, None, 0, None, 64, 8, 256, 0, -32768.0, None, None, None,
                               256, 256, 256, None)
    assert rc == -1 and "null" in _lib.last_error()
    h = ctypes.c_void_p()
    assert handle.msr_generator_create(ctypes.byref(h), 7, 64, 1, 1, 0) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MoonSRError):
        _lib.lib()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from moonsuperresolution_b200 import DEMSuperResolution, DSRConfig, CNNSpade
    with pytest.raises(_lib.MoonSRError):
        DEMSuperResolution(DSRConfig(image_size=64, stride=8, tile_size=256))
    with pytest.raises(_lib.MoonSRError):
        CNNSpade(64, 2)
