"""CPU: the C-ABI library loads and exports every symbol include/ivid_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ivid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ivid_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from ivid_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ivid_b200.h but not exported"


def test_binding_covers_header():
    from ivid_b200 import _lib
    assert set(_declared_symbols()) <= set(_lib.SIGNATURES), "ctypes binding table is missing a declared entry point"
    assert _lib.lib().ivid_version() == 100


def test_error_convention():
    """status code + thread-local message, mapped to the reference's exception types."""
    import pytest
    from ivid_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.ivid_unet_create(b'{"image_size": 32}', ctypes.byref(h))
    assert rc == _lib.IVID_ERR_INVALID_ARGUMENT and "missing" in _lib.last_error()
    with pytest.raises(AssertionError):
        _lib.check(rc)
    cfg = b'{"image_size":32,"in_channels":4,"model_channels":64,"out_channels":4,"num_res_blocks":1,"attention_resolutions":[16],"num_head_channels":32}'
    rc = L.ivid_unet_create(cfg, ctypes.byref(h))
    assert rc == _lib.IVID_ERR_NOT_IMPLEMENTED
    with pytest.raises(NotImplementedError):
        _lib.check(rc)
