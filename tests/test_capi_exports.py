"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/ia_b200.h declares;
error behaviour without a device; the product refuses to run without its CUDA library / a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|const char\*)\s+(ia_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from intrinsicavatar_b200 import capi
    capi.build()
    lib = capi.load()
    names = _declared()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ia_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with the header"
    assert lib.ia_version() == 200 and lib.ia_voxel_format() in (0, 1)


def test_argtypes_cover_every_entry_point():
    from intrinsicavatar_b200 import engine
    for n in _declared():
        if n in ("ia_last_error", "ia_version", "ia_voxel_format"):
            continue
        assert n in engine._ARGTYPES, n


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_device_is_an_error_not_a_fallback():
    from intrinsicavatar_b200 import capi
    from intrinsicavatar_b200.engine import RenderEngine
    lib = capi.load()
    h = C.c_void_p()
    lib.ia_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    rc = lib.ia_create(C.byref(h), 0)
    assert rc == -2 and not h.value                          # IA_ECUDA
    assert b"cuda" in lib.ia_last_error().lower()
    with pytest.raises(RuntimeError, match="no CPU path"):
        RenderEngine()
    assert lib.ia_create(None, 0) == -1                      # IA_EINVAL
    assert lib.ia_destroy(None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from intrinsicavatar_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "libia_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        capi.load()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "intrinsicavatar_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
