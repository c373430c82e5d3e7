"""The C-ABI library loads and exports exactly what include/ssimu2_b200.h declares; without a GPU every
entry point fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu

HEADER = os.path.join(ROOT, "include", "ssimu2_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssimu2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from turbo_metrics_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python binding knows each of them
    assert set(names) == set(_lib.SYMBOLS)


def test_version_and_strerror():
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    assert lib.ssimu2_version() >> 16 == 1
    assert lib.ssimu2_strerror(0) == b"ok"
    assert b"unsupported" in lib.ssimu2_strerror(-2)
    assert b"ticket" in lib.ssimu2_strerror(-5)


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.ssimu2_create(None, None) == -1
    cfg = _lib.Config(4, 4, 2, 0, 0, 0, 0, 0)            # smaller than 8x8 (cpu.rs:359)
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = _lib.Config(64, 64, 17, 0, 0, 0, 0, 0)         # unknown format
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    assert lib.ssimu2_flush(None) == -1
    assert lib.ssimu2_get_score(None, 0, None) == -1
    assert lib.ssimu2_destroy(None) == 0


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from turbo_metrics_b200 import PixelFormat, Ssimu2Error, Ssimulacra2
    with pytest.raises(Ssimu2Error) as e:
        Ssimulacra2(64, 64, PixelFormat.SRGB8)
    assert e.value.status == -4  # SSIMU2_E_NODEVICE
