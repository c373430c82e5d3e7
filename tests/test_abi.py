"""The C-ABI library loads and exports exactly what include/ssimu2_b200.h (the drop-in boundary) and
include/ssimu2_b200_debug.h (test / measurement hooks) declare; without a GPU every entry point fails loudly
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu

HEADERS = [os.path.join(ROOT, "include", "ssimu2_b200.h"), os.path.join(ROOT, "include", "ssimu2_b200_debug.h")]


def _declared(headers=HEADERS):
    names = set()
    for h in headers:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        names |= set(re.findall(r"\b(ssimu2_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from turbo_metrics_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 30
    public = _declared(HEADERS[:1])
    assert not any(n.startswith("ssimu2_debug") or n.endswith("_ms") for n in public), "test hooks belong in the debug header"
    assert all(n in public for n in ["ssimu2_shard_create", "ssimu2_shard_submit_host", "ssimu2_shard_get_scores",
                                     "ssimu2_stream_wait_input", "ssimu2_completed"])
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python binding knows each of them
    assert set(names) == set(_lib.SYMBOLS)


def test_version_and_strerror():
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    assert lib.ssimu2_version() >> 16 == 2
    assert lib.ssimu2_strerror(0) == b"ok"
    assert b"unsupported" in lib.ssimu2_strerror(-2)
    assert b"ticket" in lib.ssimu2_strerror(-5)


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.ssimu2_create(None, None) == -1
    from turbo_metrics_b200.ssimulacra2 import make_config
    cfg = make_config(4, 4, 2)                           # smaller than 8x8 (cpu.rs:359)
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = make_config(64, 64, 17)                        # unknown format
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = make_config(64, 64, 2, pipeline=7)             # unknown pipeline
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = make_config(64, 64, 2, flags=1 << 9)           # unknown flag
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = make_config(64, 64, 2, pipeline=1, flags=1)    # score-only exists for the product pipeline only
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -2
    cfg = make_config(64, 64, 2)
    cfg.reserved[2] = 1                                  # reserved fields must be zero
    assert lib.ssimu2_create(C.byref(h), C.byref(cfg)) == -1
    assert C.sizeof(_lib.Config) == 64 and C.sizeof(_lib.Frame) == 24
    assert lib.ssimu2_shard_create(None, None, None, 0) == -1
    assert lib.ssimu2_shard_destroy(None) == 0 and lib.ssimu2_shard_flush(None) == -1
    assert lib.ssimu2_flush(None) == -1
    assert lib.ssimu2_get_score(None, 0, None) == -1
    assert lib.ssimu2_destroy(None) == 0


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from turbo_metrics_b200 import PixelFormat, ShardedSsimulacra2, Ssimu2Error, Ssimulacra2
    with pytest.raises(Ssimu2Error) as e:
        Ssimulacra2(64, 64, PixelFormat.SRGB8)
    assert e.value.status == -4  # SSIMU2_E_NODEVICE
    with pytest.raises(Ssimu2Error) as e:   # the worker threads fail to create their handles; create reports it and joins them
        ShardedSsimulacra2(64, 64, PixelFormat.SRGB8, devices=[0, 1])
    assert e.value.status == -4


def _build_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "ssimu2_c_client")
    from turbo_metrics_b200 import _lib
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "ssimu2_c_client.c"),
                           "-L" + libdir, "-lssimu2_b200", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_plain_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    """The header is plain C (compiled here with gcc -Wall -Werror) and the library is self-contained: a C program that includes
    nothing but include/ssimu2_b200.h links against it; without a GPU it reports SSIMU2_E_NODEVICE instead of computing anything."""
    import subprocess
    exe = _build_c_client(tmp_path)
    z = tmp_path / "z.rgb"
    z.write_bytes(bytes(64 * 64 * 3))
    r = subprocess.run([exe, "64", "64", str(z), str(z)], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr and r.stdout == ""


def test_flag_and_status_constants_match_the_header():
    """The Python loader's constants are the header's (a binding written from include/ssimu2_b200.h must agree with ours)."""
    import re
    from turbo_metrics_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ssimu2_b200.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SSIMU2_FLAG_(\w+)\s+(\d+)u", hdr)}
    assert flags == {"SCORE_ONLY": _lib.FLAG_SCORE_ONLY, "NO_TIMING": _lib.FLAG_NO_TIMING, "P016_DEEP": _lib.FLAG_P016_DEEP}
    status = {m.group(1): int(m.group(2)) for m in re.finditer(r"SSIMU2_E_(\w+)\s*=\s*(-\d+)", hdr)}
    assert status == {"INVALID": -1, "UNSUPPORTED": -2, "NOMEM": -3, "NODEVICE": -4, "TICKET": -5, "INTERNAL": -6}
    fmts = {m.group(1): int(m.group(2)) for m in re.finditer(r"SSIMU2_FMT_(\w+)\s*=\s*(\d+)", hdr)}
    from turbo_metrics_b200 import PixelFormat
    assert fmts == {f.name: int(f.value) for f in PixelFormat}
