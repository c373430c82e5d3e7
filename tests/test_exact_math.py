"""The restated glibc cbrtf / powf (turbo_metrics_b200/csrc/exact_math.cuh), compiled for the HOST,
must equal this machine's libm bit for bit: that is what makes the GPU filter inputs identical to the
CPU reference's (see the header of exact_math.cuh for why anything less fails the parity bar)."""
import ctypes as C

import numpy as np
import pytest


def _run(lib, fn, x, y=None):
    out = np.empty_like(x)
    f = getattr(lib, fn)
    if y is None:
        f(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    else:
        f(x.ctypes.data_as(C.c_void_p), C.c_float(y), out.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    return out


def _bits_equal(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def test_cbrtf_exhaustive_binades(exact_math_host):
    # every float of three full binades (covers all e % 3 classes) + the opsin range edge
    for lo in [0.0037, 0.125, 0.25, 0.5]:
        start = int(np.float32(lo).view(np.uint32))
        x = np.arange(start, start + (1 << 23), dtype=np.uint32).view(np.float32)
        a, b = _run(exact_math_host, "em_cbrtf_array", x), _run(exact_math_host, "libm_cbrtf_array", x)
        assert _bits_equal(a, b).all()


def test_cbrtf_random_bit_patterns(exact_math_host):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2 ** 32, size=4_000_000, dtype=np.uint32).view(np.float32)
    a, b = _run(exact_math_host, "em_cbrtf_array", x), _run(exact_math_host, "libm_cbrtf_array", x)
    assert _bits_equal(a, b).all()
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-40, 1.0, -8.0, 27.0], np.float32)
    a, b = _run(exact_math_host, "em_cbrtf_array", special), _run(exact_math_host, "libm_cbrtf_array", special)
    assert _bits_equal(a, b).all()


@pytest.mark.parametrize("y", [float(np.float32(1.0) / np.float32(0.45)), 2.4])
def test_powf_exhaustive_eotf_range(exact_math_host, y):
    # BT.709 / sRGB EOTF arguments live in (0.07, 1.25]; take every float of [2^-4, 2)
    start, stop = int(np.float32(2 ** -4).view(np.uint32)), int(np.float32(2.0).view(np.uint32))
    for s in range(start, stop, 1 << 24):
        x = np.arange(s, min(s + (1 << 24), stop), dtype=np.uint32).view(np.float32)
        a, b = _run(exact_math_host, "em_powf_array", x, y), _run(exact_math_host, "libm_powf_array", x, y)
        assert _bits_equal(a, b).all()


def test_powf_random_bit_patterns(exact_math_host):
    rng = np.random.default_rng(4)
    x = rng.integers(0, 2 ** 32, size=2_000_000, dtype=np.uint32).view(np.float32)
    for y in [2.4, 0.5, -1.5, 30.0]:
        a, b = _run(exact_math_host, "em_powf_array", x, y), _run(exact_math_host, "libm_powf_array", x, y)
        assert _bits_equal(a, b).all()
