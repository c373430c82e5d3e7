"""Error behaviour of the C ABI on a GPU box, called through ctypes exactly as a foreign-language binding would.

The reference panics (`todo!()`, `.unwrap()`) or returns `Err(CuError)` in these situations (cuda-colorspace/src/lib.rs:45-52,
ssimulacra2-cuda/src/lib.rs:110-138); the boundary promises a negative status and no abort (include/ssimu2_b200.h,
SURVEY 8b "Errors")."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OK, E_INVALID, E_UNSUPPORTED, E_TICKET = 0, -1, -2, -5


def _cfg(L, **kw):
    c = L.Config()
    c.width, c.height, c.format, c.matrix, c.full_range, c.device = 64, 64, 2, 0, 0, 0
    c.batch, c.ring = 4, 2
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_create_rejects_bad_configurations():
    import turbo_metrics_b200._lib as L
    lib = L.lib()
    h = C.c_void_p()
    assert lib.ssimu2_create(None, C.byref(_cfg(L))) == E_INVALID
    assert lib.ssimu2_create(C.byref(h), None) == E_INVALID
    for bad in (dict(width=4), dict(height=7), dict(width=40000), dict(format=6), dict(format=-1), dict(matrix=3), dict(pipeline=2),
                dict(flags=0x80), dict(flags=1, pipeline=1), dict(flags=4)):   # (flags=4: P016_DEEP on a non-P016 format)
        assert lib.ssimu2_create(C.byref(h), C.byref(_cfg(L, **bad))) == E_UNSUPPORTED, bad
        assert not h.value
    c = _cfg(L)
    c.reserved[2] = 1
    assert lib.ssimu2_create(C.byref(h), C.byref(c)) == E_INVALID
    assert lib.ssimu2_create(C.byref(h), C.byref(_cfg(L, device=99))) == E_INVALID
    assert lib.ssimu2_destroy(None) == OK
    for code in (0, -1, -2, -3, -4, -5, -6, 700):
        assert lib.ssimu2_strerror(code)          # a message for every status, CUDA codes included


def test_calls_reject_bad_arguments_and_tickets():
    import turbo_metrics_b200._lib as L
    lib = L.lib()
    h = C.c_void_p()
    assert lib.ssimu2_create(C.byref(h), C.byref(_cfg(L))) == OK
    try:
        img = torch.randint(0, 255, (64, 64 * 3), dtype=torch.uint8, device="cuda")
        good = L.Frame()
        good.plane[0], good.pitch = img.data_ptr(), 64 * 3
        t = C.c_uint64()
        sc = C.c_double()
        # frames: null, null plane, zero pitch, pitch shorter than a row
        assert lib.ssimu2_submit(h, None, C.byref(good), None, C.byref(t)) == E_INVALID
        for plane0, pitch in ((0, 192), (img.data_ptr(), 0), (img.data_ptr(), 191)):
            bad = L.Frame()
            bad.plane[0], bad.pitch = plane0, pitch
            assert lib.ssimu2_submit(h, C.byref(good), C.byref(bad), None, C.byref(t)) == E_INVALID
            assert lib.ssimu2_submit_host(h, C.byref(bad), C.byref(good), 64 * 192, C.byref(t)) == E_INVALID
        assert lib.ssimu2_submit_host(h, C.byref(good), C.byref(good), 0, C.byref(t)) == E_INVALID
        assert lib.ssimu2_submit_batch(h, 2, None, None, None, C.byref(t)) == E_INVALID
        # nothing was accepted: no ticket exists yet
        assert lib.ssimu2_get_score(h, 0, C.byref(sc)) == E_TICKET
        assert lib.ssimu2_wait(h, 5) == E_TICKET
        assert lib.ssimu2_get_score(h, 0, None) == E_INVALID
        assert lib.ssimu2_get_norms(h, 0, None) == E_INVALID
        assert lib.ssimu2_flush(h) == OK          # an empty flush is not an error
        # a good pair still works after all that, and unknown tickets beyond it are refused
        assert lib.ssimu2_submit(h, C.byref(good), C.byref(good), None, C.byref(t)) == OK and t.value == 0
        assert lib.ssimu2_get_score(h, 0, C.byref(sc)) == OK and sc.value == 100.0
        assert lib.ssimu2_get_score(h, 1, C.byref(sc)) == E_TICKET
        norms = (C.c_double * 108)()
        assert lib.ssimu2_get_norms(h, 0, norms) == OK and max(norms) == 0.0
        assert lib.ssimu2_get_norms(h, 7, norms) == E_TICKET
    finally:
        assert lib.ssimu2_destroy(h) == OK


def test_results_expire_after_the_result_ring_wraps():
    """Scores live in a 16384-entry ring: a ticket older than that is refused, never answered with another pair's score."""
    import turbo_metrics_b200 as tm
    w = h = 16
    a = torch.randint(0, 255, (h, w * 3), dtype=torch.uint8, device="cuda")
    b = torch.randint(0, 255, (h, w * 3), dtype=torch.uint8, device="cuda")
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=256, ring=2) as m:
        first = m.compute(tm.DeviceFrame.packed(a), tm.DeviceFrame.packed(b))
        s0 = m.get_score(first)
        n = 16384 + 256
        ts = m.compute_batch([tm.DeviceFrame.packed(a)] * n, [tm.DeviceFrame.packed(b)] * n)
        last = m.get_scores(ts[-4:])
        assert np.all(last == s0)                 # same pair, same bits, whatever the slot
        with pytest.raises(tm.Ssimu2Error) as e:
            m.get_score(first)
        assert e.value.status == E_TICKET
