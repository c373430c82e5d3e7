"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): every one of the 108 per-scale / per-channel norms within 1e-4
relative, the final score within 0.01 absolute.  Because the filter inputs are reproduced bit for
bit (exact_math.cuh) the intermediate planes are additionally required to be IDENTICAL to the
oracle's, which is what keeps the norms far inside the bar.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-4   # north_star: per-scale/per-channel norms within 1e-4 relative
SCORE_ATOL = 0.01  # north_star: final score within 0.01 absolute


def _tm():
    import turbo_metrics_b200 as tm
    return tm


def _assert_norms(norms, ref_norms, score, ref_score):
    ref_norms = np.asarray(ref_norms)
    # norms of exactly 0 in the oracle (identical inputs) must be exactly 0 here too
    denom = np.maximum(np.abs(ref_norms), 1e-300)
    rel = np.abs(norms - ref_norms) / denom
    rel[ref_norms == 0] = np.abs(norms[ref_norms == 0])
    assert rel.max() <= NORM_RTOL, f"max rel norm err {rel.max():.3e} at {rel.argmax()}"
    assert abs(score - ref_score) <= SCORE_ATOL, (score, ref_score)
    return rel.max()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------
def test_device_libm_restatements_match_host_libm(exact_math_host):
    """Device cbrtf / powf == this box's libm, bit for bit, on the ranges the pipeline uses."""
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(11)
    x = np.concatenate([
        rng.uniform(0.0037, 1.01, 3_000_000).astype(np.float32),
        np.arange(int(np.float32(0.25).view(np.uint32)), int(np.float32(0.25).view(np.uint32)) + (1 << 21),
                  dtype=np.uint32).view(np.float32),
        np.array([0.0, 1.0, 0.0037930734, 8.0, 27.0], np.float32)])
    out = np.empty_like(x)
    ref = np.empty_like(x)
    fp = C.POINTER(C.c_float)
    exact_math_host.libm_cbrtf_array(x.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), C.c_size_t(x.size))
    assert lib.ssimu2_debug_math(0, x.ctypes.data_as(fp), 0.0, out.ctypes.data_as(fp), x.size) == 0
    assert np.array_equal(_bits(out), _bits(ref)), f"{(_bits(out) != _bits(ref)).sum()} cbrtf mismatches"
    pos = x > 0  # the unchecked hot-path form
    xp, out = np.ascontiguousarray(x[pos]), np.empty(int(pos.sum()), np.float32)
    assert lib.ssimu2_debug_math(5, xp.ctypes.data_as(fp), 0.0, out.ctypes.data_as(fp), xp.size) == 0
    assert np.array_equal(_bits(out), _bits(ref[pos]))
    for y in [float(np.float32(1.0) / np.float32(0.45)), 2.4]:
        xb = rng.uniform(0.07, 1.3, 3_000_000).astype(np.float32)
        out = np.empty_like(xb)
        ref = np.empty_like(xb)
        assert lib.ssimu2_debug_math(1, xb.ctypes.data_as(fp), y, out.ctypes.data_as(fp), xb.size) == 0
        exact_math_host.libm_powf_array(xb.ctypes.data_as(C.c_void_p), C.c_float(y), ref.ctypes.data_as(C.c_void_p),
                                        C.c_size_t(xb.size))
        assert np.array_equal(_bits(out), _bits(ref)), f"{(_bits(out) != _bits(ref)).sum()} powf mismatches (y={y})"
        assert lib.ssimu2_debug_math(6, xb.ctypes.data_as(fp), y, out.ctypes.data_as(fp), xb.size) == 0
        assert np.array_equal(_bits(out), _bits(ref)), "unchecked powf"


def test_device_divisions_are_ieee():
    """The hand-rolled divisions (reciprocal seed + Newton + residual correction, no special-case paths) give
    the IEEE quotient on the operand ranges the pipeline produces."""
    from turbo_metrics_b200 import _lib
    lib = _lib.lib()
    fp = C.POINTER(C.c_float)
    rng = np.random.default_rng(5)
    # f32: EOTF argument / ALPHA   (cuda-colorspace-kernel/src/lib.rs:229, srgb.rs:46)
    for alpha in (np.float32(1.0) + np.float32(5.5) * np.float32(0.018053968510807), np.float32(1.0550107)):
        x = np.concatenate([rng.uniform(0.05, 1.6, 4_000_000).astype(np.float32),
                            np.arange(int(np.float32(0.5).view(np.uint32)), int(np.float32(0.5).view(np.uint32)) + (1 << 22),
                                      dtype=np.uint32).view(np.float32)])
        out = np.empty_like(x)
        assert lib.ssimu2_debug_math(2, x.ctypes.data_as(fp), float(alpha), out.ctypes.data_as(fp), x.size) == 0
        assert np.array_equal(_bits(out), _bits(x / alpha))
    # f32: the SSIM quotient (cpu.rs:627): numerator in [-2e-3, 2.5], denominator in [5e-4, 2]
    n = 4_000_000
    nd = np.empty((n, 2), np.float32)
    nd[:, 0] = rng.uniform(-2e-3, 2.5, n) * rng.choice([1.0, 1e-2, 1e-3], n)
    nd[:, 1] = np.exp(rng.uniform(np.log(5e-4), np.log(2.0), n))
    out = np.empty(n, np.float32)
    assert lib.ssimu2_debug_math(4, nd.ctypes.data_as(fp), 0.0, out.ctypes.data_as(fp), n) == 0
    ref = nd[:, 0] / nd[:, 1]
    ok = (_bits(out) == _bits(ref)) | (np.abs(ref) < 1e-30)   # flush-to-zero territory is irrelevant here
    assert ok.all(), f"{(~ok).sum()} f32 quotient mismatches"
    # f64: the Halley step of cbrtf: num in [1.2, 3.1], den in [1.4, 3.1]
    pairs = np.empty((n, 2), np.float64)
    pairs[:, 0] = rng.uniform(1.0, 3.2, n)
    pairs[:, 1] = rng.uniform(1.2, 3.2, n)
    outd = np.empty(n, np.float64)
    assert lib.ssimu2_debug_math(3, pairs.ctypes.data_as(fp), 0.0, outd.ctypes.data_as(fp), n) == 0
    assert np.array_equal(outd.view(np.uint64), (pairs[:, 0] / pairs[:, 1]).view(np.uint64))


# ------------------------------------------------------------------------------------------
def _oracle_stages(oracle, ref_lin, dis_lin, nscales):
    """XYB planes and H-pass planes of every scale, from the oracle's own building blocks."""
    xyb, hb = [], []
    a, b = ref_lin, dis_lin
    for s in range(nscales):
        if s > 0:
            a, b = oracle.downscale_by_2(a), oracle.downscale_by_2(b)
        xa, xb = oracle.linear_to_xyb(a), oracle.linear_to_xyb(b)
        xyb.append(np.concatenate([xa, xb], axis=0))
        planes = []
        for q in (xa * xa, xb * xb, xa * xb, xa, xb):
            for c in range(3):
                planes.append(oracle.blur_horizontal(q[c]))
        hb.append(np.stack(planes))
    return xyb, hb


@pytest.mark.parametrize("w,h", [(160, 96), (203, 131)])
def test_intermediate_planes_are_bit_identical_srgb8(oracle, w, h):
    tm = _tm()
    from turbo_metrics_b200 import synth
    r, d = synth.make_pair_srgb8(w, h, frame=3, seed=5)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=1, ring=1, pipeline="split") as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg))
        score = m.get_score(t)
        norms = m.get_norms(t)
        ns = m.info().nscales
        xyb_o, hb_o = _oracle_stages(oracle, oracle.linear_from_srgb8(r.numpy()), oracle.linear_from_srgb8(d.numpy()), ns)
        for s in range(ns):
            assert np.array_equal(_bits(m.debug_read(t, 0, s)), _bits(xyb_o[s])), f"XYB planes differ at scale {s}"
            assert np.array_equal(_bits(m.debug_read(t, 1, s)), _bits(hb_o[s])), f"H-pass planes differ at scale {s}"
    so, no, nso = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
    assert nso == ns
    _assert_norms(norms, no, score, so)


@pytest.mark.parametrize("kind", ["linear", "linear_out_of_range", "linear_huge", "srgb16", "srgbf32", "srgbf32_out_of_range"])
def test_intermediate_planes_are_bit_identical_packed16_and_float(oracle, kind):
    """The fast front-end path of the linear-f32 (the reference's own `Ssimulacra2::new` input, lib.rs:48) and sRGB16 formats:
    interior regions go through it, the frame edge through the general path, and both must give the oracle's bits.  The
    out-of-range variants plant negative, -0, > 1, subnormal (and, `linear_huge`, > 1e30 and nan) samples: the regions the unchecked
    cube root cannot take must fall back to the general path, everything must still be bit-identical (for `linear_huge` the
    sums overflow in the oracle too, so only the planes are compared)."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 200, 136
    if kind == "srgb16":
        r8, d8 = synth.make_pair_srgb8(w, h, frame=4, seed=6)
        r = (r8.to(torch.int32) * 257 + 3).clamp(0, 65535).to(torch.int16)
        d = (d8.to(torch.int32) * 257 + 40).clamp(0, 65535).to(torch.int16)
        a = oracle.linear_from_srgb16(r.numpy().view(np.uint16))
        b = oracle.linear_from_srgb16(d.numpy().view(np.uint16))
        fmt = tm.PixelFormat.SRGB16
    elif kind.startswith("srgbf32"):
        r8, d8 = synth.make_pair_srgb8(w, h, frame=4, seed=6)
        r, d = r8.to(torch.float32) / 255.0, (d8.to(torch.float32) / 255.0 * 0.97 + 0.013)
        r[5, 7, :] = 0.0                              # exact zeros and the linear segment of the transfer function
        d[5, 9, :] = torch.tensor([0.0392, 0.0393, 0.0394])
        if kind == "srgbf32_out_of_range":             # negative, -0, tiny, above one, huge: the region goes to the checked path
            r[40, 50, 1] = -0.25; r[70, 10, 2] = -0.0; d[12, 150, 1] = 1.0e-38; d[90, 40, 2] = 1.5; r[100, 100, 0] = 2.0e6
        a, b = oracle.linear_from_srgbf32(r.numpy()), oracle.linear_from_srgbf32(d.numpy())
        fmt = tm.PixelFormat.SRGBF32
    else:
        r, d = synth.make_pair_linearf32(w, h, frame=4, seed=6)
        if kind != "linear":
            r, d = r.clone(), d.clone()
            r[40, 50, 1] = -0.25; r[41, 90, 0] = 1.75; r[70, 10, 2] = -0.0; d[12, 150, 1] = 1.0e-42; d[90, 40, 2] = 7.5
            if kind == "linear_huge":
                d[100, 120, 0] = 3.0e30; r[100, 30, 1] = float("nan")
        a, b = oracle.linear_from_linearf32(r.numpy()), oracle.linear_from_linearf32(d.numpy())
        fmt = tm.PixelFormat.LINEARF32
    with tm.Ssimulacra2(w, h, fmt, batch=1, ring=1, pipeline="split") as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg))
        score, norms, ns = m.get_score(t), m.get_norms(t), m.info().nscales
        xyb_o, hb_o = _oracle_stages(oracle, a, b, ns)
        for s in range(ns):
            assert np.array_equal(_bits(m.debug_read(t, 0, s)), _bits(xyb_o[s])), f"XYB planes differ at scale {s}"
            assert np.array_equal(_bits(m.debug_read(t, 1, s)), _bits(hb_o[s])), f"H-pass planes differ at scale {s}"
    if kind != "linear_huge" and np.all(np.isfinite(a)) and np.all(np.isfinite(b)):
        so, no, _ = oracle.ssimu2_linear_planar(a, b)
        _assert_norms(norms, no, score, so)


@pytest.mark.parametrize("content", ["12bit", "random16", "10bit"])
def test_p016_deep_flag_changes_the_speed_not_the_bits(oracle, content):
    """SSIMU2_FLAG_P016_DEEP: a front-end for P016 samples with more than 10 significant bits (HEVC Main12 through NVDEC) that
    evaluates the three transfer functions instead of the 10-bit memo tables.  With or without the flag, and whatever the
    content, the planes must be the oracle's bits (the oracle applies biplanar.rs:7-70 to the full 16-bit samples)."""
    tm = _tm()
    w, h, pitch, ch = 256, 168, 512, 168
    rng = np.random.default_rng(11)

    def frame():
        n = pitch // 2 * ch * 3 // 2
        if content == "random16":
            v = rng.integers(0, 65536, n, dtype=np.uint16)
        else:
            v = rng.integers(0, 1024, n, dtype=np.uint16) << 6
            if content == "12bit":
                v |= rng.integers(0, 4, n, dtype=np.uint16) << 4
        return torch.from_numpy(v.view(np.uint8).copy())
    r = frame()
    d = r.clone()
    d[::5] = frame()[::5]
    a = oracle.linear_from_yuv420(r.numpy(), pitch, ch, w, h, 16)
    b = oracle.linear_from_yuv420(d.numpy(), pitch, ch, w, h, 16)
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    rg, dg = r.cuda(), d.cuda()
    got = {}
    for deep in (False, True):
        with tm.Ssimulacra2(w, h, tm.PixelFormat.P016, batch=1, ring=1, pipeline="split", p016_deep=deep) as m:
            t = m.compute(F(rg), F(dg))
            got[deep] = (m.get_score(t), m.get_norms(t))
            ns = m.info().nscales
            xyb_o, _ = _oracle_stages(oracle, a, b, ns)
            for s in range(ns):
                assert np.array_equal(_bits(m.debug_read(t, 0, s)), _bits(xyb_o[s])), f"deep={deep}: XYB planes differ at scale {s}"
    assert got[False][0] == got[True][0] and np.array_equal(got[False][1], got[True][1])
    so, no, _ = oracle.ssimu2_linear_planar(a, b)
    _assert_norms(got[True][1], no, got[True][0], so)
    with pytest.raises(tm.Ssimu2Error) as e:     # the flag belongs to P016
        tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, p016_deep=True)
    assert e.value.status == -2


@pytest.mark.parametrize("bits", [8, 16])
def test_intermediate_planes_are_bit_identical_yuv(oracle, bits):
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 192, 108
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=2, seed=9)
    fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
    with tm.Ssimulacra2(w, h, fmt, batch=1, ring=1, pipeline="split") as m:
        rg, dg = rb.cuda(), db.cuda()
        t = m.compute(tm.DeviceFrame.yuv420(rg, pitch, ch), tm.DeviceFrame.yuv420(dg, pitch, ch))
        score, norms, ns = m.get_score(t), m.get_norms(t), m.info().nscales
        a = oracle.linear_from_yuv420(rb.numpy(), pitch, ch, w, h, bits)
        b = oracle.linear_from_yuv420(db.numpy(), pitch, ch, w, h, bits)
        xyb_o, hb_o = _oracle_stages(oracle, a, b, ns)
        for s in range(ns):
            assert np.array_equal(_bits(m.debug_read(t, 0, s)), _bits(xyb_o[s])), f"XYB planes differ at scale {s}"
            assert np.array_equal(_bits(m.debug_read(t, 1, s)), _bits(hb_o[s])), f"H-pass planes differ at scale {s}"
    so, no, _ = oracle.ssimu2_linear_planar(a, b)
    _assert_norms(norms, no, score, so)


# ------------------------------------------------------------------------------------------
CASES = [
    ("srgb8", 512, 512), ("srgb8", 64, 64), ("srgb8", 9, 300), ("srgb8", 301, 8), ("srgb8", 257, 255),
    ("nv12", 640, 360), ("nv12", 1920, 1080), ("nv12", 66, 34),
    ("p016", 640, 360), ("p016", 960, 540),
    ("linear", 320, 200), ("srgb16", 200, 120), ("srgbf32", 200, 120),
]


def _make(kind, w, h, frame, seed, oracle):
    """-> (ref_frame_builder(device tensors), oracle (score, norms, ns))"""
    tm = _tm()
    from turbo_metrics_b200 import synth
    if kind in ("nv12", "p016"):
        bits = 8 if kind == "nv12" else 16
        rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=frame, seed=seed)
        res = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, bits)
        fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
        return fmt, (lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)), rb, db, res
    if kind == "srgb8":
        r, d = synth.make_pair_srgb8(w, h, frame=frame, seed=seed)
        return tm.PixelFormat.SRGB8, tm.DeviceFrame.packed, r, d, oracle.ssimu2_srgb8(r.numpy(), d.numpy())
    if kind == "linear":
        r, d = synth.make_pair_linearf32(w, h, frame=frame, seed=seed)
        return tm.PixelFormat.LINEARF32, tm.DeviceFrame.packed, r, d, oracle.ssimu2_linearf32(r.numpy(), d.numpy())
    r8, d8 = synth.make_pair_srgb8(w, h, frame=frame, seed=seed)
    if kind == "srgb16":
        r = (r8.to(torch.int32) * 257).to(torch.int16)  # bit pattern of u16 0..65535
        d = (d8.to(torch.int32) * 257 + 13).clamp(0, 65535).to(torch.int16)
        rn, dn = r.numpy().view(np.uint16), d.numpy().view(np.uint16)
        res = oracle.ssimu2_linear_planar(oracle.linear_from_srgb16(rn), oracle.linear_from_srgb16(dn))
        return tm.PixelFormat.SRGB16, tm.DeviceFrame.packed, r, d, res
    r, d = r8.to(torch.float32) / 255.0, (d8.to(torch.float32) / 255.0 * 0.98 + 0.01)
    res = oracle.ssimu2_linear_planar(oracle.linear_from_srgbf32(r.numpy()), oracle.linear_from_srgbf32(d.numpy()))
    return tm.PixelFormat.SRGBF32, tm.DeviceFrame.packed, r, d, res


@pytest.mark.parametrize("kind,w,h", CASES)
def test_score_and_norms_match_oracle(oracle, kind, w, h):
    tm = _tm()
    fmt, mk, r, d, (so, no, nso) = _make(kind, w, h, frame=1, seed=3, oracle=oracle)
    with tm.Ssimulacra2(w, h, fmt, batch=2, ring=2) as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(mk(rg), mk(dg))
        score = m.get_score(t)
        norms = m.get_norms(t)
        assert m.info().nscales == nso
    _assert_norms(norms, no, score, so)


@pytest.mark.parametrize("w,h", [(203, 131), (640, 360), (1000, 77)])
def test_pipelines_agree(oracle, w, h):
    """The two launch pipelines ("hv": fused H+V kernel with the systolic strip hand-off; "split": separate H and V passes
    with the H-pass planes in HBM) run the same arithmetic: norms agree to accumulation-order noise, and both match the
    oracle."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    n = 5
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=17) for i in range(n)]
    dev = [(r.cuda(), d.cuda()) for r, d in pairs]
    out = {}
    for pl in ("hv", "split"):
        with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=3, ring=2, pipeline=pl) as m:
            ts = [m.compute(tm.DeviceFrame.packed(r), tm.DeviceFrame.packed(d)) for r, d in dev]
            out[pl] = [(m.get_score(t), m.get_norms(t)) for t in ts]
    for i, (r, d) in enumerate(pairs[:2]):
        so, no, _ = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
        for pl in out:
            _assert_norms(out[pl][i][1], no, out[pl][i][0], so)
    for i in range(n):
        for pl in ("split",):
            assert abs(out["hv"][i][0] - out[pl][i][0]) < 1e-6
            np.testing.assert_allclose(out["hv"][i][1], out[pl][i][1], rtol=5e-7, atol=1e-12)


def test_identical_frames_score_100():
    tm = _tm()
    from turbo_metrics_b200 import synth
    r, _ = synth.make_pair_srgb8(256, 256, frame=0, seed=1)
    with tm.Ssimulacra2(256, 256, tm.PixelFormat.SRGB8) as m:
        rg = r.cuda()
        s = m.compute_sync(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(rg))
        assert s == 100.0


def test_bt601_matrices_and_full_range(oracle):
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 320, 180
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, 8, frame=4, seed=2)
    for matrix, name in [(tm.ColorMatrix.BT601_525, "bt601_525"), (tm.ColorMatrix.BT601_625, "bt601_625")]:
        for full in (False, True):
            so, no, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, 8, name, full)
            with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, matrix=matrix, full_range=full) as m:
                rg, dg = rb.cuda(), db.cuda()
                t = m.compute(tm.DeviceFrame.yuv420(rg, pitch, ch), tm.DeviceFrame.yuv420(dg, pitch, ch))
                _assert_norms(m.get_norms(t), no, m.get_score(t), so)


def test_batches_ring_and_ticket_order(oracle):
    """37 distinct pairs through batch=4 x ring=3: tickets come back in submission order, every score
    equals the oracle's, and a partial last batch is flushed by get_score."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n = 160, 120, 37
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=21) for i in range(n)]
    expect = [oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0] for r, d in pairs]
    dev = [(r.cuda(), d.cuda()) for r, d in pairs]
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=4, ring=3) as m:
        tickets = [m.compute(tm.DeviceFrame.packed(r), tm.DeviceFrame.packed(d)) for r, d in dev]
        assert tickets == list(range(n))
        got = [m.get_score(t) for t in tickets]
        # device-side score stream holds the same values
        ptr, cap = m.scores_device()
        class _Ring:  # zero-copy view of the device score ring
            __cuda_array_interface__ = {"shape": (cap,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        torch.cuda.synchronize()
        assert torch.as_tensor(_Ring(), device="cuda")[:n].cpu().tolist() == got
    assert max(abs(a - b) for a, b in zip(got, expect)) <= SCORE_ATOL
    assert len(set(round(x, 6) for x in got)) > n // 2  # really distinct inputs


def test_host_frames_entry_point(oracle):
    """compute_from_cpu (Ssimulacra2::compute_from_cpu_srgb_sync, lib.rs:232-250) on pinned host buffers."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 640, 360
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, 16, frame=6, seed=4)
    so, no, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, 16)
    rp, dp = rb.pin_memory(), db.pin_memory()
    with tm.Ssimulacra2(w, h, tm.PixelFormat.P016, batch=2, ring=2) as m:
        ts = [m.compute_from_cpu(tm.DeviceFrame.yuv420(rp, pitch, ch), tm.DeviceFrame.yuv420(dp, pitch, ch)) for _ in range(5)]
        scores = [m.get_score(t) for t in ts]
        _assert_norms(m.get_norms(ts[-1]), no, scores[-1], so)
        # the batched entry point (ssimu2_submit_host_batch): same tickets-in-order contract, same scores
        fr, fd = tm.DeviceFrame.yuv420(rp, pitch, ch), tm.DeviceFrame.yuv420(dp, pitch, ch)
        tb = m.compute_from_cpu_batch([fr] * 5, [fd] * 5)
        assert list(tb) == list(range(ts[-1] + 1, ts[-1] + 6))
        assert all(m.get_score(t) == scores[0] for t in tb)
        tb2 = m.compute_from_cpu_batch([fr] * 3, [fd] * 3)
        np.testing.assert_array_equal(m.get_scores(tb2), np.full(3, scores[0]))
    assert all(s == scores[0] for s in scores)


def test_host_frames_in_pageable_memory_can_be_reused_at_once(oracle):
    """`compute_from_cpu_srgb_sync` takes ordinary slices (lib.rs:232-250): host frames need not be pinned, and a pageable buffer
    has been consumed when submit returns -- the caller's frame loop may decode the next frame into the same buffer."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 512, 288
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=9) for i in range(4)]
    want = [oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0] for r, d in pairs]
    buf_r, buf_d = torch.empty_like(pairs[0][0]), torch.empty_like(pairs[0][1])
    assert not buf_r.is_pinned() and not buf_d.is_pinned()
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=4, ring=2) as m:
        ts = []
        for r, d in pairs:
            buf_r.copy_(r); buf_d.copy_(d)
            ts.append(m.compute_from_cpu(tm.DeviceFrame.packed(buf_r), tm.DeviceFrame.packed(buf_d)))
            buf_r.zero_(); buf_d.fill_(255)        # clobber the buffers right after submit
        got = [m.get_score(t) for t in ts]
    assert len(set(np.round(want, 3))) == 4          # four different pairs: a mixed-up or clobbered frame would show
    np.testing.assert_allclose(got, want, rtol=0, atol=SCORE_ATOL)


def test_caller_stream_ordering(oracle):
    """Frames produced on the caller's stream right before submit are seen by the scorer."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 256, 144
    r, d = synth.make_pair_srgb8(w, h, frame=9, seed=8)
    so = oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0]
    side = torch.cuda.Stream()
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=1, ring=2) as m:
        rg = torch.zeros_like(r, device="cuda")
        dg = torch.zeros_like(d, device="cuda")
        rp, dp = r.pin_memory(), d.pin_memory()
        big = torch.empty(64 << 20, device="cuda")
        with torch.cuda.stream(side):
            for _ in range(10):
                big.normal_()           # keep the stream busy so an unordered read would see zeros
            rg.copy_(rp, non_blocking=True)
            dg.copy_(dp, non_blocking=True)
            t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg), stream=side)
        assert abs(m.get_score(t) - so) <= SCORE_ATOL


# ------------------------------------------------------------------------------------------ the benchmarked configurations
def test_4k_p016_matches_oracle(oracle):
    """BASELINE.json configs[2] (the headline bench workload): one 3840x2160 10-bit P016 pair, 108 norms + score against
    the oracle -- 60 strips x 181 bands of systolic hand-off per chain at scale 0, the largest in the product.  The pair is
    the first one bench.py times (frame 0, seed 1); it sits in the middle of a full batch of other pairs so that every
    hand-off runs under load (examples/compare.rs:42-90 is the one place the reference makes GPU and CPU meet)."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 3840, 2160
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, 16, frame=0, seed=1)
    so, no, nso = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w, h, 16)
    rb2, db2, _, _ = synth.make_pair_yuv420(w, h, 16, frame=1, seed=1, device="cuda")
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.P016, batch=8, ring=2) as m:
        rg, dg = rb.cuda(), db.cuda()
        ts = [m.compute(F(rb2), F(db2)) for _ in range(3)] + [m.compute(F(rg), F(dg))] + [m.compute(F(rb2), F(db2)) for _ in range(4)]
        t = ts[3]
        score, norms = m.get_score(t), m.get_norms(t)
        assert m.info().nscales == nso == 6
    rel = _assert_norms(norms, no, score, so)
    print(f"4K P016: score {score:.6f} (oracle {so:.6f}), max rel norm err {rel:.2e}")


def test_1080p_srgb8_matches_oracle(oracle):
    """BASELINE.json configs[0]: one 1920x1080 sRGB8 pair, seed 1."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 1920, 1080
    r, d = synth.make_pair_srgb8(w, h, frame=0, seed=1)
    so, no, _ = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=2, ring=2) as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg))
        _assert_norms(m.get_norms(t), no, m.get_score(t), so)


def test_512_batch_matches_oracle(oracle):
    """BASELINE.json configs[4]: a batch of distinct 512x512 sRGB8 pairs through ONE ssimu2_submit_batch call, every
    pair's 108 norms + score against the oracle."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w = h = 512
    n = 40
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=1) for i in range(n)]
    dev = [(r.cuda(), d.cuda()) for r, d in pairs]
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=32, ring=3) as m:
        ts = m.compute_batch([tm.DeviceFrame.packed(r) for r, _ in dev], [tm.DeviceFrame.packed(d) for _, d in dev])
        scores = m.get_scores(ts)
        norms = [m.get_norms(t) for t in ts]
    for i, (r, d) in enumerate(pairs):
        so, no, _ = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
        _assert_norms(norms[i], no, scores[i], so)
    assert len(set(round(float(x), 6) for x in scores)) > n // 2


def test_property_checks_at_full_size():
    """4K P016 (BASELINE config 3): size-independent properties on top of the oracle comparison above.
    identical -> 100; the score does not depend on the batch slot or on the neighbours in the batch;
    a more distorted frame scores lower."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 3840, 2160
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, 16, frame=0, seed=1, device="cuda")
    rb2, db2, _, _ = synth.make_pair_yuv420(w, h, 16, frame=1, seed=1, device="cuda")
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.P016, batch=4, ring=2) as m:
        t = [m.compute(F(rb), F(db)), m.compute(F(rb), F(rb)), m.compute(F(rb2), F(db2)), m.compute(F(rb), F(db)),
             m.compute(F(rb2), F(db2)), m.compute(F(rb), F(db2))]
        s = [m.get_score(x) for x in t]
    assert s[1] == 100.0
    assert s[0] == s[3] and s[2] == s[4]
    assert 0 < s[0] < 100 and 0 < s[2] < 100
    assert s[5] < min(s[0], s[2])  # unrelated frame is far worse than its own distorted version


def test_engine_compute_one_and_compute_all(oracle):
    """TurboMetrics mirror (turbo-metrics/src/lib.rs:268-433): per-pair call and the submit-ahead frame loop give
    the same ordered score stream."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n = 320, 192, 11
    pairs = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=33) for i in range(n)]
    pitch, ch = pairs[0][2], pairs[0][3]
    dev = [(r.cuda(), d.cuda()) for r, d, _, _ in pairs]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    eng = tm.TurboMetrics(w, h, tm.PixelFormat.NV12, batch=4, ring=2)
    one = [eng.compute_one(F(r), F(d)).ssimulacra2 for r, d in dev]
    res = eng.compute_all((F(r) for r, _ in dev), (F(d) for _, d in dev))
    sub = eng.compute_all((F(r) for r, _ in dev), (F(d) for _, d in dev), tm.Options(every=3, skip=1, skip_dis=1, frames=7))
    eng.close()
    allp = res.ssimulacra2.scores
    assert one == allp and res.frame_count == n and res.ssimulacra2.stats.max == max(one)
    # Options (lib.rs:385-400): ref starts at 1, dis at 2; decode counts 0, 3, 6 are scored
    idx = tm.select_frames(n, n, tm.Options(every=3, skip=1, skip_dis=1, frames=7))
    assert idx == [(1, 2), (4, 5), (7, 8)] and sub.frame_count == 3
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=1, ring=1) as m:
        assert sub.ssimulacra2.scores == [m.compute_sync(F(dev[a][0]), F(dev[b][1])) for a, b in idx]
    expect = [oracle.ssimu2_yuv420(r.numpy(), d.numpy(), pitch, ch, w, h, 8)[0] for r, d, _, _ in pairs[:3]]
    assert max(abs(a - b) for a, b in zip(one, expect)) <= SCORE_ATOL


def test_strip_handoff_is_race_free_under_load():
    """The fused kernel continues the horizontal recursion from strip to strip through global-memory records released
    per (strip, band).  A record released before all three channels have written it shows up as a score that depends on
    timing: run many 1080p pairs (30 strips x 91 bands at scale 0) through full batches of all ring slots and require
    every repetition of a pair to be bit-equal, and equal to the two-kernel pipeline that has no hand-off."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, nd, n = 1920, 1080, 4, 96
    fr = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=5, device="cuda") for i in range(nd)]
    pitch, ch = fr[0][2], fr[0][3]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=16, ring=3) as m:
        ts = [m.compute(F(fr[i % nd][0]), F(fr[i % nd][1])) for i in range(n)]
        s = [m.get_score(t) for t in ts]
        nrm = [m.get_norms(t) for t in ts[:nd]]
    for i in range(n):
        assert s[i] == s[i % nd], (i, s[i], s[i % nd])
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=4, ring=1, pipeline="split") as m:
        ts = [m.compute(F(fr[i][0]), F(fr[i][1])) for i in range(nd)]
        for i, t in enumerate(ts):
            assert abs(m.get_score(t) - s[i]) < 1e-6
            np.testing.assert_allclose(m.get_norms(t), nrm[i], rtol=5e-7, atol=1e-12)


# ------------------------------------------------------------------------------------------ frame lifetime / decoder-style producers
def test_recycled_buffer_needs_and_gets_stream_ordering(oracle):
    """frames are read when their input group is launched, not at submit.  With batch > 1 a caller that
    overwrites a buffer right after submit must order the overwrite behind ssimu2_stream_wait_input; then every score is
    right although ONE pair of device buffers serves all pairs."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n = 320, 192, 13
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=44) for i in range(n)]
    expect = [oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0] for r, d in pairs]
    pinned = [(r.pin_memory(), d.pin_memory()) for r, d in pairs]
    side = torch.cuda.Stream()
    rg, dg = torch.empty_like(pairs[0][0], device="cuda"), torch.empty_like(pairs[0][1], device="cuda")
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=8, ring=2, input_group=1) as m:
        ts = []
        with torch.cuda.stream(side):
            for r, d in pinned:
                if ts:
                    m.wait_input(ts[-1], side)        # the previous pair's front-end has consumed the buffers
                rg.copy_(r, non_blocking=True)
                dg.copy_(d, non_blocking=True)
                ts.append(m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg), stream=side))
        got = [m.get_score(t) for t in ts]
    assert max(abs(a - b) for a, b in zip(got, expect)) <= SCORE_ATOL
    assert len(set(round(x, 6) for x in got)) == n


def test_decoder_shaped_producers_with_small_surface_pools(oracle):
    """SURVEY 8f row 4 surrogate (no libnvcuvid here): two "decoders" (reference / distorted), each with its own CUDA
    stream and a pool of M surfaces in NVDEC layout (pitch-aligned Y plane, CbCr at pitch * coded_height,
    cudarse-video/src/dec.rs:299-366), M far smaller than batch x ring.  A surface is "mapped" (filled on the decoder's
    stream), handed to the scorer on that stream (turbo-metrics/src/input_video.rs:429-440), and reused only behind
    ssimu2_stream_wait_input (dec.rs:277-287: valid until unmapped) -- no host synchronisation anywhere in the loop.
    1,000 pairs; the score stream must equal a straight run over frames that all stay resident."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, bits, n_distinct, n, M = 640, 360, 8, 25, 1000, 6
    src = [synth.make_pair_yuv420(w, h, bits, frame=i, seed=77, device="cuda") for i in range(n_distinct)]
    pitch, ch = src[0][2], src[0][3]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=16, ring=3) as m:
        ts = m.compute_batch([F(src[i % n_distinct][0]) for i in range(n)], [F(src[i % n_distinct][1]) for i in range(n)])
        straight = m.get_scores(ts)
    for i in range(3):
        so = oracle.ssimu2_yuv420(src[i][0].cpu().numpy(), src[i][1].cpu().numpy(), pitch, ch, w, h, bits)[0]
        assert abs(straight[i] - so) <= SCORE_ATOL
    dec = [torch.cuda.Stream(), torch.cuda.Stream()]
    pool = [[torch.zeros_like(src[0][0]) for _ in range(M)] for _ in range(2)]
    owner = [[None] * M for _ in range(2)]      # ticket that last read each surface
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=16, ring=3, input_group=2) as m:
        tickets = []
        for i in range(n):
            k = i % M
            for side in range(2):
                with torch.cuda.stream(dec[side]):
                    if owner[side][k] is not None:
                        m.wait_input(owner[side][k], dec[side])          # "unmap": the scorer is done with the surface
                    pool[side][k].copy_(src[i % n_distinct][side], non_blocking=True)   # "decode" into the surface
            # the pair is submitted on the reference decoder's stream, which first waits for the distorted one
            dec[0].wait_stream(dec[1])
            t = m.compute(F(pool[0][k]), F(pool[1][k]), stream=dec[0])
            owner[0][k] = owner[1][k] = t
            tickets.append(t)
        got = m.get_scores(range(tickets[0], tickets[-1] + 1))
    assert np.array_equal(got, straight), f"{int((got != straight).sum())} of {n} scores differ"


def test_large_batches_in_one_submission(oracle):
    """BASELINE.json configs[4] shape: hundreds of small pairs through ONE ssimu2_submit_batch with a batch far above the
    old 32-pair limit (the frame table lives in device memory now)."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w = h = 128
    n, nd = 300, 7
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=9) for i in range(nd)]
    expect = [oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0] for r, d in pairs]
    dev = [(r.cuda(), d.cuda()) for r, d in pairs]
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8, batch=256, ring=2) as m:
        assert m.info().batch == 256
        ts = m.compute_batch([tm.DeviceFrame.packed(dev[i % nd][0]) for i in range(n)], [tm.DeviceFrame.packed(dev[i % nd][1]) for i in range(n)])
        assert m.completed() <= ts.start + 256
        got = m.get_scores(ts)
        assert m.completed() == ts.stop
    for i in range(n):
        assert got[i] == got[i % nd]
    assert max(abs(got[i] - expect[i]) for i in range(nd)) <= SCORE_ATOL


# ------------------------------------------------------------------------------------------ multi-GPU behind the C ABI
def _shard_devices():
    n = torch.cuda.device_count()
    return list(range(n)) if n >= 2 else [0, 0]      # one GPU: two workers (two handles, two threads) on the same device


def test_shard_api_host_frames_ordered_stream(oracle):
    """ssimu2_shard_*: one handle + host thread per device inside the library, global tickets in submission order.  The
    score stream of the sharded run equals the single-handle run, including a trailing partial batch and interleaved
    partial fetches."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n, nd = 256, 144, 157, 12
    pairs = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=61) for i in range(nd)]
    pitch, ch = pairs[0][2], pairs[0][3]
    pinned = [(r.pin_memory(), d.pin_memory()) for r, d, _, _ in pairs]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    refs, diss = [F(pinned[i % nd][0]) for i in range(n)], [F(pinned[i % nd][1]) for i in range(n)]
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=8, ring=2) as m:
        single = m.get_scores(m.compute_from_cpu_batch(refs, diss))
    devs = _shard_devices()
    cur = torch.cuda.current_device()
    with tm.ShardedSsimulacra2(w, h, tm.PixelFormat.NV12, devices=devs, batch=8, ring=2) as sh:
        t1 = sh.submit_host(refs[:50], diss[:50])
        a = sh.get_scores(range(t1.start, t1.start + 20))      # forces a partial-batch fetch in the middle of the stream
        t2 = sh.submit_host(refs[50:], diss[50:])
        assert t2.start == 50 and t2.stop == n
        b = sh.get_scores(range(20, n))
        assert [sh.device_of(g) for g in (0, 7, 8, 16)] == [devs[0], devs[0], devs[1 % len(devs)], devs[2 % len(devs)]]
    assert torch.cuda.current_device() == cur
    got = np.concatenate([a, b])
    assert np.array_equal(got, single), f"{int((got != single).sum())} of {n} scores differ"
    so = oracle.ssimu2_yuv420(pairs[0][0].numpy(), pairs[0][1].numpy(), pitch, ch, w, h, 8)[0]
    assert abs(got[0] - so) <= SCORE_ATOL


def test_shard_api_device_frames(oracle):
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n, nd, batch = 256, 144, 70, 5, 4
    devs = _shard_devices()
    pairs = [synth.make_pair_srgb8(w, h, frame=i, seed=62) for i in range(nd)]
    expect = [oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0] for r, d in pairs]
    with tm.ShardedSsimulacra2(w, h, tm.PixelFormat.SRGB8, devices=devs, batch=batch, ring=3) as sh:
        # a frame must live on the device its ticket is routed to
        on = {d: [(r.to(f"cuda:{d}"), x.to(f"cuda:{d}")) for r, x in pairs] for d in set(devs)}
        refs = [tm.DeviceFrame.packed(on[sh.device_of(g)][g % nd][0]) for g in range(n)]
        diss = [tm.DeviceFrame.packed(on[sh.device_of(g)][g % nd][1]) for g in range(n)]
        for d in set(devs):
            torch.cuda.synchronize(d)
        got = sh.get_scores(sh.submit_device(refs, diss))
    for g in range(n):
        assert abs(got[g] - expect[g % nd]) <= SCORE_ATOL and got[g] == got[g % nd]


# ------------------------------------------------------------------------------------------ score-only mode
@pytest.mark.parametrize("kind,w,h", CASES + [("p016", 1920, 1080)])
def test_score_only_mode_gives_the_same_bits(oracle, kind, w, h):
    """SSIMU2_FLAG_SCORE_ONLY skips the filters and the SSIM' map of every (scale, channel) whose two SSIM' weights are zero
    (X and B at scale 0 -- 30 % of the FP32 work of the 4K path; the reference computes them and multiplies by 0.0,
    ssimulacra2-cuda/src/lib.rs:586-603).  What remains is computed by the same instructions in the same order, so the score
    must be BIT-equal to the full mode; the norms are not available."""
    tm = _tm()
    fmt, mk, r, d, (so, no, nso) = _make(kind, w, h, frame=1, seed=3, oracle=oracle) if (w, h) != (1920, 1080) or kind != "p016" else \
        _make(kind, w, h, frame=1, seed=3, oracle=oracle)
    rg, dg = r.cuda(), d.cuda()
    n = 9
    with tm.Ssimulacra2(w, h, fmt, batch=4, ring=2) as m:
        full = m.get_scores(m.compute_batch([mk(rg)] * n, [mk(dg)] * n))
    with tm.Ssimulacra2(w, h, fmt, batch=4, ring=2, score_only=True) as m:
        ts = m.compute_batch([mk(rg)] * n, [mk(dg)] * n)
        lite = m.get_scores(ts)
        with pytest.raises(tm.Ssimu2Error) as e:
            m.get_norms(ts[0])
        assert e.value.status == -2          # SSIMU2_E_UNSUPPORTED
        assert m.info().flags & 1
    assert np.array_equal(full, lite), (full[0], lite[0])
    assert abs(lite[0] - so) <= SCORE_ATOL


def test_score_only_strip_handoff_under_load():
    """The lite strips have their own warp-role map and barrier counts: many 1080p pairs through full batches of all ring
    slots, every repetition bit-equal to the full mode's score."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, nd, n = 1920, 1080, 4, 96
    fr = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=5, device="cuda") for i in range(nd)]
    pitch, ch = fr[0][2], fr[0][3]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    refs, diss = [F(fr[i % nd][0]) for i in range(n)], [F(fr[i % nd][1]) for i in range(n)]
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=16, ring=3) as m:
        full = m.get_scores(m.compute_batch(refs[:nd], diss[:nd]))
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=16, ring=3, score_only=True) as m:
        s = m.get_scores(m.compute_batch(refs, diss))
    for i in range(n):
        assert s[i] == full[i % nd], (i, s[i], full[i % nd])


def test_geometry_sweep_against_the_oracle(oracle):
    """Frame sizes on and around every tiling boundary of the pipeline (32-pixel front-end regions, 64-column strips, 12-row
    bands, the 8x8 stop of the pyramid, odd sizes at every scale), sRGB8 and NV12, default pipeline, against the oracle."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    ws = [8, 9, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 191, 193, 257]
    hs = [8, 9, 11, 12, 13, 20, 23, 24, 25, 31, 33, 36, 47, 49, 64, 71, 73]
    rng = np.random.default_rng(2)
    combos = [(int(rng.choice(ws)), int(rng.choice(hs))) for _ in range(40)] + [(64, 12), (65, 13), (63, 11), (128, 24), (129, 25),
                                                                                  (32, 32), (33, 33), (257, 73), (8, 8), (9, 9)]
    worst = 0.0
    for i, (w, h) in enumerate(combos):
        if i % 3 == 2 and w >= 16 and h >= 16:
            w2, h2 = w & ~1, h & ~1                      # 4:2:0 frames have even sizes
            rb, db, pitch, ch = synth.make_pair_yuv420(w2, h2, 8, frame=i, seed=21)
            so, no, nso = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, w2, h2, 8)
            fmt, mk, r, d, ww, hh = tm.PixelFormat.NV12, (lambda t, p=pitch, c=ch: tm.DeviceFrame.yuv420(t, p, c)), rb, db, w2, h2
        else:
            r, d = synth.make_pair_srgb8(w, h, frame=i, seed=21)
            so, no, nso = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
            fmt, mk, ww, hh = tm.PixelFormat.SRGB8, tm.DeviceFrame.packed, w, h
        with tm.Ssimulacra2(ww, hh, fmt, batch=2, ring=1) as m:
            rg, dg = r.cuda(), d.cuda()
            ts = m.compute_batch([mk(rg)] * 3, [mk(dg)] * 3)
            sc = m.get_scores(ts)
            assert m.info().nscales == nso, (ww, hh)
            assert sc[0] == sc[1] == sc[2], (ww, hh)
            worst = max(worst, _assert_norms(m.get_norms(ts[0]), no, sc[0], so))
    assert worst <= NORM_RTOL


def test_fast_lite_strips_with_many_waves_do_not_deadlock():
    """Regression (found by tools/soak.py).  An H warp of k_hv only OBSERVES the XYB tile barrier of the bands of the other
    parity; it did not take part in handing the tile slot back, so in a fast strip (score-only mode, small frames, several
    waves of work items) the tile of band j + 3 could land while it was still suspended in its wait for band j: its parity
    wait then never completed and the strip chain ran into the watchdog trap (about one launch in a few hundred).  Every H
    warp now arrives on the slot's free barrier.  256x254 NV12, batch 64, random batch fills and interleaved fetches."""
    import random
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 256, 254
    fr = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=4493, device="cuda") for i in range(3)]
    pitch, ch = fr[0][2], fr[0][3]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=1, ring=1) as m:
        want = [m.compute_sync(F(a), F(b)) for a, b, _, _ in fr]
    for rep in range(250):
        rnd = random.Random(rep)
        n = rnd.randint(64, 12 * 64)
        with tm.Ssimulacra2(w, h, tm.PixelFormat.NV12, batch=64, ring=1 + rep % 3, score_only=True, input_group=(0, 3)[rep & 1]) as m:
            ts, got = [], {}
            for i in range(n):
                ts.append(m.compute(F(fr[i % 3][0]), F(fr[i % 3][1])))
                if rnd.random() < 0.1:
                    j = rnd.randrange(len(ts))
                    got[j] = m.get_score(ts[j])
            sc = m.get_scores(range(ts[0], ts[-1] + 1))
        assert all(sc[j] == want[j % 3] for j in range(n)) and all(v == want[j % 3] for j, v in got.items()), rep


def test_small_frames_follow_the_cpu_reference_scale_rule(oracle):
    """with min(width, height) < 113 the pyramid stops early (cpu.rs:359 tests the size BEFORE each downscale) and the
    108 weights are consumed densely over the scales that exist (cpu.rs:842-854).  The reference's GPU op always runs six
    scales with fixed weight offsets (ssimulacra2-cuda/src/lib.rs:61-65, 586-603) and scores such frames differently; the
    contract of this library is the CPU implementation (BASELINE.json north_star).  This pins the choice."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 160, 96
    r, d = synth.make_pair_srgb8(w, h, frame=2, seed=12)
    so, no, nso = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8) as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg))
        score, norms, info = m.get_score(t), m.get_norms(t), m.info()
    assert info.nscales == nso == 5 and (info.width[4], info.height[4]) == (10, 6)
    _assert_norms(norms, no, score, so)
    # the same norms under the GPU op's fixed indexing WEIGHT[c*36 + s*6 + k] give a clearly different score
    wts = np.load(os.path.join(os.path.dirname(__file__), "golden", "weights108.npy"))
    v = float(np.dot(wts, np.abs(norms))) * 0.9562382616834844
    v = 6.248496625763138e-5 * v ** 3 + 2.326765642916932 * v - 0.020884521182843837 * v * v
    fixed = 100.0 - 10.0 * v ** 0.6276336467831387
    assert abs(fixed - score) > 1.0, (fixed, score)


def test_sharded_engine_frame_loop(oracle):
    """ShardedTurboMetrics.compute_all: the reference's frame loop with `Options`, over all GPUs from one process, equals the
    single-GPU engine's score stream."""
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h, n = 192, 128, 45
    pairs = [synth.make_pair_yuv420(w, h, 8, frame=i, seed=71) for i in range(n)]
    pitch, ch = pairs[0][2], pairs[0][3]
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    pinned = [(r.pin_memory(), d.pin_memory()) for r, d, _, _ in pairs]
    dev = [(r.cuda(), d.cuda()) for r, d, _, _ in pairs]
    opt = tm.Options(every=2, skip=1, frames=40)
    eng = tm.TurboMetrics(w, h, tm.PixelFormat.NV12, batch=4, ring=2)
    single = eng.compute_all((F(r) for r, _ in dev), (F(d) for _, d in dev), opt)
    eng.close()
    sh = tm.ShardedTurboMetrics(w, h, tm.PixelFormat.NV12, devices=_shard_devices(), batch=4, ring=2)
    multi = sh.compute_all((F(r) for r, _ in pinned), (F(d) for _, d in pinned), opt)
    sh.close()
    assert multi.frame_count == single.frame_count == len(tm.select_frames(n, n, opt))
    assert multi.ssimulacra2.scores == single.ssimulacra2.scores
    assert multi.ssimulacra2.stats == single.ssimulacra2.stats


def test_plain_c_client_of_the_abi(oracle, tmp_path):
    """examples/ssimu2_c_client.c: a C program with no CUDA or Python in it scores a pair from host memory through the C ABI
    (and, with n_devices > 1, through ssimu2_shard_* on every GPU); same score as the oracle / the Python path."""
    import subprocess
    from test_abi import _build_c_client
    tm = _tm()
    from turbo_metrics_b200 import synth
    w, h = 320, 200
    r, d = synth.make_pair_srgb8(w, h, frame=4, seed=19)
    (tmp_path / "r.rgb").write_bytes(r.numpy().tobytes())
    (tmp_path / "d.rgb").write_bytes(d.numpy().tobytes())
    so = oracle.ssimu2_srgb8(r.numpy(), d.numpy())[0]
    exe = _build_c_client(tmp_path)
    out = subprocess.run([exe, str(w), str(h), str(tmp_path / "r.rgb"), str(tmp_path / "d.rgb")], capture_output=True, text=True, check=True)
    score = float(out.stdout.strip())
    assert abs(score - so) <= SCORE_ATOL
    with tm.Ssimulacra2(w, h, tm.PixelFormat.SRGB8) as m:
        rg, dg = r.cuda(), d.cuda()
        assert m.compute_sync(tm.DeviceFrame.packed(rg), tm.DeviceFrame.packed(dg)) == score
    n = max(2, torch.cuda.device_count())
    if torch.cuda.device_count() >= 2:
        out = subprocess.run([exe, str(w), str(h), str(tmp_path / "r.rgb"), str(tmp_path / "d.rgb"), str(n)], capture_output=True, text=True, check=True)
        assert [float(x) for x in out.stdout.split()] == [score] * n


@pytest.mark.parametrize("kind", ["nv12_random", "p016_random10", "p016_random16"])
def test_arbitrary_sample_values(oracle, kind):
    """Every code value, in and out of the nominal range (super-black / super-white luma, saturated chroma), not only what the
    synthetic generator produces; `p016_random16` has non-zero low bits in the 16-bit containers, which sends every region of
    the fast front-end path back through the general path (frontend_region_fast returns false)."""
    tm = _tm()
    w, h = 256, 160
    bits = 8 if kind.startswith("nv12") else 16
    pitch, ch = (256, 160) if bits == 8 else (512, 160)
    rng = np.random.default_rng(3)

    def frame():
        if bits == 8:
            return torch.from_numpy(rng.integers(0, 256, pitch * ch * 3 // 2, dtype=np.uint8))
        v = rng.integers(0, 1024, pitch // 2 * ch * 3 // 2, dtype=np.uint16) << 6
        if kind == "p016_random16":
            v = rng.integers(0, 65536, v.size, dtype=np.uint16)
        return torch.from_numpy(v.view(np.uint8).copy())
    r = frame()
    d = r.clone()
    d[::7] = frame()[::7]          # a distorted copy: every 7th byte replaced
    so, no, _ = oracle.ssimu2_yuv420(r.numpy(), d.numpy(), pitch, ch, w, h, bits)
    fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
    F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
    with tm.Ssimulacra2(w, h, fmt, batch=2, ring=1) as m:
        rg, dg = r.cuda(), d.cuda()
        t = m.compute(F(rg), F(dg))
        _assert_norms(m.get_norms(t), no, m.get_score(t), so)
