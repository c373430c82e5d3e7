"""The reference-design timing harness (baseline/refgpu) computes the same metric: its score agrees with the
product's to the tolerance the reference itself uses between its GPU and CPU paths (0.25, examples/compare.rs:70-74).
Its arithmetic is the reference GPU path's (hardware powf, vertical pass first, f32 tails), so this is NOT a parity test."""
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,w,h", [(8, 640, 360), (16, 960, 540), (8, 203, 131)])
def test_reference_design_scores_like_the_product(bits, w, h):
    import turbo_metrics_b200 as tm
    from turbo_metrics_b200 import synth
    from baseline.refgpu import refgpu
    if not os.path.exists(refgpu.SO_PATH):
        pytest.skip("librefgpu.so not built (NPP missing?)")
    fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
    with tm.Ssimulacra2(w, h, fmt, batch=2, ring=1) as m, refgpu.RefGpu(w, h, bits) as r:
        assert r.info()["kernel_nodes"] >= 150          # "200+" / "305 launches" per pair (ssimulacra2-cuda/README.md:9, lib.rs:26)
        for frame in range(3):
            rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=frame, seed=9, device="cuda")
            t = m.compute(tm.DeviceFrame.yuv420(rb, pitch, ch), tm.DeviceFrame.yuv420(db, pitch, ch))
            ours = m.get_score(t)
            theirs, _ = r.compute(rb, db, pitch, ch)
            assert abs(ours - theirs) <= 0.25, (frame, ours, theirs)
            same, _ = r.compute(rb, rb, pitch, ch)
            assert same == 100.0
