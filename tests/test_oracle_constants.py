"""Pins the oracle: every constant the reference's CPU path uses, checked against the reference
source when it is mounted (/root/reference, build container only) and against committed golden
copies otherwise; plus the analytic properties the reference itself asserts at build time
(ssimulacra2-cuda-kernel/build.rs:82: IIR weights sum to 1)."""
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
CPU_RS = os.path.join(REFERENCE, "crates/ssimulacra2-cuda/examples/cpu.rs")
have_ref = os.path.exists(CPU_RS)


def _rust_floats(block):
    return [float(x.replace("_", "")) for x in re.findall(r"-?\d[\d_]*\.[\d_]*(?:e-?\d+)?", block)]


def test_srgb_lut_matches_golden(oracle):
    lut = oracle.srgb8_lut()
    gold = np.load(os.path.join(GOLD, "srgb8_lut.npy"))
    assert lut.dtype == np.float32 and np.array_equal(lut.view(np.uint32), gold.view(np.uint32))
    assert lut[0] == 0.0 and lut[255] == 1.0 and np.all(np.diff(lut) > 0)


@pytest.mark.skipif(not have_ref, reason="reference source not mounted")
def test_srgb_lut_matches_reference_tables(oracle):
    lut = oracle.srgb8_lut()
    for rel, pat in [("crates/ssimulacra2-cuda/examples/cpu.rs", r"FROM_SRGB8_TABLE: \[f32; 256\] = \[(.*?)\];"),
                     ("crates/cuda-colorspace-kernel/src/srgb.rs", r"SRGB8_TO_LINEARF32_LUT: \[f32; 256\] = \[(.*?)\];"),
                     ("crates/ssimulacra2-cuda-kernel/src/srgb.rs", r"\[f32; 256\] = \[(.*?)\];")]:
        src = open(os.path.join(REFERENCE, rel)).read()
        vals = np.array(_rust_floats(re.search(pat, src, re.S).group(1)), dtype=np.float32)
        assert vals.size == 256
        assert np.array_equal(vals.view(np.uint32), lut.view(np.uint32)), rel


def test_weights_match_golden(oracle):
    w = oracle.weights()
    gold = np.load(os.path.join(GOLD, "weights108.npy"))
    assert np.array_equal(w, gold)
    assert w.shape == (108,) and (w >= 0).all() and np.count_nonzero(w) == 52


@pytest.mark.skipif(not have_ref, reason="reference source not mounted")
def test_weights_match_reference(oracle):
    w = oracle.weights()
    for rel in ["crates/ssimulacra2-cuda/examples/cpu.rs", "crates/ssimulacra2-cuda/src/lib.rs"]:
        src = open(os.path.join(REFERENCE, rel)).read()
        vals = np.array(_rust_floats(re.search(r"const WEIGHT: \[f64; 108\] = \[(.*?)\];", src, re.S).group(1)))
        assert vals.size == 108 and np.array_equal(vals, w), rel


@pytest.mark.skipif(not have_ref, reason="reference source not mounted")
def test_filter_and_opsin_constants_match_reference(oracle):
    src = open(CPU_RS).read()

    def const(name):
        m = re.search(r"const %s: f32 =\s*([^;]+);" % name, src)
        return np.float32(float(m.group(1).replace("_f32", "").replace("f32", "").replace("_", "")))
    rg = oracle.rg_constants()
    names = ["MUL_IN_1", "MUL_IN_3", "MUL_IN_5", "MUL_PREV_1", "MUL_PREV_3", "MUL_PREV_5", "MUL_PREV2_1", "MUL_PREV2_3",
             "MUL_PREV2_5"]
    for v, n in zip(rg, names):
        assert v == const(n), n
    # the vertical form uses the negated feedback constants (cpu.rs:934-939)
    for n in ["1", "3", "5"]:
        assert const("VERT_MUL_PREV_" + n) == -const("MUL_PREV_" + n)
        assert const("VERT_MUL_IN_" + n) == const("MUL_IN_" + n)
    op = oracle.opsin_constants()
    assert op[0] == const("K_M00") and op[2] == const("K_M02") and op[3] == const("K_M10")
    assert op[6] == const("K_M20") and op[7] == const("K_M21") and op[9] == const("K_B0")
    assert op[1] == np.float32(1.0) - op[2] - op[0]


def test_recursive_gaussian_is_unit_gain_symmetric_fir(oracle):
    """build.rs:82 asserts the weights sum to 1; the impulse response is the zero-padded symmetric
    9-tap kernel."""
    n = 64
    imp = np.zeros((1, n), np.float32)
    imp[0, 32] = 1.0
    out = oracle.blur_horizontal(imp)[0].astype(np.float64)
    assert abs(out.sum() - 1.0) < 2e-6
    k = out[32 - 4:32 + 5]
    assert np.allclose(k, k[::-1], atol=2e-7)
    assert np.allclose(k[4:], [0.264621, 0.212929, 0.109335, 0.036011, 0.009414], atol=2e-6)
    assert np.abs(out[:32 - 5]).max() < 1e-6 and np.abs(out[32 + 6:]).max() < 1e-6
    # vertical form gives the same response
    outv = oracle.blur_vertical(imp.T.copy())[:, 0]
    assert np.allclose(outv, out, atol=1e-6)


def test_bt709_luma_constants(oracle):
    kr, kb = oracle.matrix_kr_kb("bt709")
    assert abs(kr - 0.2126) < 2e-4 and abs(kb - 0.0722) < 2e-4  # SURVEY: kr=0.21264, kb=0.07219
    y, r, b, g1, g2 = oracle.yuv_coefficients("bt709", 8, False)
    assert y == np.float32(1.0) / np.float32(219.0)
    assert abs(r * 224 - 2 * (1 - kr)) < 1e-6 and abs(b * 224 - 2 * (1 - kb)) < 1e-6
    y16 = oracle.yuv_coefficients("bt709", 16, False)[0]
    assert y16 == np.float32(1.0) / np.float32(219 << 8)


def test_identical_images_score_100(oracle):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(64, 80, 3), dtype=np.uint8)
    s, norms, ns = oracle.ssimu2_srgb8(img, img)
    # cpu.rs:359 tests the size BEFORE downscaling: 64x80, 32x40, 16x20, 8x10 and 4x5 are all scored
    assert s == 100.0 and ns == 5
    assert np.all(norms == 0.0)


def test_score_polynomial(oracle):
    """Msssim::score on hand-made norms (cpu.rs:856-868)."""
    norms = np.zeros(108)
    assert oracle.score_from_norms(norms) == 100.0
    norms[9] = 0.5  # weight 1.1041726426657346
    s = 0.5 * 1.1041726426657346 * 0.9562382616834844
    s = 6.248496625763138e-5 * s ** 3 + 2.326765642916932 * s - 0.020884521182843837 * s * s
    expect = 100.0 - 10.0 * s ** 0.6276336467831387
    assert abs(oracle.score_from_norms(norms) - expect) < 1e-12


def test_golden_scores(oracle):
    """Oracle outputs committed from this container (tools/gen_golden.py): guards the oracle against a
    different libm / compiler on another box."""
    import torch  # noqa: F401
    from turbo_metrics_b200 import synth
    gold = np.load(os.path.join(GOLD, "oracle_cases.npz"))
    r, d = synth.make_pair_srgb8(96, 72, frame=0, seed=7)
    s, norms, _ = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
    assert np.array_equal(norms, gold["srgb8_96x72_norms"]) and s == float(gold["srgb8_96x72_score"])
    rb, db, pitch, ch = synth.make_pair_yuv420(128, 96, 8, frame=1, seed=7)
    s, norms, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, 128, 96, 8)
    assert np.array_equal(norms, gold["nv12_128x96_norms"]) and s == float(gold["nv12_128x96_score"])
    rb, db, pitch, ch = synth.make_pair_yuv420(128, 96, 16, frame=2, seed=7)
    s, norms, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, 128, 96, 16)
    assert np.array_equal(norms, gold["p016_128x96_norms"]) and s == float(gold["p016_128x96_score"])
