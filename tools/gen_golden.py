"""Writes tests/golden/{weights108.npy, oracle_cases.npz} from the oracle in THIS container.
(The reference ships no golden vectors for the path -- SURVEY.md section 4 -- so these pin the
oracle's own outputs: a different libm or compiler on another machine shows up as a test failure
instead of silently moving the parity target.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from turbo_metrics_b200 import synth  # noqa: E402

gold = os.path.join(ROOT, "tests", "golden")
np.save(os.path.join(gold, "weights108.npy"), oracle.weights())
out = {}
r, d = synth.make_pair_srgb8(96, 72, frame=0, seed=7)
s, n, _ = oracle.ssimu2_srgb8(r.numpy(), d.numpy())
out["srgb8_96x72_norms"], out["srgb8_96x72_score"] = n, s
rb, db, pitch, ch = synth.make_pair_yuv420(128, 96, 8, frame=1, seed=7)
s, n, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, 128, 96, 8)
out["nv12_128x96_norms"], out["nv12_128x96_score"] = n, s
rb, db, pitch, ch = synth.make_pair_yuv420(128, 96, 16, frame=2, seed=7)
s, n, _ = oracle.ssimu2_yuv420(rb.numpy(), db.numpy(), pitch, ch, 128, 96, 16)
out["p016_128x96_norms"], out["p016_128x96_score"] = n, s
np.savez(os.path.join(gold, "oracle_cases.npz"), **out)
print({k: (v if np.ndim(v) == 0 else v.shape) for k, v in out.items()})
