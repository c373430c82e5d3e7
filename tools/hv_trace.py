"""Timing trace of one k_hv work item (development aid; needs a library built with -DKX_TRACE, see ssimu2_kernels.cuh).
usage: SSIMU2_SO=turbo_metrics_b200/var_trace.so python tools/hv_trace.py [score_only]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import turbo_metrics_b200 as tm
import turbo_metrics_b200._lib as _l
_l.SO_PATH = os.path.abspath(os.environ['SSIMU2_SO'])
from turbo_metrics_b200 import synth
so = len(sys.argv) > 1 and sys.argv[1] == "score_only"
w, h, bits, batch = 3840, 2160, 10, 16
frames = [synth.make_pair_yuv420(w, h, bits, frame=i, seed=1, device="cuda") for i in range(8)]
pitch, ch = frames[0][2], frames[0][3]
F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
m = tm.Ssimulacra2(w, h, tm.PixelFormat.P016, batch=batch, ring=1, score_only=so)
for rep in range(2):
    ts = [m.compute(F(frames[i % 8][0]), F(frames[i % 8][1])) for i in range(32)]
    m.flush()
    sc = [m.get_score(t) for t in ts]
print("kernel ms", m.last_batch_ms())
buf = np.zeros((16, 32, 12), dtype=np.uint64)
fn = _l.lib().ssimu2_debug_hv_trace
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
t = buf.astype(np.int64)
t0 = t[t > 0].min()
r = lambda v: (v - t0) if v > 0 else -1
if not so:
    H = {0: (0, 4), 1: (1, 5), 2: (2, 6)}; Va = {0: 8, 1: 9, 2: 10}; Vb = {0: 3, 1: 7, 2: 11}
else:
    H = {1: (0, 4), 0: (1, 5)}; Va = {1: 2}; Vb = {1: 6, 0: 3, 2: 7}
print("band | tma_issue | per H set: in_full  pre_hbfree  post_hbfree  end | Va: start end | Vb: start end   (SM clocks, relative)")
for b in range(32):
    j = 60 + b
    line = f"{j:4d} | {r(t[15, b, 0]):7d} |"
    for c, (wa, wb) in H.items():
        wq = wa if (j & 1) == 0 else wb
        line += f" H{c}:" + " ".join(f"{r(t[wq, b, e]):7d}" for e in range(4)) + " |"
    for c, wv in Va.items():
        line += f" Va{c}: {r(t[wv, b, 0]):7d} {r(t[wv, b, 1]):7d} |"
    for c, wv in Vb.items():
        line += f" Vb{c}: {r(t[wv, b, 0]):7d} {r(t[wv, b, 1]):7d} |"
    print(line)
print("per sub-band, relative to the warp's band start: [pre-sync, post-sync, done] x 3, band end")
for b in range(8, 14):
    for name, d in (("Va", Va), ("Vb", Vb)):
        for c, wv in d.items():
            s0 = t[wv, b, 0]
            print(f"  band {60 + b} {name}{c}: " + " ".join(str(int(t[wv, b, e] - s0)) for e in range(2, 11)) + f" | end {int(t[wv, b, 1] - s0)}")
