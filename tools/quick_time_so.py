"""quick_time.py in score-only mode (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbo_metrics_b200 as tm
if os.environ.get('SSIMU2_SO'):   # development aid: time another build of the library
    import turbo_metrics_b200._lib as _l
    _l.SO_PATH = os.path.abspath(os.environ['SSIMU2_SO'])
from turbo_metrics_b200 import synth
w, h, bits = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch, ring, npairs, ndistinct = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
fmt = tm.PixelFormat.NV12 if bits == 8 else tm.PixelFormat.P016
frames = []
for i in range(ndistinct):
    rb, db, pitch, ch = synth.make_pair_yuv420(w, h, bits, frame=i, seed=1, device="cuda")
    frames.append((rb, db))
F = lambda t: tm.DeviceFrame.yuv420(t, pitch, ch)
m = tm.Ssimulacra2(w, h, fmt, batch=batch, ring=ring, score_only=True)
for rep in range(2):
    ts = [m.compute(F(frames[i % ndistinct][0]), F(frames[i % ndistinct][1])) for i in range(npairs)]
    m.flush()
    scores = [m.get_score(t) for t in ts]
    print(f"rep {rep}: kernel ms/batch {m.last_batch_ms()}")
print(scores[:2])
