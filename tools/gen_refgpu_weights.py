"""Writes baseline/refgpu/weights.inc from the oracle's WEIGHT table (the constants of the metric, cpu.rs:729-838)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle
oracle.build()
w = oracle.weights()
with open(os.path.join(ROOT, "baseline", "refgpu", "weights.inc"), "w") as f:
    f.write("// SSIMULACRA2 weights in [channel][scale][L1,L4][ssim,artifact,detail] order (public constants of the metric;\n"
            "// written by tools/gen_refgpu_weights.py from the oracle's table)\n")
    for i in range(0, 108, 6):
        f.write(", ".join(repr(float(x)) for x in w[i:i + 6]) + ",\n")
