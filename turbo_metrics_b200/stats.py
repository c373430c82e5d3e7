"""Score-stream consumers of the reference, restated:

* `Stats.compute`   -- quick-stats/src/lib.rs:22-97 `full::Stats::compute(&[f64])`.
* `format_frame` / `format_results` -- turbo-metrics-cli/src/output.rs:42-142 for `--output json-lines | json | csv`
  with only the ssimulacra2 metric enabled (serde skips the `None` metrics, lib.rs:114-123).
"""
from __future__ import annotations

import json
import math
from dataclasses import asdict, dataclass
from typing import List, Sequence


@dataclass
class Stats:
    min: float
    max: float
    mean: float
    var: float          # population variance
    sample_var: float
    stddev: float
    sample_stddev: float
    p1: float
    p5: float
    p50: float
    p95: float
    p99: float

    @staticmethod
    def compute(values: Sequence[float]) -> "Stats":
        values = [float(x) for x in values]
        assert values, "Stats::compute indexes sorted[0] (quick-stats/src/lib.rs:25)"
        srt = sorted(values)
        n = len(values)
        mean = _sum_in_order(srt) / n   # lib.rs:27: sum of the SORTED values
        var = _compute_var(values, mean, False)
        svar = _compute_var(values, mean, True)
        return Stats(srt[0], srt[-1], mean, var, svar, math.sqrt(var), math.sqrt(svar),
                     _percentile(srt, 1.0), _percentile(srt, 5.0), _percentile(srt, 50.0), _percentile(srt, 95.0),
                     _percentile(srt, 99.0))


def _sum_in_order(xs):
    s = 0.0
    for x in xs:
        s += x
    return s


def _compute_var(values, mean, sample):
    """lib.rs:79-97: 0.0 for fewer than 2 values; accumulation in the ORIGINAL order."""
    if len(values) < 2:
        return 0.0
    v = 0.0
    for s in values:
        x = s - mean
        v += x * x
    return v / (len(values) - 1 if sample else len(values))


def _percentile(srt, pct):
    """lib.rs:55-77: linear interpolation between closest ranks."""
    if len(srt) == 1:
        return srt[0]
    if pct == 100.0:
        return srt[-1]
    rank = (pct / 100.0) * (len(srt) - 1)
    lrank = math.floor(rank)
    d = rank - lrank
    n = int(lrank)
    return srt[n] + (srt[n + 1] - srt[n]) * d


def _rust_f64(x: float) -> str:
    """`{}` of an f64 as Rust / serde_json print it: shortest round-trip digits, always with a fraction."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    r = repr(float(x))
    if "e" in r or "E" in r:  # Rust never uses exponent notation for Display
        r = format(float(x), "f").rstrip("0")
        if r.endswith("."):
            r += "0"
    return r


def format_frame(score: float, fmt: str = "json-lines") -> str:
    """One frame's output (output.rs:42-77)."""
    if fmt == "json-lines":
        return '{"ssimulacra2":' + _rust_f64(score) + "}"
    if fmt == "csv":
        s = _rust_f64(score)
        return s[:-2] if s.endswith(".0") else s  # `write!("{}", x)` prints 100 for 100.0
    raise ValueError(f"no per-frame output in format {fmt!r}")


def csv_header() -> str:
    return "ssimulacra2"  # output.rs:27-37


def format_results(scores: Sequence[float], fmt: str = "json-lines") -> str:
    """Final block (output.rs:79-142): json-lines prints `MetricsStats` (frame_count + stats per metric)."""
    st = asdict(Stats.compute(scores))
    if fmt == "json-lines":
        body = ",".join(f'"{k}":{_rust_f64(v)}' for k, v in st.items())
        return '{"frame_count":%d,"ssimulacra2":{%s}}' % (len(scores), body)
    if fmt == "json":
        return json.dumps({"frame_count": len(scores), "ssimulacra2": {"scores": list(map(float, scores)), "stats": st}}, indent=2)
    raise ValueError(fmt)


def frame_rows(scores: Sequence[float], fmt: str) -> List[str]:
    rows = [csv_header()] if fmt == "csv" else []
    return rows + [format_frame(s, fmt) for s in scores]
