"""Score-stream consumers of the reference, restated:

* `Stats.compute`   -- quick-stats/src/lib.rs:22-97 `full::Stats::compute(&[f64])`.
* `format_frame` / `format_results` -- turbo-metrics-cli/src/output.rs:42-142 for `--output json-lines | json | csv`
  with only the ssimulacra2 metric enabled (serde skips the `None` metrics, lib.rs:114-123).
"""
from __future__ import annotations

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


def _digits(x: float):
    """Shortest round-trip decimal digits of a finite non-zero f64 (what ryu / Rust's Grisu+Dragon produce; Python's repr is the
    same digit string) -> (sign, digits, kk) with value = 0.digits x 10^kk."""
    r = repr(abs(float(x)))
    mant, _, exp = r.partition("e")
    e = int(exp) if exp else 0
    ip, _, fp = mant.partition(".")
    if fp == "0":
        fp = ""
    digs = (ip + fp).lstrip("0")
    lead = len(ip + fp) - len((ip + fp).lstrip("0"))   # zeros stripped in front (0.00xyz)
    kk = len(ip) - lead + e
    digs = digs.rstrip("0") or "0"
    return ("-" if x < 0 or (x == 0 and math.copysign(1.0, x) < 0) else ""), digs, kk


def _rust_f64(x: float) -> str:
    """`{}` (Display) of an f64 in Rust: shortest round-trip digits, never an exponent, no forced fraction
    (100.0 prints `100`) -- what the CSV writer uses (output.rs:56-61 `write!(&mut fmt_buffer, "{}", x)`)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == 0:
        return "-0" if math.copysign(1.0, x) < 0 else "0"
    sign, d, kk = _digits(x)
    if kk <= 0:
        return sign + "0." + "0" * (-kk) + d
    if kk >= len(d):
        return sign + d + "0" * (kk - len(d))
    return sign + d[:kk] + "." + d[kk:]


def _serde_f64(x: float) -> str:
    """An f64 as serde_json writes it (ryu's pretty format): shortest round-trip digits, `100.0` for integral values,
    plain decimals for 1e-5 <= |x| < 1e16, otherwise `1.5e-7` / `1e16` style exponents; non-finite values become `null`."""
    if x != x or x in (float("inf"), float("-inf")):
        return "null"
    if x == 0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign, d, kk = _digits(x)
    n = len(d)
    if n <= kk <= 16:
        return sign + d + "0" * (kk - n) + ".0"
    if 0 < kk <= 16:
        return sign + d[:kk] + "." + d[kk:]
    if -5 < kk <= 0:
        return sign + "0." + "0" * (-kk) + d
    e = kk - 1
    return sign + (d if n == 1 else d[0] + "." + d[1:]) + "e" + str(e)


def _debug_f64(x: float) -> str:
    """`{:?}` of an f64 in Rust (the `{:#?}` of `Stats`, output.rs:83-93): like Display but integral values keep `.0` and
    magnitudes below 1e-5 or from 1e16 up use exponent notation."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == 0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign, d, kk = _digits(x)
    if abs(x) < 1e-4 or abs(x) >= 1e16:   # core::fmt::float::float_to_general_debug
        e = kk - 1
        return sign + (d if len(d) == 1 else d[0] + "." + d[1:]) + "e" + str(e)
    r = _rust_f64(abs(x))
    return sign + (r if "." in r else r + ".0")


STAT_FIELDS = ["min", "max", "mean", "var", "sample_var", "stddev", "sample_stddev", "p1", "p5", "p50", "p95", "p99"]


def format_frame(score: float, fmt: str = "json-lines") -> str:
    """One frame's output (output.rs:42-77): serde_json of `FrameScores` (the `None` metrics are skipped, lib.rs:114-123)
    or one CSV field."""
    if fmt == "json-lines":
        return '{"ssimulacra2":' + _serde_f64(score) + "}"
    if fmt == "csv":
        return _rust_f64(score)
    if fmt in ("default", "json"):
        return ""          # output.rs:44-49: nothing per frame
    raise ValueError(f"unknown output format {fmt!r}")


def csv_header() -> str:
    return "ssimulacra2"  # output.rs:27-37


def format_results(scores: Sequence[float], fmt: str = "json-lines") -> str:
    """Final block (output.rs:79-142)."""
    st = asdict(Stats.compute(scores))
    assert list(st) == STAT_FIELDS
    if fmt == "default":      # println!("SSIMULACRA2: {:#?}", results.stats)
        return "SSIMULACRA2: Stats {\n" + "".join(f"    {k}: {_debug_f64(v)},\n" for k, v in st.items()) + "}"
    if fmt == "json-lines":   # serde_json::to_string(&MetricsStats::from(results)), lib.rs:86-110
        body = ",".join(f'"{k}":{_serde_f64(v)}' for k, v in st.items())
        return '{"frame_count":%d,"ssimulacra2":{%s}}' % (len(scores), body)
    if fmt == "json":         # serde_json::to_string_pretty(&MetricsResults), lib.rs:56-84: 2-space indent, one array item per line
        lines = ["{", f'  "frame_count": {len(scores)},', '  "ssimulacra2": {']
        if len(scores):
            lines += ['    "scores": ['] + [f"      {_serde_f64(float(v))}" + ("," if i + 1 < len(scores) else "") for i, v in enumerate(scores)]
            lines += ["    ],"]
        else:
            lines += ['    "scores": [],']
        lines += ['    "stats": {']
        items = list(st.items())
        lines += [f'      "{k}": {_serde_f64(v)}' + ("," if i + 1 < len(items) else "") for i, (k, v) in enumerate(items)]
        lines += ["    }", "  }", "}"]
        return "\n".join(lines)
    if fmt == "csv":          # header + one row per frame (output.rs:103-139)
        return "\n".join(frame_rows(scores, "csv"))
    raise ValueError(fmt)


def frame_rows(scores: Sequence[float], fmt: str) -> List[str]:
    rows = [csv_header()] if fmt == "csv" else []
    return rows + [format_frame(s, fmt) for s in scores]
