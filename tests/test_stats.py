"""quick-stats `Stats::compute` and the CLI's per-frame / final output rows."""
import json
import math

import numpy as np

from turbo_metrics_b200.stats import Stats, format_frame, format_results, frame_rows


def test_stats_match_numpy_definitions():
    rng = np.random.default_rng(0)
    v = (80 + 10 * rng.standard_normal(301)).tolist()
    s = Stats.compute(v)
    a = np.array(v)
    assert s.min == a.min() and s.max == a.max()
    assert math.isclose(s.mean, a.mean(), rel_tol=1e-14)
    assert math.isclose(s.var, a.var(), rel_tol=1e-12) and math.isclose(s.sample_var, a.var(ddof=1), rel_tol=1e-12)
    assert math.isclose(s.stddev, a.std(), rel_tol=1e-12)
    for p, got in [(1, s.p1), (5, s.p5), (50, s.p50), (95, s.p95), (99, s.p99)]:
        assert math.isclose(got, np.percentile(a, p), rel_tol=1e-12)   # numpy's default is the same linear rule


def test_stats_edge_cases_follow_reference():
    s = Stats.compute([42.5])
    assert (s.min, s.max, s.mean, s.var, s.sample_var, s.p1, s.p99) == (42.5, 42.5, 42.5, 0.0, 0.0, 42.5, 42.5)
    s = Stats.compute([1.0, 3.0])
    assert s.var == 1.0 and s.sample_var == 2.0 and s.p50 == 2.0 and s.p1 == 1.02


def test_number_formats_follow_rust():
    """serde_json (ryu), `{}` and `{:?}` of f64 -- the three printers the CLI uses (turbo-metrics-cli/src/output.rs:42-142).
    Expected strings are what Rust prints for these values (ryu's pretty format: decimals for 1e-5 <= |x| < 1e16)."""
    from turbo_metrics_b200.stats import _debug_f64, _rust_f64, _serde_f64
    cases = [  # value, serde_json, Display, Debug
        (100.0, "100.0", "100", "100.0"), (80.6539462204559, "80.6539462204559", "80.6539462204559", "80.6539462204559"),
        (-3.25, "-3.25", "-3.25", "-3.25"), (0.0, "0.0", "0", "0.0"), (0.5, "0.5", "0.5", "0.5"),
        (1e-5, "0.00001", "0.00001", "1e-5"), (1.25e-7, "1.25e-7", "0.000000125", "1.25e-7"), (1e-7, "1e-7", "0.0000001", "1e-7"),
        (0.00012, "0.00012", "0.00012", "0.00012"), (1e16, "1e16", "10000000000000000", "1e16"),
        (123456789012.0, "123456789012.0", "123456789012", "123456789012.0"), (1.5e300, "1.5e300", "15" + "0" * 299, "1.5e300"),
    ]
    for v, sj, disp, dbg in cases:
        assert _serde_f64(v) == sj, (v, _serde_f64(v))
        assert _rust_f64(v) == disp, (v, _rust_f64(v))
        assert _debug_f64(v) == dbg, (v, _debug_f64(v))
        if sj != "null":
            assert float(_serde_f64(v)) == v and float(_rust_f64(v)) == v     # all of them round-trip
    assert _serde_f64(float("nan")) == "null" and _rust_f64(float("inf")) == "inf"


def test_output_rows():
    assert format_frame(80.65394622045590, "json-lines") == '{"ssimulacra2":80.6539462204559}'
    assert format_frame(100.0, "json-lines") == '{"ssimulacra2":100.0}'
    assert format_frame(100.0, "csv") == "100" and format_frame(-3.25, "csv") == "-3.25"
    assert frame_rows([1.5, 2.0], "csv") == ["ssimulacra2", "1.5", "2"]
    out = json.loads(format_results([80.0, 90.0, 100.0], "json-lines"))
    assert out["frame_count"] == 3 and out["ssimulacra2"]["p50"] == 90.0 and out["ssimulacra2"]["min"] == 80.0
    assert list(out["ssimulacra2"]) == ["min", "max", "mean", "var", "sample_var", "stddev", "sample_stddev", "p1", "p5",
                                       "p50", "p95", "p99"]


def test_final_blocks_have_the_reference_layout():
    """`--output json` = serde_json::to_string_pretty(&MetricsResults) (output.rs:96-98; struct layout
    turbo-metrics/src/lib.rs:56-84: frame_count, then per metric {scores, stats}, `None` metrics skipped);
    `--output default` = `{:#?}` of quick_stats::full::Stats (output.rs:83-93)."""
    js = format_results([80.0, 90.5], "json")
    assert js == """{
  "frame_count": 2,
  "ssimulacra2": {
    "scores": [
      80.0,
      90.5
    ],
    "stats": {
      "min": 80.0,
      "max": 90.5,
      "mean": 85.25,
      "var": 27.5625,
      "sample_var": 55.125,
      "stddev": 5.25,
      "sample_stddev": 7.424621202458749,
      "p1": 80.105,
      "p5": 80.525,
      "p50": 85.25,
      "p95": 89.975,
      "p99": 90.395
    }
  }
}"""
    assert json.loads(js)["ssimulacra2"]["scores"] == [80.0, 90.5]
    d = format_results([80.0, 90.5], "default")
    assert d.splitlines()[0] == "SSIMULACRA2: Stats {" and d.splitlines()[1] == "    min: 80.0," and d.splitlines()[-1] == "}"
    assert "    sample_stddev: 7.424621202458749," in d.splitlines()
    assert format_results([1.5, 2.0], "csv") == "ssimulacra2\n1.5\n2"
