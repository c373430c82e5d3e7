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


def test_output_rows():
    assert format_frame(80.65394622045590, "json-lines") == '{"ssimulacra2":80.6539462204559}'
    assert format_frame(100.0, "json-lines") == '{"ssimulacra2":100.0}'
    assert format_frame(100.0, "csv") == "100" and format_frame(-3.25, "csv") == "-3.25"
    assert frame_rows([1.5, 2.0], "csv") == ["ssimulacra2", "1.5", "2"]
    out = json.loads(format_results([80.0, 90.0, 100.0], "json-lines"))
    assert out["frame_count"] == 3 and out["ssimulacra2"]["p50"] == 90.0 and out["ssimulacra2"]["min"] == 80.0
    assert list(out["ssimulacra2"]) == ["min", "max", "mean", "var", "sample_var", "stddev", "sample_stddev", "p1", "p5",
                                       "p50", "p95", "p99"]
