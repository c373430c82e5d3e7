"""The bench line of record (profiles/r2_bench_4k.json, written by `python bench.py` on a B200) carries every key the bench
contract names, with consistent arithmetic.  CPU-only: guards the contract against drift when bench.py is edited."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def line():
    return json.load(open(os.path.join(ROOT, "profiles", "r2_bench_4k.json")))


def test_base_contract_keys(line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["unit"] == "pairs/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert line["data"] == "synthetic" and "workload" in line["config"] and "model" not in line["config"]
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0
    # value = pairs of the timed region / its duration
    pairs = line["config"]["pairs_per_step_per_gpu"] * line["n_gpus"] * line["steps"]
    assert abs(line["value"] - pairs / (line["ms_per_step"] * line["steps"] / 1e3)) / line["value"] < 1e-6
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in line["clocks"]
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_e2e_block(line):
    e = line["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    # both frames of every pair cross the link: 2 x 24.9 MB (4K P016) per pair
    assert e["h2d_bytes_per_step"] == 2 * (3840 * 2 * 2160 * 3 // 2) * line["config"]["pairs_per_step_per_gpu"]
    assert 0 < e["value"] < line["value"]                   # host buffers cannot beat device-resident frames
    assert abs(e["h2d_gbs"] - e["value"] * e["h2d_bytes_per_step"] / line["config"]["pairs_per_step_per_gpu"] / 1e9) < 1e-6


def test_roofline_and_cpu_baseline_blocks(line):
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    # achieved = B_alg x measured pairs/s per GPU (the contract figure)
    assert abs(r["achieved"] - r["alg_bytes_per_pair"] * line["value"] / line["n_gpus"] / 1e9) / r["achieved"] < 1e-9
    assert r["alg_bytes_per_pair"] == 1_625_195_520 and r["io_bytes_per_pair"] == 49_766_408
    assert "STALE" not in r["traffic_source"]               # the committed ncu capture is of the committed kernels
    for k in ("k_frontend2", "k_hv", "k_finalize"):
        assert r["kernels"][k]["ms_per_launch"] > 0
    c = line["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference") and c["unit"] == line["unit"]


def test_parity_and_sub_results(line):
    p = line["parity"]
    assert p["dscore"] <= 0.01 and p["max_rel_norm"] <= 1e-4
    for wl in ("1080p", "512", "1080p_srgb8"):
        w = line["workloads"][wl]
        assert w["value"] > 0 and w["e2e"]["value"] > 0 and w["parity"]["dscore"] <= 0.01 and w["parity"]["max_rel_norm"] <= 1e-4
    so = line["score_only_mode"]
    assert so["scores_bit_equal_to_full_mode"] is True and so["value"] > line["value"]


def test_committed_traffic_matches_the_committed_kernels():
    """profiles/traffic_r2.json is keyed to a hash of the kernel sources: a kernel edit without a new capture must be visible."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
    assert t["4k"]["source_hash"] == bench.kernel_source_hash()
