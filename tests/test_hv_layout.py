"""Host-side check of k_hv's shared-memory layout: with the lane map, plane order and ones-row offsets that
ssimu2_kernels.cuh actually contains, every 128-bit access of the horizontal scan costs the ideal 4 wavefronts
(profiles/: 7 per store and 5-6 per load before the layout was chosen this way).  The model covers the H scan of the full
role map only: the 64-bit column accesses of the V warps, the mu hand-off ring and the HB warps of lite strips are not modelled
(ncu still counts ~20 M bank conflicts per launch of 16 4K pairs against 604 M shared-memory wavefronts, 3 %)."""
import os

from conftest import ROOT
from tools import banksim

SRC = os.path.join(ROOT, "turbo_metrics_b200", "csrc", "ssimu2_kernels.cuh")


def test_lane_map_is_a_bijection_onto_quantities_and_row_pairs():
    P = banksim.kernel_params(SRC)
    assert sorted(P["order"]) == [0, 1, 2, 3, 4]
    for q in range(5):
        assert sorted(P["rowmap"][q]) == [0, 1, 2, 3, 4, 5]
    assert sorted(P["slot"].values()) == [0, 1, 2, 3, 4]


def test_scan_accesses_are_conflict_free():
    P = banksim.kernel_params(SRC)
    for ch in range(3):
        r = banksim.simulate(P, ch)
        assert all(v == 4 for v in r.values()), (ch, r)


def test_model_sees_the_conflicts_of_the_naive_layout():
    P = banksim.kernel_params(SRC)
    naive = dict(P, order=[0, 1, 2, 3, 4], rowmap={q: list(range(6)) for q in range(5)}, slot={q: q for q in range(5)},
                 ones={"A": {0: 0, 1: 0, 2: 0}, "B": {0: 0, 1: 0, 2: 0}})
    r = banksim.simulate(naive, 0, ones_base=banksim.kernel_params(SRC)["kXR"] * 0 + 1000000)
    assert r["sa"] == 7 and r["xa"] == 6   # what ncu measured on the earlier kernel (profiles/r1_v6_k_hv_ncu_summary.md era)
