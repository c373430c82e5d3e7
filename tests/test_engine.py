"""Host logic of the TurboMetrics mirror without a GPU: the frame loop's `Options` semantics
(turbo-metrics/src/lib.rs:385-400) and its submit-ahead / collect-in-order behaviour, against a fake scorer."""
import itertools

import pytest

import turbo_metrics_b200 as tm
from turbo_metrics_b200.engine import MetricsResults, Options, TurboMetrics, select_frames


class _FakeScorer:
    """Stands in for Ssimulacra2: the 'score' of a pair encodes which frames were paired."""

    def __init__(self):
        self.submitted, self.fetched, self.max_inflight = [], [], 0

    def compute(self, fref, fdis, stream=None):
        self.submitted.append((fref, fdis))
        self.max_inflight = max(self.max_inflight, len(self.submitted) - len(self.fetched))
        return len(self.submitted) - 1

    def get_score(self, t):
        assert t == len(self.fetched), "scores must be collected in submission order"
        self.fetched.append(t)
        a, b = self.submitted[t]
        return 1000.0 * a + b

    def flush(self):
        pass


def _engine(window):
    eng = TurboMetrics.__new__(TurboMetrics)
    eng.ssimulacra2, eng.window = _FakeScorer(), window
    return eng


@pytest.mark.parametrize("n_ref,n_dis", [(20, 20), (20, 13), (5, 40)])
@pytest.mark.parametrize("opt", [Options(), Options(every=3), Options(skip=2, skip_ref=1), Options(skip_dis=4, frames=6),
                                 Options(every=2, skip=1, skip_ref=2, skip_dis=1, frames=9), Options(every=4, frames=1),
                                 Options(skip=50)])
def test_compute_all_follows_the_reference_frame_selection(n_ref, n_dis, opt):
    eng = _engine(window=6)
    res = eng.compute_all(range(n_ref), range(n_dis), opt)
    expect = select_frames(n_ref, n_dis, opt)
    assert isinstance(res, MetricsResults) and res.frame_count == len(expect)
    if expect:
        assert res.ssimulacra2.scores == [1000.0 * a + b for a, b in expect]
        assert res.ssimulacra2.stats.min == min(res.ssimulacra2.scores)
    else:
        assert res.ssimulacra2 is None
    assert eng.ssimulacra2.max_inflight <= 6


def test_compute_all_consumes_generators_lazily_and_stops_at_frames():
    eng = _engine(window=4)
    seen = []

    def src():
        for i in itertools.count():
            seen.append(i)
            yield i
    res = eng.compute_all(src(), itertools.count(), Options(frames=5))
    assert res.frame_count == 5 and max(seen) <= 6     # an endless source is not read past the stop condition


def test_select_frames_hand_checked():
    # decode_count 0 is always scored; with every = 3: 0, 3, 6, ...; `frames` bounds the decode count (lib.rs:390-394)
    assert select_frames(10, 10, Options(every=3)) == [(0, 0), (3, 3), (6, 6), (9, 9)]
    assert select_frames(10, 10, Options(every=3, frames=7)) == [(0, 0), (3, 3), (6, 6)]
    assert select_frames(10, 10, Options(skip=2, skip_ref=1, frames=3)) == [(3, 2), (4, 3), (5, 4)]
    assert tm.FrameScores(ssimulacra2=1.0).ssimulacra2 == 1.0


class _FakeSharded:
    def __init__(self):
        self.pairs, self.fetched_upto, self.max_inflight = [], 0, 0

    def submit_host(self, refs, diss):
        t0 = len(self.pairs)
        self.pairs.extend(zip(refs, diss))
        self.max_inflight = max(self.max_inflight, len(self.pairs) - self.fetched_upto)
        return range(t0, len(self.pairs))

    def get_scores(self, tickets):
        import numpy as np
        assert tickets.start == self.fetched_upto, "ordered fetch"
        self.fetched_upto = tickets.stop
        return np.array([1000.0 * a + b for a, b in self.pairs[tickets.start:tickets.stop]])

    def flush(self):
        pass


@pytest.mark.parametrize("opt", [Options(), Options(every=3, skip=2), Options(frames=17), Options(skip=100)])
def test_sharded_frame_loop_follows_the_same_selection(opt):
    from turbo_metrics_b200.engine import ShardedTurboMetrics
    eng = ShardedTurboMetrics.__new__(ShardedTurboMetrics)
    eng.sharded, eng.chunk, eng.window = _FakeSharded(), 8, 24
    res = eng.compute_all(range(70), range(64), opt)
    expect = select_frames(70, 64, opt)
    assert res.frame_count == len(expect)
    assert (res.ssimulacra2.scores if expect else []) == [1000.0 * a + b for a, b in expect]
    assert eng.sharded.max_inflight <= 24 + 8
