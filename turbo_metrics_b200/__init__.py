"""turbo_metrics_b200 -- B200-native SSIMULACRA2 frame-pair scorer.

Drop-in for the SSIMULACRA2 hot path of Gui-Yom/turbo-metrics (`Ssimulacra2::compute` on device
frames -> per-frame score stream).  The compute path is libssimu2_b200.so (hand-written CUDA for
sm_100a behind the C ABI of include/ssimu2_b200.h); this package is the host-side mirror of the
reference's operator interface.  (The directory is spelled with an underscore because
`turbo-metrics_b200` is not an importable Python name.)
"""
from ._lib import SO_PATH, Ssimu2Error  # noqa: F401
from .engine import (FrameScores, MetricAggregate, MetricsResults, Options, ShardedTurboMetrics, TurboMetrics,  # noqa: F401
                     gather_scores, select_frames, shard_range)
from .stats import Stats  # noqa: F401
from .ssimulacra2 import ColorMatrix, DeviceFrame, PixelFormat, ShardedSsimulacra2, Ssimulacra2  # noqa: F401

__all__ = ["Ssimulacra2", "ShardedSsimulacra2", "MetricsResults", "MetricAggregate", "PixelFormat", "ColorMatrix", "DeviceFrame", "Ssimu2Error", "SO_PATH", "TurboMetrics", "ShardedTurboMetrics",
           "FrameScores", "Options", "select_frames", "shard_range", "gather_scores", "Stats"]
