"""Small utilities: device timing and clock sampling used by the benchmarks."""
from .timing import cuda_time_ms, percentile, ClockSampler

__all__ = ["cuda_time_ms", "percentile", "ClockSampler"]
