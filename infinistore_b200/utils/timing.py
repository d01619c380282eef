from __future__ import annotations

import os
import threading
from typing import Callable, List, Sequence


def cuda_time_ms(fn: Callable[[], None], device, iters: int = 10, warmup: int = 3) -> List[float]:
    """Per-call device time of `fn` with CUDA events on the current stream (ms, sorted)."""
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    out = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(device))
        fn()
        e1.record(torch.cuda.current_stream(device))
        e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)


def percentile(sorted_values: Sequence[float], q: float) -> float:
    if not sorted_values:
        return float("nan")
    i = min(len(sorted_values) - 1, max(0, int(round(q / 100.0 * (len(sorted_values) - 1)))))
    return sorted_values[i]


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason sampler (NVML in-process) for timed regions."""

    def __init__(self, index: int, period_s: float = 0.1):
        super().__init__(daemon=True)
        self.period = period_s
        self.samples: List[int] = []
        self.reasons = set()
        self.max_mhz = 0
        self._stop_ev = threading.Event()
        self._n = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = int(vis.split(",")[index]) if vis else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._n = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._n = None

    def run(self):
        n = self._n
        while n is not None and not self._stop_ev.is_set():
            try:
                self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                                  ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                                  ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                                  ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_ev.wait(self.period)

    def stop(self) -> dict:
        self._stop_ev.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz or None,
                "reasons": sorted(self.reasons), "samples": len(s)}
