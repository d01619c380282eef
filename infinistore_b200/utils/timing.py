from __future__ import annotations

import os
import subprocess
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
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.  Uses
    NVML in-process (a `nvidia-smi` subprocess every 200 ms takes driver-wide locks long
    enough to perturb a launch-latency-sensitive loop); falls back to nvidia-smi."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = 0
        self._stop_ev = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis:
                idx = int(vis.split(",")[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
             "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5
        ).stdout.strip().split(",")
        if len(out) >= 6:
            self.samples.append(int(float(out[0])))
            self.max_mhz = int(float(out[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown",
                                "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                if v.strip().lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_ev.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:  # noqa: BLE001
                pass
            self._stop_ev.wait(0.1 if self._nvml is not None else 0.5)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz or None,
                "reasons": sorted(self.reasons), "samples": len(s),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}
