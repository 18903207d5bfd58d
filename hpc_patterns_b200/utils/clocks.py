"""Sample SM clocks / throttle reasons with ``nvidia-smi`` while a timed region runs.

Follows the profiling recipe: the sampler starts before the timed region, is stopped
after it, and the summary (median SM MHz under load, max SM MHz, active throttle
reasons) is attached to every reported number.
"""
from __future__ import annotations

import shutil
import statistics
import subprocess
import tempfile
from typing import Dict, List, Optional

_FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")
_REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]


class ClockSampler:
    def __init__(self, gpu_index: Optional[int] = None, period_ms: int = 100):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self._proc: Optional[subprocess.Popen] = None
        self._file = None

    def start(self) -> "ClockSampler":
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return self
        self._file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        cmd = [exe, f"--query-gpu={_FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms)]
        if self.gpu_index is not None:
            cmd += ["-i", str(self.gpu_index)]
        try:
            self._proc = subprocess.Popen(cmd, stdout=self._file, stderr=subprocess.DEVNULL)
        except OSError:
            self._proc = None
        return self

    def stop(self) -> Dict:
        if self._proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._proc.terminate()
        try:
            self._proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self._proc.kill()
        self._file.flush()
        self._file.seek(0)
        return summarize(self._file.read().splitlines())


def summarize(lines: List[str]) -> Dict:
    sm: List[float] = []
    sm_max: List[float] = []
    power: List[float] = []
    reasons = set()
    for line in lines:
        parts = [p.strip() for p in line.split(",")]
        if len(parts) < 8:
            continue
        try:
            sm.append(float(parts[1]))
            sm_max.append(float(parts[2]))
            power.append(float(parts[3]))
        except ValueError:
            continue
        for name, val in zip(_REASONS, parts[4:8]):
            if val.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    # "Under load" = samples in the upper half of the observed power range.
    lo, hi = min(power), max(power)
    loaded = [c for c, p in zip(sm, power) if p >= lo + 0.5 * (hi - lo)] or sm
    return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(sm_max),
            "power_w_max": hi, "reasons": sorted(reasons), "samples": len(sm)}
