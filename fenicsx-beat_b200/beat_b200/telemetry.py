"""Monitors with the reference's interface (src/beat/telemetry.py:15-136).

``track_time`` labels are the reference's; on the device path wall-clock deltas around asynchronous
launches mean nothing, so :class:`PerformanceMonitor` additionally receives CUDA-event stage totals from
the context (``record_device_stages``) and reports those under the same label names.
"""

from __future__ import annotations

import abc
import json
import logging
import time
from contextlib import contextmanager
from pathlib import Path
from typing import Dict, Union

logger = logging.getLogger(__name__)


class BaseMonitor(abc.ABC):
    @abc.abstractmethod
    @contextmanager
    def track_time(self, name: str):
        yield

    @abc.abstractmethod
    def record_ksp(self, ksp) -> None:
        pass

    @abc.abstractmethod
    def advance_step(self, t0: float, t1: float) -> None:
        pass


class _NoTimer:
    """Re-usable no-op context manager (a generator-based one costs microseconds per step)."""

    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_TIMER = _NoTimer()


class NullMonitor(BaseMonitor):
    def track_time(self, name: str):
        return _NO_TIMER

    def record_ksp(self, ksp) -> None:
        pass

    def advance_step(self, t0: float, t1: float) -> None:
        pass


class PerformanceMonitor(BaseMonitor):
    """Accumulates timings and KSP statistics; logs every ``log_frequency`` steps; JSON summary on rank 0."""

    def __init__(self, log_frequency: int = 1, comm=None):
        from .fem import comm_world

        self.log_frequency = log_frequency
        self.comm = comm if comm is not None else comm_world()
        self.step_counter = 0
        self.timings: Dict[str, float] = {}
        self.ksp_total_iterations = 0
        self.ksp_max_iterations = 0
        self.ksp_last_iterations = 0
        self.ksp_last_residual_norm = 0.0
        self.ksp_last_converged_reason = 0

    @contextmanager
    def track_time(self, name: str):
        tic = time.perf_counter()
        try:
            yield
        finally:
            toc = time.perf_counter()
            self.timings[name] = self.timings.get(name, 0.0) + (toc - tic)

    def record_device_stages(self, stages: dict) -> None:
        """CUDA-event totals (seconds) for labels, e.g. {"ode_step": .., "pde_step": ..}."""
        for name, sec in stages.items():
            self.timings[name] = self.timings.get(name, 0.0) + float(sec)

    def record_ksp(self, ksp) -> None:
        try:
            iterations = int(ksp.getIterationNumber())
            self.ksp_last_iterations = iterations
            self.ksp_total_iterations += iterations
            self.ksp_max_iterations = max(self.ksp_max_iterations, iterations)
            self.ksp_last_residual_norm = float(ksp.getResidualNorm())
            self.ksp_last_converged_reason = int(ksp.getConvergedReason())
        except Exception:  # mirrors the reference's tolerance of a failing KSP query
            pass

    def advance_step(self, t0: float, t1: float) -> None:
        self.step_counter += 1
        if self.log_frequency <= 0 or self.step_counter % self.log_frequency != 0:
            return
        timing_text = ", ".join(f"{name}={value:.6f}s" for name, value in self.timings.items())
        logger.info(
            f"PDE step timing step={self.step_counter}, "
            f"t=({t0:.5f}, {t1:.5f}), "
            f"ksp_iterations={self.ksp_last_iterations}, "
            f"ksp_residual_norm={self.ksp_last_residual_norm:.6e}, "
            f"ksp_converged_reason={self.ksp_last_converged_reason}, "
            f"{timing_text}",
        )

    def display_summary(self) -> None:
        if self.comm.rank != 0:
            return
        lines = ["\n" + "=" * 50, f"{'PERFORMANCE SUMMARY':^50}", "=" * 50]
        lines.append(f"Total Steps:           {self.step_counter}")
        lines.append(f"KSP Total Iterations:  {self.ksp_total_iterations}")
        lines.append(f"KSP Max Iterations:    {self.ksp_max_iterations}")
        lines.append("-" * 50)
        lines.append(f"{'Metric':<35} | {'Time (s)':>10}")
        lines.append("-" * 50)
        for name, duration in sorted(self.timings.items(), key=lambda kv: kv[1], reverse=True):
            lines.append(f"{name:<35} | {duration:>10.4f}")
        lines.append("=" * 50 + "\n")
        logger.info("\n".join(lines))

    def save_summary(self, filepath: Union[str, Path]) -> None:
        if self.comm.rank != 0:
            return
        data = {
            "total_steps": self.step_counter,
            "ksp": {"total_iterations": self.ksp_total_iterations, "max_iterations": self.ksp_max_iterations},
            "timings": self.timings,
        }
        filepath = Path(filepath)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        with open(filepath, "w") as f:
            json.dump(data, f, indent=4)
        logger.info(f"Performance summary saved to {filepath}")
