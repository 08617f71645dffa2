"""Monitors with the interface of the reference's ``beat.telemetry`` (src/beat/telemetry.py:15-136): the three classes,
their methods, the public counters (``timings``, ``step_counter``, ``ksp_*``), the tokens of the per-step log line and
the keys of the JSON summary are what the reference's callers and tests rely on, so those are kept.

Behind that interface the device path needs something different: a wall-clock interval around an asynchronous kernel
launch measures the launch, not the work.  :class:`PerformanceMonitor` therefore keeps two ledgers - host intervals
(``track_time``) and device stage totals measured with CUDA events by the context (``record_device_stages``) - and reports
both under the reference's label names; a device figure replaces the host figure of the same label.
"""

from __future__ import annotations

import abc
import json
import logging
import time
from dataclasses import asdict, dataclass
from pathlib import Path

logger = logging.getLogger(__name__)


class BaseMonitor(abc.ABC):
    """What a solver calls on its monitor during a step."""

    @abc.abstractmethod
    def track_time(self, name: str):
        """Context manager charging the enclosed host interval to ``name``."""

    @abc.abstractmethod
    def record_ksp(self, ksp) -> None:
        """Called after a linear solve with an object that answers the PETSc KSP getters."""

    @abc.abstractmethod
    def advance_step(self, t0: float, t1: float) -> None:
        """Called once per finished step."""


class _Idle:
    """The one context manager every NullMonitor hands out (entering a fresh generator per label costs microseconds,
    which is visible at 60 us per step)."""

    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_IDLE = _Idle()


class NullMonitor(BaseMonitor):
    def track_time(self, name: str):
        return _IDLE

    def record_ksp(self, ksp) -> None:
        return None

    def advance_step(self, t0: float, t1: float) -> None:
        return None


class _Interval:
    """Host stopwatch for one label; adds to the ledger when the ``with`` block ends, also on an exception."""

    __slots__ = ("ledger", "label", "started")

    def __init__(self, ledger: dict, label: str):
        self.ledger, self.label, self.started = ledger, label, 0.0

    def __enter__(self):
        self.started = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.ledger[self.label] = self.ledger.get(self.label, 0.0) + (time.perf_counter() - self.started)
        return False


@dataclass
class _KspLedger:
    total: int = 0
    largest: int = 0
    last: int = 0
    last_norm: float = 0.0
    last_reason: int = 0

    def add(self, iterations: int, norm: float, reason: int) -> None:
        self.total += iterations
        self.largest = max(self.largest, iterations)
        self.last, self.last_norm, self.last_reason = iterations, norm, reason


class PerformanceMonitor(BaseMonitor):
    def __init__(self, log_frequency: int = 1, comm=None):
        if comm is None:
            from .fem import comm_world

            comm = comm_world()
        self.log_frequency = log_frequency
        self.comm = comm
        self.step_counter = 0
        self.timings: dict[str, float] = {}         # label -> seconds (host intervals; device totals where recorded)
        self.device_labels: set[str] = set()        # labels whose figure comes from CUDA events
        self._ksp = _KspLedger()

    # the reference exposes the solver statistics as plain attributes
    ksp_total_iterations = property(lambda self: self._ksp.total)
    ksp_max_iterations = property(lambda self: self._ksp.largest)
    ksp_last_iterations = property(lambda self: self._ksp.last)
    ksp_last_residual_norm = property(lambda self: self._ksp.last_norm)
    ksp_last_converged_reason = property(lambda self: self._ksp.last_reason)

    def track_time(self, name: str):
        if name in self.device_labels:  # the device figure is the authoritative one for this label
            return _IDLE
        return _Interval(self.timings, name)

    def record_device_stages(self, stages: dict) -> None:
        """CUDA-event totals in seconds, e.g. ``{"ode_step": ..., "pde_step": ...}`` since the previous call."""
        for label, seconds in stages.items():
            if label not in self.device_labels:
                self.device_labels.add(label)
                self.timings[label] = 0.0  # drop whatever host time was charged before the first device figure
            self.timings[label] += float(seconds)

    def record_ksp(self, ksp) -> None:
        try:
            numbers = (int(ksp.getIterationNumber()), float(ksp.getResidualNorm()), int(ksp.getConvergedReason()))
        except Exception:  # a solver that cannot answer is not a reason to stop the run
            return
        self._ksp.add(*numbers)

    def advance_step(self, t0: float, t1: float) -> None:
        self.step_counter += 1
        every = self.log_frequency
        if every > 0 and self.step_counter % every == 0:
            logger.info(self._step_line(t0, t1))

    def _step_line(self, t0: float, t1: float) -> str:
        k = self._ksp
        fields = [f"PDE step timing step={self.step_counter}", f"t=({t0:.5f}, {t1:.5f})", f"ksp_iterations={k.last}",
                  f"ksp_residual_norm={k.last_norm:.6e}", f"ksp_converged_reason={k.last_reason}"]
        fields += [f"{label}={seconds:.6f}s" for label, seconds in self.timings.items()]
        return ", ".join(fields)

    def _table(self) -> str:
        rule, thin = "=" * 50, "-" * 50
        head = [rule, "PERFORMANCE SUMMARY".center(50), rule, f"{'Total Steps:':<23}{self.step_counter}",
                f"{'KSP Total Iterations:':<23}{self._ksp.total}", f"{'KSP Max Iterations:':<23}{self._ksp.largest}", thin,
                f"{'Metric':<35} | {'Time (s)':>10}", thin]
        slowest_first = sorted(self.timings, key=self.timings.get, reverse=True)
        body = [f"{label + (' [device]' if label in self.device_labels else ''):<35} | {self.timings[label]:>10.4f}" for label in slowest_first]
        return "\n" + "\n".join(head + body + [rule]) + "\n"

    def display_summary(self) -> None:
        if self.comm.rank == 0:
            logger.info(self._table())

    def save_summary(self, filepath) -> None:
        if self.comm.rank != 0:
            return
        target = Path(filepath)
        target.parent.mkdir(parents=True, exist_ok=True)
        ksp = asdict(self._ksp)
        summary = {"total_steps": self.step_counter,
                   "ksp": {"total_iterations": ksp["total"], "max_iterations": ksp["largest"]},
                   "timings": self.timings, "device_timed": sorted(self.device_labels)}
        target.write_text(json.dumps(summary, indent=4))
        logger.info(f"Performance summary saved to {target}")
