"""Single-cell pre-pacing on the device: the reference's ``beat.single_cell.get_steady_state``
(src/beat/single_cell.py:86-156) with the same arguments, cache files and return value.

The reference JIT-compiles ``fun`` with numba and loops ``y[:] = fun(states=y, t=t, parameters=p, dt=dt)`` over
``times = arange(0, BCL, dt)`` for ``nbeats`` beats (:42-67).  Here ``fun`` is a device model handle and the loop
drives the cell-model kernel (one launch per step, queued asynchronously; the host only waits when it reads tracked
values).  Extension: ``init_states`` may be (num_states, n_cells) and ``parameters`` (num_parameters, n_cells) to pace
several parameter sets - e.g. the three transmural cell types - in the same launches.
"""

from __future__ import annotations

import hashlib
import logging
from pathlib import Path

import numpy as np

from .device_model import DeviceODE
from .odesolver import ODESystemSolver

logger = logging.getLogger(__name__)


def compute_hash(fun: DeviceODE, init_states: np.ndarray, parameters: np.ndarray, nbeats: int = 200, BCL: float = 1000.0,
                 dt: float = 0.05) -> str:
    """Cache key of a pacing run (single_cell.py:70-83; the reference hashes the byte code of ``fun``, here the model
    tag, the scheme and the operation counts of the generated kernel stand for the code)."""
    h = hashlib.md5()
    h.update(f"{fun.model_tag}/{fun.scheme}/{sorted(fun.op_counts.items())}".encode())
    for item in (np.asarray(init_states), np.asarray(parameters), nbeats, BCL, dt):
        h.update(str(item).encode())
    return h.hexdigest()


def get_steady_state(fun: DeviceODE, init_states: np.ndarray, parameters: np.ndarray, outdir: Path, nbeats: int = 200, BCL: int = 1000,
                     save_every_ms: float = 1.0, dt: float = 0.05, track_indices: list[int] | None = None) -> np.ndarray:
    if not isinstance(fun, DeviceODE):
        raise TypeError("fun must be a device model handle (beat_b200.models.<model>.<scheme>); there is no CPU fallback")
    outdir = Path(outdir)
    key = compute_hash(fun=fun, init_states=init_states, parameters=parameters, nbeats=nbeats, BCL=BCL, dt=dt)
    fname = outdir / f"steady_states_{key}.npy"
    if fname.is_file():
        return np.load(fname)
    outdir.mkdir(exist_ok=True, parents=True)
    logger.info(f"Computing steady state with {nbeats} beats.")

    y0 = np.asarray(init_states, dtype=np.float64)
    single = y0.ndim == 1
    states = np.ascontiguousarray(y0.reshape(fun.num_states, -1))
    solver = ODESystemSolver(fun, states.copy(), parameters)
    ctx = solver._ctx
    solver._mirror.flush()
    solver._sync_parameters()
    times = np.arange(0.0, BCL, dt)
    track_values = None
    if track_indices is not None:
        save_freq = int(np.ceil(save_every_ms / dt))
        rows = int(np.ceil(len(times) / save_freq) * nbeats)
        idx = np.asarray(track_indices, dtype=np.int64)
        track_values = np.zeros((rows, len(idx)) if single else (rows, len(idx), states.shape[1]))
        scratch = np.empty_like(states)
        k = 0
    for _ in range(nbeats):
        for j, t in enumerate(times.tolist()):
            if track_values is not None and j % save_freq == 0:
                ctx.ode_get_states(scratch)  # the state BEFORE the step at time t (single_cell.py:47-53)
                track_values[k] = scratch[idx, 0] if single else scratch[idx]
                k += 1
            ctx.ode_step(t, dt)
    y = ctx.ode_get_states(np.empty_like(states))
    ctx.close()
    y = y[:, 0] if single else y
    if track_values is not None:
        np.save(outdir / f"tracked_values_{key}.npy", track_values)
        _plot_tracked(outdir / f"tracked_values_{key}.png", track_values if single else track_values[..., 0], times, save_freq, BCL, nbeats,
                      save_every_ms)
    np.save(fname, y)
    return y


def _plot_tracked(path: Path, track_values: np.ndarray, times: np.ndarray, save_freq: int, BCL: float, nbeats: int,
                  save_every_ms: float) -> None:
    """The reference's overview figure (single_cell.py:139-149); skipped when matplotlib is not installed."""
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        logger.warning("Matplotlib not installed, plotting not available.")
        return
    rows, n = track_values.shape
    last = int(np.ceil(BCL // save_every_ms))
    fig, ax = plt.subplots(n, 2, sharex="col", sharey="row", squeeze=False)
    for i in range(n):
        ax[i, 0].plot(np.linspace(0, BCL * nbeats, rows), track_values[:, i])
        ax[i, 1].plot(times[::save_freq][-last:], track_values[-last:, i])
    fig.tight_layout()
    fig.savefig(path)
    plt.close(fig)
