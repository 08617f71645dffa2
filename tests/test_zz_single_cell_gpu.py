"""Single-cell pre-pacing (reference src/beat/single_cell.py:42-156) on the device against the
same loop over the oracle's NumPy cell model."""

import numpy as np
import pytest

import _problems as P
from beat_b200 import single_cell
from beat_b200.models import tp06

pytestmark = pytest.mark.gpu


def _oracle_pacing(om, y0, params, nbeats, BCL, dt, track, save_freq):
    y = y0.reshape(len(y0), 1).copy()
    rows = []
    for _ in range(nbeats):
        for j, t in enumerate(np.arange(0.0, BCL, dt)):
            if j % save_freq == 0:
                rows.append(y[track, 0].copy())
            y = om.generalized_rush_larsen(y, t, dt, params)
    return y[:, 0], np.array(rows)


def test_get_steady_state_matches_the_oracle_loop(tmp_path):
    om = P.oracle_model("tp06")
    y0 = om.init_state_values()
    params = om.init_parameter_values(stim_start=1.0, stim_period=6.0)  # a stimulus inside every (short) beat
    nbeats, BCL, dt = 2, 6.0, 0.05
    track = [om.state_index("V"), om.state_index("Ca_i")]
    y = single_cell.get_steady_state(tp06.generalized_rush_larsen, y0, params, tmp_path, nbeats=nbeats, BCL=BCL, save_every_ms=1.0, dt=dt,
                                     track_indices=track)
    want, rows = _oracle_pacing(om, y0, params, nbeats, BCL, dt, track, int(np.ceil(1.0 / dt)))
    assert y.shape == y0.shape
    assert np.abs(y - want).max() <= 1e-8 * np.abs(want).max()
    assert abs(y[om.state_index("V")] - y0[om.state_index("V")]) > 1.0  # the cell was actually paced
    key = single_cell.compute_hash(tp06.generalized_rush_larsen, y0, params, nbeats, BCL, dt)
    tracked = np.load(tmp_path / f"tracked_values_{key}.npy")
    assert tracked.shape == rows.shape and np.abs(tracked - rows).max() <= 1e-8 * np.abs(rows).max()
    # second call: served from the cache file (single_cell.py:108-109)
    again = single_cell.get_steady_state(tp06.generalized_rush_larsen, y0, params, tmp_path, nbeats=nbeats, BCL=BCL, dt=dt)
    assert np.array_equal(again, y)


def test_several_parameter_sets_in_one_run(tmp_path):
    om = P.oracle_model("tp06")
    y0 = om.init_state_values()
    p = om.init_parameter_values(stim_start=1.0, stim_period=6.0)
    params = np.repeat(p[:, None], 3, axis=1)
    params[om.parameter_index("g_Ks")] *= [1.0, 0.25, 1.0]  # a mid-myocardial-like variant in the middle
    states = np.repeat(y0[:, None], 3, axis=1)
    y = single_cell.get_steady_state(tp06.generalized_rush_larsen, states, params, tmp_path, nbeats=1, BCL=6.0, dt=0.05)
    assert y.shape == (len(y0), 3)
    assert np.array_equal(y[:, 0], y[:, 2]) and not np.array_equal(y[:, 0], y[:, 1])
    want, _ = _oracle_pacing(om, y0, p, 1, 6.0, 0.05, [0], 10**9)
    assert np.abs(y[:, 0] - want).max() <= 1e-8 * np.abs(want).max()
