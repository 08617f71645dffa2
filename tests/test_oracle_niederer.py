"""Pins the TP06 + Kuhn-slab part of the oracle on the only reference output that exercises it: the
activation-time table printed in /root/reference/demos/niederer_benchmark.py:315-325.

What the oracle reproduces, and what it does not (also in DESIGN.md "Oracle"):
  * dx=0.5, dt=0.05 at the demo's own solver tolerance: P1 exact; the far corners arrive 0.15-0.7 ms
    (0.9-2 %) EARLIER than the published row - not within one dt (0.05 ms).
  * the gap is first order in dt: with the linear solve tightened it is <= 0.15 ms at dt=0.01 and <= 0.12 ms
    at dt=0.005 (0.2 %), so mesh, mass/stiffness, conductivities, stimulus and cell model are the reference's;
    what differs is an O(dt) detail of the time integration that cannot be resolved without running
    gotranx/PETSc (neither installable here).  The table is pinned with those honest tolerances.
"""
import numpy as np

import _problems as P
from oracle import niederer as N


def _diff(dx, dt, **kw):
    act = N.run(dx, dt, T=70.0, tp06=P.oracle_model("tp06"), **kw)
    got = np.array([act[k] for k in N.POINTS])
    assert (got >= 0).all(), act
    return got - np.array(N.PUBLISHED[(dx, dt)])


def test_activation_times_dx05_dt005_demo_tolerance():
    d = _diff(0.5, 0.05)
    assert abs(d[0]) < 1e-9          # P1 (stimulated corner): 1.25 ms exactly as published
    assert np.abs(d).max() <= 0.75   # ms; 2 % of the 35-59 ms arrival times
    assert (d <= 1e-9).all()         # systematically not later than the reference


def test_activation_times_converge_to_published_row_dt001():
    d = _diff(0.5, 0.01, rtol=1e-10)
    assert np.abs(d).max() <= 0.16, d  # ms (0.3 %)
