"""Stencil-dictionary coverage of a Niederer slab partition (CPU only): how many owned rows of one rank repeat one of the
64 most frequent stencils (mono_csr_row_patterns) - the rows the experimental MONO_PDE_DICT kernel serves from shared memory.

    python tools/stencil_stats.py --dx 0.05 --ranks 8 --rank 3
"""

from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--dx", type=float, default=0.2)
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--rank", type=int, default=0)
    args = ap.parse_args(argv)

    import numpy as np

    from beat_b200 import fem
    from beat_b200._lib import csr_row_patterns

    n = [int(round(L / args.dx)) for L in (20.0, 7.0, 3.0)]
    t0 = time.perf_counter()
    mesh = fem.create_box(fem.Comm(args.rank, args.ranks), [np.zeros(3), np.array([20.0, 7.0, 3.0])], n)
    csr = fem.assemble_p1_local(mesh, np.diag([9.5298e-4, 1.2576e-4, 1.2576e-4]))
    t1 = time.perf_counter()
    pat, rep, cnt = csr_row_patterns(*csr, max_patterns=64)
    t2 = time.perf_counter()
    rows = pat.size
    print(f"dx={args.dx} rank {args.rank}/{args.ranks}: {rows} owned rows, {csr[0][-1] / rows:.2f} nnz/row; "
          f"{len(rep)} stencils kept, {100.0 * (pat != 255).mean():.3f} % of the rows covered "
          f"(most frequent: {100.0 * cnt[0] / rows:.2f} %); mesh + assembly {t1 - t0:.1f} s, classification {t2 - t1:.2f} s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
