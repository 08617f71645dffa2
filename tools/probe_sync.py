import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, _p)
import numpy as np
import beat_b200.niederer as nied
ksp = sys.argv[1] if len(sys.argv) > 1 else "pipecg"
solver, info = nied.setup(dx=0.2, probes=False, ksp_type=ksp)
ctx = solver.pde._ctx
print("sync us:", ctx.bench_grid_sync(1000))
t, dt = 0.0, 0.01
for _ in range(300):
    solver.step((t, t + dt)); t += dt
ctx.debug_timeline(True, False)
for rep in range(3):
    ctx.split_step(t, t + dt, 1.0); t += dt
    st = ctx.debug_timeline(True, True)
    d = np.diff(np.array(st, dtype=np.int64)) / 1e3
    print("its", ctx.ksp_info()[0], "total us %.1f" % ((st[-1] - st[0]) / 1e3), " phases us:", np.round(d, 2).tolist())
