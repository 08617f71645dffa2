import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import beat_b200.niederer as nied
solver, info = nied.setup(dx=0.2, probes=False, ksp_type="auto")
ode, pde, ctx = solver.ode, solver.pde, solver.pde._ctx
dt, t = 0.01, 0.0
for _ in range(20):
    solver.step((t, t + dt)); t += dt
host_v = np.array(pde.state.x.array_ro)
# CPU cost of enqueueing one fused step (queue empty before, not waiting for the GPU)
ctx.sync()
a = time.perf_counter()
for _ in range(10):
    ctx.split_step(t, t + dt, 1.0); t += dt
b = time.perf_counter()
ctx.sync()
print("enqueue cost of ctx.split_step: %.1f us" % ((b - a) / 10 * 1e6))
ctx.sync()
a = time.perf_counter()
for _ in range(10):
    solver.step((t, t + dt)); t += dt
b = time.perf_counter()
ctx.sync()
print("enqueue cost of solver.step: %.1f us" % ((b - a) / 10 * 1e6))
def loop(K):
    global t
    for k in range(K):
        ode.v_ode.x.array[:] = host_v
        ode.from_dolfin()
        solver.step((t, t + dt))
        host_v[:] = pde.state.x.array_ro
        t += dt
loop(50)
a = time.perf_counter(); loop(500); print("e2e us/step", (time.perf_counter() - a) / 500 * 1e6)
pr = cProfile.Profile(); pr.enable(); loop(300); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
