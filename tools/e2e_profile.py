"""Where the end-to-end step time goes (host side): times each statement of bench.py's e2e loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import beat_b200.niederer as nied

solver, info = nied.setup(dx=0.2, probes=False, ksp_type="pipecg")
ode, pde, ctx = solver.ode, solver.pde, solver.pde._ctx
dt = 0.01
t = 0.0
for _ in range(20):
    solver.step((t, t + dt)); t += dt
host_v = np.array(pde.state.x.array_ro)
acc = np.zeros(5)
K = 500
for k in range(K):
    a = time.perf_counter()
    ode.v_ode.x.array[:] = host_v
    b = time.perf_counter()
    ode.from_dolfin()
    c = time.perf_counter()
    solver.step((t, t + dt))
    d = time.perf_counter()
    host_v[:] = pde.state.x.array_ro
    e = time.perf_counter()
    t += dt
    acc += [b - a, c - b, d - c, e - d, e - a]
print("us per step: assign %.1f from_dolfin %.1f step(enqueue) %.1f readback %.1f total %.1f" % tuple(acc / K * 1e6))
# raw costs
a = time.perf_counter()
for _ in range(1000): ctx.sync()
print("ctx.sync() us", (time.perf_counter() - a) * 1e3)
x = np.empty(info["n_owned"])
a = time.perf_counter()
for _ in range(1000): ctx.get_v(x)
print("get_v us", (time.perf_counter() - a) * 1e3)
a = time.perf_counter()
for _ in range(1000): ctx.set_v_ode(x)
print("set_v_ode us", (time.perf_counter() - a) * 1e3)
a = time.perf_counter()
for _ in range(1000): host_v[:] = x
print("numpy copy us", (time.perf_counter() - a) * 1e3)
a = time.perf_counter()
for _ in range(300):
    ctx.split_step(t, t + dt, 1.0); t += dt
ctx.sync()
print("split_step back-to-back us", (time.perf_counter() - a) / 300 * 1e6)
a = time.perf_counter()
for _ in range(300):
    ctx.split_step(t, t + dt, 1.0); t += dt; ctx.sync()
print("split_step + sync us", (time.perf_counter() - a) / 300 * 1e6)
