"""Per-phase times of one persistent PDE kernel launch (globaltimer stamps of CTA 0)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import beat_b200.niederer as nied
dx = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
ksp = sys.argv[2] if len(sys.argv) > 2 else "cg"
pc = sys.argv[3] if len(sys.argv) > 3 else None
solver, info = nied.setup(dx=dx, probes=False, ksp_type=ksp, pc_type=pc)
ctx = solver.pde._ctx
n = info["n_owned"]
t, dt = 0.0, 0.01
for _ in range(10):
    solver.step((t, t + dt)); t += dt
ctx.debug_timeline(True, False)
if os.environ.get("TIMELINE_DBG"):  # ablation switches of the ring kernel, applied to the measured launches only
    os.environ["MONO_RING_DBG"] = os.environ["TIMELINE_DBG"]
for rep in range(2):
    ctx.split_step(t, t + dt, 1.0); t += dt
    st = np.array(ctx.debug_timeline(True, True), dtype=np.int64)
    d = np.diff(st) / 1e3
    print("rows", n, "its", ctx.ksp_info()[0], "total us %.1f" % ((st[-1] - st[0]) / 1e3))
    if ksp == "cg" and os.environ.get("MONO_PDE_TAGGED_STREAM") is None and not getattr(solver.pde, "_resident", False):
        # pde_cg_stream_kernel (dictionary rows): stamps after rhs, red0, then per iteration spmv, red, axpy, red, pupd, barrier
        names = ["rhs", "red0"] + ["spmv", "red", "axpy", "red", "pupd", "barr"] * 20
        bytes_row = {"rhs": 49, "spmv": 25, "axpy": 56, "pupd": 32}
    elif ksp == "cg":
        names = ["rhs", "red0"] + ["spmv", "red", "axpy", "red", "pupd"] * 20
        bytes_row = {"rhs": 244, "spmv": 212, "axpy": 56, "pupd": 48}
    else:
        names = ["p0+app0", "p1"] + ["upd+post", "m,n=A m", "wait"] * 20
        bytes_row = {}
    for nm, us in zip(names, d):
        gbs = bytes_row.get(nm, 0) * n / (us * 1e-6) / 1e9 if nm in bytes_row else 0
        print("  %-5s %8.1f us %8.0f GB/s" % (nm, us, gbs))
