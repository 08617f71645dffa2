"""Worker of tests/test_multi_gpu.py: one process per GPU (launched with torch.distributed.run).

Every rank advances its x-slab of the Niederer problem through the public API; the persistent PDE kernels of the
ranks exchange boundary values and dot products through peer memory.  Each rank then repeats the run on the WHOLE
slab on its own GPU (single-rank context) and compares its owned AND ghost dofs with it.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_case(nied, fem, comm, dx, dt, nsteps, theta, ksp, x0_prev, rtol, pc=None, lv=None):
    if lv is not None:  # synthetic LV shell: phi-sector partition (periodic neighbours), cell-wise tensor, facet stimulus
        from beat_b200 import lv_ellipsoid

        solver, info = lv_ellipsoid.setup(n=tuple(lv), comm=comm, rtol=rtol, ksp_type=ksp, pc_type=pc, initial_guess_previous=x0_prev)
    else:
        solver, info = nied.setup(dx=dx, comm=comm, rtol=rtol, ksp_type=ksp, initial_guess_previous=x0_prev, probes=False, pc_type=pc)
    solver.theta = theta
    t = 0.0
    for _ in range(nsteps):
        solver.step((t, t + dt))
        t += dt
    v = np.array(solver.pde.state.x.array_ro)
    states = np.array(solver.ode.full_values)
    its, _ = solver.pde._ctx.ksp_totals()
    reason = solver.pde._ctx.ksp_info()[2]
    l2g = info["mesh"].index_map.local_to_global
    n_owned = info["n_owned"]
    if comm.size > 1:
        import torch.distributed as dist

        dist.barrier()  # nobody unmaps / frees exchange buffers while a peer may still be stepping
    solver.pde._ctx.close()
    return v, states, its, reason, l2g, n_owned


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["MONO_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import beat_b200.niederer as nied
    from beat_b200 import fem

    cases = json.loads(sys.argv[1])
    out = []
    ok = True
    for case in cases:
        dx, dt, nsteps, theta, ksp, x0_prev, rtol, tol = (case[k] for k in ("dx", "dt", "nsteps", "theta", "ksp", "x0_prev", "rtol", "tol"))
        for key, val in case.get("env", {}).items():
            os.environ[key] = val
        pc, lv = case.get("pc"), case.get("lv")
        v, s, its, reason, l2g, n_owned = run_case(nied, fem, fem.Comm(rank, world), dx, dt, nsteps, theta, ksp, x0_prev, rtol, pc, lv)
        vg, sg, itsg, reasong, _, _ = run_case(nied, fem, fem.COMM_SELF, dx, dt, nsteps, theta, ksp, x0_prev, rtol, pc, lv)
        for key in case.get("env", {}):
            os.environ.pop(key, None)
        scale = np.abs(vg).max()
        err_owned = float(np.abs(v[:n_owned] - vg[l2g[:n_owned]]).max() / scale)
        err_ghost = float(np.abs(v[n_owned:] - vg[l2g[n_owned:]]).max() / scale) if len(v) > n_owned else 0.0
        sscale = np.abs(sg).max(axis=1, keepdims=True)
        err_states = float((np.abs(s - sg[:, l2g]) / sscale).max())
        good = reason > 0 and reasong > 0 and max(err_owned, err_ghost, err_states) <= tol and abs(its - itsg) <= max(2, 0.02 * itsg)
        ok &= good
        out.append(dict(case=case, rank=rank, n_owned=int(n_owned), n_ghost=int(len(v) - n_owned), err_owned=err_owned,
                        err_ghost=err_ghost, err_states=err_states, its=int(its), its_single=int(itsg), reason=int(reason), ok=bool(good)))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for r in range(world):
        if r == rank:
            for o in out:
                print("MGPU", json.dumps(o), flush=True)
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
