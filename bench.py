"""bench.py - monodomain node-steps/s of the operator-split step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one `MonodomainSplittingSolver.step((t, t+dt))` over the whole mesh: TP06 generalized
Rush-Larsen ODE update + Crank-Nicolson diffusion solve (Jacobi-PCG, PETSc-default rtol 1e-5, zero initial
guess) + S1 stimulus.

Default workload (every N) = BASELINE.json configs[3], the strong-scaling case the north_star's targets are
stated on: Niederer slab 20x7x3 mm refined to dx=0.025 mm (27 237 681 nodes, the "~30M dofs" of that config),
anisotropic fibre conductivity, dt=0.01 ms, split over the N GPUs along x (`"scaling": "strong"`), with the
halo exchange and the CG reductions between ranks inside the step.  The same line carries, under
`"secondary"`, BASELINE.json configs[1] (dx=0.2 mm, 58 176 nodes per GPU, the latency-bound case; weak
scaling: one 20 mm block per GPU) measured in the same process right after.  `--workload` / `--scaling`
select any other case (e.g. niederer_dx0.016 = the 103.8 M-dof slab, lv_ellipsoid_100M).

Printed JSON line (rank 0):
  value        node-steps/s, states resident in HBM, CUDA events around every step, L2 flushed between
               steps (the flush is outside the events); max over ranks.
  warm_l2      same loop without the flush (what a real time loop sees: the 58k-node working set lives in L2).
  e2e          the same steps through the public Python API with HOST buffers each step: upload of the
               membrane potential from pinned host memory (v_ode.x.array -> from_dolfin), solver.step,
               download of pde.state.x.array.
  roofline     dominant kernel of the step (by device time), algorithmic bytes / CUDA-event time.
  cpu_baseline oracle C port (OpenMP, all host cores) on a bounded sample of the same workload (an x-block
               of the same slab at the same dx / dt / solver settings; node-steps/s does not depend on the
               block length).
  selfcheck    correctness evidence of THIS run: CG iteration totals equal on all ranks, ghost values equal to
               their owners' (NCCL exchange vs the in-kernel one), global checksums of V (compare across N:
               strong scaling solves the same problem), and - on the secondary workload - a bit-exact repeat.
`--impl reference` times that CPU port alone (the reference's dolfinx/PETSc stack cannot be installed
in this image: see DESIGN.md) on the same workload string and metric, with all host cores whatever
OMP_NUM_THREADS says; under torchrun only rank 0 works.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "monodomain node-steps/sec, TP06 Niederer slab"
UNIT = "node-steps/s"
WORKLOADS = {
    # name: (dx, dt)
    "niederer_dx0.2": (0.2, 0.01),
    "niederer_dx0.1": (0.1, 0.01),
    "niederer_dx0.05": (0.05, 0.01),
    "niederer_dx0.5": (0.5, 0.01),
    "niederer_dx0.025": (0.025, 0.01),   # 27.2 M dofs (BASELINE config 4's "~30M")
    "niederer_dx0.016": (0.016, 0.01),   # 103.8 M dofs: the north_star's 100M-dof slab (several GPUs)
}
# BASELINE config 5 (synthetic LV shell, endocardial surface stimulus, three transmural layers): (n_r, n_mu, n_phi), dt
LV_WORKLOADS = {
    "lv_ellipsoid_40k": ((6, 48, 128), 0.01),
    "lv_ellipsoid_320k": ((12, 128, 192), 0.01),
    "lv_ellipsoid_1.4M": ((16, 256, 320), 0.01),
    "lv_ellipsoid_13M": ((24, 256, 2048), 0.01),    # 13.2 M dofs: fits one GPU, the per-GPU share of the 100 M shell
    "lv_ellipsoid_100M": ((48, 512, 4096), 0.01),   # 103.0 M dofs: BASELINE config 5 at its named size (8 GPUs, --scaling strong;
                                                    # set-up ~2 min and ~23 GB of host memory per rank)
}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi samples DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------- workload text
SLAB = (20.0, 7.0, 3.0)
SOLVER_TXT = "TP06 GRL1, Godunov split + CN diffusion (Jacobi-PCG, PETSc-default rtol 1e-5, x0=0)"


def slab_nodes(dx: float, Lx: float) -> int:
    nx, ny, nz = (int(round(l / dx)) for l in (Lx, SLAB[1], SLAB[2]))
    return (nx + 1) * (ny + 1) * (nz + 1)


def workload_config(workload: str, scaling: str, world: int) -> dict:
    """The `config` object: IDENTICAL in both arms (the driver compares them), so it only says what is solved and how a
    run of it is timed - not what a particular arm chose (Krylov driver, kernel mode: those are in `solver` / `roofline`)."""
    if workload in LV_WORKLOADS:
        (n_r, n_mu, n_phi), dt = LV_WORKLOADS[workload]
        if scaling == "weak":
            n_phi *= world
        txt = (f"{workload}: synthetic LV shell ({n_r}x{n_mu}x{n_phi} hexahedra, Kuhn tets, cell-wise Bishop conductivity "
               f"tensor, ENDO surface stimulus, 3 layers), {SOLVER_TXT}, dt={dt} ms")
        nodes = None
    else:
        dx, dt = WORKLOADS[workload]
        Lx = SLAB[0] * world if scaling == "weak" else SLAB[0]
        nodes = slab_nodes(dx, Lx)
        txt = (f"{workload}: Niederer slab {Lx:g}x7x3 mm, dx={dx} mm, {nodes} nodes, anisotropic fibre conductivity, "
               f"{SOLVER_TXT}, dt={dt} ms")
    return {"workload": txt, "nodes": nodes, "scaling": scaling,
            "l2": "GPU arm: L2 flushed (256 MiB memset) between timed steps, flush outside the CUDA events",
            "parallelism": f"{world} rank(s), one per GPU, {'phi-sector' if workload in LV_WORKLOADS else 'x-slab'} partition "
                           f"({scaling} scaling); reference arm: host cores of the same box on an x-block sample"}


# ------------------------------------------------------------------------------------- CPU oracle leg
def cpu_sample_length(dx: float, target_nodes: float = 1.8e6) -> float:
    """Length of the x-block of the slab the CPU port is timed on (the full 20 mm when that is small enough): set-up of the
    NumPy/SciPy oracle assembly is ~15 s per million nodes, so the sample is bounded at ~1.8 M nodes."""
    plane = (int(round(SLAB[1] / dx)) + 1) * (int(round(SLAB[2] / dx)) + 1)
    nx = int(min(round(SLAB[0] / dx), max(16, target_nodes // plane - 1)))
    return nx * dx


def cpu_problem(dx: float, dt: float, length_x: float = 20.0):
    """Oracle-side arrays of the workload (NumPy/SciPy assembly of oracle/fem.py)."""
    import _problems as P
    from oracle import cport

    prob = P.niederer_slab(dx, L=(length_x, 7.0, 3.0))
    om = P.oracle_model("tp06")
    n = prob["mass"].shape[0]
    params = om.init_parameter_values(stim_amplitude=0.0)
    y0 = om.init_state_values()
    sp = cport.SplitProblem(prob["mass"], prob["stiff"], prob["C_m"], 0.5, dt,
                            [(prob["stim_load"], 0.0, 2.0, prob["stim_amp"])], rtol=1e-5)
    states = np.ascontiguousarray(np.repeat(y0[:, None], n, axis=1))
    return sp, states, params, om.state_index("V"), n


def cpu_run(dx: float, dt: float, steps: int, warmup: int, budget_s: float = 25.0):
    """Times `steps` split steps of the oracle C port on ALL host cores (bounded by budget_s)."""
    from oracle import cport

    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cport.set_threads(ncpu)  # torchrun exports OMP_NUM_THREADS=1: the baseline must not inherit that
    Ls = cpu_sample_length(dx)
    t_setup = time.perf_counter()
    sp, states, params, vidx, n = cpu_problem(dx, dt, Ls)
    setup_s = time.perf_counter() - t_setup
    cores = cport.num_threads()
    t = 0.0
    if warmup:
        sp.split_steps("tp06", "generalized_rush_larsen", vidx, states, params, t, warmup)
        t += warmup * dt
    done, its, t_begin = 0, 0, time.perf_counter()
    chunk = max(1, min(steps, 20))
    while done < steps:
        k = min(chunk, steps - done)
        _, i = sp.split_steps("tp06", "generalized_rush_larsen", vidx, states, params, t, k)
        t += k * dt
        done += k
        its += i
        if time.perf_counter() - t_begin > budget_s:
            break
    wall = time.perf_counter() - t_begin
    whole = abs(Ls - SLAB[0]) < 1e-9
    return {"value": n * done / wall, "unit": UNIT, "cores": cores, "kind": "port", "steps": done, "nodes": n,
            "ms_per_step": 1e3 * wall / done, "cg_iterations_per_step": its / done, "setup_s": setup_s,
            "sample": f"{done} full split steps of {'the whole' if whole else f'the first {Ls:g} mm (x-block incl. the stimulus corner) of the'} "
                      f"20x7x3 mm slab at dx={dx}, dt={dt}: {n} nodes, by oracle/c/oracle_step.c (OpenMP, {cores} threads = all host "
                      f"cores); the dolfinx/PETSc reference itself is not installable here"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload in LV_WORKLOADS:
        print(json.dumps({"impl": "reference", "unavailable": "the CPU port covers the slab workloads only"}), flush=True)
        return 0
    dx, dt = WORKLOADS[args.workload]
    r = cpu_run(dx, dt, args.steps, args.warmup, budget_s=150.0)
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args.workload, args.scaling, max(1, args.gpus)),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "cg_iterations_per_step": r["cg_iterations_per_step"], "sample_nodes": r["nodes"], "setup_s": r["setup_s"],
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- GPU leg
class Dist:
    """The rank plumbing of one bench process (torch.distributed over NCCL; nothing on the data path)."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def _reduce(self, xs, op):
        if self.world == 1:
            return [float(x) for x in xs]
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return [float(v) for v in t.tolist()]

    def max(self, *xs):
        r = self._reduce(xs, None if self.world == 1 else self.dist.ReduceOp.MAX)
        return r[0] if len(r) == 1 else r

    def min(self, *xs):
        r = self._reduce(xs, None if self.world == 1 else self.dist.ReduceOp.MIN)
        return r[0] if len(r) == 1 else r

    def sum(self, *xs):
        r = self._reduce(xs, None if self.world == 1 else self.dist.ReduceOp.SUM)
        return r[0] if len(r) == 1 else r

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def selfcheck(D: Dist, solver, ctx, n_owned: int) -> dict:
    """Correctness evidence of this run (no oracle: the parity tests are tests/ -m gpu): every rank ran the same number of CG
    iterations, the ghosts the persistent kernel refreshed equal what an NCCL send/recv of the owners' values delivers,
    and global checksums of V that can be compared between the N = 1, 2, 4, 8 lines of a strong-scaling series."""
    it_tot, solves = ctx.ksp_totals()
    its, rnorm, reason = ctx.ksp_info()
    it_lo, it_hi = D.min(float(it_tot)), D.max(float(it_tot))
    v = np.array(solver.pde.state.x.array_ro)              # owned + ghost as the kernel left them
    ghosts_equal = True
    if D.world > 1:
        ctx.halo_refresh_nccl()                            # owners' values through ncclSend/Recv
        v2 = np.empty_like(v)
        ctx.get_v(v2)
        ghosts_equal = bool(np.array_equal(v[n_owned:], v2[n_owned:]))
    vo = v[:n_owned]
    s1, s2 = D.sum(float(vo.sum()), float(np.square(vo).sum()))
    vmin, vmax = D.min(float(vo.min())), D.max(float(vo.max()))
    ok = it_lo == it_hi and bool(D.min(1.0 if ghosts_equal else 0.0) == 1.0) and bool(np.isfinite(s2)) and reason > 0
    return {"ok": bool(ok), "cg_iterations_total": int(it_hi), "cg_iterations_equal_on_all_ranks": it_lo == it_hi,
            "solves": int(solves), "last_reason": int(reason), "ghosts_equal_owners": bool(D.min(1.0 if ghosts_equal else 0.0) == 1.0),
            "v_sum": s1, "v_sumsq": s2, "v_min_mV": vmin, "v_max_mV": vmax}


def measure(D: Dist, args, workload: str, scaling: str, K: int, W: int, full: bool) -> dict:
    """One workload: set-up, warm-up, the timed regions, the roofline figures.  `full` adds the e2e region and the x0 = v_
    extra; the secondary workload runs the device-timed regions and the repeat check only."""
    import beat_b200.niederer as nied
    from beat_b200 import fem

    rank, world = D.rank, D.world
    comm = fem.Comm(rank, world)
    t_setup = time.perf_counter()
    is_lv = workload in LV_WORKLOADS
    if is_lv:
        from beat_b200 import lv_ellipsoid

        (n_r, n_mu, n_phi), dt = LV_WORKLOADS[workload]
        if scaling == "weak":
            n_phi *= world  # finer in phi: one sector of the same size per GPU
        dx = 0.0
        solver, info = lv_ellipsoid.setup(n=(n_r, n_mu, n_phi), comm=comm, ksp_type=args.ksp, pc_type=None if args.pc == "auto" else args.pc,
                                          initial_guess_previous=args.x0 == "previous")
    else:
        dx, dt = WORKLOADS[workload]
        Lx = SLAB[0] * world if scaling == "weak" else SLAB[0]
        solver, info = nied.setup(dx=dx, comm=comm, L=(Lx, 7.0, 3.0), probes=False, ksp_type=args.ksp,
                                  initial_guess_previous=args.x0 == "previous", pc_type=None if args.pc == "auto" else args.pc)
    ctx = solver.pde._ctx
    ksp_used, pc_used = solver.pde.ksp_type_used, solver.pde.pc_type_used
    n_global, n_owned = info["n_global"], info["n_owned"]
    setup_s = time.perf_counter() - t_setup
    barrier = lambda: D.barrier(ctx)  # noqa: E731

    # ---- warm-up -------------------------------------------------------------------------------
    t = 0.0
    solver.step((t, t + dt))  # first step through the API: flushes host mirrors, builds A/B for this dt
    t += dt
    for _ in range(max(W, 3) - 1):
        ctx.split_step(t, t + dt, 1.0)
        t += dt
    barrier()

    # ---- timed region 1: K steps, device-resident, L2 flushed between steps ----------------------
    it0, _ = ctx.ksp_totals()
    l0 = ctx.launch_count()
    clocks = ClockSampler(D.local_rank).start() if rank == 0 else None
    barrier()
    for k in range(K):
        ctx.l2_flush()
        ctx.event_record(2 * k)
        ctx.split_step(t, t + dt, 1.0)
        ctx.event_record(2 * k + 1)
        t += dt
    barrier()
    ms_flushed = sum(ctx.event_elapsed_ms(2 * k, 2 * k + 1) for k in range(K))
    launches = ctx.launch_count() - l0
    it1, _ = ctx.ksp_totals()
    iters_per_step = (it1 - it0) / K
    ms_flushed = D.max(ms_flushed)

    # ---- timed region 2: the same K steps back to back (warm L2) -----------------------------------
    barrier()
    ctx.event_record(2 * K)
    for k in range(K):
        ctx.split_step(t, t + dt, 1.0)
        t += dt
    ctx.event_record(2 * K + 1)
    barrier()
    ms_warm = D.max(ctx.event_elapsed_ms(2 * K, 2 * K + 1))

    # ---- extra: the same steps with the solve started from v_ instead of zero (not the headline: the reference
    #      runs PETSc's default zero initial guess; same convergence test, so the answer is at least as accurate) ----
    fast = None
    if full and args.x0 == "zero" and not args.no_extras:
        rtol_, atol_, max_it_, pc_, norm_, _ = solver.pde._solver_settings()
        ctx.pde_config(float(solver.pde.C_m), float(solver.pde.parameters["theta"]), rtol_, atol_, max_it_, pc_, norm_, 1)
        for _ in range(3):
            ctx.split_step(t, t + dt, 1.0)
            t += dt
        itf0, _ = ctx.ksp_totals()
        barrier()
        for k in range(K):
            ctx.l2_flush()
            ctx.event_record(2 * K + 10 + 2 * k)
            ctx.split_step(t, t + dt, 1.0)
            ctx.event_record(2 * K + 11 + 2 * k)
            t += dt
        barrier()
        ms_fast = D.max(sum(ctx.event_elapsed_ms(2 * K + 10 + 2 * k, 2 * K + 11 + 2 * k) for k in range(K)))
        itf1, _ = ctx.ksp_totals()
        fast = {"value": n_global * K / (ms_fast * 1e-3), "ms_per_step": ms_fast / K, "cg_iterations_per_step": (itf1 - itf0) / K,
                "note": "initial_guess_previous=True (x0 = v_), same rtol/convergence test; L2 flushed between steps"}
        ctx.pde_config(float(solver.pde.C_m), float(solver.pde.parameters["theta"]), rtol_, atol_, max_it_, pc_, norm_, 0)
        ctx.split_step(t, t + dt, 1.0)
        t += dt
        barrier()

    # ---- per-stage device times (separate pass: the stage events add launch gaps) --------------------
    ctx.stage_timing(True)
    ctx.stage_times_ms(reset=True)
    ks = min(K, 200)
    for k in range(ks):
        ctx.l2_flush()
        ctx.split_step(t, t + dt, 1.0)
        t += dt
    st = ctx.stage_times_ms(reset=True)
    ctx.stage_timing(False)
    ode_ms, pde_ms = D.max(st["ode_ms"] / ks, st["pde_ms"] / ks)

    # ---- timed region 3: end to end through the public API with host buffers ----------------------
    e2e = None
    ode, pde = solver.ode, solver.pde
    if full:
        import torch

        npts = int(ode.v_ode.x.array_ro.size)  # owned + ghost dofs (num_points is a per-marker method on the multi-region solver)
        torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // max(1, world))))  # (torchrun pins OMP to 1 thread)
        ke = min(K, 500)
        t_out, t_in = torch.from_numpy(pde.state.x.array), torch.from_numpy(ode.v_ode.x.array)  # views of the two page-locked mirrors
        solver.step((t, t + dt))  # (the writable views above marked both mirrors host-dirty: settle that outside the timed region)
        t += dt
        barrier()
        t_e0 = time.perf_counter()
        ctx.event_record(2 * K + 2)
        for k in range(ke):
            out = pde.state.x.array_ro               # D2H of the result into its page-locked mirror (what the demos read every step)
            vin = ode.v_ode.x.array_wo               # the host OWNS V between steps: it writes the whole page-locked input mirror ...
            t_in.copy_(t_out)                        # (... here: the V it just read; torch's multi-threaded host copy of vin[:] = out)
            ode.from_dolfin()                        # ... H2D + states[v_index] <- v_ode (odesolver.py:168-170)
            solver.step((t, t + dt))
            t += dt
        out = pde.state.x.array_ro                   # the last result
        ctx.event_record(2 * K + 3)
        barrier()
        e2e_wall_ms = (time.perf_counter() - t_e0) * 1e3
        e2e_ms = D.max(max(ctx.event_elapsed_ms(2 * K + 2, 2 * K + 3), e2e_wall_ms))
        e2e = {"value": n_global * ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * npts,
               "d2h_bytes_per_step": 8 * npts, "steps": ke, "ms_per_step": e2e_ms / ke,
               "path": "out=pde.state.x.array_ro [D2H]; ode.v_ode.x.array_wo[:]=out [host copy into the pinned input mirror]; "
                       "ode.from_dolfin() [H2D]; solver.step((t,t+dt))"}
    clk = clocks.stop() if clocks else None
    check = selfcheck(D, solver, ctx, n_owned)

    # ---- bit-exact repeat (small meshes: the states fit a host array): the same K steps from the same state twice ----
    if not full and not is_lv and n_owned * info["num_states"] * 8 <= (1 << 28):
        def rerun():
            ode.values[:] = np.repeat(np.asarray(nied.tp06.init_state_values(**nied.IC))[:, None], ode.values.shape[1], axis=1)
            ode.to_dolfin()
            ode.ode_to_pde()
            pde.assign_previous()
            tt = 0.0
            for _ in range(K):
                solver.step((tt, tt + dt))
                tt += dt
            vv = np.array(pde.state.x.array_ro)[:n_owned]
            its_, _ = ctx.ksp_totals()
            return vv, its_
        try:
            va, ia = rerun()
            vb, ib = rerun()
            same = bool(np.array_equal(va, vb)) and (ib - ia) == (ia - check["cg_iterations_total"])
            check["repeat_bit_exact"] = bool(D.min(1.0 if same else 0.0) == 1.0)
            check["ok"] = check["ok"] and check["repeat_bit_exact"]
        except Exception as exc:  # the repeat is evidence, not the measurement: report, do not hide
            check["repeat_bit_exact"] = None
            check["repeat_error"] = f"{type(exc).__name__}: {exc}"

    # ---- roofline denominators ----------------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dfma_tflops = ctx.bench_dfma()

    nnz_row = solver.pde._nnz_per_row if hasattr(solver.pde, "_nnz_per_row") else 15.0
    dict_info = ctx.pde_dictionary_info()
    # algorithmic bytes per owned row (DESIGN.md "K2+K4"): SELL entries 12 B (fp64 value + int32 column); rows served by the
    # stencil dictionary read 1 pattern byte instead of their entries (so `achieved` never counts bytes that DRAM did not move)
    cover = dict_info["rows_covered"] if dict_info["active"] else 0.0
    mat_b = (1.0 - cover) * 12.0 * nnz_row + cover * 1.0
    rhs_b = mat_b + 16.0
    init_b = 48.0
    iter_b = mat_b + 96.0
    pde_bytes = n_owned * (rhs_b + init_b + iter_b * iters_per_step)
    ns = info["num_states"]
    ode_bytes = n_owned * (2.0 * 8.0 * ns + 8.0)
    ode_flop = n_owned * nied.tp06.generalized_rush_larsen.fp64_instr_per_node() * 2.0
    no_traffic = "not measured in this run (ncu cannot run inside the timed bench); per-launch dram bytes of the same kernels: profiles/README.md"
    roof_pde = {"kernel": f"pde_{ksp_used}_kernel (RHS SpMV + stimulus + {pc_used}-PCG, one persistent launch)", "bound": "hbm",
                "achieved": pde_bytes / (pde_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": pde_bytes / (pde_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "traffic_note": no_traffic, "peak_source": peak_src,
                "ms_per_launch": pde_ms, "algorithmic_bytes_per_launch": pde_bytes, "bytes_per_row": {"rhs": rhs_b, "init": init_b, "iteration": iter_b}}
    roof_ode = {"kernel": "ode_kernel_regions<tp06_grl1>" if is_lv else "ode_kernel_uniform<tp06_grl1>", "bound": "fp64", "achieved": ode_flop / (ode_ms * 1e-3) / 1e12,
                "peak": dfma_tflops, "unit": "TFLOP/s (fp64-pipe instructions x2)", "frac": ode_flop / (ode_ms * 1e-3) / 1e12 / dfma_tflops,
                "traffic": None, "traffic_note": no_traffic, "peak_source": "mono_bench_dfma (measured DFMA rate, this run)", "ms_per_launch": ode_ms,
                "hbm_gbs": ode_bytes / (ode_ms * 1e-3) / 1e9}
    dominant = roof_pde if pde_ms >= ode_ms else roof_ode

    rec = {
        "value": n_global * K / (ms_flushed * 1e-3), "ms_per_step": ms_flushed / K, "nodes": n_global, "nodes_rank0": n_owned,
        "solver": {"ksp": ksp_used, "pc": pc_used, "x0": args.x0, "matrix_dictionary": dict_info},
        "warm_l2": {"value": n_global * K / (ms_warm * 1e-3), "ms_per_step": ms_warm / K, "note": "same K steps back to back, no L2 flush"},
        "x0_previous": fast, "e2e": e2e, "gpu_launches": int(launches),
        "stages": {"ode_ms_per_step": ode_ms, "pde_ms_per_step": pde_ms, "cg_iterations_per_step": iters_per_step},
        "roofline": {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_note", "kernel", "peak_source")},
        "roofline_stages": {"pde": roof_pde, "ode": roof_ode},
        "clocks": clk, "setup_s": setup_s, "selfcheck": check, "dt": dt, "dx": dx,
    }
    # collective teardown: nobody frees its exchange buffers while a peer may still store into them
    barrier()
    del solver, ode, pde
    ctx.close()
    D.barrier()
    return rec


def run_b200(args):
    D = Dist(args)
    K, W = args.steps, args.warmup
    if args.matrix_dict is not None:
        os.environ["MONO_PDE_DICT"] = "1" if args.matrix_dict else "0"  # read by mono_pde_set_matrices
    main = measure(D, args, args.workload, args.scaling, K, W, full=True)
    secondary = None
    if args.secondary != "none" and args.secondary != args.workload:
        try:
            s = measure(D, args, args.secondary, "weak", K, W, full=False)
            secondary = {"config": workload_config(args.secondary, "weak", D.world), "scaling": "weak",
                         **{k: s[k] for k in ("value", "ms_per_step", "nodes", "solver", "warm_l2", "gpu_launches", "stages", "roofline",
                                              "roofline_stages", "selfcheck", "setup_s")}}
        except Exception as exc:  # never lose the primary line to the extra record
            secondary = {"error": f"{type(exc).__name__}: {exc}"}
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": D.world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, args.scaling, D.world),
        "nodes": main["nodes"], "nodes_rank0": main["nodes_rank0"], "solver": main["solver"],
        "warm_l2": main["warm_l2"], "x0_previous": main["x0_previous"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
        "stages": main["stages"], "roofline": main["roofline"], "roofline_stages": main["roofline_stages"], "clocks": main["clocks"],
        "setup_s": main["setup_s"], "selfcheck": main["selfcheck"], "v_max_mV": main["selfcheck"]["v_max_mV"], "secondary": secondary,
    }
    if D.rank == 0:
        if D.world == 1 and not args.no_cpu_baseline and args.workload not in LV_WORKLOADS:
            r = cpu_run(main["dx"], main["dt"], steps=2000, warmup=2, budget_s=20.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="niederer_dx0.025", choices=sorted(WORKLOADS) + sorted(LV_WORKLOADS))
    ap.add_argument("--secondary", default="niederer_dx0.2", choices=["none"] + sorted(WORKLOADS),
                    help="second workload measured in the same run and reported under \"secondary\" (weak scaling, device-timed regions only)")
    ap.add_argument("--ksp", default="auto", choices=["auto", "cg", "pipecg"],
                    help="Krylov driver of the diffusion solve (PETSc names; auto = pipecg while the CG vectors fit in shared "
                         "memory, cg beyond - same iterates in exact arithmetic)")
    ap.add_argument("--pc", default="auto", choices=["auto", "jacobi", "chebyshev"],
                    help="preconditioner (auto: the reference's hypre request mapped to the fastest native one for the mesh size)")
    ap.add_argument("--x0", default="zero", choices=["zero", "previous"],
                    help="initial guess of the diffusion solve: zero = PETSc default (as the reference runs), previous = v_")
    ap.add_argument("--matrix-dict", dest="matrix_dict", action="store_true", default=None,
                    help="force the stencil dictionary of the streaming KSPCG kernel on (MONO_PDE_DICT=1); bit-identical results")
    ap.add_argument("--no-matrix-dict", dest="matrix_dict", action="store_false", help="force it off (MONO_PDE_DICT=0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the x0=v_ extra measurement")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N>1: strong = the 20 mm slab split over the GPUs, weak = one 20 mm block per GPU (slab grows along x)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
