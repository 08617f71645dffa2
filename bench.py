"""bench.py - monodomain node-steps/s of the operator-split step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one `MonodomainSplittingSolver.step((t, t+dt))` over the whole mesh: TP06 generalized
Rush-Larsen ODE update + Crank-Nicolson diffusion solve (Jacobi-PCG, PETSc-default rtol 1e-5, zero initial
guess) + S1 stimulus.  Workload at N=1 = BASELINE.json configs[1]: Niederer slab 20x7x3 mm, dx=0.2 mm,
58 176 nodes, dt=0.01 ms.  At N>1 the slab is extended along x (20*N mm), one 20 mm block per rank
(weak scaling), with the halo exchange and the CG reductions between ranks inside the step.

Printed JSON line (rank 0):
  value        node-steps/s, states resident in HBM, CUDA events around every step, L2 flushed between
               steps (the flush is outside the events); max over ranks.
  warm_l2      same loop without the flush (what a real time loop sees: the 58k-node working set lives in L2).
  e2e          the same steps through the public Python API with HOST buffers each step: upload of the
               membrane potential from pinned host memory (v_ode.x.array -> from_dolfin), solver.step,
               download of pde.state.x.array.
  roofline     dominant kernel of the step (by device time), algorithmic bytes / CUDA-event time.
  cpu_baseline oracle C port (OpenMP, all host cores) on a bounded sample of the same workload.
`--impl reference` times that CPU port alone (the reference's dolfinx/PETSc stack cannot be installed
in this image: see DESIGN.md), on the same workload/metric.
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


# ------------------------------------------------------------------------------------- CPU oracle leg
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
    """Times `steps` split steps of the oracle C port on all host cores (bounded by budget_s)."""
    from oracle import cport

    sp, states, params, vidx, n = cpu_problem(dx, dt)
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
    return {"value": n * done / wall, "unit": UNIT, "cores": cores, "kind": "port", "steps": done, "nodes": n,
            "ms_per_step": 1e3 * wall / done, "cg_iterations_per_step": its / done,
            "sample": f"{done} full split steps of the {n}-node slab (dx={dx}, dt={dt}) by oracle/c/oracle_step.c "
                      f"(OpenMP, {cores} threads); the dolfinx/PETSc reference itself is not installable here"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dx, dt = WORKLOADS[args.workload]
    r = cpu_run(dx, dt, args.steps, args.warmup, budget_s=150.0)
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{args.workload}: Niederer slab 20x7x3 mm, dx={dx} mm, {r['nodes']} nodes, TP06 GRL1, "
                               f"CN diffusion (Jacobi-PCG rtol 1e-5), dt={dt} ms", "nodes": r["nodes"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "cg_iterations_per_step": r["cg_iterations_per_step"],
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- GPU leg
def run_b200(args):
    import torch
    import torch.distributed as dist

    import beat_b200.niederer as nied
    from beat_b200 import fem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = fem.Comm(rank, world)
    K, W = args.steps, args.warmup
    if args.matrix_dict:
        os.environ["MONO_PDE_DICT"] = "1"  # read by mono_pde_set_matrices
    t_setup = time.perf_counter()
    is_lv = args.workload in LV_WORKLOADS
    if is_lv:
        from beat_b200 import lv_ellipsoid

        (n_r, n_mu, n_phi), dt = LV_WORKLOADS[args.workload]
        if args.scaling == "weak":
            n_phi *= world  # finer in phi: one sector of the same size per GPU
        dx, Lx = 0.0, 0.0
        solver, info = lv_ellipsoid.setup(n=(n_r, n_mu, n_phi), comm=comm, ksp_type=args.ksp, pc_type=None if args.pc == "auto" else args.pc,
                                          initial_guess_previous=args.x0 == "previous")
        geom_txt = f"synthetic LV shell ({n_r}x{n_mu}x{n_phi} hexahedra, Kuhn tets, cell-wise Bishop conductivity tensor, ENDO surface stimulus, 3 layers)"
    else:
        dx, dt = WORKLOADS[args.workload]
        Lx = 20.0 * world if args.scaling == "weak" else 20.0
        solver, info = nied.setup(dx=dx, comm=comm, L=(Lx, 7.0, 3.0), probes=False, ksp_type=args.ksp,
                                  initial_guess_previous=args.x0 == "previous", pc_type=None if args.pc == "auto" else args.pc)
        geom_txt = f"Niederer slab {Lx:g}x7x3 mm, dx={dx} mm"
    ctx = solver.pde._ctx
    args.ksp = solver.pde.ksp_type_used
    n_global, n_owned = info["n_global"], info["n_owned"]
    setup_s = time.perf_counter() - t_setup

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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
    clocks = ClockSampler(local_rank).start() if rank == 0 else None
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
    ms_flushed = max_over_ranks(ms_flushed)

    # ---- timed region 2: the same K steps back to back (warm L2) -----------------------------------
    barrier()
    ctx.event_record(2 * K)
    for k in range(K):
        ctx.split_step(t, t + dt, 1.0)
        t += dt
    ctx.event_record(2 * K + 1)
    barrier()
    ms_warm = max_over_ranks(ctx.event_elapsed_ms(2 * K, 2 * K + 1))

    # ---- extra: the same steps with the solve started from v_ instead of zero (not the headline: the reference
    #      runs PETSc's default zero initial guess; same convergence test, so the answer is at least as accurate) ----
    fast = None
    if args.x0 == "zero" and not args.no_extras:
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
        ms_fast = max_over_ranks(sum(ctx.event_elapsed_ms(2 * K + 10 + 2 * k, 2 * K + 11 + 2 * k) for k in range(K)))
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
    ode_ms, pde_ms = st["ode_ms"] / ks, st["pde_ms"] / ks

    # ---- timed region 3: end to end through the public API with host buffers ----------------------
    ode, pde = solver.ode, solver.pde
    npts = int(ode.v_ode.x.array_ro.size)  # owned + ghost dofs (num_points is a per-marker method on the multi-region solver)
    host_v = np.array(pde.state.x.array_ro)  # D2H
    ke = min(K, 500)
    barrier()
    t_e0 = time.perf_counter()
    ctx.event_record(2 * K + 2)
    for k in range(ke):
        ode.v_ode.x.array[:] = host_v            # host owns V: written into the pinned mirror ...
        ode.from_dolfin()                        # ... H2D + states[v_index] <- v_ode (odesolver.py:168-170)
        solver.step((t, t + dt))
        host_v[:] = pde.state.x.array_ro         # D2H of the result the demos read every step
        t += dt
    ctx.event_record(2 * K + 3)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t_e0) * 1e3
    e2e_ms = max_over_ranks(max(ctx.event_elapsed_ms(2 * K + 2, 2 * K + 3), e2e_wall_ms))
    clk = clocks.stop() if clocks else None
    v_max = max_over_ranks(float(host_v.max()))

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

    mesh = info["mesh"]
    nnz_row = solver.pde._nnz_per_row if hasattr(solver.pde, "_nnz_per_row") else 15.0
    # algorithmic bytes per owned row (DESIGN.md "K2+K4"): SELL entries 12 B (fp64 value + int32 column)
    rhs_b = 12.0 * nnz_row + 16.0
    init_b = 48.0
    iter_b = 12.0 * nnz_row + 96.0
    pde_bytes = n_owned * (rhs_b + init_b + iter_b * iters_per_step)
    ns = info["num_states"]
    ode_bytes = n_owned * (2.0 * 8.0 * ns + 8.0)
    ode_flop = n_owned * nied.tp06.generalized_rush_larsen.fp64_instr_per_node() * 2.0
    roof_pde = {"kernel": f"pde_{args.ksp}_kernel (RHS SpMV + stimulus + Jacobi-PCG, one persistent launch)", "bound": "hbm",
                "achieved": pde_bytes / (pde_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": pde_bytes / (pde_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                "ms_per_launch": pde_ms, "algorithmic_bytes_per_launch": pde_bytes}
    roof_ode = {"kernel": "ode_kernel_regions<tp06_grl1>" if is_lv else "ode_kernel_uniform<tp06_grl1>", "bound": "fp64", "achieved": ode_flop / (ode_ms * 1e-3) / 1e12,
                "peak": dfma_tflops, "unit": "TFLOP/s (fp64-pipe instructions x2)", "frac": ode_flop / (ode_ms * 1e-3) / 1e12 / dfma_tflops,
                "traffic": None, "peak_source": "mono_bench_dfma (measured DFMA rate, this run)", "ms_per_launch": ode_ms,
                "hbm_gbs": ode_bytes / (ode_ms * 1e-3) / 1e9}
    try:  # DRAM bytes per launch of the same kernels from the committed ncu --set full captures (profiles/)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tr = json.load(fh).get(f"{args.workload}/{args.ksp}", {}) if world == 1 else {}
        roof_pde["traffic"], roof_ode["traffic"] = tr.get("pde"), tr.get("ode")
    except (OSError, ValueError):
        pass
    dominant = roof_pde if pde_ms >= ode_ms else roof_ode

    value = n_global * K / (ms_flushed * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_flushed / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {geom_txt}, {n_global} nodes "
                               f"({n_owned} owned by rank 0), TP06 GRL1, Godunov split + CN diffusion "
                               f"({solver.pde.pc_type_used}-preconditioned {args.ksp}, rtol 1e-5, x0={'0' if args.x0 == 'zero' else 'v_'}), dt={dt} ms", "nodes": n_global,
                   "l2": "flushed (256 MiB memset) between timed steps; flush outside the CUDA events",
                   "parallelism": f"{'phi-sector' if is_lv else 'x-slab'} partition, {world} rank(s), one per GPU"},
        "warm_l2": {"value": n_global * K / (ms_warm * 1e-3), "ms_per_step": ms_warm / K,
                    "note": "same K steps back to back, no L2 flush"},
        "x0_previous": fast,
        "e2e": {"value": n_global * ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * npts,
                "d2h_bytes_per_step": 8 * npts, "steps": ke, "ms_per_step": e2e_ms / ke,
                "path": "v_ode.x.array[:]=host_v; ode.from_dolfin(); solver.step((t,t+dt)); host_v[:]=pde.state.x.array"},
        "gpu_launches": int(launches),
        "stages": {"ode_ms_per_step": ode_ms, "pde_ms_per_step": pde_ms, "cg_iterations_per_step": iters_per_step},
        "roofline": {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "peak_source")},
        "roofline_stages": {"pde": roof_pde, "ode": roof_ode},
        "clocks": clk, "setup_s": setup_s, "v_max_mV": v_max,
    }
    if args.matrix_dict:
        line["config"]["matrix_dictionary"] = ctx.pde_dictionary_info()
        # dictionary rows read 1 byte of pattern id instead of 12 z bytes of SELL entries: the roofline above still uses the
        # SELL byte count, so `achieved` can exceed what DRAM actually moved - compare ms_per_step, not frac
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and not is_lv:
            r = cpu_run(dx, dt, steps=2000, warmup=2, budget_s=20.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="niederer_dx0.2", choices=sorted(WORKLOADS) + sorted(LV_WORKLOADS))
    ap.add_argument("--ksp", default="auto", choices=["auto", "cg", "pipecg"],
                    help="Krylov driver of the diffusion solve (PETSc names; auto = pipecg while the CG vectors fit in shared "
                         "memory, cg beyond - same iterates in exact arithmetic)")
    ap.add_argument("--pc", default="auto", choices=["auto", "jacobi", "chebyshev"],
                    help="preconditioner (auto: the reference's hypre request mapped to the fastest native one for the mesh size)")
    ap.add_argument("--x0", default="zero", choices=["zero", "previous"],
                    help="initial guess of the diffusion solve: zero = PETSc default (as the reference runs), previous = v_")
    ap.add_argument("--matrix-dict", action="store_true",
                    help="EXPERIMENTAL: stencil dictionary for the matrix stream of the streaming KSPCG kernel (MONO_PDE_DICT=1); "
                         "bit-identical results, not yet measured")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the x0=v_ extra measurement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one 20 mm block per GPU (slab grows along x), strong = the 20 mm slab split over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
