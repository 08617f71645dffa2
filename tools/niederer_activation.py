"""The Niederer benchmark run itself (demos/niederer_benchmark.py:233-289) on 1..8 GPUs: activation times of P1..P9 from
the device-side probes, no field read-back per step.  This is the north_star's target run at its full size:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/niederer_activation.py --dx 0.016 --dt 0.01 --T 45        # 103.8 M dofs

    python tools/niederer_activation.py --dx 0.1 --dt 0.05 --T 45       # one GPU, a published row (:323)

Prints one JSON line: activation times, the published row for (dx, dt) when there is one, node-steps/s of the whole run.
A probe that lies on a partition boundary is registered by both ranks; the merged time is the one any rank reports.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))

PUBLISHED = {  # demos/niederer_benchmark.py:317-325, (dx, dt) -> P1..P9 in ms
    (0.5, 0.05): [1.25, 51.1, 34.9, 58.9, 14.1, 49.5, 34, 56.65, 26.05],
    (0.5, 0.01): [1.22, 50.85, 33.96, 58.05, 13.98, 49.36, 33.07, 55.91, 25.64],
    (0.5, 0.005): [1.215, 50.775, 33.825, 57.96, 13.97, 49.345, 32.945, 55.825, 25.595],
    (0.2, 0.05): [1.25, 29.7, 32.9, 40.2, 9.55, 30, 32.95, 39.9, 18.9],
    (0.2, 0.01): [1.24, 29.09, 31.25, 38.66, 9.34, 29.4, 31.29, 38.42, 18.14],
    (0.2, 0.005): [1.235, 29.015, 31.05, 38.475, 9.315, 29.32, 31.08, 38.235, 18.045],
    (0.1, 0.05): [1.25, 26.85, 33.3, 40.35, 8.4, 27.5, 33.85, 40.55, 18.95],
    (0.1, 0.01): [1.23, 25.64, 31.46, 38.08, 8.03, 26.24, 31.94, 38.21, 17.95],
    (0.1, 0.005): [1.225, 25.5, 31.26, 37.81, 7.99, 26.09, 31.72, 37.93, 17.835],
}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--dx", type=float, default=0.2)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--T", type=float, default=45.0, help="ms; the last corner of the published rows activates before 41 ms for dx <= 0.2")
    ap.add_argument("--ksp", default="auto", choices=["auto", "cg", "pipecg"])
    ap.add_argument("--rtol", type=float, default=None, help="default: PETSc's 1e-5, as the demo runs it")
    ap.add_argument("--chunk", type=int, default=500, help="steps per mono_split_solve call (progress / early stop granularity)")
    ap.add_argument("--no-matrix-dict", action="store_true", help="disable the stencil dictionary (MONO_PDE_DICT=0)")
    args = ap.parse_args(argv)

    import numpy as np
    import torch
    import torch.distributed as dist

    from beat_b200 import fem, niederer

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.no_matrix_dict:
        os.environ["MONO_PDE_DICT"] = "0"
    t_setup = time.perf_counter()
    solver, info = niederer.setup(dx=args.dx, comm=fem.Comm(rank, world), rtol=args.rtol, ksp_type=args.ksp, probes=True)
    ctx = solver.pde._ctx
    names = list(niederer.POINTS)
    mine = {name: pid for name, pid in info["probe_ids"].items() if pid is not None}
    setup_s = time.perf_counter() - t_setup

    def merged_times() -> np.ndarray:
        # EVERY rank drains its own stream before the torch collective below: a rank without probes would otherwise launch
        # the NCCL kernel while its persistent PDE kernels (cooperative, one CTA with all registers per SM) are still queued
        # behind it - they cannot become resident next to the NCCL CTAs, NCCL waits for the peers, the peers wait in
        # cudaStreamSynchronize for kernels that wait for this rank: a deadlock (seen on 8 GPUs, r02n).
        ctx.sync()
        local = np.full(len(names), -1.0)
        if mine:
            act = solver.pde.activation_times()
            for name, pid in mine.items():
                local[names.index(name)] = act[pid]
        if world > 1:
            t = torch.tensor(local, dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # -1 = not activated / not on this rank
            local = t.cpu().numpy()
        return local

    nsteps = int(round(args.T / args.dt))
    ctx.sync()
    if world > 1:
        dist.barrier()
    t_run = time.perf_counter()
    done, t = 0, 0.0
    act = merged_times()
    while done < nsteps:
        n = min(args.chunk, nsteps - done)
        solver.solve_on_device(t, args.dt, n)
        done += n
        t = done * args.dt
        act = merged_times()  # synchronises
        if (act >= 0).all():
            break
    ctx.sync()
    if world > 1:
        dist.barrier()
    run_s = time.perf_counter() - t_run
    its, solves = ctx.ksp_totals()
    if rank == 0:
        pub = PUBLISHED.get((round(args.dx, 6), round(args.dt, 6)))
        line = {
            "workload": f"Niederer slab 20x7x3 mm, dx={args.dx} mm, dt={args.dt} ms, TP06 GRL1 + CN, {info['n_global']} dofs, {world} GPU(s)",
            "activation_times_ms": dict(zip(names, [round(float(a), 6) for a in act])),
            "published_ms": dict(zip(names, pub)) if pub else None,
            "max_abs_diff_to_published_ms": float(np.abs(act - np.array(pub)).max()) if pub else None,
            "steps": done, "run_s": run_s, "setup_s": setup_s, "node_steps_per_s": info["n_global"] * done / run_s,
            "cg_iterations_per_step": its / max(solves, 1), "ksp": solver.pde.ksp_type_used, "pc": solver.pde.pc_type_used,
            "matrix_dictionary": ctx.pde_dictionary_info(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
