"""BASELINE config 3: ODE-only throughput sweep (no diffusion) on N independent nodes, TP06 and ToR-ORd, GRL1.

    python tools/ode_sweep.py [--models tp06,torord] [--nodes 1e6,1e7,1e8] [--steps 50] [--warmup 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ode_sweep.py ...   (N GPUs: the nodes
        are split evenly over the ranks - no communication, the stage is embarrassingly parallel; time = max over ranks)

States = IC x (1 + 1e-3 U(-1,1)), V ~ U(-90, 40) mV, numpy default_rng(1234); shared parameters;
dt = 0.01.  One JSON line per (model, N): node-steps/s, achieved fp64 instruction rate vs the DFMA peak measured in
the same process, and state traffic in GB/s vs MEASURED_PEAKS.json.
"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fenicsx-beat_b200")):
    sys.path.insert(0, p)
import numpy as np
from beat_b200._lib import Context

MODEL_ID = {"fhn": 0, "tp06": 1, "torord": 2}
VNAME = {"fhn": "v", "tp06": "V", "torord": "v"}


# what ties the cell model's arithmetic to the reference
PIN = {"tp06": "published Niederer activation times (demos/niederer_benchmark.py:315-325) through the oracle",
       "fhn": "README.md:58-89 closed form (forward Euler); GRL1: device <-> oracle only",
       "torord": "UNPINNED: no reference test or demo output exercises ToR-ORd; device <-> oracle only"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="tp06,torord")
    ap.add_argument("--nodes", default="1e6,1e7")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--scheme", default="generalized_rush_larsen")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for tag in args.models.split(","):
        hm = importlib.import_module(f"beat_b200.models.{tag}")
        dev = getattr(hm, args.scheme)
        y0 = hm.init_state_values()
        prm = hm.init_parameter_values()
        ns = len(y0)
        for ntxt in args.nodes.split(","):
            n_total = int(float(ntxt))
            n = n_total // world + (1 if rank < n_total % world else 0)
            ctx = Context(local if world > 1 else int(os.environ.get("MONO_DEVICE", "0")))
            dfma = ctx.bench_dfma()
            ctx.ode_create(MODEL_ID[tag], dev.scheme_id, n, hm.state_index(VNAME[tag]), ns)
            rng = np.random.default_rng(1234)
            chunk = min(n, 1 << 20)  # the same perturbed block repeated: keeps host memory small at 1e8
            blk = np.repeat(y0[:, None], chunk, axis=1) * (1 + 1e-3 * rng.uniform(-1, 1, (ns, chunk)))
            blk[hm.state_index(VNAME[tag])] = rng.uniform(-90.0, 40.0, chunk)
            for row in range(ns):
                ctx.ode_set_state_row(row, np.ascontiguousarray(np.resize(blk[row], n)))
            ctx.ode_set_params(prm, dev.derived(prm))
            t, dt = 0.0, 0.01
            for _ in range(args.warmup):
                ctx.ode_step(t, dt); t += dt
            ctx.sync()
            if dist is not None:
                dist.barrier()
            ctx.timer_start(0)
            for _ in range(args.steps):
                ctx.ode_step(t, dt); t += dt
            ctx.timer_stop(0)
            ms = ctx.timer_elapsed_ms(0) / args.steps
            if dist is not None:
                tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                ms = float(tmax.item())
            v = ctx.ode_get_state_row(hm.state_index(VNAME[tag]))
            instr = dev.fp64_instr_per_node()
            flops = 2.0 * instr * n / (ms * 1e-3) / 1e12   # per GPU (rank 0's share)
            gbs = (2 * 8 * ns) * n / (ms * 1e-3) / 1e9
            if rank == 0:
                print(json.dumps({"workload": f"ode_only {tag} {args.scheme}", "n_gpus": world, "nodes": n_total, "ms_per_step": ms,
                                  "node_steps_per_s": n_total / (ms * 1e-3), "per_gpu": {"fp64_tflops_equiv": flops, "dfma_peak_tflops": dfma,
                                  "fp64_frac": flops / dfma, "state_gbs": gbs, "hbm_frac": gbs / hbm}, "fp64_instr_per_node": instr,
                                  "finite": bool(np.isfinite(v).all()), "variant": os.environ.get("MONO_ODE_VARIANT", "auto"),
                                  "reference_pin": PIN[tag]}), flush=True)
            ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
