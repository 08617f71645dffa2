"""Static analysis of the generated cell-model programs (CPU only): weighted work, critical path and the parallelism a
single node's update offers - the numbers behind the "small meshes are latency bound" paragraph of DESIGN.md section 2 and
the starting point for splitting one node across several lanes.

    python tools/ode_dag_analysis.py [tp06 torord fhn]        (reads the .ode files below $MONO_ODES, default /root/reference/odes)

Weights = fp64-pipe instructions of each operation as emitted (ode_math.cuh / libdevice): add/mul/neg/select 1, div 7,
rcp 6, exp 17, log 25, sqrt 12, pow = exp + log + 1.  The expansions of div/exp/log are themselves dependent chains, so
the same weights serve as latencies in units of one dependent fp64 instruction."""

from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))

from codegen import generate, ir  # noqa: E402
from codegen.program import build_program  # noqa: E402

W = {"add": 1, "sub": 1, "mul": 1, "neg": 0, "div": 7, "pow": 43, "cond": 1, "lt": 1, "gt": 1, "le": 1, "ge": 1, "eq": 1, "ne": 1, "and": 0,
     "or": 0, "num": 0, "sym": 0}
CALL = {"exp": 17, "log": 25, "sqrt": 12, "floor": 1, "abs": 0}


def weight(n) -> int:
    if n.kind == "call":
        return CALL[n.value]
    if n.kind == "div" and ir.is_num(n.args[0], 1.0):
        return 6
    return W[n.kind]


def analyse(prog):
    env = dict(prog.body)
    uniform = {name for name, _ in prog.uniform}
    depth: dict[int, int] = {}
    seen: dict[int, object] = {}

    def walk(n) -> int:
        d = depth.get(id(n))
        if d is not None:
            return d
        if n.kind == "sym":
            d = walk(env[n.value]) if n.value in env else 0  # states, parameters, hoisted constants: ready at t = 0
        else:
            d = weight(n) + max((walk(a) for a in n.args), default=0)
            seen[id(n)] = n
        depth[id(n)] = d
        return d

    cp_out = [walk(e) for e in prog.outputs]
    work = sum(weight(n) for n in seen.values())
    return {"work": work, "critical_path": max(cp_out), "parallelism": work / max(cp_out), "per_output_depth": cp_out,
            "n_uniform": len(uniform)}


def main(argv):
    tags = argv or ["tp06", "torord", "fhn"]
    for tag in tags:
        model, _ = generate.load_model(tag, os.environ.get("MONO_ODES", "/root/reference/odes"))
        for scheme in ("forward_explicit_euler", "generalized_rush_larsen"):
            prog = build_program(model, scheme)
            r = analyse(prog)
            order = sorted(zip(r["per_output_depth"], model.states), reverse=True)[:4]
            print(f"{tag:7s} {scheme:24s} work {r['work']:6d}  critical path {r['critical_path']:5d}  work/path {r['parallelism']:5.1f}"
                  f"  deepest: " + ", ".join(f"{s}={d}" for d, s in order))


if __name__ == "__main__":
    main(sys.argv[1:])
