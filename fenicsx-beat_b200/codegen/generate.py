"""Generate the committed artefacts for every supported cell model.

    python -m codegen.generate [--odes /root/reference/odes]          (run from fenicsx-beat_b200/)

Outputs (all committed; the GPU box has no /root/reference):
  csrc/generated/<tag>.cuh            device code, forward Euler + GRL1
  beat_b200/models/<tag>.py           host metadata + device handles (gotranx-module-like surface)
  csrc/generated/manifest.json        model ids, sizes, op counts, sha256 of the .ode inputs

Model ids are part of the C ABI (include/mono_abi.h: MONO_MODEL_*).
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from codegen.emit import emit_cuda, emit_host_module  # noqa: E402
from codegen.odefile import BUILTIN_ODES, load_ode, parse_ode  # noqa: E402
from codegen.program import SCHEMES, build_program, op_counts  # noqa: E402

# tag -> (model id in the C ABI, path below the odes/ root or None for built-in text)
MODELS = {
    "fhn": (0, None),
    "tp06": (1, "tentusscher_panfilov_2006/tentusscher_panfilov_2006_epi_cell.ode"),
    "torord": (2, "torord/ToRORd_dynCl_endo.ode"),
    "simple": (3, None),
}


def load_model(tag: str, odes_root: str):
    mid, rel = MODELS[tag]
    if rel is None:
        name, text = BUILTIN_ODES[tag]
        return parse_ode(text, name), hashlib.sha256(text.encode()).hexdigest()
    path = os.path.join(odes_root, rel)
    with open(path, "rb") as fh:
        digest = hashlib.sha256(fh.read()).hexdigest()
    return load_ode(path), digest


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--odes", default="/root/reference/odes")
    ap.add_argument("--check", action="store_true", help="fail if the committed artefacts differ")
    args = ap.parse_args(argv)

    manifest = {}
    outputs: dict[str, str] = {}
    for tag, (mid, _rel) in MODELS.items():
        model, digest = load_model(tag, args.odes)
        progs = [build_program(model, s) for s in SCHEMES]
        outputs[os.path.join(PKG, "csrc", "generated", f"{tag}.cuh")] = emit_cuda(progs, tag)
        outputs[os.path.join(PKG, "beat_b200", "models", f"{tag}.py")] = emit_host_module(progs, tag, mid)
        manifest[tag] = {
            "model_id": mid,
            "source": model.name,
            "source_sha256": digest,
            "num_states": len(model.states),
            "num_parameters": len(model.parameters),
            "states": model.state_names,
            "parameters": model.parameter_names,
            "schemes": {
                p.scheme: {
                    "rush_larsen_states": p.rl_states,
                    "forward_euler_states": p.fe_states,
                    "num_derived": len(p.uniform),
                    "op_counts": op_counts(p),
                }
                for p in progs
            },
        }
    outputs[os.path.join(PKG, "csrc", "generated", "manifest.json")] = json.dumps(manifest, indent=1) + "\n"

    rc = 0
    for path, text in outputs.items():
        if args.check:
            old = open(path).read() if os.path.exists(path) else None
            if old != text:
                print(f"STALE: {path}")
                rc = 1
        else:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as fh:
                fh.write(text)
            print(f"wrote {path} ({len(text)} bytes)")
    return rc


if __name__ == "__main__":
    raise SystemExit(main())
