#!/bin/bash
# Streaming KSPCG kernel variants on the streaming sizes (1 GPU): rings (default), dictionary through the L1, tagged exchange.
# usage: bash tools/stream_round.sh TAG [workloads...]
T=${1:-r02s}; shift; O=gpurun_out; mkdir -p $O
WL=${@:-"niederer_dx0.05 niederer_dx0.025"}
for w in $WL; do
  for mode in ring l1 tagged; do
    unset MONO_PDE_TAGGED_STREAM MONO_PDE_NO_RING
    [ $mode = tagged ] && export MONO_PDE_TAGGED_STREAM=1
    [ $mode = l1 ] && export MONO_PDE_NO_RING=1
    timeout 900 python bench.py --workload $w --secondary none --steps 20 --warmup 5 --no-cpu-baseline --no-extras \
      > $O/${T}_bench_${w}_$mode.json 2> $O/${T}_bench_${w}_$mode.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_${w}_$mode.json").read().strip().splitlines()[-1])
    print("$w $mode", "%.4g node-steps/s"%d["value"], "%.4f ms/step"%d["ms_per_step"], d["stages"], "pde frac %.3f"%d["roofline_stages"]["pde"]["frac"], d["selfcheck"]["ok"], d["solver"]["matrix_dictionary"]["active"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print("$w $mode ERR", e); print(open("$O/${T}_bench_${w}_$mode.err").read()[-1500:])
PY
  done
done
unset MONO_PDE_TAGGED_STREAM MONO_PDE_NO_RING
