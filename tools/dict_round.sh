#!/bin/bash
# The stencil dictionary on / off with the ROUND-1 streaming kernel (tagged exchange, MONO_PDE_TAGGED_STREAM=1) on a 1-GPU
# box - the measurement that made the dictionary the default (profiles/r02d_*) - then one ncu capture of that kernel.
# (tools/stream_round.sh compares the current streaming kernel's modes.)
# usage: bash tools/dict_round.sh TAG     (outputs under gpurun_out/TAG_*)
T=${1:-r02d}; O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_zz_input_validation_gpu.py tests/test_zzz_full_size_properties_gpu.py -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
for w in niederer_dx0.1 niederer_dx0.05 niederer_dx0.025; do
  for mode in sell dict; do
    extra="--no-matrix-dict"; [ $mode = dict ] && extra="--matrix-dict"
    export MONO_PDE_TAGGED_STREAM=1
    [ $w = niederer_dx0.1 ] && export MONO_PDE_STREAM=1 || unset MONO_PDE_STREAM   # 442 k rows would run resident
    timeout 900 python bench.py --workload $w --secondary none --ksp cg --steps 30 --warmup 5 --no-cpu-baseline --no-extras $extra \
      > $O/${T}_bench_${w}_$mode.json 2> $O/${T}_bench_${w}_$mode.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_${w}_$mode.json").read().strip().splitlines()[-1])
    print("$w $mode", "%.4g node-steps/s"%d["value"], "%.4f ms/step"%d["ms_per_step"], d["stages"], d["solver"]["matrix_dictionary"])
except Exception as e:
    print("$w $mode ERR", e); print(open("$O/${T}_bench_${w}_$mode.err").read()[-800:])
PY
  done
done
unset MONO_PDE_STREAM
export MONO_PDE_TAGGED_STREAM=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pde_cg -s 6 -c 1 -f -o $O/${T}_pde_cg_dict \
  python bench.py --workload niederer_dx0.05 --secondary none --ksp cg --steps 6 --warmup 3 --no-cpu-baseline --no-extras --matrix-dict > $O/${T}_ncu_pde_dict.log 2>&1
tail -1 $O/${T}_ncu_pde_dict.log | cut -c1-200
