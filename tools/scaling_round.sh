#!/bin/bash
# Scaling pass on an N-GPU box.  usage: bash tools/scaling_round.sh TAG NGPUS
T=${1:-r01s}; NG=${2:-8}; O=gpurun_out; mkdir -p $O
run() {  # n, extra bench args..., output name
  local n=$1; local name=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $O/${T}_$name.json 2> $O/${T}_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > $O/${T}_$name.json 2> $O/${T}_$name.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_$name.json").read().strip().splitlines()[-1])
    print("$name", "n=%d"%d["n_gpus"], "%.4g node-steps/s"%d["value"], "%.4f ms/step"%d["ms_per_step"], "warm %.4f"%d["warm_l2"]["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["stages"])
except Exception as e:
    print("$name ERR", e); print(open("$O/${T}_$name.err").read()[-1500:])
PY
}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "$NG" > $O/${T}_mgpu_pytest.log 2>&1; tail -3 $O/${T}_mgpu_pytest.log
for n in 1 2 4 8; do [ $n -le $NG ] && run $n weak_dx0.2_n$n --workload niederer_dx0.2 --scaling weak --secondary none --steps 300 --warmup 10 --no-cpu-baseline; done
STRONG_N=${STRONG_N:-"1 2 4 8"}
for n in $STRONG_N; do [ $n -le $NG ] && run $n strong_dx0.05_n$n --workload niederer_dx0.05 --scaling strong --secondary none --steps 30 --warmup 5 --no-cpu-baseline --no-extras; done
for n in $STRONG_N; do [ $n -le $NG ] && [ $n -gt 1 ] && run $n strong_dx0.025_n$n --workload niederer_dx0.025 --scaling strong --secondary none --steps 20 --warmup 5 --no-cpu-baseline --no-extras; done
for n in 1 8; do [ $n -le $NG ] && run $n weak_lv320k_n$n --workload lv_ellipsoid_320k --scaling weak --secondary none --steps 100 --warmup 5 --no-cpu-baseline --no-extras; done
for n in 1 8; do [ $n -le $NG ] && run $n strong_lv1.4M_n$n --workload lv_ellipsoid_1.4M --scaling strong --secondary none --steps 50 --warmup 5 --no-cpu-baseline --no-extras; done
# BASELINE config 3 on all GPUs: ODE-only, nodes split over the ranks
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29877 tools/ode_sweep.py \
  --models tp06,torord --nodes 1e8 --steps 20 > $O/${T}_ode_sweep_n$NG.jsonl 2> $O/${T}_ode_sweep_n$NG.err; cut -c1-220 $O/${T}_ode_sweep_n$NG.jsonl
