#!/bin/bash
# The north_star's target on an 8-GPU box: the 100 M-dof Niederer slab (dx = 0.016 mm, 103.8 M dofs), throughput with and
# without the stencil dictionary, then the full activation-time run from the device-side probes.
# usage: bash tools/northstar_round.sh TAG [NGPUS]        (outputs under gpurun_out/TAG_*; about 10 minutes of box time)
T=${1:-r02n}; NG=${2:-8}; O=gpurun_out; mkdir -p $O
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for mode in sell dict; do
  extra=""; [ $mode = dict ] && extra="--matrix-dict"
  timeout 900 bash -c "$(declare -f tr); NG=$NG; tr bench.py --gpus $NG --workload niederer_dx0.016 --scaling strong --steps 10 --warmup 3 --no-cpu-baseline --no-extras $extra" \
    > $O/${T}_strong_dx0.016_n${NG}_$mode.json 2> $O/${T}_strong_dx0.016_n${NG}_$mode.err
  tail -c 700 $O/${T}_strong_dx0.016_n${NG}_$mode.json; echo
done
# full run to T = 45 ms (4500 steps); add --matrix-dict once the dictionary has been measured faster
timeout 1500 bash -c "$(declare -f tr); NG=$NG; tr tools/niederer_activation.py --dx 0.016 --dt 0.01 --T 45" \
  > $O/${T}_activation_dx0.016_n${NG}.json 2> $O/${T}_activation_dx0.016_n${NG}.err
tail -c 900 $O/${T}_activation_dx0.016_n${NG}.json; echo
# BASELINE config 5 at its named size: the 103 M-dof LV shell, 12.9 M dofs per GPU (set-up ~2 min per rank)
timeout 1500 bash -c "$(declare -f tr); NG=$NG; tr bench.py --gpus $NG --workload lv_ellipsoid_100M --scaling strong --steps 10 --warmup 3 --no-cpu-baseline --no-extras" \
  > $O/${T}_strong_lv100M_n${NG}.json 2> $O/${T}_strong_lv100M_n${NG}.err
tail -c 700 $O/${T}_strong_lv100M_n${NG}.json; echo
# the published dx = 0.1 rows on one GPU, for the convergence of the activation times towards the fine mesh
for dt in 0.05 0.01; do
  timeout 600 python tools/niederer_activation.py --dx 0.1 --dt $dt --T 45 > $O/${T}_activation_dx0.1_dt$dt.json 2> $O/${T}_activation_dx0.1_dt$dt.err
  tail -c 600 $O/${T}_activation_dx0.1_dt$dt.json; echo
done
