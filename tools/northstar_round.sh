#!/bin/bash
# The north_star's target runs.  usage: bash tools/northstar_round.sh TAG NGPUS   (outputs under gpurun_out/TAG_*)
#   NGPUS = 8 (or 4, 2): strong scaling of the 27 M-dof slab (bench default) and of the 100 M-dof slab (dx = 0.016 mm,
#           103.8 M dofs), the full activation-time run of the 100 M-dof slab from the device-side probes, the dx = 0.1
#           activation run (to compare with the 1-GPU one), the LV shell weak-scaled to ~105 M dofs.
#   NGPUS = 1: the same workloads on one GPU (the strong-scaling baselines; dx = 0.1 activation runs at dt = 0.05 / 0.01).
T=${1:-r02n}; NG=${2:-8}; O=gpurun_out; mkdir -p $O
free -g | head -2; nproc
if [ "$NG" -gt 1 ]; then
  run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
else
  run() { timeout $1 python "${@:2}"; }
fi
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if "value" in d:
        print(sys.argv[1], "%.4g node-steps/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], d["stages"], "selfcheck", d["selfcheck"]["ok"],
              "v_sum %.9g" % d["selfcheck"]["v_sum"], "setup %.0f s" % d["setup_s"], d["solver"]["matrix_dictionary"],
              "pde frac %.3f ode frac %.3f" % (d["roofline_stages"]["pde"]["frac"], d["roofline_stages"]["ode"]["frac"]))
        if d.get("secondary") and "value" in d["secondary"]:
            s = d["secondary"]; print("   secondary %.4g node-steps/s %.4f ms/step" % (s["value"], s["ms_per_step"]), s["stages"], s["solver"]["pc"], s["selfcheck"]["ok"])
    else:
        print(sys.argv[1], json.dumps(d)[:900])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
# 1. bench default: 27 M-dof slab, strong scaling (+ the 58 k-node-per-GPU secondary)
run 600 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench27M_n$NG.json 2> $O/${T}_bench27M_n$NG.err; show $O/${T}_bench27M_n$NG.json
# 2. the 100 M-dof slab
run 900 bench.py --gpus $NG --workload niederer_dx0.016 --secondary none --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench100M_n$NG.json 2> $O/${T}_bench100M_n$NG.err; show $O/${T}_bench100M_n$NG.json
# 3. activation times: dx = 0.1 (published rows; 1 GPU also at dt = 0.05), and on several GPUs the full run of the 100 M-dof slab
run 400 tools/niederer_activation.py --dx 0.1 --dt 0.01 --T 45 > $O/${T}_activation_dx0.1_dt0.01_n$NG.json 2> $O/${T}_activation_dx0.1_dt0.01_n$NG.err; show $O/${T}_activation_dx0.1_dt0.01_n$NG.json
if [ "$NG" -eq 1 ]; then
  run 400 tools/niederer_activation.py --dx 0.1 --dt 0.05 --T 45 > $O/${T}_activation_dx0.1_dt0.05_n$NG.json 2> $O/${T}_activation_dx0.1_dt0.05_n$NG.err; show $O/${T}_activation_dx0.1_dt0.05_n$NG.json
else
  run 900 tools/niederer_activation.py --dx 0.016 --dt 0.01 --T 45 > $O/${T}_activation_dx0.016_n$NG.json 2> $O/${T}_activation_dx0.016_n$NG.err; show $O/${T}_activation_dx0.016_n$NG.json
fi
# 4. BASELINE config 5: the LV shell, 13.2 M dofs per GPU (weak scaling; ~105 M dofs on 8 GPUs; set-up ~2 min and ~23 GB of host memory per rank)
if [ "$(free -g | awk '/Mem:/{print $7}')" -ge $((28 * NG)) ]; then
  run 900 bench.py --gpus $NG --workload lv_ellipsoid_13M --scaling weak --secondary none --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_lv13Mweak_n$NG.json 2> $O/${T}_lv13Mweak_n$NG.err; show $O/${T}_lv13Mweak_n$NG.json
else
  echo "LV run skipped: not enough host memory"
fi
