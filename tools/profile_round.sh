#!/bin/bash
# One measurement pass on a 1-GPU box: tests, bench (both arms, default workload = 27 M-dof slab), the dx = 0.2 slab with ncu
# launch list + full captures, ODE sweep, other slabs, LV shells.
# usage: bash tools/profile_round.sh TAG     (outputs under gpurun_out/TAG_*)
T=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 600 $O/${T}_bench.json
timeout 400 python bench.py --impl reference --steps 50 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; tail -c 400 $O/${T}_bench_reference.json
# ncu launch list of the same command (after it ran clean above)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches.csv \
  python bench.py --workload niederer_dx0.2 --scaling weak --secondary none --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_ncu_launches.log 2>&1; tail -1 $O/${T}_ncu_launches.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pde_pipecg -s 10 -c 1 -f -o $O/${T}_pde_pipecg \
  python bench.py --workload niederer_dx0.2 --scaling weak --secondary none --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_ncu_pde.log 2>&1; tail -1 $O/${T}_ncu_pde.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ode_kernel -s 10 -c 1 -f -o $O/${T}_ode_tp06 \
  python bench.py --workload niederer_dx0.2 --scaling weak --secondary none --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_ncu_ode.log 2>&1; tail -1 $O/${T}_ncu_ode.log | cut -c1-200
timeout 600 python tools/ode_sweep.py --models tp06,torord --nodes 1e6,1e7,1e8 --steps 20 > $O/${T}_ode_sweep.jsonl 2> $O/${T}_ode_sweep.err; cut -c1-230 $O/${T}_ode_sweep.jsonl
for w in niederer_dx0.1 niederer_dx0.05 niederer_dx0.025; do
  timeout 900 python bench.py --workload $w --secondary none --steps 30 --warmup 5 --no-cpu-baseline > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "%.4g"%d["value"], d["ms_per_step"], d["stages"], "pde frac %.3f ode frac %.3f"%(d["roofline_stages"]["pde"]["frac"], d["roofline_stages"]["ode"]["frac"]), "x0prev", d["x0_previous"] and d["x0_previous"]["value"], "setup", d["setup_s"])
except Exception as e:
    print("$w ERR", e); print(open("$O/${T}_bench_$w.err").read()[-800:])
PY
done
for w in lv_ellipsoid_320k lv_ellipsoid_1.4M; do
  timeout 900 python bench.py --workload $w --secondary none --steps 100 --warmup 5 > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "%.4g"%d["value"], d["ms_per_step"], d["stages"], "pde frac %.3f ode frac %.3f"%(d["roofline_stages"]["pde"]["frac"], d["roofline_stages"]["ode"]["frac"]), "setup", d["setup_s"])
except Exception as e:
    print("$w ERR", e); print(open("$O/${T}_bench_$w.err").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pde_cg -s 6 -c 1 -f -o $O/${T}_pde_cg_stream \
  python bench.py --workload niederer_dx0.05 --secondary none --steps 6 --warmup 3 --no-cpu-baseline --no-extras > $O/${T}_ncu_pde_stream.log 2>&1; tail -1 $O/${T}_ncu_pde_stream.log | cut -c1-200
