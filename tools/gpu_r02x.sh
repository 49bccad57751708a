#!/bin/bash
# r02x: validation of the round's final state on one GPU: full GPU suite, reference arm, full bench line, launch list
# of the bench command, smoke, config-3 sweep with 2 / 4 / 6 host threads
TAG=r02x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
( time timeout 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err ) 2>&1 | tail -3
python tools/bench_summary.py $TAG < gpurun_out/bench_n1_$TAG.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r02x.json').read().strip().splitlines()[-1])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_by_transfer_mode"])
print("plugin", d["e2e_plugin"]["single"]["ms_per_call"], d["e2e_plugin"]["batch"]["ms_per_call"])
for k,v in d["configs"].items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","parity_frac_within_1e-4","error","value_one_scene_at_a_time","block_wall_s")})
PY
Q="--no-cpu-baseline --no-aux --no-plugin --configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 $Q > gpurun_out/b_launch_$TAG.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
for T in 2 4 6; do
  OMGB_SWEEP_THREADS=$T python tools/bench_configs.py config3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['config3']
print('threads', d.get('host_threads'), 'value', d.get('value'), 'seq', d.get('value_one_scene_at_a_time'), 'same', d.get('concurrent_equals_sequential'), d.get('error'))"
done
