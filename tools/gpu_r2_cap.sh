#!/bin/bash
# round 2: ncu captures exported as CSV on the box (the .ncu-rep files are too large to travel), the traffic JSON of this
# session, then the default bench line that quotes it, the reference arm and the launch list.  usage: gpu_r2_cap.sh TAG
T=${1:-r2y}; O=gpurun_out; mkdir -p $O; S=/tmp/ncu_$T; mkdir -p $S
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span|kge_small_sort' -s 16 -c 8 \
  -o $S/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_rank_tc_kernel' -s 1 -c 1 \
  -o $S/prof_rank python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_fullr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span' -s 12 -c 6 \
  -o $S/prof_cfg5 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_full5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_dim_partial|kge_dim_backward|kge_reduce_apply_group' -s 9 -c 3 \
  -o $S/prof_dim python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_full_dim.log 2>&1
for n in prof prof_rank prof_cfg5 prof_dim; do
  ncu -i $S/$n.ncu-rep --page raw --csv > $O/${T}_${n}_raw.csv 2>/dev/null
done
ncu -i $S/prof.ncu-rep --page source --csv -k regex:'kge_small_sort' > $O/${T}_prof_sort_source.csv 2>/dev/null
python tools/ncu_traffic.py $T $O/traffic_session.json cfg3=$S/prof.ncu-rep cfg3=$S/prof_rank.ncu-rep cfg5=$S/prof_cfg5.ncu-rep cfg5w8=$S/prof_dim.ncu-rep > $O/${T}_traffic.log 2>&1
cp $O/traffic_session.json $O/${T}_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kge_|RadixSort|DeviceSelect' -c 400 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --rank-steps 1 --no-sub > $O/${T}_ncu_bench.log 2>&1
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $O/${T}_smoke.log; cat $O/${T}_smoke.log
ls -la $O | tail -20; du -sh $O
python - <<PY
import json
d = json.loads(open("$O/${T}_bench_default.json").read().strip().splitlines()[-1])
print("cfg3 ms/step", d["ms_per_step"], "warm", d["ms_per_step_warm"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["phases_ms"])
print("roofline", d["roofline"]["frac"], d["roofline"]["step_frac"], d["roofline"]["traffic"], d["roofline"]["traffic_source"])
print("rank", d["rank"]["value"], d["rank"]["ms_per_step"], d["rank"]["roofline"]["frac"], d["rank"]["mrr"], d["rank_parity"])
print("cfg5", {k: d["cfg5"].get(k) for k in ("value", "ms_per_step", "value_warm_l2", "rank_parity", "first_step_loss")})
print("others", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value")) for k, v in d["others"].items()})
print("cpu", d.get("cpu_baseline"))
print("ref", open("$O/${T}_bench_ref.json").read()[:600])
PY
