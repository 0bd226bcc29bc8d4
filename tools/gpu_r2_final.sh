#!/bin/bash
# round 2, final 1-GPU session: parity tests, ncu captures (traffic of THIS session feeds the bench line), the default bench
# line (cfg3 + cfg5 sub-record + others), the reference arm, single-config lines, ncu launch list.  usage: gpu_r2_final.sh TAG
T=${1:-r2z}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $O/${T}_smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span' -s 12 -c 6 \
  -o $O/${T}_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_rank_tc_kernel' -s 1 -c 1 \
  -o $O/${T}_prof_rank python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_fullr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span' -s 12 -c 6 \
  -o $O/${T}_prof_cfg5 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_full5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_dim_partial|kge_dim_backward|kge_reduce_apply_group' -s 9 -c 3 \
  -o $O/${T}_prof_dim python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_full_dim.log 2>&1
python tools/ncu_traffic.py $T $O/traffic_session.json cfg3=$O/${T}_prof.ncu-rep cfg3=$O/${T}_prof_rank.ncu-rep cfg5=$O/${T}_prof_cfg5.ncu-rep cfg5w8=$O/${T}_prof_dim.ncu-rep > $O/${T}_traffic.log 2>&1
cp $O/traffic_session.json $O/${T}_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --rank-steps 1 --no-sub > $O/${T}_ncu_bench.log 2>&1
for kv in KGE_FWD_MAXCTAS=4 KGE_FWD_MAXCTAS=3; do
  env $kv timeout 200 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ab5_${kv}.json 2> $O/${T}_ab5_${kv}.err
done
for kv in KGE_APPLY_SPLIT=1 KGE_APPLY_SPLIT=0 KGE_FWD_MAXCTAS=4 KGE_FWD_MAXCTAS=5; do
  env $kv timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-rank --no-sub > $O/${T}_ab_${kv}.json 2> $O/${T}_ab_${kv}.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${T}_ab*_*.json")) + ["$O/${T}_bench_default.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-46s flushed %.4f warm %.4f e2e %.4f" % (f.split("/")[-1], d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k: round(v, 4) for k, v in d["roofline"]["phases_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
d = json.loads(open("$O/${T}_bench_default.json").read().strip().splitlines()[-1])
print("roofline", d["roofline"]["frac"], d["roofline"]["step_frac"], d["roofline"]["traffic"], d["roofline"]["traffic_source"])
print("rank", d["rank"]["value"], d["rank"]["ms_per_step"], d["rank"]["roofline"]["frac"], d["rank"]["mrr"], d["rank_parity"])
print("cfg5", {k: d["cfg5"].get(k) for k in ("value", "ms_per_step", "value_warm_l2", "rank_parity", "first_step_loss")})
print("others", {k: (v.get("value"), v.get("ms_per_step"), v.get("e2e", {}).get("value")) for k, v in d["others"].items()})
print("cpu", d.get("cpu_baseline"))
PY
tail -3 $O/${T}_pytest.log; cat $O/${T}_smoke.log
