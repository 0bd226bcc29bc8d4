#!/bin/bash
# One GPU-box session: parity tests, bench lines (cfg3 default + others), ncu launch list and full captures.
# usage: tools/gpu_session.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${TAG}_clocks.csv &
SMI=$!
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
for c in cfg1 cfg2 cfg4 cfg5; do
  timeout 600 python bench.py --workload $c --steps 30 --warmup 5 --no-cpu > $O/${TAG}_bench_$c.json 2> $O/${TAG}_bench_$c.err
done
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --rank-steps 1 > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span' -s 12 -c 6 \
  -o $O/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-rank > $O/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_rank_tc_kernel|kge_rank_sweep' -s 1 -c 2 \
  -o $O/${TAG}_prof_rank python bench.py --steps 3 --warmup 3 --no-cpu --rank-steps 1 > $O/${TAG}_ncu_fullr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_fwd_bwd|kge_reduce_apply|kge_span' -s 12 -c 6 \
  -o $O/${TAG}_prof_cfg5 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-rank > $O/${TAG}_ncu_full5.log 2>&1
# TransE distance sweep (cfg1) and the A/B knobs: one line each
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kge_rank_sweep2 -s 1 -c 1 \
  -o $O/${TAG}_prof_sweep2 python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu --rank-steps 1 > $O/${TAG}_ncu_sweep2.log 2>&1
for kv in KGE_PIPELINE=0 KGE_SPAN_WARPS=8 KGE_FWD_MAXCTAS=4 KGE_FWD_MAXCTAS=5 KGE_FWD_PIPE=0 KGE_REDUCE_STAGED=1; do
  env $kv timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu --no-rank > $O/${TAG}_ab_${kv}.json 2> $O/${TAG}_ab_${kv}.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${TAG}_ab_*.json")) + ["$O/${TAG}_bench_cfg3.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-46s flushed %.4f warm %.4f e2e %.4f" % (f.split("/")[-1], d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/${TAG}_pytest.log
head -c 600 $O/${TAG}_bench_cfg3.json
