#!/bin/bash
# round 2, session B (1 GPU): L2-prefetch A/B on the reduction, ncu of the dimension-sharded kernels, the new bench.py
O=gpurun_out; mkdir -p $O; T=r2b
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest_all.log
for W in 8 4 2; do
  timeout 300 python tools/dim_probe.py --workload cfg5 --world $W --steps 15 > $O/${T}_probe_cfg5_w$W.json 2> $O/${T}_probe_cfg5_w$W.err
done
KGE_APPLY_PREFETCH=0 timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 > $O/${T}_probe_cfg5_w8_nopf.json 2> $O/${T}_probe_cfg5_w8_nopf.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 8 --steps 30 > $O/${T}_probe_cfg3_w8.json 2> $O/${T}_probe_cfg3_w8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/${T}_launches_probe_w8.csv \
  python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_dim_partial|kge_dim_backward|kge_reduce_apply_group' -s 9 -c 3 \
  -o $O/${T}_prof_dim python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_full_dim.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
KGE_APPLY_PREFETCH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu --no-rank > $O/${T}_bench_cfg3_nopf.json 2> $O/${T}_bench_cfg3_nopf.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
tail -3 $O/${T}_pytest_all.log; cat $O/${T}_probe_*.json; tail -c 1500 $O/${T}_bench_default.err; head -c 1500 $O/${T}_bench_default.json
