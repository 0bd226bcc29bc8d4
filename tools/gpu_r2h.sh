#!/bin/bash
# round 2, session H (1 GPU): sorted-order phase 1; prefetch A/B on the latency-bound group reduction
O=gpurun_out; mkdir -p $O; T=r2h
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${T}_pytest_all.log
for W in 8 4 2; do
  for S in 0 1; do
    timeout 300 python tools/dim_probe.py --workload cfg5 --world $W --steps 15 --sorted $S > $O/${T}_probe_cfg5_w${W}_s$S.json 2> $O/${T}_probe_cfg5_w${W}_s$S.err
  done
done
KGE_APPLY_PREFETCH=1 timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --sorted 1 > $O/${T}_probe_cfg5_w8_s1_pf.json 2> $O/${T}_probe_cfg5_w8_s1_pf.err
timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --sorted 1 --chunks 4 --pipeline 1 > $O/${T}_probe_cfg5_w8_s1_c4p.json 2> $O/${T}_probe_cfg5_w8_s1_c4p.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 8 --steps 30 --sorted 1 > $O/${T}_probe_cfg3_w8_s1.json 2> $O/${T}_probe_cfg3_w8_s1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/${T}_launches_probe_w8.csv \
  python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 --sorted 1 > $O/${T}_ncu_probe.log 2>&1
tail -4 $O/${T}_pytest_all.log; cat $O/${T}_probe_*.json
