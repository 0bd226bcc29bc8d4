#!/bin/bash
# round 2, session I (1 GPU): sorted-order phase 1 after the division fix, ncu of it and of the onesweep sort
O=gpurun_out; mkdir -p $O; T=r2i
( timeout 600 python -m pytest tests/test_gpu_dim_sharded.py -x -q 2>&1 | tail -5 ) > $O/${T}_pytest_dim.log
for S in 0 1; do
  timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --sorted $S > $O/${T}_probe_cfg5_w8_s$S.json 2> $O/${T}_probe_cfg5_w8_s$S.err
done
timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --sorted 1 --chunks 4 --pipeline 1 > $O/${T}_probe_cfg5_w8_s1_c4p.json 2> $O/${T}_probe_cfg5_w8_s1_c4p.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/${T}_launches_probe_w8.csv \
  python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 --sorted 1 > $O/${T}_ncu_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_dim_sorted_partial|DeviceRadixSortOnesweep' -s 4 -c 4 \
  -o $O/${T}_prof_sorted python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 --sorted 1 > $O/${T}_ncu_full_sorted.log 2>&1
tail -3 $O/${T}_pytest_dim.log; cat $O/${T}_probe_*.json
