#!/bin/bash
# round 2, session E (1 GPU): branch-free group reduction
O=gpurun_out; mkdir -p $O; T=r2g
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${T}_pytest_all.log
for W in 8 4 2; do
  timeout 300 python tools/dim_probe.py --workload cfg5 --world $W --steps 15 > $O/${T}_probe_cfg5_w$W.json 2> $O/${T}_probe_cfg5_w$W.err
done
timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --chunks 2 --pipeline 1 > $O/${T}_probe_cfg5_w8_c2p.json 2> $O/${T}_probe_cfg5_w8_c2p.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 8 --steps 30 > $O/${T}_probe_cfg3_w8.json 2> $O/${T}_probe_cfg3_w8.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu --no-rank > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_reduce_apply_group' -s 3 -c 1 \
  -o $O/${T}_prof_group python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_full_group.log 2>&1
tail -4 $O/${T}_pytest_all.log; cat $O/${T}_probe_*.json
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${T}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phases_ms"].items()})
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".err")).read()[-800:])
PY
