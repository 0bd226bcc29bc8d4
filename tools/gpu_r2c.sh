#!/bin/bash
# round 2, session C (1 GPU): slot-synchronous group reduction, look-ahead prefetch in the phase kernels, fp16-split ranking
O=gpurun_out; mkdir -p $O; T=r2c
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${T}_pytest_all.log
for W in 8 4 2; do
  timeout 300 python tools/dim_probe.py --workload cfg5 --world $W --steps 15 > $O/${T}_probe_cfg5_w$W.json 2> $O/${T}_probe_cfg5_w$W.err
done
KGE_APPLY_PREFETCH=0 timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 > $O/${T}_probe_cfg5_w8_nopf.json 2> $O/${T}_probe_cfg5_w8_nopf.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 8 --steps 30 > $O/${T}_probe_cfg3_w8.json 2> $O/${T}_probe_cfg3_w8.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
KGE_RANK_TF32=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu > $O/${T}_bench_cfg3_tf32.json 2> $O/${T}_bench_cfg3_tf32.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu --no-sub > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_dim_partial|kge_dim_backward|kge_reduce_apply_group' -s 9 -c 3 \
  -o $O/${T}_prof_dim python tools/dim_probe.py --workload cfg5 --world 8 --steps 3 > $O/${T}_ncu_full_dim.log 2>&1
tail -4 $O/${T}_pytest_all.log; cat $O/${T}_probe_*.json
python - <<PY
import json
for f in ["cfg3","cfg3_tf32","cfg5","cfg2"]:
    try:
        d=json.loads(open("$O/${T}_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phases_ms"].items()}, "rank", d.get("rank",{}).get("ms_per_step"), d.get("rank",{}).get("roofline",{}).get("frac"), d.get("rank",{}).get("mrr"))
    except Exception as e:
        print(f,"ERR",e); print(open("$O/${T}_bench_%s.err"%f).read()[-800:])
PY
