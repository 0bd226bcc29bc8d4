#!/bin/bash
# round 2: the load-list reduction for wide rows (kge_apply_wide.cu) -- GPU parity suite with it on, A/B lines against the
# warp-per-chunk kernel on every config, one ncu capture of it (source page exported here).  usage: gpu_r2_wide.sh TAG
T=${1:-r2w}; O=gpurun_out; mkdir -p $O; S=/tmp/ncu_$T; mkdir -p $S
( KGE_APPLY_WIDE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest_wide1.log
tail -3 $O/${T}_pytest_wide1.log
for wl in cfg3 cfg1 cfg2 cfg4 cfg5; do
  for kv in "KGE_APPLY_WIDE=0" "KGE_APPLY_WIDE=1" "KGE_APPLY_WIDE=1 KGE_WIDE_UB=4"; do
    tag=$(echo $kv | tr ' =' '__')
    env $kv timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-rank --no-sub > $O/${T}_ab_${wl}_${tag}.json 2> $O/${T}_ab_${wl}_${tag}.err
  done
done
KGE_APPLY_WIDE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kge_reduce_apply_wide' -s 4 -c 2 \
  -o $S/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_full.log 2>&1
ncu -i $S/prof.ncu-rep --page raw --csv > $O/${T}_prof_wide_raw.csv 2>/dev/null
ncu -i $S/prof.ncu-rep --page source --csv > $O/${T}_prof_wide_source.csv 2>/dev/null
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${T}_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-52s flushed %.4f warm %.4f e2e %.4f" % (f.split("/")[-1], d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k: round(v, 4) for k, v in d["roofline"]["phases_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-400:])
PY
