#!/bin/bash
# round 2: span kernel without the memset / with the 32-ary run-end search, small sort on: GPU suite + every config.
# usage: gpu_r2_ab3.sh TAG
T=${1:-r2p}; O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
for wl in cfg3 cfg1 cfg2 cfg4 cfg5; do
  env KGE_NOP=1 timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-rank --no-sub > $O/${T}_ab_${wl}.json 2> $O/${T}_ab_${wl}.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kge_small_sort|RadixSort|kge_emit|kge_fwd_bwd|kge_reduce|kge_span|kge_loss' -c 300 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_ncu_bench.log 2>&1
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${T}_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-32s flushed %.4f warm %.4f e2e %.4f" % (f.split("/")[-1], d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k: round(v, 4) for k, v in d["roofline"]["phases_ms"].items()}, round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-400:])
PY
