#!/bin/bash
# round 2: span kernel with 8 partial rows in flight per warp + A/B knobs on the small configs.  usage: gpu_r2_ab2.sh TAG
T=${1:-r2v}; O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
for wl in cfg3 cfg1 cfg2 cfg4; do
  for kv in "KGE_NOP=1" "KGE_APPLY_SPLIT=4" "KGE_SPAN_WARP=1" "KGE_FWD_SPLIT=2" "KGE_FWD_SPLIT=2 KGE_FWD_MAXCTAS=4"; do
    tag=$(echo $kv | tr ' =' '__')
    env $kv timeout 200 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-rank --no-sub > $O/${T}_ab_${wl}_${tag}.json 2> $O/${T}_ab_${wl}_${tag}.err
  done
done
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${T}_ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-52s flushed %.4f warm %.4f e2e %.4f" % (f.split("/")[-1], d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k: round(v, 4) for k, v in d["roofline"]["phases_ms"].items()}, round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-400:])
PY
