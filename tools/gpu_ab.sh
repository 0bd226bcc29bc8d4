#!/bin/bash
# A/B of the staged (bulk-copy) kernel variants + parity tests.  usage: tools/gpu_ab.sh TAG [workloads...]
TAG=${1:-ab}; shift
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
for c in ${@:-cfg3 cfg5 cfg1 cfg2 cfg4}; do
  for v in 10 00; do
    KGE_FWD_PIPE=${v:0:1} KGE_REDUCE_STAGED=${v:1:1} timeout 300 python bench.py --workload $c --steps 30 --warmup 5 --no-cpu --no-rank > $O/${TAG}_${c}_v$v.json 2> $O/${TAG}_${c}_v$v.err
    python - <<PY
import json
f="$O/${TAG}_${c}_v$v"
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1]); r=d["roofline"]["phases_ms"]
    print("$c pipe/staged=$v ms/step %.4f warm %.4f e2e_ms %.4f"%(d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k:round(v,4) for k,v in r.items()})
except Exception as e:
    print("$c $v ERR", e, open(f+".err").read()[-600:])
PY
  done
done
