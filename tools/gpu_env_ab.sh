#!/bin/bash
# bench the training step under several settings of one env variable.  usage: tools/gpu_env_ab.sh TAG VAR "v1 v2 .." [workloads...]
TAG=$1; VAR=$2; VALS=$3; shift 3
O=gpurun_out; mkdir -p $O
for c in ${@:-cfg3}; do
  for v in $VALS; do
    env $VAR=$v timeout 300 python bench.py --workload $c --steps 30 --warmup 5 --no-cpu --no-rank > $O/${TAG}_${c}_$v.json 2> $O/${TAG}_${c}_$v.err
    python - <<PY
import json
f="$O/${TAG}_${c}_$v"
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1]); r=d["roofline"]["phases_ms"]
    print("$c $VAR=$v ms/step %.4f warm %.4f e2e_ms %.4f"%(d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k:round(v,4) for k,v in r.items()})
except Exception as e:
    print("$c $v ERR", e, open(f+".err").read()[-600:])
PY
  done
done
