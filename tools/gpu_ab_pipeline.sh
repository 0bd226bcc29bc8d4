#!/bin/bash
# Parity suite + A/B of the software-pipelined step (KGE_PIPELINE=1/0) on the given workloads (default cfg3).
TAG=${1:-p}; shift
O=gpurun_out; mkdir -p $O
( timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/${TAG}_pytest.log
tail -8 $O/${TAG}_pytest.log
for c in ${@:-cfg3}; do
 for v in 1 0; do
  KGE_PIPELINE=$v timeout 200 python bench.py --workload $c --steps 50 --warmup 5 --no-cpu --no-rank > $O/${TAG}_${c}_pipe$v.json 2> $O/${TAG}_${c}_pipe$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${c}_pipe$v.json").read().strip().splitlines()[-1])
    print("$c KGE_PIPELINE=$v ms/step flushed %.4f warm %.4f e2e %.4f sync %.4f" % (d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"], d["e2e"]["synchronous"]["ms_per_step"]))
except Exception as e:
    print("$c pipe=$v ERR", e, open("$O/${TAG}_${c}_pipe$v.err").read()[-800:])
PY
 done
done
