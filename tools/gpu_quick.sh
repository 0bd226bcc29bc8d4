#!/bin/bash
# parity tests + default bench (+ optional extra bench args).  usage: tools/gpu_quick.sh TAG [workloads...]
TAG=${1:-q}; shift
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
for c in ${@:-cfg3}; do
  timeout 300 python bench.py --workload $c --steps 30 --warmup 5 --no-cpu --no-rank > $O/${TAG}_$c.json 2> $O/${TAG}_$c.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$c.json").read().strip().splitlines()[-1]); r=d["roofline"]["phases_ms"]
    print("$c ms/step %.4f warm %.4f e2e_ms %.4f"%(d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"]), {k:round(v,4) for k,v in r.items()})
except Exception as e:
    print("$c ERR", e, open("$O/${TAG}_$c.err").read()[-600:])
PY
done
