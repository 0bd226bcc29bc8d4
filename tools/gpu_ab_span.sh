#!/bin/bash
# A/B of the span/hub reduction's warps per CTA (KGE_SPAN_WARPS=8|32) + the hub / parity tests with 32.
TAG=${1:-h}; shift
O=gpurun_out; mkdir -p $O
( KGE_SPAN_WARPS=32 timeout 300 python -m pytest tests -m gpu -q -k "hub or bench_shapes or pipelined or golden or optimizer or edge" 2>&1 | tail -8 ) > $O/${TAG}_pytest32.log
tail -3 $O/${TAG}_pytest32.log
for c in ${@:-cfg3}; do
 for v in 8 32; do
  KGE_SPAN_WARPS=$v timeout 200 python bench.py --workload $c --steps 50 --warmup 5 --no-cpu --no-rank > $O/${TAG}_${c}_span$v.json 2> $O/${TAG}_${c}_span$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_${c}_span$v.json").read().strip().splitlines()[-1]); ph=d["roofline"]["phases_ms"]
    print("$c KGE_SPAN_WARPS=$v ms/step flushed %.4f warm %.4f e2e %.4f | span_hub %.4f reduce_apply %.4f" % (d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"], ph["span_hub"], ph["reduce_apply"]))
except Exception as e:
    print("$c span=$v ERR", e, open("$O/${TAG}_${c}_span$v.err").read()[-800:])
PY
 done
done
