#!/bin/bash
# Parity suite + A/B of the TransE distance sweep (second-generation kernel vs KGE_SWEEP_V1=1) on cfg1.
# usage: tools/gpu_ab_sweep.sh TAG
TAG=${1:-s}
O=gpurun_out; mkdir -p $O
( timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
if ! grep -q " passed" $O/${TAG}_pytest.log || grep -q "failed" $O/${TAG}_pytest.log; then
  ( KGE_SWEEP_V1=1 timeout 300 python -m pytest tests -m gpu -q -k "rank or Rank or subset or non_linearity" 2>&1 | tail -15 ) > $O/${TAG}_pytest_v1.log
  echo "--- with KGE_SWEEP_V1=1:"; tail -4 $O/${TAG}_pytest_v1.log
fi
for v in 0 1; do
  KGE_SWEEP_V1=$v timeout 200 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu --rank-steps 3 > $O/${TAG}_cfg1_v1is$v.json 2> $O/${TAG}_cfg1_v1is$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_cfg1_v1is$v.json").read().strip().splitlines()[-1]); r=d["rank"]
    print("cfg1 KGE_SWEEP_V1=$v rank %.3f M test triples/s, sweep %.3f ms, frac %.3f, mrr %.5f | train ms/step %.4f" % (r["value"]/1e6, r["roofline"]["kernel_ms"], r["roofline"]["frac"], r["mrr"], d["ms_per_step"]))
except Exception as e:
    print("cfg1 v1=$v ERR", e, open("$O/${TAG}_cfg1_v1is$v.err").read()[-800:])
PY
done
