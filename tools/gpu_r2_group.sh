#!/bin/bash
# A/B of KGE_RANK_GROUP (CTAs that share a tile range of the tc ranking sweep): tc tests with 4, rank lines with 1 / 4 / 2
T=${1:-r2re}; O=gpurun_out; mkdir -p $O
( KGE_RANK_GROUP=4 timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc" 2>&1 | tail -4 ) > $O/${T}_pytest_g4.log
for g in 1 4 2; do
  KGE_RANK_GROUP=$g timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 3 > $O/${T}_rank_g$g.json 2> $O/${T}_rank_g$g.err
done
cat $O/${T}_pytest_g4.log
python - <<PY
import json
for g in (1, 4, 2):
    try:
        d = json.loads(open("$O/${T}_rank_g%d.json" % g).read().strip().splitlines()[-1]); r = d["rank"]
        print("G=%d rank ms %.4f sweep %.4f M/s %.2f frac %.3f" % (g, r["ms_per_step"], r["roofline"]["kernel_ms"], r["value"] / 1e6, r["roofline"]["frac"]), d["rank_parity"]["sha1"], r.get("clocks"))
    except Exception as e:
        print(g, "ERR", e)
PY
