#!/bin/bash
# round 2, last 1-GPU call: smoke(), the default bench line and the reference arm on the same box with the final code
T=${1:-r2rc}; O=gpurun_out; mkdir -p $O
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $O/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
cat $O/${T}_smoke.log; tail -c 600 $O/${T}_bench_default.err
python - <<PY
import json
d = json.loads(open("$O/${T}_bench_default.json").read().strip().splitlines()[-1])
r = d["rank"]
print("train ms %.4f warm %.4f e2e %.4f frac %.3f" % (d["ms_per_step"], d["ms_per_step_warm"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), d["clocks"])
print("rank ms %.4f M/s %.2f e2e %.2f frac %.3f" % (r["ms_per_step"], r["value"] / 1e6, r["e2e"]["value"] / 1e6, r["roofline"]["frac"]), r.get("clocks"), d["rank_parity"]["sha1"])
print("cfg5", d["cfg5"]["ms_per_step"], d["cfg5"]["e2e"]["ms_per_step"], "others", {k: v.get("ms_per_step") for k, v in d["others"].items()})
print("cpu", d["cpu_baseline"]["value"], d["rank"]["cpu_baseline"]["value"])
f = json.loads(open("$O/${T}_bench_ref.json").read().strip().splitlines()[-1])
print("ref", f["value"], f["rank"]["value"], "same config:", f["config"] == d["config"])
PY
