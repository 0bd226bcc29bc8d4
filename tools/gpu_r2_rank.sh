#!/bin/bash
# round 2, ranking session (1 GPU): the full GPU tier on the new tc sweep (k-tail trim, flat tile split, fused |Q| maximum,
# vectorised operand split), the tc ranking tests again with 64-byte k-blocks (4-stage ring), an A/B of the two ring shapes,
# the launch lists of both, one --set full capture of the winner and the default bench line with it.
# usage: gpu_r2_rank.sh TAG
T=${1:-r2ra}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest.log
( KGE_RANK_SW=64 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_features.py -m gpu -x -q -k "tc or rank" 2>&1 | tail -8 ) > $O/${T}_pytest_sw64.log
for sw in 128 64; do
  KGE_RANK_SW=$sw timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-sub --rank-steps 3 > $O/${T}_rank_sw$sw.json 2> $O/${T}_rank_sw$sw.err
  KGE_RANK_SW=$sw timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kge_rank|kge_absmax|kge_f16' -c 24 --csv \
    --log-file $O/${T}_launches_rank_sw$sw.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_rank_sw$sw.log 2>&1
done
WIN=$(python - <<PY
import json
def ok(f, lg):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        good = d["rank_parity"]["sha1"] == "a4057acc1c272926"
        return d["rank"]["ms_per_step"] if good else 1e9
    except Exception:
        return 1e9
sw64_tests = "passed" in open("$O/${T}_pytest_sw64.log").read() and "failed" not in open("$O/${T}_pytest_sw64.log").read()
a, b = ok("$O/${T}_rank_sw128.json", 0), ok("$O/${T}_rank_sw64.json", 0)
print(64 if (sw64_tests and b < a) else 128)
PY
)
echo "winner KGE_RANK_SW=$WIN" > $O/${T}_winner.txt
export KGE_RANK_SW=$WIN
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'kge_rank_tc_kernel' -s 1 -c 1 \
  -o $O/${T}_prof_rank python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_fullr.log 2>&1
ncu -i $O/${T}_prof_rank.ncu-rep --page raw --csv > $O/${T}_prof_rank_raw.csv 2>/dev/null
python tools/ncu_traffic.py $T $O/traffic_session.json cfg3=$O/${T}_prof_rank.ncu-rep > $O/${T}_traffic.log 2>&1
cp $O/traffic_session.json $O/${T}_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
kill $SMI
python - <<PY
import json
for f in ("$O/${T}_rank_sw128.json", "$O/${T}_rank_sw64.json", "$O/${T}_bench_default.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["rank"]
        print(f.split("/")[-1], "train ms %.4f frac %.3f | rank ms %.4f sweep %.4f M/s %.2f frac %.3f e2e %.2f mrr %.9f" % (
            d["ms_per_step"], d["roofline"]["frac"], r["ms_per_step"], r["roofline"]["kernel_ms"], r["value"] / 1e6, r["roofline"]["frac"],
            r["e2e"]["value"] / 1e6, r["mrr"]), d["rank_parity"]["sha1"])
    except Exception as e:
        print(f, "ERR", e)
try:
    d = json.loads(open("$O/${T}_bench_default.json").read().strip().splitlines()[-1])
    print("cfg5", {k: d["cfg5"].get(k) for k in ("value", "ms_per_step", "rank_parity")})
    print("others", {k: (v.get("ms_per_step"), v.get("e2e", {}).get("ms_per_step")) for k, v in d["others"].items()})
except Exception as e:
    print("ERR", e)
PY
cat $O/${T}_winner.txt; tail -3 $O/${T}_pytest.log; tail -3 $O/${T}_pytest_sw64.log
grep -h "kge_" $O/${T}_launches_rank_sw128.csv | awk -F'","' '{print $5, $NF}' | head -12
grep -h "kge_" $O/${T}_launches_rank_sw64.csv | awk -F'","' '{print $5, $NF}' | head -12
