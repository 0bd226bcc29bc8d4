#!/bin/bash
# round 2, second ranking session (1 GPU): full GPU tier on the 8-epilogue-warp tc sweep, A/B of the epilogue width and the
# ring shape on one box, launch list + one --set full capture of the default, the default bench line.
# usage: gpu_r2_rank2.sh TAG
T=${1:-r2rb}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest.log
for kv in "KGE_RANK_EW=8 KGE_RANK_SW=64" "KGE_RANK_EW=4 KGE_RANK_SW=64" "KGE_RANK_EW=8 KGE_RANK_SW=128"; do
  tag=$(echo $kv | tr -d ' ' | sed 's/KGE_RANK_//g; s/=//g')
  env $kv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-sub --rank-steps 3 > $O/${T}_rank_$tag.json 2> $O/${T}_rank_$tag.err
done
( KGE_RANK_SW=128 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc" 2>&1 | tail -4 ) > $O/${T}_pytest_sw128.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kge_rank|kge_absmax|kge_f16' -c 24 --csv \
  --log-file $O/${T}_launches_rank.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_rank.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'kge_rank_tc_kernel' -s 1 -c 1 \
  -o $O/${T}_prof_rank python bench.py --steps 3 --warmup 3 --no-cpu --no-sub --rank-steps 1 > $O/${T}_ncu_fullr.log 2>&1
ncu -i $O/${T}_prof_rank.ncu-rep --page raw --csv > $O/${T}_prof_rank_raw.csv 2>/dev/null
python tools/ncu_traffic.py $T $O/traffic_session.json cfg3=$O/${T}_prof_rank.ncu-rep > $O/${T}_traffic.log 2>&1
cp $O/traffic_session.json $O/${T}_traffic.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
kill $SMI
python - <<PY
import glob, json
for f in sorted(glob.glob("$O/${T}_rank_*.json")) + ["$O/${T}_bench_default.json"]:
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
tail -3 $O/${T}_pytest.log; tail -2 $O/${T}_pytest_sw128.log
grep -h "kge_" $O/${T}_launches_rank.csv | awk -F'","' '{print $5, $NF}' | head -12
