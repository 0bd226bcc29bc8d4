#!/bin/bash
# round 2, session K (1 GPU): higher occupancy targets (partial 6 CTAs, group reduction 10 CTAs)
O=gpurun_out; mkdir -p $O; T=r2l
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/${T}_pytest_all.log
run() { name=$1; shift; env "$@" timeout 300 python tools/dim_probe.py --workload $WL --world $W --steps 15 $EXTRA > $O/${T}_probe_$name.json 2> $O/${T}_probe_$name.err; }
WL=cfg5; W=8; EXTRA=""; run w8 X=1
EXTRA="--pipeline 1"; run w8_p X=1
EXTRA="--pipeline 1 --chunks 4"; run w8_p_c4 X=1
W=4; EXTRA="--pipeline 1"; run w4_p X=1
W=2; run w2_p X=1
WL=cfg3; W=8; run cfg3_w8_p X=1
W=2; run cfg3_w2_p X=1
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank --no-sub > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
tail -3 $O/${T}_pytest_all.log
for f in $O/${T}_probe_*.json; do echo -n "$(basename $f) "; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['phases_ms'].items()})" 2>/dev/null || tail -2 ${f%.json}.err; done
python -c "
import json
d=json.loads(open('$O/${T}_bench_cfg5.json').read().strip().splitlines()[-1]); print('cfg5 N=1 ms/step', d['ms_per_step'], 'warm', d['ms_per_step_warm'], d['roofline']['phases_ms'])"
