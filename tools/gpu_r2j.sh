#!/bin/bash
# round 2, session J (1 GPU): does the radix sort co-run with the phase kernels when they leave room on the SMs?
O=gpurun_out; mkdir -p $O; T=r2j
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 $EXTRA > $O/${T}_probe_$name.json 2> $O/${T}_probe_$name.err
}
EXTRA=""
run base X=1
run dim3 KGE_DIM_MAXCTAS=3
run dim4 KGE_DIM_MAXCTAS=4
run dim2 KGE_DIM_MAXCTAS=2
EXTRA="--pipeline 1"
run p_base X=1
run p_app5 KGE_APPLY_MAXCTAS=5
run p_app4 KGE_APPLY_MAXCTAS=4
run p_app6 KGE_APPLY_MAXCTAS=6
run p_app5_dim3 KGE_APPLY_MAXCTAS=5 KGE_DIM_MAXCTAS=3
EXTRA="--pipeline 1 --chunks 2"
run p_c2_app5 KGE_APPLY_MAXCTAS=5
for f in $O/${T}_probe_*.json; do echo -n "$(basename $f) "; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['phases_ms'].items()})" 2>/dev/null || tail -2 ${f%.json}.err; done
