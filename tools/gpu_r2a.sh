#!/bin/bash
# round 2, session A (1 GPU): parity of the dimension-sharded kernels + per-rank probes of the W-GPU step
O=gpurun_out; mkdir -p $O; T=r2a
( timeout 600 python -m pytest tests/test_gpu_dim_sharded.py -x -q 2>&1 | tail -25 ) > $O/${T}_pytest_dim.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${T}_pytest_all.log
for W in 8 4 2; do
  timeout 300 python tools/dim_probe.py --workload cfg5 --world $W --steps 15 > $O/${T}_probe_cfg5_w$W.json 2> $O/${T}_probe_cfg5_w$W.err
done
timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 --chunks 2 --pipeline 1 > $O/${T}_probe_cfg5_w8_c2p.json 2> $O/${T}_probe_cfg5_w8_c2p.err
KGE_APPLY_GROUP=0 timeout 300 python tools/dim_probe.py --workload cfg5 --world 8 --steps 15 > $O/${T}_probe_cfg5_w8_nogroup.json 2> $O/${T}_probe_cfg5_w8_nogroup.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 8 --steps 30 > $O/${T}_probe_cfg3_w8.json 2> $O/${T}_probe_cfg3_w8.err
timeout 300 python tools/dim_probe.py --workload cfg3 --world 2 --steps 30 > $O/${T}_probe_cfg3_w2.json 2> $O/${T}_probe_cfg3_w2.err
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu --no-rank > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
tail -5 $O/${T}_pytest_dim.log; tail -3 $O/${T}_pytest_all.log; cat $O/${T}_probe_*.json
