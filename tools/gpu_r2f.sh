#!/bin/bash
# round 2, session F (2 GPUs): sharded parity tests + the multi-GPU bench line (cfg3 main record + cfg5 sub-record)
O=gpurun_out; mkdir -p $O; T=${1:-r2f}; N=${2:-2}
nvidia-smi -L > $O/${T}_gpus.txt
( timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -40 ) > $O/${T}_pytest_multi.log
tail -5 $O/${T}_pytest_multi.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err
tail -c 3000 $O/${T}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_n$N.json").read().strip().splitlines()[-1])
    print("cfg3 N=$N ms/step %.4f value %.3e e2e %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["phases_ms_max_over_ranks"], d["rank_parity"], d["first_step_loss"], d["first_step_loss_single_gpu"])
    print("rank", {k:d["rank"][k] for k in ("value","ms_per_step","mrr")})
    c=d.get("cfg5",{})
    print("cfg5", {k:c.get(k) for k in ("value","ms_per_step","ms_per_step_warm","phases_ms_max_over_ranks","rank_parity","first_step_loss","first_step_loss_single_gpu","error")}, c.get("e2e"))
except Exception as e:
    print("ERR", e)
PY
